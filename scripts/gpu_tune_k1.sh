#!/bin/bash
# K1 tuning sweep: rebuild with different settings on the GPU box and time the grid kernel.
mkdir -p gpurun_out
for cfg in "4 2" "4 4" "5 4" "3 4"; do
  set -- $cfg
  touch contextual_rs_b200/csrc/posterior_grid.cu
  make -C contextual_rs_b200/csrc -j8 EXTRA="-DCRS_GRID_MIN_BLOCKS=$1 -DCRS_PHASE_A_UNROLL=$2" > /dev/null 2>&1 || echo "build failed $cfg"
  for wl in c2 c5 c4; do
    python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t_$wl.json 2> gpurun_out/t_$wl.err || tail -3 gpurun_out/t_$wl.err
    python -c "
import json; d=json.load(open('gpurun_out/t_$wl.json')); print('minblocks $1 unroll $2 $wl grid_ms %.4f frac %.3f value %.1f' % (d['kernel_ms']['posterior_grid'], d['roofline']['frac'], d['value']))"
  done
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
