#!/bin/bash
# sweep the grid kernel's occupancy / ILP knobs on config 4 (and config 2)
for cfg in "4 4" "5 4" "6 4" "3 4" "4 2" "5 2" "6 2"; do
  set -- $cfg
  touch contextual_rs_b200/csrc/posterior_grid.cu
  make -C contextual_rs_b200/csrc -j16 EXTRA="-DCRS_GRID_MIN_BLOCKS=$1 -DCRS_PHASE_A_UNROLL=$2" > /tmp/mk.log 2>&1 || { echo build failed; tail -3 /tmp/mk.log; }
  r4=$(python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline --no-pcs | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3f ms' % d['kernel_ms']['posterior_grid'])")
  r2=$(python bench.py --workload c2 --steps 200 --warmup 5 --no-cpu-baseline --no-pcs | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f us' % (d['kernel_ms']['posterior_grid']*1e3))")
  echo "minblocks=$1 unroll=$2: c4 grid $r4, c2 grid $r2"
done
touch contextual_rs_b200/csrc/posterior_grid.cu
make -C contextual_rs_b200/csrc -j16 > /dev/null 2>&1
