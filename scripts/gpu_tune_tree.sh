#!/bin/bash
# rebuild pcs_tree.cu with different pool parameters on the box and time the three PCS shapes
mkdir -p gpurun_out
export FLAT=0
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "inverse_normal or pcs_tree" 2>&1 | tail -2
run() {
  for a in "c3 f32" "c5 f32" "c4 f32 16384"; do timeout 300 python scripts/pcs_tree_run.py 65536 $a 2>&1 | tail -1 | awk '{print $1, $7, $8, $9}' | tr '\n' ' '; done; echo
}
echo "baseline (queue 96, min active 20)"; run
for v in "-DCRS_TREE_MIN_ACTIVE=12" "-DCRS_TREE_MIN_ACTIVE=28" "-DCRS_TREE_QUEUE=64" "-DCRS_TREE_QUEUE=160 -DCRS_TREE_MIN_ACTIVE=24" "-DCRS_TREE_MIN_ACTIVE=8"; do
  touch contextual_rs_b200/csrc/pcs_tree.cu
  make -C contextual_rs_b200/csrc EXTRA="$v" > gpurun_out/tune_build.log 2>&1 || { echo "build failed $v"; tail -3 gpurun_out/tune_build.log; continue; }
  echo "variant $v"; run
done
