#!/bin/bash
mkdir -p gpurun_out
export FLAT=0
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "pcs" 2>&1 | tail -2
run() {
  for a in "c3 f32" "c5 f32" "c4 f32 16384" "c4 f64 16384"; do timeout 300 python scripts/pcs_tree_run.py 65536 $a 2>&1 | tail -1 | awk '{print $1, $7, $8}' | tr '\n' ' '; done; echo
}
echo "default (min active 28)"; run
for v in "-DCRS_TREE_MIN_ACTIVE=32"; do
  touch contextual_rs_b200/csrc/pcs_tree.cu
  make -C contextual_rs_b200/csrc EXTRA="$v" > gpurun_out/tune_build.log 2>&1 || { echo "build failed $v"; tail -3 gpurun_out/tune_build.log; continue; }
  echo "variant $v"; run
done
