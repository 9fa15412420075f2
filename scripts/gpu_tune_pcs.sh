#!/bin/bash
for mb in 4 5 6; do
  touch contextual_rs_b200/csrc/pcs_philox.cu
  make -C contextual_rs_b200/csrc -j8 EXTRA="-DCRS_PCS_MIN_BLOCKS=$mb" > /dev/null 2>&1 || echo build failed
  for cfg in c3 c4; do echo "minblocks=$mb: $(python scripts/pcs_run.py 65536 $cfg | tail -1)"; done
done
touch contextual_rs_b200/csrc/pcs_philox.cu
make -C contextual_rs_b200/csrc -j8 > /dev/null 2>&1
