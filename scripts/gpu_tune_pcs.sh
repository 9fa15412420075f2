#!/bin/bash
for cfg in "4 4096 3" "4 2048 4" "4 2048 5" "4 2048 6" "4 1024 6" "4 1024 8"; do
  set -- $cfg
  touch contextual_rs_b200/csrc/pcs_philox.cu
  make -C contextual_rs_b200/csrc -j8 EXTRA="-DCRS_PCS_ROUND=$1 -DCRS_PCS_CHUNK=$2 -DCRS_PCS_MIN_BLOCKS=$3" > /dev/null 2>&1 || echo build failed
  echo "round=$1 chunk=$2 minblocks=$3: $(python scripts/pcs_run.py 65536)"
done
