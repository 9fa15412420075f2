#!/bin/bash
touch contextual_rs_b200/csrc/ocba_scan.cu
make -C contextual_rs_b200/csrc -j8 EXTRA="-DCRS_SCAN_EXPERIMENT=1" > /dev/null 2>&1 || echo build failed
echo "experiment (loads + argmax + sum only):"; python scripts/scan_run.py 65536 1000 f64 2>&1 | head -1; python scripts/scan_run.py 65536 1000 f32 2>&1 | head -1
touch contextual_rs_b200/csrc/ocba_scan.cu
make -C contextual_rs_b200/csrc -j8 > /dev/null 2>&1
echo "full:"; python scripts/scan_run.py 65536 1000 f64 2>&1 | head -1
