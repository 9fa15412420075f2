#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ocba_scan_tma_kernel -s 3 -c 1 -o gpurun_out/prof_scan64 -f python scripts/scan_run.py 65536 1000 f64 > gpurun_out/ncu_scan64.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:ocba_scan_tma_kernel -s 3 -c 1 -o gpurun_out/prof_scan32 -f python scripts/scan_run.py 65536 1000 f32 > gpurun_out/ncu_scan32.log 2>&1; echo rc=$?
