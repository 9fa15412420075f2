#!/bin/bash
mkdir -p gpurun_out
for a in "65536 1000 f64" "65536 1000 f32" "16384 1000 f64" "16384 1000 f32" "32768 512 f64" "32768 700 f32" "32768 700 f64" "32768 300 f64" "32768 1024 f64" "8192 264 f32"; do python scripts/scan_run.py $a 2>&1 | tail -2 | tr '\n' ' '; echo; done
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
