#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "recorded or transformed_train or posterior_and_decision or wsc_fixture or condition_on or sequential_loop or fit_modellist" > gpurun_out/pytest_trace.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_trace.log
