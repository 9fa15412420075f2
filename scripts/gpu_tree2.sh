#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pcs" > gpurun_out/pytest_tree.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_tree.log
timeout 300 python scripts/pcs_tree_run.py 65536 c3 f32 2>&1 | tail -4
timeout 300 python scripts/pcs_tree_run.py 65536 c5 f32 2>&1 | tail -4
timeout 300 python scripts/pcs_tree_run.py 65536 c4 f32 16384 2>&1 | tail -4
timeout 600 python tests/tools/pcs_bias.py c5 192 22 2>&1 | tail -6
