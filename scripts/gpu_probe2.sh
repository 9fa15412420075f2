#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "pcs or PCS" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for cfg in c3 c5 c4; do python scripts/pcs_run.py 65536 $cfg 2>&1 | tail -1; done
python scripts/pcs_run.py 8192 c4 2>&1 | tail -1
python scripts/pcs_run.py 1024 c5 2>&1 | tail -1
if [ "$1" = "ncu" ]; then
CMDP="python scripts/pcs_run.py 65536 c4"
ncu --set full --clock-control none --import-source on -k regex:pcs_counts_kernel -s 2 -c 1 -o gpurun_out/prof_pcs4 -f $CMDP > gpurun_out/ncu_pcs4.log 2>&1
echo "ncu pcs rc=$?"
fi
