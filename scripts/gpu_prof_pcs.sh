#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/pcs_run.py 16384"
$CMD > gpurun_out/plain_pcs.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pcs_counts_kernel -s 3 -c 1 -o gpurun_out/prof_pcs -f $CMD > gpurun_out/ncu_pcs.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/plain_pcs.log | tail -1
