#!/bin/bash
mkdir -p gpurun_out
CMD4="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-flush"
$CMD4 > gpurun_out/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'posterior_grid_kernel|ocba_scan_kernel' -s 8 -c 2 -o gpurun_out/prof_c4 -f $CMD4 > gpurun_out/ncu_c4_full.log 2>&1
echo "ncu full c4 rc=$?"; tail -3 gpurun_out/ncu_c4_full.log
