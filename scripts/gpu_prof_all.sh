#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-pcs"
$CMD > gpurun_out/plain_c2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_c2.log 2>&1
echo "ncu launches rc=$?"
CMDF="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pcs --no-flush"
$CMDF > gpurun_out/plain_c2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'posterior_grid_kernel|ocba_scan|gp_prepare' -s 9 -c 3 -o gpurun_out/prof_c2 -f $CMDF > gpurun_out/ncu_c2_full.log 2>&1
echo "ncu full c2 rc=$?"
CMD4="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-flush"
$CMD4 > gpurun_out/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'posterior_grid_kernel|ocba_scan' -s 6 -c 2 -o gpurun_out/prof_c4 -f $CMD4 > gpurun_out/ncu_c4_full.log 2>&1
echo "ncu full c4 rc=$?"
CMDP="python scripts/pcs_run.py 65536 c3"
$CMDP > gpurun_out/plain_pcs.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pcs_counts_kernel -s 2 -c 1 -o gpurun_out/prof_pcs -f $CMDP > gpurun_out/ncu_pcs.log 2>&1
echo "ncu pcs rc=$?"
