#!/bin/bash
# r02a profiles: launch list of the default bench + ncu --set full of every kernel of the headline step.
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sub"
$CMD > gpurun_out/r02a_plain.json 2> gpurun_out/r02a_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02a_launches.csv $CMD > gpurun_out/r02a_ncu.log 2>&1
echo "launch list rc=$?"
RUN="python scripts/headline_run.py 1 c4"
$RUN > gpurun_out/r02a_headline_run.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'pcs_cdf|posterior_grid_kernel|gp_prepare|ocba_scan' -s 5 -c 5 -o gpurun_out/r02a_headline -f $RUN > gpurun_out/r02a_ncu_full.log 2>&1
echo "full rc=$?"; cat gpurun_out/r02a_headline_run.log
