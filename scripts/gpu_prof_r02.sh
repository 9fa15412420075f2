#!/bin/bash
# Round-2 profile set: launch list of the default bench command + ncu --set full of every kernel of the
# headline step (config 4 decision + 2^16-sample PCS on one GPU).  Each ncu pass runs only after the same
# command exited 0 without ncu.  usage: bash scripts/gpu_prof_r02.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sub"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "launch list rc=$?"
RUN="python scripts/headline_run.py 1 c4"
$RUN > gpurun_out/${TAG}_headline_run.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'pcs_cdf|posterior_grid_kernel|gp_prepare|ocba_scan' -s 5 -c 5 -o gpurun_out/${TAG}_headline -f $RUN > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full rc=$?"; cat gpurun_out/${TAG}_headline_run.log
