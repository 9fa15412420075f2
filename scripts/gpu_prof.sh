#!/bin/bash
# usage: gpu_prof.sh <workload> <kernel-regex> <tag>
WL=${1:-c4}; RX=${2:-posterior_grid_kernel}; TAG=${3:-$WL}
mkdir -p gpurun_out
CMD="python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-flush"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$RX" -s ${SKIP:-8} -c ${COUNT:-1} -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$TAG.log
