#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (both arms), ncu launch list + full capture of the hot kernels.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref_c2.json
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
for wl in c3 c5 c4; do
  python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$wl.json')); print('$wl value %.1f e2e %.1f cpu %.3f' % (d['value'], d['e2e']['value'], d['cpu_baseline']['value']), d['kernel_ms'], 'grid frac %.3f scan frac %.3f' % (d['roofline']['frac'], d['roofline_scan']['frac']))"
done
if [ "${1:-}" = "noprof" ]; then exit 0; fi
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
