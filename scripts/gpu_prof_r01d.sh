#!/bin/bash
# r01d profiling round: launch list of the default bench + full ncu captures of every hot kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_c2.json 2> gpurun_out/plain_c2.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv $CMD > gpurun_out/ncu_c2.log 2>&1
echo "ncu launches rc=$?"
CMDF="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pcs --no-flush"
$CMDF > gpurun_out/plain_c2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'posterior_grid_kernel|ocba_scan|gp_prepare' -s 9 -c 3 -o gpurun_out/prof_c2 -f $CMDF > gpurun_out/ncu_c2_full.log 2>&1
echo "ncu full c2 rc=$?"
CMD4="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-flush --no-pcs"
$CMD4 > gpurun_out/plain_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'posterior_grid_kernel|ocba_scan' -s 6 -c 2 -o gpurun_out/prof_c4 -f $CMD4 > gpurun_out/ncu_c4_full.log 2>&1
echo "ncu full c4 rc=$?"
export FLAT=0
CMDP="python scripts/pcs_tree_run.py 65536 c3 f32"
$CMDP > gpurun_out/plain_tree_c3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pcs_tree_kernel -s 2 -c 1 -o gpurun_out/prof_tree_c3 -f $CMDP > gpurun_out/ncu_tree_c3.log 2>&1
echo "ncu tree c3 rc=$?"; tail -1 gpurun_out/plain_tree_c3.log
CMDQ="python scripts/pcs_tree_run.py 65536 c4 f32 16384"
$CMDQ > gpurun_out/plain_tree_c4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pcs_tree_kernel -s 2 -c 1 -o gpurun_out/prof_tree_c4 -f $CMDQ > gpurun_out/ncu_tree_c4.log 2>&1
echo "ncu tree c4 rc=$?"; tail -1 gpurun_out/plain_tree_c4.log
for wl in c3 c5 c4; do
  python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
done
python bench.py --workload loop --steps 10 --warmup 3 > gpurun_out/bench_loop.json 2> gpurun_out/bench_loop.err; echo "bench loop rc=$?"
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err; echo "ref rc=$?"
