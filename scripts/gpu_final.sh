#!/bin/bash
mkdir -p gpurun_out
for a in "65536 1000 f64" "65536 1000 f32" "16384 1000 f64" "16384 1000 f32" "32768 512 f64" "32768 700 f32" "32768 300 f64" "8192 264 f32"; do python scripts/scan_run.py $a 2>&1 | tail -2 | tr '\n' ' '; echo; done
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err; echo "ref rc=$?"
for wl in c3 c5 c4; do
  python bench.py --workload $wl --steps 20 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
done
python bench.py --workload loop --steps 10 --warmup 3 > gpurun_out/bench_loop.json 2> gpurun_out/bench_loop.err; echo "bench loop rc=$?"
