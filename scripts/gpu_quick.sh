#!/bin/bash
# quick GPU check: parity tests + kernel timings
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for wl in ${1:-c2 c3 c5 c4}; do
  python bench.py --workload $wl --steps 200 --warmup 5 --no-cpu-baseline --no-pcs > gpurun_out/q_$wl.json 2> gpurun_out/q_$wl.err || tail -5 gpurun_out/q_$wl.err
  python - <<PY
import json
d = json.load(open("gpurun_out/q_$wl.json"))
print("$wl", "value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v*1e3,1) for k,v in d["kernel_ms"].items()}, "grid frac %.3f scan frac %.3f" % (d["roofline"]["frac"], d["roofline_scan"]["frac"]))
PY
done
