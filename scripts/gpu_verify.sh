#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_c2.json"))
print("value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), "pcs ms %.2f frac %.3f" % (d["pcs"]["ms_per_estimate"], d["pcs"]["roofline"]["frac"]), "headline %.2f" % d["headline"]["ms_per_step"])
print(d["pcs"].get("c1_joint_sampler"))
PY
