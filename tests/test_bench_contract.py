"""The part of bench.py's contract that can be checked without a GPU: the reference arm prints
exactly ONE JSON line on stdout (libraries' banners go to stderr) with the required keys, honours
--steps / --warmup and prints the same `config` object the GPU arm prints for the same flags."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *flags],
                         capture_output=True, text=True, timeout=900, cwd=ROOT, env={**os.environ, "NCCL_DEBUG": "VERSION"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def _gpu_arm_config(name, world):
    sys.path.insert(0, ROOT)
    import importlib
    fd1 = os.dup(1)  # bench.py points fd 1 at stderr on import: undo it for pytest's capture
    try:
        bench = importlib.import_module("bench")
    finally:
        os.dup2(fd1, 1)
        os.close(fd1)
    args = argparse.Namespace(shard=False, partition="contexts", no_flush=False)
    return bench.workload_config(name, world, args)


def test_reference_arm_prints_one_json_line():
    d = _run("--steps", "2", "--warmup", "1")
    assert d["impl"] == "reference" and d["metric"] == "gp_c_ocba_decisions_per_sec" and d["unit"] == "decisions/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert (d["steps"], d["warmup"]) == (2, 1)  # the flags are honoured, not overridden
    assert d["config"]["workload"].startswith("c4 + PCS:") and d["data"] == "synthetic" and d["dtype"] == "f64"
    assert d["config"] == _gpu_arm_config("headline", 1)
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "decisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_config2():
    d = _run("--workload", "c2", "--steps", "3", "--warmup", "0")
    assert (d["steps"], d["warmup"]) == (3, 0) and d["config"] == _gpu_arm_config("c2", 1)
    assert d["config"]["workload"].startswith("c2:") and d["config"]["pcs_samples"] == 0
