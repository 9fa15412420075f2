"""Live measurement of the pipe peaks used as roofline denominators for the compute-bound
kernels (FP64 / FP32 FMA): MEASURED_PEAKS.json only records HBM copy bandwidth and bf16 GEMM."""
from __future__ import annotations

import torch

from . import _lib


def measure_fma_peak(kind: int, device, iters: int = 4096, blocks: int = 148 * 8, reps: int = 5) -> float:
    """TFLOP/s of a pure FMA micro-kernel (kind 0 = fp64 FMA, 1 = fp32 FMA, 2 = fp64 mma.sync m8n8k4), best of ``reps``, CUDA events."""
    lib = _lib.get()
    out = torch.zeros(2, dtype=torch.float64, device=device)
    best = 0.0
    with torch.cuda.device(device):
        stream = _lib.stream_ptr(torch.device(device))
        for r in range(reps + 1):
            s = torch.cuda.Event(enable_timing=True)
            e = torch.cuda.Event(enable_timing=True)
            s.record()
            _lib.check(lib.crs_peak_probe(kind, iters, blocks, out.data_ptr(), stream), "crs_peak_probe")
            e.record()
            e.synchronize()
            if r:  # first launch is the warm-up
                per_thread = 16 * 512 / 32.0 if kind == 2 else 64 * 2  # flops per thread per iteration
                best = max(best, blocks * 256.0 * iters * per_thread / (s.elapsed_time(e) * 1e-3) / 1e12)
    return best


def measure_pcs_draw_peak(device, iters: int = 4096, blocks: int = 148 * 8, reps: int = 5) -> float:
    """N(0,1) draws/s of the bare PCS draw sequence (Philox4x32-10 + Box-Muller + 4 FMA + max on
    registers; ``crs_peak_probe`` kind 5) — the roofline denominator of ``crs_pcs_counts``."""
    lib = _lib.get()
    out = torch.zeros(2, dtype=torch.float32, device=device)
    best = 0.0
    with torch.cuda.device(device):
        stream = _lib.stream_ptr(torch.device(device))
        for r in range(reps + 1):
            s = torch.cuda.Event(enable_timing=True)
            e = torch.cuda.Event(enable_timing=True)
            s.record()
            _lib.check(lib.crs_peak_probe(5, iters, blocks, out.data_ptr(), stream), "crs_peak_probe")
            e.record()
            e.synchronize()
            if r:
                best = max(best, blocks * 256.0 * iters * 4 / (s.elapsed_time(e) * 1e-3))
    return best


def measure_pcs_tree_rates(device, iters: int = 2048, blocks: int = 148 * 8, reps: int = 4) -> dict:
    """Steps/s of the bare arithmetic of the max-tree PCS kernel (``crs_pcs_counts_tree``) on
    registers — ``crs_peak_probe`` kind 6: one root test (Philox4x32-10, Box-Muller pair, one
    inverse normal, two leaf tests, one node bound); kind 7: one node expansion (Philox4x32-10,
    three inverse normals, three leaf tests, three node bounds).  No memory, queueing or control
    flow: the roofline denominators of that kernel."""
    lib = _lib.get()
    out = torch.zeros(2, dtype=torch.float32, device=device)
    rates = {}
    with torch.cuda.device(device):
        stream = _lib.stream_ptr(torch.device(device))
        for name, kind in (("root", 6), ("expand", 7), ("direct", 5)):
            best = 0.0
            for r in range(reps + 1):
                s = torch.cuda.Event(enable_timing=True)
                e = torch.cuda.Event(enable_timing=True)
                s.record()
                _lib.check(lib.crs_peak_probe(kind, iters, blocks, out.data_ptr(), stream), "crs_peak_probe")
                e.record()
                e.synchronize()
                if r:
                    best = max(best, blocks * 256.0 * iters / (s.elapsed_time(e) * 1e-3))
            rates[name] = best
    return rates


def pcs_tree_roofline(roots: float, expansions: float, seconds: float, n_gpus: int, rates: dict,
                      directs: float = 0.0) -> dict:
    """Roofline entry of the tree kernel: the time the executed root tests, direct-arm tests and node
    expansions would take as bare register arithmetic (``measure_pcs_tree_rates``) over the measured time."""
    ideal = (roots / rates["root"] + expansions / rates["expand"] + directs / rates["direct"]) / n_gpus
    steps = roots + expansions + directs
    return {"kernel": "pcs_tree_kernel", "bound": "issue", "achieved": steps / seconds / n_gpus / 1e9,
            "peak": steps / ideal / n_gpus / 1e9, "unit": "G steps/s", "frac": ideal / seconds, "traffic": None,
            "root_tests": roots, "direct_arm_tests": directs, "node_expansions": expansions,
            "peak_source": "measured in this run by crs_peak_probe kinds 6 / 5 / 7: the bare root-test, direct-arm-test "
                           "(Philox + 2 Box-Muller + 4 FMA + max) and node-expansion sequences on registers "
                           "(%.1f / %.1f / %.1f G steps/s per GPU), weighted by the executed mix; achieved = executed "
                           "steps/s per GPU" % (rates["root"] / 1e9, rates["direct"] / 1e9, rates["expand"] / 1e9)}
