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


def measure_pcs_cdf_rates(device, iters: int = 2048, blocks: int = 148 * 8, reps: int = 4) -> dict:
    """Rates of the bare arithmetic of the cdf PCS kernels (``crs_pcs_counts_cdf``) on registers /
    shared memory — ``crs_peak_probe`` kind 8: one sampling call (Philox4x32-10, one Box-Muller pair,
    two cubic table evaluations, two integer compares = TWO samples) -> calls/s; kind 9: table terms of
    the build (one FMA, one ln Phi look-up in shared memory, one add) -> terms/s.  No global memory,
    queueing or control flow: the roofline denominators of those kernels."""
    lib = _lib.get()
    out = torch.zeros(2, dtype=torch.float32, device=device)
    rates = {}
    with torch.cuda.device(device):
        stream = _lib.stream_ptr(torch.device(device))
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        # kind 8 runs `blocks` CTAs of 256 threads; kind 9 (the build's 128 KB table) one 1,024-thread CTA per SM
        for name, kind, threads, per_iter in (("call", 8, blocks * 256.0, 1.0), ("term", 9, sms * 1024.0, 4.0)):
            best = 0.0
            for r in range(reps + 1):
                s = torch.cuda.Event(enable_timing=True)
                e = torch.cuda.Event(enable_timing=True)
                s.record()
                _lib.check(lib.crs_peak_probe(kind, iters, blocks, out.data_ptr(), stream), "crs_peak_probe")
                e.record()
                e.synchronize()
                if r:
                    best = max(best, threads * iters * per_iter / (s.elapsed_time(e) * 1e-3))
            rates[name] = best
    return rates


def pcs_cdf_roofline(kernel: str, work: float, seconds: float, n_gpus: int, rates: dict) -> dict:
    """Roofline entry of one of the two cdf kernels: ``work`` = Philox calls executed by
    ``pcs_cdf_sample_kernel`` (all ranks) or table terms evaluated by ``pcs_cdf_build_kernel``;
    peak = the bare sequence's measured rate per GPU."""
    key, unit = ("call", "G sampling calls/s (2 samples each)") if kernel == "pcs_cdf_sample_kernel" else ("term", "G table terms/s")
    achieved = work / seconds / n_gpus / 1e9 if seconds > 0 else 0.0
    peak = rates[key] / 1e9
    return {"kernel": kernel, "bound": "issue", "achieved": achieved, "peak": peak, "unit": unit,
            "frac": achieved / peak if peak > 0 else None, "traffic": None, "avg_launch_ms": seconds * 1e3,
            "work_per_launch_all_ranks": work,
            "peak_source": "measured in this run by crs_peak_probe kind %d: the kernel's bare per-%s instruction sequence "
                           "on registers / shared memory, no global memory or control flow (%.1f G/s per GPU)"
                           % (8 if key == "call" else 9, key, peak)}
