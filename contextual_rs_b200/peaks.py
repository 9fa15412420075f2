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
