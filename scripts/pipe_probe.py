"""Are the FP64 FMA and FP64 MMA (DMMA) paths separate pipes?  Times crs_peak_probe kinds 0 (DFMA only),
2 (DMMA only) and 10 (even warps DFMA, odd warps DMMA, same per-warp work) with identical launch shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contextual_rs_b200 import _lib
lib = _lib.get()
out = torch.zeros(2, dtype=torch.float64, device="cuda:0")
st = _lib.stream_ptr(torch.device("cuda:0"))
for kind, name in ((0, "DFMA all warps"), (2, "DMMA all warps"), (10, "DFMA even / DMMA odd warps")):
    best = 1e9
    for r in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); _lib.check(lib.crs_peak_probe(kind, 4096, 148 * 8, out.data_ptr(), st)); e.record(); torch.cuda.synchronize()
        if r: best = min(best, s.elapsed_time(e))
    print(f"kind {kind:2d} {name}: {best:.3f} ms")
