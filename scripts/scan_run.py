"""Time the scan kernel alone on config-4-shaped random tables (no model needed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contextual_rs_b200 import _lib
lib = _lib.get()
n_c, K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dt = torch.float64 if (len(sys.argv) <= 3 or sys.argv[3] == "f64") else torch.float32
dev = "cuda:0"
mean = torch.randn(n_c, K, dtype=dt, device=dev)
var = torch.rand(n_c, K, dtype=dt, device=dev) + 0.1
out = {k: torch.zeros(n_c, dtype=torch.int32, device=dev) for k in ("best", "zarm", "ties")}
minz = torch.zeros(n_c, dtype=dt, device=dev)
dec = torch.zeros(64, dtype=torch.uint8, device=dev)
ws = torch.zeros(int(lib.crs_scan_workspace_bytes()), dtype=torch.uint8, device=dev)
def run():
    _lib.check(lib.crs_ocba_scan(mean.data_ptr(), var.data_ptr(), n_c, K, K, _lib.DTYPES[dt], 0, None, None, None,
                                 out["best"].data_ptr(), minz.data_ptr(), out["zarm"].data_ptr(), out["ties"].data_ptr(),
                                 dec.data_ptr(), ws.data_ptr(), None))
for _ in range(3): run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20): run()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 20
w = 8 if dt == torch.float64 else 4
print(f"scan n_c={n_c} K={K} {dt}: {ms*1e3:.1f} us, {2*n_c*K*w/ms/1e6:.0f} GB/s")
# correctness vs torch
best = mean.argmax(dim=-1)
assert torch.equal(out["best"].long(), best)
mb = mean.gather(1, best[:, None]); vb = var.gather(1, best[:, None])
z = (mb - mean).pow(2) / (vb + var); z.scatter_(1, best[:, None], float("inf"))
assert torch.equal(minz, z.min(dim=-1).values), "min zeta mismatch"
print("ok")
