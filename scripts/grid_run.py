"""Time crs_posterior_grid alone on a config (CUDA events, L2 flushed) and checksum the tables.
usage: grid_run.py [config] [dtype]   (env CRS_GRID_DIRECT=0/1 selects the kernel variant)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contextual_rs_b200 as crs
from contextual_rs_b200.synthetic import make_config
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
dt = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else torch.float64
desc, ctx = make_config(name)
model = crs.B200ModelListGP(desc, device="cuda:0", dtype=dt)
ctx_d = ctx.to("cuda:0", dt)
ws = model.private_workspace(ctx.shape[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
ms = []
for i in range(6):
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); model.posterior_tables(ctx_d, ws["mean"], ws["var"], arms=(0, model.num_arms)); e.record(); torch.cuda.synchronize()
    ms.append(s.elapsed_time(e))
print(f"{name} {dt} DIRECT={os.environ.get('CRS_GRID_DIRECT', 'auto')}: grid {min(ms[1:]):.3f} ms (mean {sum(ms[1:]) / 5:.3f}); "
      f"checksum mean {float(ws['mean'].double().sum()):.12e} var {float(ws['var'].double().sum()):.12e}")
