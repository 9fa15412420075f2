"""Run the PCS kernel alone (default: config 3, 256 arms x 8,192 contexts, fp32) — profiling target.
usage: pcs_run.py [S] [config] [dtype]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contextual_rs_b200 as crs
from contextual_rs_b200.generalized_pcs import pcs_counts
from contextual_rs_b200.synthetic import make_config
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
name = sys.argv[2] if len(sys.argv) > 2 else "c3"
dt = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
desc, ctx = make_config(name)
model = crs.B200ModelListGP(desc, device="cuda:0", dtype=dt)
p = model.posterior(ctx.to(dt))
mean, var = p.mean.contiguous(), p.variance.contiguous()
st = torch.zeros(4, dtype=torch.int64, device="cuda:0")
for _ in range(2):
    c = pcs_counts(model, mean, var, S, seed=1, stats=st)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); c = pcs_counts(model, mean, var, S, seed=1, stats=st); e.record(); torch.cuda.synchronize()
d, k = st.cpu().tolist()[:2]
print(f"{name} {dt} S={S}: {s.elapsed_time(e):.3f} ms, executed {d:.3e} draws ({d/(d+k):.4f}), {d/s.elapsed_time(e)/1e6:.1f} G draws/s, "
      f"pcs mean {float(c.float().mean())/S:.4f} min {float(c.min())/S:.4f}")
