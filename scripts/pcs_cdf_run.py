"""Time the cdf PCS kernels against the max-tree kernel on a config's posterior tables.
usage: pcs_cdf_run.py [S] [config] [dtype] [n_ctx_limit]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contextual_rs_b200 as crs
from contextual_rs_b200.generalized_pcs import pcs_counts
from contextual_rs_b200.synthetic import make_config
S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
name = sys.argv[2] if len(sys.argv) > 2 else "c3"
dt = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
lim = int(sys.argv[4]) if len(sys.argv) > 4 else 0
only = os.environ.get("ONLY")
desc, ctx = make_config(name)
if lim:
    ctx = ctx[:lim]
model = crs.B200ModelListGP(desc, device="cuda:0", dtype=dt)
p = model.posterior(ctx.to(dt))
mean, var = p.mean.contiguous(), p.variance.contiguous()
st = torch.zeros(8, dtype=torch.int64, device="cuda:0")
n = S * mean.shape[0]
res = {}
for sampler in ((only,) if only else ("cdf", "tree")):
    for _ in range(2):
        c = pcs_counts(model, mean, var, S, seed=1, stats=st, sampler=sampler)
    torch.cuda.synchronize()
    ms = []
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); c = pcs_counts(model, mean, var, S, seed=1, stats=st, sampler=sampler); e.record(); torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    res[sampler] = c.double() / S
    v = st.cpu().tolist()
    print(f"{name} {dt} S={S} n_c={mean.shape[0]} K={mean.shape[1]} {sampler}: {min(ms):.3f} ms ({n / min(ms) / 1e6:.1f} G samples/s), "
          f"stats {v[:5]}, pcs mean {float(res[sampler].mean()):.5f}")
if len(res) == 2:
    d = res["cdf"] - res["tree"]
    pq = res["tree"].clamp(1e-4, 1 - 1e-4)
    se = (2 * pq * (1 - pq) / S).sqrt()
    print(f"cdf - tree: mean diff {float(d.mean()):+.2e} (se {float(se.pow(2).sum().sqrt() / len(d)):.1e}), standardised rms {float((d / se).pow(2).mean().sqrt()):.3f} max {float((d / se).abs().max()):.2f}")
