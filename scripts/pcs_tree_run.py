"""Time the max-tree PCS kernel against the flat one on a config's posterior tables.
usage: pcs_tree_run.py [S] [config] [dtype] [n_ctx_limit]"""
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
flat = os.environ.get("FLAT", "1") == "1"
desc, ctx = make_config(name)
if lim:
    ctx = ctx[:lim]
model = crs.B200ModelListGP(desc, device="cuda:0", dtype=dt)
p = model.posterior(ctx.to(dt))
mean, var = p.mean.contiguous(), p.variance.contiguous()
st = torch.zeros(4, dtype=torch.int64, device="cuda:0")
res = {}
for sampler in (("tree", "flat") if flat else ("tree",)):
    for _ in range(2):
        c = pcs_counts(model, mean, var, S, seed=1, stats=st, sampler=sampler)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); c = pcs_counts(model, mean, var, S, seed=1, stats=st, sampler=sampler); e.record(); torch.cuda.synchronize()
    a, b, dct = st.cpu().tolist()[:3]
    ms = s.elapsed_time(e)
    res[sampler] = c.double() / S
    n = S * mean.shape[0]
    if sampler == "tree":
        print(f"{name} {dt} S={S} n_c={mean.shape[0]} K={mean.shape[1]} tree: {ms:.3f} ms, roots {a:.3e} direct {dct:.3e} expansions {b:.3e} ({b/a:.3f}/sample), "
              f"{n/ms/1e6:.2f} G samples/s, pcs mean {float(res[sampler].mean()):.5f}")
    else:
        print(f"{name} {dt} S={S} flat: {ms:.3f} ms, executed {a:.3e} draws ({a/4/n:.2f} calls/sample), pcs mean {float(res[sampler].mean()):.5f}")
if flat:
    d = res["tree"] - res["flat"]
    pq = res["flat"].clamp(1e-4, 1 - 1e-4)
    se = (2 * pq * (1 - pq) / S).sqrt()
    print(f"tree - flat: mean diff {float(d.mean()):+.2e} (se {float(se.pow(2).sum().sqrt() / len(d)):.1e}), standardised rms {float((d / se).pow(2).mean().sqrt()):.3f} max {float((d / se).abs().max()):.2f}")
