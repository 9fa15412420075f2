"""Diagnostic: error of the CUDA posterior grid against the fp64 oracle over ALL rows of a config."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import contextual_rs_b200 as crs
from contextual_rs_b200.synthetic import make_config
from oracle import gp_oracle
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
desc, ctx = make_config(name)
model = crs.B200ModelListGP(desc, device="cuda:0", dtype=torch.float64)
p = model.posterior(ctx)
mu, var = gp_oracle.posterior_grid_batched(desc, ctx, chunk=1024)
em = (p.mean.cpu() - mu).abs()
ev = (p.variance.cpu() - var).abs()
print("mean scale", float(mu.abs().max()), "var scale", float(var.max()))
print("mean err: max %.3e  99.9%% %.3e  median %.3e" % (float(em.max()), float(em.flatten().kthvalue(int(em.numel() * 0.999)).values), float(em.median())))
print("var  err: max %.3e  median %.3e" % (float(ev.max()), float(ev.median())))
flat = em.flatten().topk(10)
K = mu.shape[1]
for v, i in zip(flat.values.tolist(), flat.indices.tolist()):
    c, k = i // K, i % K
    X = desc["train_X"][k]
    dmin = float((X - ctx[c]).abs().sum(-1).min())
    print(f"  err {v:.3e} ctx {c} arm {k}: mean {float(mu[c, k]):+.4f} var {float(var[c, k]):.3e} noise {float(desc['noise'][k]):.2e} "
          f"os {float(desc['outputscale'][k]):.2f} min L1 dist to train {dmin:.2e} ystd {float(desc['train_Y'][k].std()):.3f}")
# per-arm worst error vs condition number of K^
worst_arm = em.max(dim=0).values
top = worst_arm.topk(5)
for v, k in zip(top.values.tolist(), top.indices.tolist()):
    m = gp_oracle.model_from_description(desc).models[k]
    Xs, L, alpha = m._prepare()
    Kh = L @ L.t()
    print(f"  arm {k}: worst mean err {v:.3e}, cond(K^) {float(torch.linalg.cond(Kh)):.3e}, |alpha|max {float(alpha.abs().max()):.3e}")
