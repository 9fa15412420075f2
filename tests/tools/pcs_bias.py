"""Bias check of the two PCS streams against exact quadrature: many samples on a few hundred contexts.
usage: pcs_bias.py [config] [n_ctx] [log2 S]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import contextual_rs_b200 as crs
from contextual_rs_b200.generalized_pcs import pcs_counts
from contextual_rs_b200.synthetic import make_config
from oracle.pcs_oracle import pcs_exact_quadrature
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
n_ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 256
S = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 22)
desc, ctx = make_config(name)
sel = torch.linspace(0, ctx.shape[0] - 1, n_ctx).long()
model = crs.B200ModelListGP(desc, device="cuda:0", dtype=torch.float32)
p = model.posterior(ctx[sel].to(torch.float32))
mean, var = p.mean.contiguous(), p.variance.contiguous()
t0 = time.time()
exact = pcs_exact_quadrature(mean.double().cpu().numpy(), var.double().cpu().numpy())
print(f"quadrature {time.time() - t0:.1f} s; mean PCS {exact.mean():.6f}")
se = np.sqrt(np.maximum(exact * (1 - exact), 1e-12) / S)
for sampler in ("tree", "flat"):
    for seed in (1, 2):
        c = pcs_counts(model, mean, var, S, seed=seed, sampler=sampler).cpu().numpy() / S
        d = c - exact
        print(f"{name} {sampler} seed {seed}: mean diff {d.mean():+.3e} (se {np.sqrt((se ** 2).sum()) / n_ctx:.2e}) standardised mean {np.mean(d / se):+.3f} (se {1 / np.sqrt(n_ctx):.3f}) rms {np.sqrt(np.mean((d / se) ** 2)):.3f}")
