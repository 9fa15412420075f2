"""Interpolation bias of the cdf PCS stream's table: per context, the difference between the kernel's
tabulated F_c(z) (cubic on 125 cells + guard slots + exact sharp factors; oracle/pcs_cdf_ref.py restates
it) and the exact product prod_j Phi(a_j z + b_j), integrated against the standard normal density over the
sampler's range |z| <= 5.77 — i.e. the bias of the PCS(c) the cdf stream estimates, before any Monte-Carlo
noise.  Run on the oracle's posterior tables of a slice of every BASELINE configuration.

usage: python tests/tools/cdf_design.py [contexts per config = 48]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from contextual_rs_b200.synthetic import make_config  # noqa: E402
from oracle import gp_oracle, pcs_cdf_ref  # noqa: E402


def context_bias(mean_row, var_row, nodes=40001):
    plan = pcs_cdf_ref.context_plan(mean_row, var_row)
    z = np.linspace(-5.77, 5.77, nodes)
    w = np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi) * (z[1] - z[0])
    w[0] *= 0.5
    w[-1] *= 0.5
    tab = pcs_cdf_ref.success_probability(plan, z.astype(np.float32))
    exact = pcs_cdf_ref.exact_success_probability(plan, z.astype(np.float32))
    return float(((tab - exact) * w).sum()), float((exact * w).sum()), float(np.abs(tab - exact).max())


def main(n_ctx=48):
    worst = 0.0
    for name in ("c1", "c2", "c3", "c5", "c4"):
        desc, ctx = make_config(name)
        sel = np.linspace(0, ctx.shape[0] - 1, min(n_ctx, ctx.shape[0])).astype(int)
        mu, var = gp_oracle.posterior_grid_batched(desc, ctx[sel])
        for dt in (np.float64, np.float32):
            res = np.array([context_bias(mu[i].numpy().astype(dt), var[i].numpy().astype(dt)) for i in range(len(sel))])
            print(f"{name} {np.dtype(dt).name}: {len(sel)} contexts, PCS range [{res[:, 1].min():.4f}, {res[:, 1].max():.4f}]; "
                  f"bias of PCS(c): mean {res[:, 0].mean():+.2e}, max |.| {np.abs(res[:, 0]).max():.2e}; max pointwise |F_tab - F| {res[:, 2].max():.2e}")
            worst = max(worst, float(np.abs(res[:, 0]).max()))
    print(f"worst |bias| over all contexts: {worst:.2e}")
    return worst


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 48)
