"""CPU simulation of the max-tree sampler with arg-max-leaf shortcut (design aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.special import ndtri_exp
from contextual_rs_b200.synthetic import make_config
from oracle.gp_oracle import posterior_grid_batched
from oracle.pcs_oracle import pcs_exact_quadrature

cfgname = sys.argv[1] if len(sys.argv) > 1 else "c4"
n_ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 6
S = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
A = int(sys.argv[4]) if len(sys.argv) > 4 else 4
desc, ctx = make_config(cfgname)
sel = torch.linspace(0, ctx.shape[0] - 1, n_ctx).long()
mean, var = posterior_grid_batched(desc, ctx[sel])
mean, var = mean.numpy(), var.numpy()
quad = pcs_exact_quadrature(mean, var)
rng = np.random.default_rng(1)
T = dict(exp=0, n=0, succ_exp=0, nsucc=0, fail_exp=0)
for c in range(n_ctx):
    mu, sd = mean[c], np.sqrt(var[c])
    K = len(mu)
    b = int(np.argmax(mu))
    mu = mu - mu[b]
    threat = mu / np.sqrt(sd ** 2 + sd[b] ** 2)
    threat[b] = -np.inf
    order = np.argsort(-threat)[:-1] if os.environ.get("SORT", "1") == "1" else np.delete(np.arange(K), b)
    m, s = mu[order], sd[order]
    Kk = len(m)
    depth = max(1, int(np.ceil(np.log(Kk) / np.log(A) - 1e-9)))
    NL = A ** depth
    mt = np.concatenate([m, np.full(NL - Kk, -np.inf)]); st = np.concatenate([s, np.zeros(NL - Kk)])
    mumax, sdmax, sdmin = {depth: mt}, {depth: st}, {depth: np.where(mt == -np.inf, np.inf, st)}
    for l in range(depth - 1, -1, -1):
        mumax[l] = mumax[l + 1].reshape(-1, A).max(1)
        sdmax[l] = sdmax[l + 1].reshape(-1, A).max(1)
        sdmin[l] = sdmin[l + 1].reshape(-1, A).min(1)
    def node_bound(l, n, z):
        sg = sdmax[l][n] if z >= 0 else (sdmin[l][n] if np.isfinite(sdmin[l][n]) else 0.0)
        return mumax[l][n] + sg * z
    nexp = np.zeros(S, int); succ = np.zeros(S, bool)
    for si in range(S):
        yb = sd[b] * rng.standard_normal()
        L = np.log(rng.random()) / NL
        J = rng.integers(NL)
        z = ndtri_exp(L)
        cnt = 0
        fail = mt[J] + st[J] * z >= yb
        stack = []
        if not fail and node_bound(0, 0, z) >= yb:
            stack.append((0, 0, L, J, z))
        while stack and not fail:
            l, n, L, J, z = stack.pop()
            cnt += 1
            mch = A ** (depth - l - 1)
            jpath = (J // mch) % A
            new = []
            for k in range(A):
                cn = n * A + k
                if k == jpath:
                    Lc, Jc, zc = L, J, z
                else:
                    Lc = L + np.log(rng.random()) / mch
                    Jc = cn * mch + rng.integers(mch)
                    zc = ndtri_exp(Lc)
                    if mt[Jc] + st[Jc] * zc >= yb:
                        fail = True
                        break
                if l + 1 < depth and node_bound(l + 1, cn, zc) >= yb:
                    new.append((l + 1, cn, Lc, Jc, zc))
            stack.extend(reversed(new))
        nexp[si] = cnt; succ[si] = not fail
    se = np.sqrt(max(quad[c] * (1 - quad[c]), 1e-9) / S)
    print(f"ctx {c}: PCS tree {succ.mean():.4f} quad {quad[c]:.4f} (dev {(succ.mean()-quad[c])/se:+.1f} se) | expansions/sample all {nexp.mean():.2f} succ {nexp[succ].mean() if succ.any() else 0:.1f} fail {nexp[~succ].mean():.2f} | fail at root {(nexp[~succ]==0).mean():.3f}")
    T["exp"] += nexp.sum(); T["n"] += S
print("mean Philox calls per sample (1 root + expansions): %.2f" % (1 + T["exp"] / T["n"]))
