"""Design aid: expansions per sample of the max-tree walk under different node bounds."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from scipy.special import ndtri_exp
from contextual_rs_b200.synthetic import make_config
from oracle.gp_oracle import posterior_grid_batched

cfgname = sys.argv[1]; n_ctx = int(sys.argv[2]); S = int(sys.argv[3])
BOUND = sys.argv[4]  # simple | chord
T = int(sys.argv[5])  # top-threat arms drawn directly at the root
A = 4
desc, ctx = make_config(cfgname)
sel = torch.linspace(0, ctx.shape[0] - 1, n_ctx).long()
mean, var = posterior_grid_batched(desc, ctx[sel])
mean, var = mean.numpy(), var.numpy()
rng = np.random.default_rng(1)
STEP = float(os.environ.get("STEP", "1")); ZG = np.arange(0, 7.0 + STEP, STEP)
pre_pass = 0; pre_tot = 0; exp_nophx = 0; tot_exp = 0; tot_n = 0; tot_succ = 0; exp_succ = 0; exp_fail = 0; root_fail = 0
for c in range(n_ctx):
    mu, sd = mean[c], np.sqrt(var[c])
    K = len(mu); b = int(np.argmax(mu)); mu = mu - mu[b]
    threat = mu / np.sqrt(sd ** 2 + sd[b] ** 2); threat[b] = -np.inf
    order = np.argsort(-threat)[:-1]
    top, rest = order[:T], order[T:]
    m, s = mu[rest], sd[rest]
    Kk = len(m)
    depth = max(1, int(np.ceil(np.log(Kk) / np.log(A) - 1e-9))); NL = A ** depth
    mt = np.concatenate([m, np.full(NL - Kk, -np.inf)]); st = np.concatenate([s, np.zeros(NL - Kk)])
    mumax, sdmax, env = {depth: mt}, {depth: st}, {}
    envl = mt[:, None] + st[:, None] * ZG[None, :]
    env[depth] = envl
    for l in range(depth - 1, -1, -1):
        mumax[l] = mumax[l + 1].reshape(-1, A).max(1); sdmax[l] = sdmax[l + 1].reshape(-1, A).max(1)
        env[l] = env[l + 1].reshape(-1, A, len(ZG)).max(1)
    def node_bound(l, n, z):
        if z < 0: return mumax[l][n]
        if BOUND == "simple": return mumax[l][n] + sdmax[l][n] * z
        k = min(int(z / STEP), len(ZG) - 2); f = z / STEP - k
        return env[l][n, k] * (1 - f) + env[l][n, k + 1] * f
    for si in range(S):
        yb = sd[b] * rng.standard_normal()
        fail = bool((mu[top] + sd[top] * rng.standard_normal(len(top)) >= yb).any()) if T else False
        L = np.log(rng.random()) / NL; J = rng.integers(NL); z = ndtri_exp(L)
        cnt = 0
        fail = fail or (mt[J] + st[J] * z >= yb)
        if fail: root_fail += 1
        stack = []
        if not fail and node_bound(0, 0, z) >= yb: stack.append((0, 0, L, J, z))
        while stack and not fail:
            l, n, L, J, z = stack.pop(); cnt += 1
            mch = A ** (depth - l - 1); jpath = (J // mch) % A; new = []
            npass = 0
            for k in range(A):
                cn = n * A + k
                if k != jpath:
                    pre_tot += 1
                    pv = node_bound(l + 1, cn, z) if l + 1 < depth else (mt[cn] + st[cn] * z)
                    if pv >= yb: pre_pass += 1; npass += 1
            if npass == 0: exp_nophx += 1
            for k in range(A):
                cn = n * A + k
                if k == jpath: Lc, Jc, zc = L, J, z
                else:
                    Lc = L + np.log(rng.random()) / mch; Jc = cn * mch + rng.integers(mch); zc = ndtri_exp(Lc)
                    if mt[Jc] + st[Jc] * zc >= yb: fail = True; break
                if l + 1 < depth and node_bound(l + 1, cn, zc) >= yb: new.append((l + 1, cn, Lc, Jc, zc))
            stack.extend(reversed(new))
        tot_exp += cnt; tot_n += 1
        if fail: exp_fail += cnt
        else: tot_succ += 1; exp_succ += cnt
print(f"pretest: {pre_pass/max(pre_tot,1):.3f} of non-path children pass; {pre_pass/max(tot_exp,1):.2f} per expansion; expansions without any passing child {exp_nophx/max(tot_exp,1):.3f}")
print(f"{cfgname} bound={BOUND} T={T}: expansions/sample {tot_exp/tot_n:.2f} | PCS {tot_succ/tot_n:.4f} | per success {exp_succ/max(tot_succ,1):.1f} | per failure {exp_fail/max(tot_n-tot_succ,1):.2f} | failed at root {root_fail/max(tot_n-tot_succ,1):.3f}")
