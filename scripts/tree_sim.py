"""CPU simulation: cost (Philox calls per sample) of the flat threat-ordered PCS sampler vs a
lazy max-tree sampler, on config-4/5-shaped posteriors.  Design aid only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.special import ndtri_exp, ndtri
from contextual_rs_b200.synthetic import make_config
from oracle.gp_oracle import posterior_grid_batched

cfgname = sys.argv[1] if len(sys.argv) > 1 else "c4"
n_ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 6
S = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
ARITY = int(sys.argv[4]) if len(sys.argv) > 4 else 4
desc, ctx = make_config(cfgname)
sel = torch.linspace(0, ctx.shape[0] - 1, n_ctx).long()
mean, var = posterior_grid_batched(desc, ctx[sel])
mean, var = mean.numpy(), var.numpy()
rng = np.random.default_rng(0)
ZMAX = 5.7683
tot = dict(flat=0, tree=0, tree_succ=0, flat_succ=0, nsucc=0, n=0)
for c in range(n_ctx):
    mu, sd = mean[c], np.sqrt(var[c])
    K = len(mu)
    b = int(np.argmax(mu))
    mu = mu - mu[b]
    threat = mu / np.sqrt(sd ** 2 + sd[b] ** 2)
    threat[b] = -np.inf
    order = np.argsort(-threat)  # b last
    order = order[:-1]
    m, s = mu[order], sd[order]
    # screening
    keep = m + ZMAX * s >= -ZMAX * sd[b]
    m, s = m[keep], s[keep]
    Kk = len(m)
    zb = rng.standard_normal(S)
    yb = sd[b] * zb
    # ---- flat: blocks of 4 in threat order; stop at first failure; success bound by suffix max of m + ZMAX s
    nb = (Kk + 3) // 4
    pad = nb * 4 - Kk
    mp = np.concatenate([m, np.full(pad, -np.inf)]); sp = np.concatenate([s, np.ones(pad)])
    ub = (mp + ZMAX * sp).reshape(nb, 4).max(1)
    suf = np.maximum.accumulate(ub[::-1])[::-1]  # max over blocks >= i
    z = rng.standard_normal((S, nb * 4))
    y = mp + sp * z
    fail_blk = (y.reshape(S, nb, 4).max(2) >= yb[:, None])
    first_fail = np.where(fail_blk.any(1), fail_blk.argmax(1), nb)
    # bound: stop before block i if suf[i] < yb
    bound_stop = np.array([np.searchsorted(-suf, -ybi, side="left") for ybi in yb])  # first i with suf[i] < yb -> since suf decreasing
    # suf is non-increasing; first index where suf[i] < yb:
    bound_stop = np.array([int(np.argmax(suf < ybi)) if (suf < ybi).any() else nb for ybi in yb])
    visited = np.minimum(first_fail + 1, bound_stop)
    visited = np.minimum(visited, nb)
    succ = first_fail >= bound_stop
    succ |= (first_fail == nb)
    flat_calls = visited + 1  # + own draw
    # ---- tree
    depth = int(np.ceil(np.log(max(Kk, 2)) / np.log(ARITY)))
    L_leaves = ARITY ** depth
    mt = np.concatenate([m, np.full(L_leaves - Kk, -np.inf)]); st = np.concatenate([s, np.full(L_leaves - Kk, np.nan)])
    # per level node stats
    mumax, sdmax, sdmin = [None] * (depth + 1), [None] * (depth + 1), [None] * (depth + 1)
    mumax[depth] = mt; sdmax[depth] = np.where(np.isnan(st), 0, st); sdmin[depth] = np.where(np.isnan(st), np.inf, st)
    for l in range(depth - 1, -1, -1):
        mumax[l] = mumax[l + 1].reshape(-1, ARITY).max(1)
        sdmax[l] = sdmax[l + 1].reshape(-1, ARITY).max(1)
        sdmin[l] = sdmin[l + 1].reshape(-1, ARITY).min(1)
    # generate full tree of log-CDF values
    Ls = [np.log(rng.random((S, 1))) / L_leaves]
    for l in range(depth):
        nn = ARITY ** l
        mchild = ARITY ** (depth - l - 1)
        U = rng.random((S, nn, ARITY))
        arg = rng.integers(0, ARITY, size=(S, nn))
        child = Ls[l][:, :, None] + np.log(U) / mchild
        child[np.arange(S)[:, None], np.arange(nn)[None, :], arg] = Ls[l]
        Ls.append(child.reshape(S, nn * ARITY))
    def bound(l):
        zz = ndtri_exp(Ls[l])
        sg = np.where(zz >= 0, sdmax[l][None, :], np.where(np.isinf(sdmin[l]), 0, sdmin[l])[None, :])
        return mumax[l][None, :] + sg * zz
    live = [None] * (depth + 1)
    live[0] = bound(0) >= yb[:, None]
    for l in range(1, depth + 1):
        live[l] = np.repeat(live[l - 1], ARITY, axis=1) & (bound(l) >= yb[:, None])
    tree_succ = ~live[depth].any(1)
    # expansions = live internal nodes (levels 0..depth-1); upper bound for failing samples
    expansions_all = sum(live[l].sum(1) for l in range(depth))
    # DFS for failing samples: children in order, stop at first failing leaf. exact count via python
    exp_dfs = expansions_all.copy()
    fail_idx = np.where(~tree_succ)[0]
    for si in fail_idx[:600]:
        cnt = 0
        stack = [(0, 0)]
        found = False
        while stack and not found:
            l, n = stack.pop()
            if not live[l][si, n]:
                continue
            if l == depth:
                found = True
                break
            cnt += 1
            ch = [(l + 1, n * ARITY + k) for k in range(ARITY)]
            # visit highest-bound child first: push in reverse order of preference; here: largest L first
            ch.sort(key=lambda t: Ls[t[0]][si, t[1]])
            stack.extend(ch)
        exp_dfs[si] = cnt
    dfs_mean_fail = exp_dfs[fail_idx[:600]].mean() if len(fail_idx) else 0
    tree_calls_succ = expansions_all[tree_succ] + 1
    print(f"ctx {c}: K kept {Kk}/{K-1}, PCS flat {succ.mean():.4f} tree {tree_succ.mean():.4f} | flat calls/sample: all {flat_calls.mean():.1f} succ {flat_calls[succ].mean() if succ.any() else 0:.1f} fail {flat_calls[~succ].mean():.1f}"
          f" | tree expansions: succ {tree_calls_succ.mean() if tree_succ.any() else 0:.1f} fail(all live) {expansions_all[~tree_succ].mean():.1f} fail(dfs) {dfs_mean_fail:.1f}")
    tot["flat"] += flat_calls.sum(); tot["n"] += S
    tot["tree"] += tree_calls_succ.sum() + (dfs_mean_fail + 1) * (~tree_succ).sum()
print("mean Philox calls per sample: flat %.1f  tree %.1f" % (tot["flat"] / tot["n"], tot["tree"] / tot["n"]))
