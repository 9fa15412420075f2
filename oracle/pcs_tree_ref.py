"""The "max-tree" PCS stream, restated on the CPU by BRUTE FORCE (every node and every leaf of the
tree is generated for every sample).  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).

What is estimated is the ModelListGP branch of ``estimate_current_generalized_pcs``
(``contextual_rs/generalized_pcs.py:441-475``): per context, the fraction of posterior samples in
which the predicted-best arm b beats every other arm, ``func_I = (delta > 0)``
(``experiments/wsc_experiments/main.py:453``).  The reference draws with ``torch.randn`` inside
``posterior.rsample`` (``generalized_pcs.py:449-451``); the stream below is the build's own
definition (kernel: ``contextual_rs_b200/csrc/pcs_tree.cu``), equal in law to K - 1 independent
N(0,1) draws per (sample, context):

  * The K - 1 competitors are ordered by decreasing threat (mu_j - mu_b) / sqrt(sd_j^2 + sd_b^2)
    (IEEE fp32, ties by arm id).  The first five ("direct" arms) are drawn directly; the other
    K - 6 are the leaves of a complete 4-ary tree of depth D (4^D >= K - 6) in that order; unused
    leaves (and unused direct slots when K < 6) have mu = -inf.
  * A node that covers m leaves carries (t, J): t = -ln P(max of its m iid normals <= its value),
    J = the leaf that attains the maximum.  The maximum of m iid N(0,1) has CDF Phi(z)^m, so
    t_root = E / m with E ~ Exp(1) and J_root uniform.  Given a node's (t, J), the child that
    contains J inherits (t, J); every other child (m_c leaves) has the maximum of m_c normals
    truncated at the parent's value, t_c = t + E' / m_c, and a uniform J_c among its leaves.
    At the leaves m = 1 and z_leaf = zfun(t_leaf), Phi(zfun(t)) = exp(-t).
  * Philox4x32-10 (``philox_ref``), key = seed, counters
        (sample, context, 0, offset)            -> w0, w1: Box-Muller (z_b, z of direct arm 0), y_b = sd_b z_b;
                                                  w2: E of the root, w3 & (4^D - 1): J of the root
        (sample, context, 0xffffffff, offset)   -> Box-Muller (w0, w1), (w2, w3): z of direct arms 1, 2, 3, 4
        (sample, context, 1 + node id, offset)  -> w0, w1, w2: E' of the three children that do not
                                                  contain J (in child order); 10-bit fields of w3: their J
    node id = (4^level - 1) / 3 + index;  E(w) = -ln u(w), u(w) = ((w >> 9) + 1/2) 2^-23.
  * success  <=>  y_b > max_j (mu_j + sd_j z_j)   (strict).

The kernel walks the same tree lazily: it tests y_J at every node it creates (a failure ends the
sample) and expands a node only if an upper bound of max_j (mu_j + sd_j z(t)) over the node's
leaves (chords of that convex envelope) reaches y_b — exact because zfun is non-increasing in t.  Counts must agree with this brute-force restatement up to MUFU-vs-libm
borderline cases.
"""
from __future__ import annotations

import numpy as np

from .philox_ref import philox4x32_10, box_muller, _unit

F32 = np.float32
LN2 = F32(0.6931471805599453)
LOG2E = F32(1.4426950408889634)
W_MAX = F32(27.0)
W_SPLIT = F32(6.25)
CEN_MID = F32(3.125)
TAIL_MID = F32(3.848076105117798)
# sqrt(2) erfinv(x) / x  as a polynomial in (w - CEN_MID), w = -ln(4 p (1 - p)) <= W_SPLIT, and in
# (sqrt(w) - TAIL_MID) for W_SPLIT <= w <= W_MAX (Chebyshev least-squares fits, fp32-rounded)
_SQ2 = np.sqrt(2.0)
CEN = [F32(_SQ2 * c) for c in (1.6536552906036377, 0.24015966057777405, -0.006036306265741587,
                               -0.000742710311897099, 0.00018829306645784527, -1.314547989750281e-05,
                               -1.6936899100983283e-06, 3.2441943176308996e-07)]
TAIL = [F32(_SQ2 * c) for c in (3.6864500045776367, 1.0090477466583252, 0.0017072835471481085,
                                -0.001114317448809743, 0.00042597835999913514, -0.00014658486179541796,
                                0.00011464921408332884, -0.00010220367403235286, 3.313723573228344e-05)]
EXPM1 = [F32(c) for c in (1.0, -1.0 / 2, 1.0 / 6, -1.0 / 24, 1.0 / 120)]  # (1 - e^-t) / t, t < 1/8
MAX_LEAVES = 4096


def _fma(a, b, c):
    """fp32 fused multiply-add (the product of two fp32 is exact in fp64)."""
    return (np.asarray(a, dtype=np.float64) * np.asarray(b, dtype=np.float64)
            + np.asarray(c, dtype=np.float64)).astype(F32)


def _horner(coef, v):
    acc = np.full(np.shape(v), coef[-1], dtype=F32)
    for c in coef[-2::-1]:
        acc = _fma(acc, v, c)
    return acc


def zfun32(t) -> np.ndarray:
    """z with Phi(z) = exp(-t), t >= 0, evaluated with the kernel's fp32 operation sequence."""
    t = np.asarray(t, dtype=F32)
    upper = t < LN2                                   # P > 1/2: z > 0
    P = np.exp2((t * -LOG2E).astype(F32)).astype(F32)
    with np.errstate(over="ignore", invalid="ignore"):  # the series is only selected for t < 1/8
        q = np.where(t < F32(0.125), (_horner(EXPM1, t) * t).astype(F32), (F32(1) - P).astype(F32))
    p = np.where(upper, q, P).astype(F32)             # the smaller tail probability
    a = _fma(-p, p, p)                                # p (1 - p)
    with np.errstate(divide="ignore"):
        w = _fma(np.log2(a).astype(F32), -LN2, F32(-1.3862943611198906))  # -ln(4 p (1-p))
    w = np.minimum(w, W_MAX)  # w >= -1e-6 (rounding); the NaN square root of such a w is never selected
    ec = _horner(CEN, (w - CEN_MID).astype(F32))
    with np.errstate(invalid="ignore"):
        et = _horner(TAIL, (np.sqrt(w).astype(F32) - TAIL_MID).astype(F32))
    e = np.where(w < W_SPLIT, ec, et)
    x = _fma(F32(-2), p, F32(1))
    z = (e * x).astype(F32)
    return np.where(upper, z, -z).astype(F32)


def exp_log2(w: np.ndarray) -> np.ndarray:
    """log2 u(w), u(w) = ((w >> 9) + 1/2) 2^-23 in (0, 1) — the Exp(1) variate is -ln 2 times this."""
    return np.log2(_unit(w) + F32(2.0 ** -24)).astype(F32)


NUM_DIRECT = 5           # most threatening competitors, drawn outside the tree
DIRECT_NODE = 0xFFFFFFFF  # Philox counter word of the call that draws direct arms 1..4
GRID = F32(1.8)          # spacing of the envelope table g(0), g(1.8), .. g(7.2)
EMPTY = F32(-1e30)


def tree_depth(K: int) -> int:
    D = 1
    while 4 ** D < K - 1 - NUM_DIRECT:
        D += 1
    return D


def tree_plan(mean_row: np.ndarray, var_row: np.ndarray):
    """Competitors in threat order and the per-node bounds.  Returns dict(best, sd_b, D, top_mu [5],
    top_sd [5], leaf_mu, leaf_sd, order, levels) with levels[l] = ``[4^l, 5]`` envelope tables
    g(z) = max over the node's leaves of mu_j + sd_j z at z = 0, 1.8, .. 7.2 (EMPTY if none)."""
    K = mean_row.shape[0]
    b = int(np.argmax(mean_row))
    mu = (mean_row - mean_row[b]).astype(F32)          # centred in the grid dtype, then fp32
    sd = np.sqrt(var_row.astype(np.float64)).astype(F32)
    s = ((sd * sd).astype(F32) + (sd[b] * sd[b]).astype(F32)).astype(F32)
    s = (s + F32(1e-30)).astype(F32)
    key = (mu / np.sqrt(s).astype(F32)).astype(F32)
    arms = np.array([a for a in range(K) if a != b], dtype=np.int64)
    order = arms[np.lexsort((arms, -key[arms]))]       # decreasing key, ties: lower arm first
    D = tree_depth(K)
    NL = 4 ** D
    leaf_mu = np.full(NL, -np.inf, dtype=F32)
    leaf_sd = np.zeros(NL, dtype=F32)
    nleaf = max(K - 1 - NUM_DIRECT, 0)
    leaf_mu[:nleaf] = mu[order[NUM_DIRECT:]]
    leaf_sd[:nleaf] = sd[order[NUM_DIRECT:]]
    real = (np.arange(NL) < nleaf).reshape(-1, 1)
    top_mu = np.full(NUM_DIRECT, -np.inf, dtype=F32)
    top_sd = np.zeros(NUM_DIRECT, dtype=F32)
    nd = min(NUM_DIRECT, K - 1)
    top_mu[:nd] = mu[order[:nd]]
    top_sd[:nd] = sd[order[:nd]]
    zk = (np.arange(5, dtype=np.float64) * float(GRID)).astype(F32).reshape(1, -1)
    g = np.where(real, _fma(leaf_sd.reshape(-1, 1), zk, leaf_mu.reshape(-1, 1)), EMPTY).astype(F32)  # [NL, 5]
    levels = [None] * D
    for l in range(D - 1, -1, -1):
        g = g.reshape(-1, 4, 5).max(axis=1)
        if l == D - 1:  # head-room for the rounding of the chord lines
            g = (g + (np.maximum(np.abs(g[:, :1]), np.abs(g[:, 4:])) * F32(2.0 ** -18)).astype(F32)).astype(F32)
        levels[l] = g
    return dict(best=b, sd_b=sd[b], D=D, top_mu=top_mu, top_sd=top_sd, leaf_mu=leaf_mu,
                leaf_sd=leaf_sd, order=order, levels=levels)


def node_bound(g: np.ndarray, z) -> float:
    """Upper bound of a node's envelope at z: the maximum of the four chord lines of its table
    (the kernel's ``node_bound``; z < 0 -> g(0))."""
    zp = F32(max(F32(z), F32(0)))
    inv = F32(1.0 / float(GRID))
    best = -np.inf
    for k in range(4):
        sl = F32(F32(g[k + 1] - g[k]) * inv)
        a = g[0] if k == 0 else _fma(-sl, F32(k * float(GRID)), g[k])
        best = max(best, float(_fma(sl, zp, a)))
    return best


def tree_leaf_normals(num_samples: int, context: int, D: int, seed: int, offset: int, first_sample: int = 0):
    """Brute force: (z_b [S], z_direct [S, 5], z_leaf [S, 4^D]) of the stream for one context."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    NL = 4 ** D
    smp = (np.arange(num_samples, dtype=np.uint64) + np.uint64(first_sample)).astype(np.uint32).reshape(-1, 1)
    w0, w1, w2, w3 = philox4x32_10(smp, np.uint32(context), np.uint32(0), np.uint32(offset), k0, k1)
    zb, zd0 = (z.reshape(-1) for z in box_muller(w0, w1))
    v0, v1, v2, v3 = philox4x32_10(smp, np.uint32(context), np.uint32(DIRECT_NODE), np.uint32(offset), k0, k1)
    zd = np.stack([zd0] + [z.reshape(-1) for z in box_muller(v0, v1) + box_muller(v2, v3)], axis=1)
    t = (exp_log2(w2) * F32(-float(LN2) / NL)).astype(F32)       # [S, 1]
    J = (w3 & np.uint32(NL - 1)).astype(np.int64)
    for l in range(D):
        nn = 4 ** l
        sh = 2 * (D - l - 1)
        m_c = 1 << sh
        c_m = F32(-float(LN2) / m_c)
        ids = np.uint32((4 ** l - 1) // 3) + np.arange(nn, dtype=np.uint32)
        w0, w1, w2, w3 = philox4x32_10(smp, np.uint32(context), np.uint32(1) + ids.reshape(1, -1),
                                       np.uint32(offset), k0, k1)
        kp = (J >> sh) & 3                                        # child that contains J
        tc = np.empty((num_samples, nn, 4), dtype=F32)
        Jc = np.empty((num_samples, nn, 4), dtype=np.int64)
        words = (w0, w1, w2)
        for k in range(4):
            r = np.where(k < kp, k, k - 1)                        # rank among the other three children
            word = np.where(r == 0, words[0], np.where(r == 1, words[1], words[2]))
            t_new = _fma(exp_log2(word), c_m, t)
            J_new = (np.arange(nn).reshape(1, -1) * 4 + k) * m_c + ((w3.astype(np.int64) >> (10 * np.clip(r, 0, 2))) & (m_c - 1))
            own = kp == k
            tc[:, :, k] = np.where(own, t, t_new)
            Jc[:, :, k] = np.where(own, J, J_new)
        t = tc.reshape(num_samples, nn * 4)
        J = Jc.reshape(num_samples, nn * 4)
    return zb, zd, zfun32(t)


def pcs_counts_tree_ref(mean: np.ndarray, variance: np.ndarray, num_samples: int, seed: int,
                        offset: int = 0, ctx_offset: int = 0, chunk: int = 4096) -> np.ndarray:
    """Correct-selection counts per context under the max-tree stream.  mean/variance: ``[n_c, K]``."""
    n_c, K = mean.shape
    assert 2 <= K <= MAX_LEAVES + 1 + NUM_DIRECT
    counts = np.zeros(n_c, dtype=np.int64)
    for c in range(n_c):
        pl = tree_plan(mean[c], variance[c])
        for s0 in range(0, num_samples, chunk):
            S = min(chunk, num_samples - s0)
            zb, zd, z = tree_leaf_normals(S, c + ctx_offset, pl["D"], seed, offset, first_sample=s0)
            yb = (pl["sd_b"] * zb).astype(F32)
            ytop = _fma(pl["top_sd"].reshape(1, -1), zd, pl["top_mu"].reshape(1, -1))
            y = _fma(pl["leaf_sd"].reshape(1, -1), z, pl["leaf_mu"].reshape(1, -1))
            counts[c] += int(((y.max(axis=1) < yb) & (ytop.max(axis=1) < yb)).sum())
    return counts


def pcs_counts_tree_lazy(mean: np.ndarray, variance: np.ndarray, num_samples: int, seed: int,
                         offset: int = 0, ctx_offset: int = 0):
    """The kernel's lazy walk (per-sample Python loop — small cases only).  Returns (counts, number of
    Philox calls), to check that pruning changes no outcome and to predict the kernel's work."""
    n_c, K = mean.shape
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    counts = np.zeros(n_c, dtype=np.int64)
    calls = 0
    for c in range(n_c):
        pl = tree_plan(mean[c], variance[c])
        D, NL, lm, ls, lv = pl["D"], 4 ** pl["D"], pl["leaf_mu"], pl["leaf_sd"], pl["levels"]
        ctx = np.uint32(c + ctx_offset)

        def bound(l, i, z):
            return node_bound(lv[l][i], z)

        smp = np.arange(num_samples, dtype=np.uint32)
        w0, w1, w2, w3 = philox4x32_10(smp, ctx, np.uint32(0), np.uint32(offset), k0, k1)
        zb, zd0 = box_muller(w0, w1)
        yb_all = (pl["sd_b"] * zb).astype(F32)
        ytop_all = _fma(pl["top_sd"][0], zd0, pl["top_mu"][0])
        t_all = (exp_log2(w2) * F32(-float(LN2) / NL)).astype(F32)
        z_all = zfun32(t_all)
        J_all = (w3 & np.uint32(NL - 1)).astype(np.int64)
        calls += num_samples
        for s in range(num_samples):
            yb, t, J, z = float(yb_all[s]), t_all[s], int(J_all[s]), z_all[s]
            fail = float(ytop_all[s]) >= yb or float(_fma(ls[J], z, lm[J])) >= yb
            if not fail:  # direct arms 1..4: one more Philox call for every sample that survives the root test
                calls += 1
                v = philox4x32_10(np.uint32(s), ctx, np.uint32(DIRECT_NODE), np.uint32(offset), k0, k1)
                zd = np.array([z_[0] for z_ in box_muller(v[0].reshape(1), v[1].reshape(1)) + box_muller(v[2].reshape(1), v[3].reshape(1))])
                fail = bool((_fma(pl["top_sd"][1:], zd, pl["top_mu"][1:]) >= yb).any())
            stack = []
            if not fail and bound(0, 0, z) >= yb:
                stack.append((0, 0, t, J, z))
            while stack and not fail:
                l, i, t, J, z = stack.pop()
                calls += 1
                sh = 2 * (D - l - 1)
                m_c = 1 << sh
                c_m = F32(-float(LN2) / m_c)
                w = philox4x32_10(np.uint32(s), ctx, np.uint32(1 + (4 ** l - 1) // 3 + i), np.uint32(offset), k0, k1)
                kp = (J >> sh) & 3
                for k in range(4):
                    ci = 4 * i + k
                    if k == kp:
                        tc, Jc, zc = t, J, z
                    else:
                        r = k if k < kp else k - 1
                        tc = _fma(exp_log2(w[r].reshape(1)), c_m, t)[0]
                        Jc = ci * m_c + ((int(w[3]) >> (10 * r)) & (m_c - 1))
                        zc = zfun32(np.array([tc]))[0]
                        if float(_fma(ls[Jc], zc, lm[Jc])) >= yb:
                            fail = True
                            break
                    if l + 1 < D and bound(l + 1, ci, zc) >= yb:
                        stack.append((l + 1, ci, tc, Jc, zc))
            counts[c] += 0 if fail else 1
    return counts, calls
