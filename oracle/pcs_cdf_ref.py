"""The "cdf" PCS stream, restated on the CPU with numpy.  TEST INFRASTRUCTURE ONLY (see
``oracle/__init__``).

What is estimated is the ModelListGP branch of ``estimate_current_generalized_pcs``
(``contextual_rs/generalized_pcs.py:441-475``): per context, the fraction of posterior samples in
which the predicted-best arm b beats every other arm, ``func_I = (delta > 0)``
(``experiments/wsc_experiments/main.py:453``).  The reference draws all K normals of a sample with
``torch.randn`` inside ``posterior.rsample`` (``generalized_pcs.py:449-451``); the stream below is the
build's own definition (kernels: ``contextual_rs_b200/csrc/pcs_cdf.cu``).  Arms are independent, so
conditionally on the best arm's draw ``y_b = mu_b + sd_b z`` the sample succeeds with probability

    F_c(z) = prod_{j != b} Phi(a_j z + b_j),     a_j = sd_b / sd_j,   b_j = (mu_b - mu_j) / sd_j,

and ``1{u < F_c(z)}`` with ``z ~ N(0,1)``, ``u ~ U(0,1)`` has exactly the law of the reference's
indicator (a Bernoulli(PCS(c)) draw): ``counts[c] ~ Binomial(S, PCS(c))`` either way.

  * range: ``lo = max(max_j (-6 - b_j) / a_j, -5.8)``, ``hi = min(max_j (7 - b_j) / a_j, 5.8)`` (fp32, the
    kernel's operation order); ``F = 0`` below ``lo`` (one factor is below Phi(-6)), ``F = 1`` above
    ``hi`` (every factor is above Phi(7)); ``hi = lo + 1e-3`` if the range is empty.
  * live range: ``ln F`` summed over ALL competitors is non-decreasing in z; two 8-point evaluations (nine
    equal parts of ``[lo, hi]``, then five equal parts of the two bracketing intervals) move ``lo`` up to the
    last point with ``ln F < -23`` and ``hi`` down to the first with ``ln F > -2e-8`` (skipped when
    ``hi - lo <= 29 h0``).  ``F = 0`` / ``F = 1`` outside the live range to 1e-10 / 2e-8.
  * table: ``G = 32 R - 3`` cells (``R = 1..4`` quarters of the grid: 29 / 61 / 93 / 125), the fewest with
    ``h = (hi - lo) / G <= h0 = 11.6 / 125``; competitors with ``a_j h > 1`` whose factor
    is not yet 1 at ``lo`` (``a_j lo + b_j < 7``) are SHARP: they stay out of the table (at most 8 per
    context, in arm order) and are evaluated exactly per sample.  The other factors are multiplied at
    the grid points ``z_g = lo + (g - 1) h``, ``g = 0..G+2``, and each cell carries the 4-point Lagrange
    cubic through its four surrounding grid values.
  * Philox4x32-10 (``philox_ref``), key = seed, counter ``(call, context, 0xCDF00001, offset)``:
    ``(w0, w1)`` -> Box-Muller pair ``(z1, z2)`` = the best arm's normals of samples ``2 call`` and
    ``2 call + 1``; the uniforms are ``w2`` and ``w3`` as 32-bit integers:
    success  <=>  ``w < floor(F(z) 2^32)`` (saturating).

``ln Phi`` is evaluated exactly here (scipy) where the kernel uses a 1,024-cell cubic table (error
< 5e-9), and ``exp`` exactly where the kernel uses MUFU.EX2: F differs by ~1e-6 relative, so counts
agree with the kernel up to that fraction of borderline samples.
"""
from __future__ import annotations

import numpy as np
from scipy.special import log_ndtr

from .philox_ref import philox4x32_10, box_muller

F32 = np.float32
G_CELLS = 125                      # most cells of a table
MIN_CELLS = 29                     # fewest (one quarter of the grid points)
H0 = F32(11.6 / 125.0)             # design spacing (a full-range table)
LN_DEAD = -23.0                    # ln F below this: F = 0 below the live range
LN_FULL = -2e-8                    # ln F above this: F = 1 above it
Z_MAX = F32(5.8)
X_LO = F32(6.0)
X_HI = F32(7.0)
SHARP = F32(1.0)
MAX_SHARP = 8
TAG = 0xCDF00001


def context_plan(mean_row: np.ndarray, var_row: np.ndarray) -> dict:
    """Range, sharp list and cubic table of one context (``mean_row``/``var_row`` in the grid dtype)."""
    b = int(np.argmax(mean_row))
    mu = np.delete(mean_row - mean_row[b], b).astype(F32)        # centred in the grid dtype, then fp32
    if var_row.dtype == np.float32:
        sd_all = np.sqrt(var_row).astype(F32)
    else:
        sd_all = np.sqrt(var_row.astype(np.float64)).astype(F32)
    sd_b, sd = sd_all[b], np.delete(sd_all, b)
    r = (sd / sd_b).astype(F32)
    a = (sd_b / sd).astype(F32)
    bb = (-mu / sd).astype(F32)
    zlo = np.max(((-X_LO - bb).astype(F32) * r).astype(F32))
    zhi = np.max(((X_HI - bb).astype(F32) * r).astype(F32))
    lo = max(zlo, -Z_MAX)
    hi = min(zhi, Z_MAX)
    if not hi > F32(lo + F32(1e-3)):
        hi = F32(lo + F32(1e-3))
    lo, hi = F32(lo), F32(hi)
    # live range: ln F (all competitors) is non-decreasing; two 8-point evaluations bracket where it leaves
    # LN_DEAD (F = 0 below) and reaches LN_FULL (F = 1 above) — the kernel's search, with exact ln Phi
    if F32(hi - lo) > F32(MIN_CELLS * H0):
        a64, b64 = a.astype(np.float64), bb.astype(np.float64)

        def ln_f(zs):
            x = a64[None, :] * np.asarray(zs, dtype=np.float64)[:, None] + b64[None, :]
            return log_ndtr(np.clip(x, -9.0, 40.0)).sum(axis=1)

        def leading(flags):
            n = 0
            for f in flags:
                if not f:
                    break
                n += 1
            return n

        d1 = F32(F32(hi - lo) * F32(1.0 / 9.0))
        z1 = [F32(np.float64(p + 1) * np.float64(d1) + np.float64(lo)) for p in range(8)]
        lf = ln_f(z1)
        nlo, nhi = leading(lf < LN_DEAD), leading(lf < LN_FULL)
        l0 = F32(np.float64(nlo) * np.float64(d1) + np.float64(lo))
        u1 = hi if nhi == 8 else F32(np.float64(nhi + 1) * np.float64(d1) + np.float64(lo))
        u0 = lo if F32(u1 - d1) < lo else (F32(8.0 * np.float64(d1) + np.float64(lo)) if nhi == 8 else F32(u1 - d1))
        d2l, d2u = F32(d1 * F32(0.2)), F32(F32(u1 - u0) * F32(0.2))
        z2 = [F32(np.float64(p + 1) * np.float64(d2l) + np.float64(l0)) for p in range(4)] + \
             [F32(np.float64(p + 1) * np.float64(d2u) + np.float64(u0)) for p in range(4)]
        lf2 = ln_f(z2)
        mlo, mhi = leading(lf2[:4] < LN_DEAD), leading(lf2[4:] < LN_FULL)
        lo2 = F32(np.float64(mlo) * np.float64(d2l) + np.float64(l0))
        hi2 = u1 if mhi == 4 else F32(np.float64(mhi + 1) * np.float64(d2u) + np.float64(u0))
        lo = max(lo, lo2)
        hi = min(hi, max(hi2, F32(lo + F32(1e-3))))
        lo, hi = F32(lo), F32(hi)
    R = (int(np.ceil(F32(F32(hi - lo) * F32(1.0 / H0)))) + 3 + 31) >> 5
    R = min(max(R, 1), 4)
    cells = 32 * R - 3
    h = F32(F32(hi - lo) / F32(cells))
    inv_h = F32(F32(cells) / F32(hi - lo))
    alive = (a * lo + bb).astype(F32) < X_HI                     # (the kernel's fma differs in the last ulp only)
    sharp_idx = np.nonzero(((a * h).astype(F32) > SHARP) & alive)[0]
    overflow = len(sharp_idx) > MAX_SHARP
    sharp_idx = sharp_idx[:MAX_SHARP]
    smooth = np.ones(len(a), dtype=bool)
    smooth[sharp_idx] = False
    g = np.arange(cells + 3, dtype=np.float64)
    z = (np.float64(lo) + (g - 1.0) * np.float64(h)).astype(F32).astype(np.float64)
    x = a[smooth].astype(np.float64)[None, :] * z[:, None] + bb[smooth].astype(np.float64)[None, :]
    lnF = np.where(x < float(X_HI), log_ndtr(np.minimum(x, 40.0)), 0.0).sum(axis=1)
    Fg = np.exp(np.maximum(lnF, -80.0)).astype(F32).astype(np.float64)
    p0, p1, p2, p3 = Fg[0:cells], Fg[1:cells + 1], Fg[2:cells + 2], Fg[3:cells + 3]
    coef = np.stack([p1, -p0 / 3 - p1 / 2 + p2 - p3 / 6, (p0 + p2) / 2 - p1, (p3 - p0) / 6 + (p1 - p2) / 2], axis=1)
    return {"best": b, "lo": lo, "hi": hi, "h": h, "inv_h": inv_h, "cells": cells, "coef": coef.astype(F32),
            "sharp_a": a[sharp_idx], "sharp_b": bb[sharp_idx], "overflow": overflow, "a": a, "b": bb}


def success_probability(plan: dict, z: np.ndarray) -> np.ndarray:
    """The kernel's F(z): table coordinate ``t' = fma(z, inv_h, t0)`` with ``t0 = fma(-lo, inv_h, 1)``,
    clamped to [0, 127); slot 0 is the guard below the range (F = 0), slots 1..125 the cubic cells (fp32
    Horner), slot 126 the guard above it (F = 1); times the sharp competitors' exact factors."""
    z = z.astype(F32)
    lo64, ih64 = np.float64(plan["lo"]), np.float64(plan["inv_h"])
    t0 = F32(1.0 - lo64 * ih64)                                  # one rounding, like the kernel's fma
    t = (z.astype(np.float64) * ih64 + np.float64(t0)).astype(F32)
    tc = np.minimum(np.maximum(t, F32(0.0)), F32(plan["cells"] + 2 - 0.001))
    i = tc.astype(np.int32)
    f = (tc - i.astype(F32)).astype(F32)
    guard_lo = np.zeros((1, 4), dtype=F32)
    guard_hi = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]], dtype=F32)
    c = np.concatenate([guard_lo, plan["coef"], guard_hi], axis=0)[i]
    F = ((c[:, 3] * f + c[:, 2]).astype(F32) * f + c[:, 1]).astype(F32)
    F = (F * f + c[:, 0]).astype(F32).astype(np.float64)
    for a, b in zip(plan["sharp_a"], plan["sharp_b"]):
        F = F * np.exp(log_ndtr(np.float64(a) * z.astype(np.float64) + np.float64(b)))
    return F


def exact_success_probability(plan: dict, z: np.ndarray) -> np.ndarray:
    """prod_j Phi(a_j z + b_j) without any table (float64) — what the interpolant approximates."""
    x = plan["a"].astype(np.float64)[None, :] * z.astype(np.float64)[:, None] + plan["b"].astype(np.float64)[None, :]
    return np.exp(log_ndtr(x).sum(axis=1))


def pcs_counts_cdf_ref(mean: np.ndarray, variance: np.ndarray, num_samples: int, seed: int, offset: int = 0,
                       ctx_offset: int = 0, exact: bool = False):
    """Counts per context under the cdf stream; ``exact=True`` replaces the table interpolant by the
    exact F (same draws) — the difference between the two is the interpolation's effect on the counts.
    Returns ``(counts, margin)``; ``margin[c]`` = number of samples whose uniform lies within 1e-5 of
    F(z) in relative terms or 1e-6 absolute (candidates for a kernel-vs-restatement flip)."""
    n_c = mean.shape[0]
    ncalls = (num_samples + 1) // 2
    call = np.arange(ncalls, dtype=np.uint32)
    counts = np.zeros(n_c, dtype=np.int64)
    margin = np.zeros(n_c, dtype=np.int64)
    for c in range(n_c):
        plan = context_plan(mean[c], variance[c])
        w0, w1, w2, w3 = philox4x32_10(call, np.uint32(c + ctx_offset), np.uint32(TAG), np.uint32(offset),
                                       seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        z1, z2 = box_muller(w0, w1)
        z = np.stack([z1, z2], axis=1).reshape(-1)[:num_samples]
        w = np.stack([w2, w3], axis=1).reshape(-1)[:num_samples].astype(np.float64)
        F = exact_success_probability(plan, z) if exact else success_probability(plan, z)
        thr = np.floor(np.clip(F, 0.0, 1.0) * 4294967296.0)
        counts[c] = int((w < np.minimum(thr, 4294967295.0)).sum())
        margin[c] = int((np.abs(w - thr) <= np.maximum(1e-5 * thr, 1e-6 * 4294967296.0)).sum())
    return counts, margin
