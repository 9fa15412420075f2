"""Philox4x32-10 and the PCS normal-variate stream, restated on the CPU with numpy integer
arithmetic.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).

Philox4x32-10 is the published counter-based generator of Salmon et al., "Parallel random
numbers: as easy as 1, 2, 3" (SC'11) as shipped in Random123 / cuRAND / PyTorch-CUDA.  It is
pinned by the Random123 known-answer vectors in ``tests/test_oracle_golden.py``.

Stream layout used by the PCS kernel (``contextual_rs_b200/csrc/pcs_philox.cu``) — the build's
own definition, the reference draws with ``torch.randn`` inside ``posterior.rsample``
(``contextual_rs/generalized_pcs.py:449-451``):

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (sample, context, arm >> 2, offset)
    words (w0, w1, w2, w3) -> the N(0,1) draws of that sample at arms 4*(arm>>2) + 0..3:
        u(w)  = (w >> 9) * 2^-23                          in [0, 1), exact in fp32
        (z0, z1) = BoxMuller(u(w0) + 2^-24, u(w1)),  (z2, z3) = BoxMuller(u(w2) + 2^-24, u(w3))
        BoxMuller(a, b) = sqrt(-2 ln a) * (cos 2 pi b, sin 2 pi b)
    |z| <= sqrt(-2 ln 2^-24) = 5.7683 — the bound the kernel's exact screening relies on.
Keying by (sample, arm block) makes every (sample, context) pair independent of the order in
which the kernel visits arms, which is what lets it stop a sample at its first failure.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
Z_MAX = float(np.sqrt(-2.0 * np.log(2.0 ** -24)))  # 5.768...


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10.  c0..c3: broadcastable uint32 arrays; k0,k1: python ints.
    Returns four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3)))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _unit(w: np.ndarray) -> np.ndarray:
    return (w >> np.uint32(9)).astype(np.float32) * np.float32(2.0 ** -23)


def box_muller(wa: np.ndarray, wb: np.ndarray):
    a = _unit(wa) + np.float32(2.0 ** -24)
    b = _unit(wb)
    rad = np.sqrt(np.float32(-2.0) * np.log(a)).astype(np.float32)
    ang = (np.float32(2.0 * np.pi) * b).astype(np.float32)
    return (rad * np.cos(ang)).astype(np.float32), (rad * np.sin(ang)).astype(np.float32)


def pcs_normals(num_samples: int, context: int, arms: np.ndarray, seed: int, offset: int) -> np.ndarray:
    """N(0,1) draws ``z[arm, sample]`` (float32) for one context under the layout above."""
    arms = np.asarray(arms, dtype=np.uint32).reshape(-1)
    blocks = np.unique(arms >> np.uint32(2))
    smp = np.arange(num_samples, dtype=np.uint32).reshape(1, -1)
    w0, w1, w2, w3 = philox4x32_10(smp, np.uint32(context), blocks.reshape(-1, 1), np.uint32(offset),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    z0, z1 = box_muller(w0, w1)
    z2, z3 = box_muller(w2, w3)
    zblk = np.stack([z0, z1, z2, z3], axis=1)  # [block, arm & 3, sample]
    pos = np.searchsorted(blocks, arms >> np.uint32(2))
    return zblk[pos, arms & np.uint32(3), :]


def pcs_counts_philox_ref(mean: np.ndarray, variance: np.ndarray, num_samples: int, seed: int,
                          offset: int = 0, arm_offset: int = 0) -> np.ndarray:
    """Correct-selection counts per context under the shared Philox stream, mimicking the
    kernel's arithmetic: best = first argmax of the mean row, means centred on the best arm in
    the grid dtype then cast to fp32, y = fma(sigma, z, mu) in fp32, success iff y_best >
    max_{j != best} y_j (strict, the reference's func_I = (delta > 0),
    ``experiments/wsc_experiments/main.py:453``).  mean/variance: ``[n_c, K]``."""
    n_c, K = mean.shape
    counts = np.zeros(n_c, dtype=np.int64)
    arms = np.arange(K, dtype=np.uint32) + np.uint32(arm_offset)
    for c in range(n_c):
        b = int(np.argmax(mean[c]))
        mu = (mean[c] - mean[c, b]).astype(np.float32).reshape(-1, 1)
        sd = np.sqrt(variance[c]).astype(np.float32).reshape(-1, 1)
        z = pcs_normals(num_samples, c, arms, seed, offset)
        y = (sd.astype(np.float64) * z.astype(np.float64) + mu.astype(np.float64)).astype(np.float32)
        yb = y[b]
        rest = np.delete(y, b, axis=0)
        counts[c] = int((yb > rest.max(axis=0)).sum())
    return counts
