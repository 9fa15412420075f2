"""CPU checks of the max-tree PCS stream restatement (oracle/pcs_tree_ref.py): the inverse-normal
sequence against scipy, the lazy walk against brute force, and the estimator against quadrature."""
import numpy as np
import pytest
from scipy.special import ndtri_exp

from oracle import pcs_oracle, pcs_tree_ref, philox_ref


def test_zfun32_accuracy_and_monotonicity():
    rng = np.random.default_rng(0)
    t = np.concatenate([10 ** rng.uniform(-11, -3, 100000), rng.uniform(1e-3, 0.7, 100000),
                        rng.uniform(0.6, 5, 100000), rng.uniform(5, 25, 100000)]).astype(np.float32)
    err = np.abs(pcs_tree_ref.zfun32(t) - ndtri_exp(-t.astype(np.float64)))
    assert err.max() < 3e-6  # |z| up to 6.7: a few fp32 ulps
    grid = (10 ** np.linspace(-11, 1.8, 1000001)).astype(np.float32)
    assert not (np.diff(pcs_tree_ref.zfun32(grid)) > 0).any()  # pruning relies on this
    assert np.isfinite(pcs_tree_ref.zfun32(np.array([0.0, 1e-30, 64.0, 1e9], dtype=np.float32))).all()


def test_tree_plan_order_and_bounds():
    rng = np.random.default_rng(1)
    mean = rng.normal(size=23)
    var = rng.uniform(0.1, 2.0, size=23)
    mean[7] = mean[3]  # exact tie in the key when variances match too
    var[7] = var[3]
    pl = pcs_tree_ref.tree_plan(mean, var)
    b = int(np.argmax(mean))
    assert pl["best"] == b and pl["D"] == 3 and len(pl["order"]) == 22 and b not in pl["order"]
    key = (mean - mean[b]) / np.sqrt(var + var[b])
    assert np.all(np.diff(key[pl["order"]]) <= 1e-6)
    i3, i7 = list(pl["order"]).index(3), list(pl["order"]).index(7)
    assert i7 == i3 + 1  # ties: lower arm id first
    nd = pcs_tree_ref.NUM_DIRECT
    assert np.array_equal(pl["top_mu"], (mean[pl["order"][:nd]] - mean[b]).astype(np.float32))
    assert pl["leaf_mu"][0] == np.float32(mean[pl["order"][nd]] - mean[b])
    # the chord table bounds the envelope max_j (mu_j + sd_j z) of every node from above
    nl = 22 - nd
    lm, ls = pl["leaf_mu"][:nl].astype(np.float64), pl["leaf_sd"][:nl].astype(np.float64)
    for z in np.linspace(-3, 6.9, 100):
        assert pcs_tree_ref.node_bound(pl["levels"][0][0], z) >= (lm + ls * max(z, 0)).max()  # root
        for i in range(4):
            seg = slice(16 * i, min(16 * i + 16, nl))
            if seg.start < nl:
                assert pcs_tree_ref.node_bound(pl["levels"][1][i], z) >= (lm[seg] + ls[seg] * z).max()
    assert (pl["levels"][2][5:] < -9e29).all()  # nodes without a real leaf can never win
    small = pcs_tree_ref.tree_plan(mean[:4], var[:4])  # K - 1 = 3 competitors: all direct, the tree is empty
    assert np.isneginf(small["top_mu"][3:]).all() and np.isneginf(small["leaf_mu"]).all() and small["D"] == 1


@pytest.mark.parametrize("K", [2, 5, 6, 7, 17, 40, 300])
def test_lazy_walk_equals_brute_force(K):
    rng = np.random.default_rng(K)
    mean = rng.normal(size=(2, K))
    var = rng.uniform(0.2, 2.0, size=(2, K))
    S = 1500
    brute = pcs_tree_ref.pcs_counts_tree_ref(mean, var, S, seed=12345, offset=2, ctx_offset=5)
    lazy, calls = pcs_tree_ref.pcs_counts_tree_lazy(mean, var, S, seed=12345, offset=2, ctx_offset=5)
    assert np.array_equal(brute, lazy)
    assert 2 * S <= calls <= 2 * S * (2 + (4 ** pcs_tree_ref.tree_depth(K) - 1) // 3)


def test_tree_stream_is_unbiased():
    """Same law as K - 1 iid normals: standardised deviations from quadrature have mean 0, rms 1."""
    rng = np.random.default_rng(2)
    dev = []
    for K in (3, 17, 100):
        mean = 0.5 * rng.normal(size=(40, K))
        var = rng.uniform(0.2, 2.0, size=(40, K))
        S = 8192
        cnt = pcs_tree_ref.pcs_counts_tree_ref(mean, var, S, seed=999)
        exact = pcs_oracle.pcs_exact_quadrature(mean, var)
        dev.append((cnt / S - exact) / np.sqrt(exact * (1 - exact) / S))
    dev = np.concatenate(dev)
    assert abs(dev.mean()) < 4 / np.sqrt(dev.size) and 0.8 < np.sqrt((dev ** 2).mean()) < 1.2
    # split sample ranges and context offsets only re-key the stream
    a = pcs_tree_ref.pcs_counts_tree_ref(mean[:3], var[:3], 3000, seed=4, offset=1, ctx_offset=9, chunk=1024)
    b = pcs_tree_ref.pcs_counts_tree_ref(mean[1:3], var[1:3], 3000, seed=4, offset=1, ctx_offset=10, chunk=3000)
    assert np.array_equal(a[1:], b)
