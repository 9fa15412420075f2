"""GPU parity tests proper: the CUDA path (through the C ABI of libcrs_b200.so) against the CPU
oracle on the same seeded inputs and against the committed golden fixtures.

Stated tolerances (BASELINE.json north_star): selected (arm, context) and best-arm indices
bit-exact; posterior mean / variance within 1e-10 in fp64 and 1e-5 in fp32, *relative to the
table's scale* (max |mean|, resp. max variance = the prior scale s^2 o) — a pure element-wise
relative bound is meaningless where the mean crosses zero or in fp32 where var = o - ||v||^2
cancels — plus an element-wise relative bound on the variance in fp64; PCS within
5 MC standard errors of exact quadrature and count-identical to the CPU Philox restatement up to
borderline draws (MUFU vs libm transcendental rounding).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import gp_oracle, levi_oracle, ocba_oracle, pcs_oracle, pcs_tree_ref, philox_ref

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def crs():
    import contextual_rs_b200 as crs
    from contextual_rs_b200 import _lib
    _lib.get()
    return crs


def _scaled_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())


CASES = [
    # K, n_c, d, n, kernel, train
    (10, 10, 1, 20, "matern52", "full"),      # BASELINE config 1
    (3, 4, 2, 10, "matern52", "rows"),        # shapes of the reference's test_gao_modellist
    (100, 1024, 4, 20, "matern52", "rows"),   # BASELINE config 2
    (100, 1024, 4, 20, "rbf", "rows"),
    (17, 333, 3, 27, "matern52", "rows"),     # d not a power of two, ragged tiles
    (9, 100, 5, 40, "rbf", "rows"),           # n in (32, 64] -> NB=64 variant
    (2, 1, 1, 1, "matern52", "rand"),         # minimum sizes: K=2, one context, one observation
    (5, 130, 8, 56, "matern52", "rand"),
    (4, 64, 16, 12, "rbf", "rand"),           # maximum supported dimension
]


@pytest.mark.parametrize("K,n_c,d,n,kernel,train", CASES)
def test_posterior_and_decision_fp64(crs, K, n_c, d, n, kernel, train):
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=K, n_c=n_c, d=d, n=n, kernel=kernel, train=train, seed=1)
    model = crs.B200ModelListGP(desc, device=DEV, dtype=torch.float64)
    model.check_status()
    ref = gp_oracle.model_from_description(desc)
    rp, p = ref.posterior(ctx), model.posterior(ctx)
    assert p.mean.shape == (n_c, K) and p.variance.shape == (n_c, K)
    assert _scaled_err(p.mean, rp.mean) < 1e-10
    assert _scaled_err(p.variance, rp.variance) < 1e-10
    torch.testing.assert_close(p.variance.cpu(), rp.variance, rtol=1e-9, atol=0)
    # batch x 1 x d input (MinZeta's shape) gives batch x 1 x K
    p3 = model.posterior(ctx.unsqueeze(1))
    assert p3.mean.shape == (n_c, 1, K) and torch.equal(p3.mean.squeeze(1), p.mean)
    # model.train_inputs are bit-identical to the oracle's (needed by the exact-match count)
    for (a,), (b,) in zip(model.train_inputs, ref.train_inputs):
        assert torch.equal(a.cpu(), b)
    assert torch.equal(crs.empirical_best_arms(model, ctx).cpu(), ocba_oracle.best_arms_ref(ref, ctx))
    for kw in ({}, {"use_kde": True, "kernel_scale": 1.0}, {"infer_p": True}, {"randomize_ties": False}):
        arm, x = crs.gao_modellist(model, ctx, **kw)
        rarm, rx = ocba_oracle.gao_modellist_ref(ref, ctx, **kw)
        assert arm.dtype == torch.long and arm.dim() == 0
        assert int(arm) == int(rarm), kw
        assert torch.equal(x, rx) and x.data_ptr() == rx.data_ptr(), kw  # a view of context_set
    # context_set already on the GPU: same decision, outputs on the GPU
    arm_d, x_d = crs.gao_modellist(model, ctx.to(DEV))
    rarm, rx = ocba_oracle.gao_modellist_ref(ref, ctx)
    assert arm_d.device.type == "cuda" and int(arm_d) == int(rarm) and torch.equal(x_d.cpu(), rx)


@pytest.mark.parametrize("K,n_c,d,n,kernel", [(100, 1024, 4, 20, "matern52"), (20, 257, 8, 50, "rbf")])
def test_posterior_fp32(crs, K, n_c, d, n, kernel):
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=K, n_c=n_c, d=d, n=n, kernel=kernel, seed=2)
    model = crs.B200ModelListGP(desc, device=DEV, dtype=torch.float32)
    rp = gp_oracle.model_from_description(desc).posterior(ctx)  # fp64 truth
    p = model.posterior(ctx.float())
    assert p.mean.dtype == torch.float32
    assert _scaled_err(p.mean, rp.mean) < 1e-5
    assert _scaled_err(p.variance, rp.variance) < 1e-5
    # decision logic on the GPU's own fp32 tables is index-exact against the oracle's logic
    class Tab:
        models = [None] * K
        train_inputs = model.train_inputs
        def posterior(self, X):
            return type("P", (), {"mean": p.mean.cpu(), "variance": p.variance.cpu()})()
    tab = Tab()
    tab.train_inputs = [(t[0].cpu(),) for t in model.train_inputs]
    for kw in ({}, {"use_kde": True, "kernel_scale": 1.0}):
        arm, x = crs.gao_modellist(model, ctx.float(), **kw)
        rarm, rx = ocba_oracle.gao_modellist_ref(tab, ctx.float(), **kw)
        assert int(arm) == int(rarm) and torch.equal(x, rx)


def test_wsc_fixture_replay(crs, golden):
    """Real training data written by the reference (0000_ML_Gao.pt): ragged per-arm n as the
    reference's own decision sequence is replayed, Branin-scale targets exercising Standardize."""
    g = golden("wsc_branin_seed0.json")
    X = torch.tensor(g["X"], dtype=torch.float64)
    Y = torch.tensor(g["Y"], dtype=torch.float64).view(-1, 1)
    ctx = torch.tensor(g["context_map"], dtype=torch.float64)
    K = g["num_arms"]
    gen = torch.Generator().manual_seed(3)
    hyp = dict(lengthscale=0.3 + 0.5 * torch.rand(K, 1, generator=gen, dtype=torch.float64),
               outputscale=0.5 + torch.rand(K, generator=gen, dtype=torch.float64),
               noise=torch.full((K,), 1e-2, dtype=torch.float64),
               mean_const=torch.zeros(K, dtype=torch.float64), kernel="matern52")
    for upto in (200, 263, 400):
        desc = dict(hyp)
        desc["train_X"] = [X[:upto][X[:upto, 0] == k][:, 1:] for k in range(K)]
        desc["train_Y"] = [Y[:upto][X[:upto, 0] == k] for k in range(K)]
        assert len({t.shape[0] for t in desc["train_X"]}) > (1 if upto > 200 else 0)
        model = crs.B200ModelListGP(desc, device=DEV)
        ref = gp_oracle.model_from_description(desc)
        rp, p = ref.posterior(ctx), model.posterior(ctx)
        assert _scaled_err(p.mean, rp.mean) < 1e-10 and _scaled_err(p.variance, rp.variance) < 1e-10
        for kw in ({}, {"use_kde": True, "kernel_scale": 1.0}):
            arm, x = crs.gao_modellist(model, ctx, **kw)
            rarm, rx = ocba_oracle.gao_modellist_ref(ref, ctx, **kw)
            assert int(arm) == int(rarm) and torch.equal(x, rx)


def test_reference_recorded_run_through_the_product(crs, golden):
    """The reference's recorded run (0000_ML_Gao.pt) replayed END TO END through the product, the
    way ``experiments/wsc_experiments/main.py:345-402`` drives it: ``fit_modellist_with_reuse``
    every 10 iterations (batched device L-BFGS on ``crs_gp_mll``), ``condition_on_observations`` in
    between, ``gao_modellist(model, context_map, randomize_ties, kernel_scale=1.0)``.  The decisions
    must track both the reference's recorded ones and the CPU oracle's replay."""
    from contextual_rs_b200.fit import fit_modellist_with_reuse
    g = golden("wsc_branin_seed0.json")
    trace = golden("reference_trace_ml_gao.json")
    X = torch.tensor(g["X"], dtype=torch.float64)
    Y = torch.tensor(g["Y"], dtype=torch.float64).view(-1, 1)
    ctx = torch.tensor(g["context_map"], dtype=torch.float64)
    n0, K, N = g["num_init"], g["num_arms"], 40
    torch.manual_seed(0)
    model, hits_ref, hits_oracle, best_hits, pcs_dev = None, 0, 0, 0, []
    w = torch.tensor(trace["pcs_weights"], dtype=torch.float64)
    rho = lambda P: (P * w.to(P.device).view(-1, 1).expand_as(P)).mean(dim=-2)  # main.py:245-249
    arm_set = torch.arange(K, dtype=torch.float64).view(-1, 1)
    for i in range(N):
        Xi, Yi = X[: n0 + i], Y[: n0 + i]
        if i % 10 == 0:
            model = fit_modellist_with_reuse(Xi, Yi, K, model, device=DEV)
        else:
            model.condition_on_observations(int(Xi[-1, 0]), Xi[-1:, 1:], Yi[-1:])
        assert model.num_observations() == [int((Xi[:, 0] == k).sum()) for k in range(K)]
        arm, x = crs.gao_modellist(model, ctx, True, kernel_scale=1.0)
        got = (int(arm), float(x.reshape(-1)[0]))
        ref = (int(X[n0 + i, 0]), float(X[n0 + i, 1]))
        hits_ref += got[0] == ref[0] and abs(got[1] - ref[1]) < 1e-12
        hits_oracle += got == tuple(trace["oracle_raw_train_inputs"][i])
        # best-arm readout (main.py:459-466) against the recorded all_maximizers
        best = crs.empirical_best_arms(model, ctx).cpu()
        best_hits += int((best == torch.tensor(trace["reference_best_arms"][i])).sum())
        if i % 5 == 0:  # PCS estimate (main.py:447-455): 2^16 samples here, 64 in the recorded run
            est = crs.estimate_current_generalized_pcs(model, arm_set, ctx, 1 << 16, None,
                                                       lambda V: (V > 0).to(torch.float64), rho, seed=i, reuse_grid=True)
            p_c = np.array(trace["oracle_pcs_exact"][i])
            se64 = np.sqrt((w.numpy() ** 2 * p_c * (1 - p_c) / 64).sum()) / K
            pcs_dev.append((trace["reference_pcs_estimates"][i] - float(est)) / se64)
    assert best_hits >= 0.93 * N * 10, best_hits
    assert abs(np.mean(pcs_dev)) < 1.5 and np.abs(pcs_dev).max() < 4.0, pcs_dev
    assert hits_ref >= 30, (hits_ref, N)        # the CPU oracle reproduces 35 of these 40
    assert hits_oracle >= 34, (hits_oracle, N)  # same pipeline, two optimisers (device L-BFGS vs scipy L-BFGS-B)


def test_reference_recorded_levi_run_through_the_product(crs, golden):
    """Row f2 end to end: the recorded LEVI run replayed through the product (device fit +
    conditioning + ``discrete_levi`` on ``crs_levi_scan``) tracks the recorded decisions and the
    oracle's replay."""
    from contextual_rs_b200.fit import fit_modellist_with_reuse
    g = golden("reference_trace_levi_new.json")
    X = torch.tensor(g["X"], dtype=torch.float64)
    Y = torch.tensor(g["Y"], dtype=torch.float64).view(-1, 1)
    ctx = torch.tensor(g["context_map"], dtype=torch.float64).view(-1, 1)
    w = torch.tensor(g["weights"], dtype=torch.float64)
    n0, K, N = g["num_init"], g["num_arms"], 40
    model, hits_ref, hits_oracle = None, 0, 0
    for i in range(N):
        Xi, Yi = X[: n0 + i], Y[: n0 + i]
        if i % 10 == 0:
            model = fit_modellist_with_reuse(Xi, Yi, K, model, device=DEV)
        else:
            model.condition_on_observations(int(Xi[-1, 0]), Xi[-1:, 1:], Yi[-1:])
        arm, x = crs.discrete_levi(model, ctx, w)
        got = (int(arm), float(x.reshape(-1)[0]))
        hits_ref += got[0] == int(X[n0 + i, 0]) and abs(got[1] - float(X[n0 + i, 1])) < 1e-12
        hits_oracle += got == tuple(g["oracle"][i])
    assert hits_ref >= 23, (hits_ref, N)        # the CPU oracle reproduces 29 of these 40
    assert hits_oracle >= 32, (hits_oracle, N)  # two optimisers (device L-BFGS vs scipy L-BFGS-B)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_decide_host_with_pinned_contexts(crs, dtype):
    """``crs_gao_decide_host`` with a mapped pinned context buffer and a prepare range: the H2D copy of
    the context set rides in the prepare launch (extra CTAs) and the grid kernel's dependency wait
    covers it.  Same decision as with pageable memory (plain copy) and as the oracle, also after the
    host buffer is rewritten in place between calls."""
    from contextual_rs_b200.contextual_rs_strategies import decide
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=24, n_c=700, d=4, n=14, seed=9)
    model = crs.B200ModelListGP(desc, device=DEV, dtype=dtype)
    ref = gp_oracle.model_from_description(desc, dtype=dtype)
    host = ctx.to(dtype).pin_memory()
    for trial in range(3):
        rarm, rx = ocba_oracle.gao_modellist_ref(ref, host.clone(), randomize_ties=False)
        d_pin = decide(model, host, False, 0, 2.0, prepare_arms=(0, 24))
        got_pin = (d_pin.next_arm, d_pin.context)
        d_page = decide(model, host.clone(), False, 0, 2.0, prepare_arms=(0, 24))  # pageable: cudaMemcpyAsync path
        assert got_pin == (d_page.next_arm, d_page.context)
        if dtype == torch.float64:  # fp32 tables may break near-ties differently from torch's fp32 arithmetic
            assert got_pin[0] == int(rarm) and torch.equal(host[got_pin[1]], rx)
        host.copy_(torch.rand(700, 4, generator=torch.Generator().manual_seed(trial), dtype=torch.float64).to(dtype))


def test_transformed_train_inputs_mode(crs):
    """``train_inputs="transformed"`` (BoTorch >= 0.4 eval mode): step 2(b) sees the Normalize-d
    inputs; ``"raw"`` (default) the raw ones.  Both equal the oracle in the same mode."""
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=10, n_c=10, d=1, n=20, kernel="matern52", train="full", seed=4)
    for mode in ("raw", "transformed"):
        d2 = dict(desc)
        d2["train_inputs"] = mode
        model = crs.B200ModelListGP(d2, device=DEV)
        ref = gp_oracle.model_from_description(d2)
        assert model.train_inputs_mode == mode and model.description()["train_inputs"] == mode
        for (a,), (b,) in zip(model.train_inputs, ref.train_inputs):
            assert torch.equal(a.cpu(), b)
        raw_equal = all(torch.equal(a.cpu(), x.to(torch.float64)) for (a,), x in zip(model.train_inputs, desc["train_X"]))
        assert raw_equal == (mode == "raw")
        for kw in ({}, {"use_kde": True, "kernel_scale": 1.0}):
            arm, x = crs.gao_modellist(model, ctx, **kw)
            rarm, rx = ocba_oracle.gao_modellist_ref(ref, ctx, **kw)
            assert int(arm) == int(rarm) and torch.equal(x, rx), (mode, kw)
        m32 = crs.B200ModelListGP(d2, device=DEV, dtype=torch.float32, train_inputs=mode)
        ref32 = gp_oracle.model_from_description(d2, dtype=torch.float32)
        for (a,), (b,) in zip(m32.train_inputs, ref32.train_inputs):
            assert torch.equal(a.cpu(), b)


def _run_tables(crs, mean, var, model, ctx, p_mode=0, kernel_scale=2.0):
    """scan + select on caller-supplied tables through the C ABI."""
    from contextual_rs_b200 import _lib
    lib = _lib.get()
    n_c, K = mean.shape
    dt = _lib.DTYPES[mean.dtype]
    out = {k: torch.zeros(n_c, dtype=torch.int32, device=DEV) for k in ("best", "zarm", "ties")}
    out["minz"] = torch.zeros(n_c, dtype=mean.dtype, device=DEV)
    dec = torch.zeros(64, dtype=torch.uint8, device=DEV)
    ws = torch.zeros(int(lib.crs_scan_workspace_bytes()), dtype=torch.uint8, device=DEV)
    _lib.check(lib.crs_ocba_scan(mean.data_ptr(), var.data_ptr(), n_c, K, K, dt, 0, None, None, None,
                                 out["best"].data_ptr(), out["minz"].data_ptr(), out["zarm"].data_ptr(),
                                 out["ties"].data_ptr(), dec.data_ptr(), ws.data_ptr(), None))
    if model is not None:
        _lib.check(lib.crs_ocba_select(C.byref(model.state), ctx.data_ptr(), mean.data_ptr(), var.data_ptr(),
                                       K, 0, p_mode, kernel_scale, dec.data_ptr(), None))
    torch.cuda.synchronize()
    host = dec.cpu().numpy().tobytes()
    d = _lib.Decision.from_buffer_copy(host)
    if model is not None:  # the fused scan + select launch must give the identical decision
        dec2 = torch.zeros(64, dtype=torch.uint8, device=DEV)
        out2 = {k: torch.zeros_like(v) for k, v in out.items()}
        _lib.check(lib.crs_ocba_scan_select(C.byref(model.state), ctx.data_ptr(), mean.data_ptr(), var.data_ptr(),
                                            n_c, K, 0, p_mode, kernel_scale, out2["best"].data_ptr(),
                                            out2["minz"].data_ptr(), out2["zarm"].data_ptr(), out2["ties"].data_ptr(),
                                            dec2.data_ptr(), ws.data_ptr(), None))
        torch.cuda.synchronize()
        assert dec2.cpu().numpy().tobytes() == host
        for k in out:
            assert torch.equal(out[k], out2[k]), k
    return d, out, dec


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_mock_tables_golden(crs, golden, dtype):
    """The reference's MockModel tables (test/test_contextual_rs_strategies.py:169-176)."""
    g = golden("mock_tables.json")
    mean = torch.tensor(g["mean"], dtype=dtype, device=DEV)
    var = torch.tensor(g["variance"], dtype=dtype, device=DEV)
    ctx = torch.linspace(0, 1, 3, dtype=dtype).unsqueeze(-1)
    zeta = torch.tensor(g["zeta"], dtype=dtype)
    for case in g["decisions"]:
        rows = [torch.cat([ctx[case["next_context"]].view(1, 1).repeat(c, 1), torch.full((3, 1), 0.123, dtype=dtype)])
                for c in case["counts"]]
        desc = {"train_X": rows, "train_Y": [torch.zeros(r.shape[0], 1, dtype=dtype) for r in rows],
                "lengthscale": torch.ones(3, 1), "outputscale": torch.ones(3), "noise": torch.ones(3),
                "mean_const": torch.zeros(3), "normalize": False, "standardize": False}
        model = crs.B200ModelListGP(desc, device=DEV, dtype=dtype)
        d, out, _ = _run_tables(crs, mean, var, model, ctx.to(DEV))
        assert out["best"].cpu().tolist() == g["best_arms"]
        torch.testing.assert_close(out["minz"].cpu(), zeta.min(dim=-1).values, rtol=0, atol=0)
        assert (d.context, d.next_arm, d.tie_count) == (case["next_context"], case["next_arm"], 1)
        assert d.min_zeta == float(zeta.min())


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_gao_iid_known_answer_tables(crs, golden, dtype):
    """zeta scan on the tables behind the reference's pinned C-OCBA answer (arm 2, context 1)."""
    g = golden("gao_iid_seed5.json")
    c = g["cases"][dtype]
    dt = getattr(torch, dtype)
    means, vars_, nobs = (torch.tensor(c[k], dtype=dt) for k in ("means", "vars", "num_obs"))
    p = nobs / nobs.sum()
    mean_t = means.t().contiguous().to(DEV)
    var_t = (vars_ / p).t().contiguous().to(DEV)
    d, out, _ = _run_tables(crs, mean_t, var_t, None, None)
    zeta, order = ocba_oracle.zeta_table(mean_t.cpu(), var_t.cpu())
    flat = int(torch.argmin(zeta.view(-1)))
    assert d.context == flat // 2 == g["expected"][1]
    assert d.zeta_arm == int(order[d.context, flat % 2 + 1])
    assert d.min_zeta == float(zeta.min())
    assert 2 in (d.best_arm, d.zeta_arm)


def test_tie_randomisation_contract(crs):
    """strategies.py:147-154: ties are counted exactly, the r-th tied entry in flat order is
    selectable, and the global RNG is consumed only when ties exist."""
    from contextual_rs_b200 import _lib
    lib = _lib.get()
    mean = torch.tensor([[3.0, 1.0, 2.0], [0.5, 3.5, 2.5], [3.0, 2.0, 1.0], [9.0, 0.0, 1.0]], dtype=torch.float64, device=DEV)
    var = torch.ones(4, 3, dtype=torch.float64, device=DEV)
    d, out, dec = _run_tables(crs, mean, var, None, None)
    zeta, order = ocba_oracle.zeta_table(mean.cpu(), var.cpu())
    tied = (zeta.view(-1) == zeta.min()).nonzero().flatten().tolist()
    assert d.tie_count == len(tied) == 3 and d.min_zeta == 0.5
    assert (d.context, d.zeta_arm) == (0, 2)
    seen = []
    for r in range(3):
        _lib.check(lib.crs_ocba_resolve_tie(mean.data_ptr(), var.data_ptr(), 4, 3, 3, 0, out["best"].data_ptr(),
                                            out["minz"].data_ptr(), out["ties"].data_ptr(), r, dec.data_ptr(), None))
        torch.cuda.synchronize()
        dd = _lib.Decision.from_buffer_copy(dec.cpu().numpy().tobytes())
        seen.append((dd.context, dd.zeta_arm))
    expect = [(f // 2, int(order[f // 2, f % 2 + 1])) for f in tied]
    assert seen == expect
    # end to end: tied GP posteriors (two identical arms), RNG consumed iff randomize_ties
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=3, n_c=6, d=1, n=6, seed=9)
    for key in ("train_X", "train_Y"):
        desc[key] = [desc[key][0], desc[key][1], desc[key][1]]
    for key in ("lengthscale", "outputscale", "noise", "mean_const"):
        desc[key] = desc[key][[0, 1, 1]]
    model = crs.B200ModelListGP(desc, device=DEV)
    torch.manual_seed(5)
    s0 = torch.get_rng_state()
    crs.gao_modellist(model, ctx, randomize_ties=False)
    assert torch.equal(s0, torch.get_rng_state())
    ref = gp_oracle.model_from_description(desc)
    zt, _ = ocba_oracle.zeta_table(ref.posterior(ctx).mean, ref.posterior(ctx).variance)
    has_ties = int((zt.view(-1) == zt.min()).sum()) > 1
    crs.gao_modellist(model, ctx, randomize_ties=True)
    assert torch.equal(s0, torch.get_rng_state()) != has_ties


def test_condition_on_observations_matches_oracle(crs):
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=6, n_c=40, d=2, n=9, seed=4)
    model = crs.B200ModelListGP(desc, device=DEV)
    ref = gp_oracle.model_from_description(desc)
    gen = torch.Generator().manual_seed(0)
    for it in range(12):  # crosses the n_cap = 24 boundary for arm 1 -> storage regrows
        arm = 1 if it % 2 == 0 else int(torch.randint(6, (1,), generator=gen))
        x = ctx[int(torch.randint(40, (1,), generator=gen))].view(1, -1)
        y = torch.randn(1, 1, generator=gen, dtype=torch.float64)
        model.condition_on_observations(arm, x, y)
        ref.models[arm] = ref.models[arm].condition_on_observations(x, y)
        rp, p = ref.posterior(ctx), model.posterior(ctx)
        assert _scaled_err(p.mean, rp.mean) < 1e-10 and _scaled_err(p.variance, rp.variance) < 1e-10
        arm_g, x_g = crs.gao_modellist(model, ctx)
        arm_r, x_r = ocba_oracle.gao_modellist_ref(ref, ctx)
        assert int(arm_g) == int(arm_r) and torch.equal(x_g, x_r)
    assert model.num_observations()[1] >= 15


def test_continuous_context_parity(crs):
    from contextual_rs_b200.synthetic import make_problem
    desc, _ = make_problem(K=8, n_c=16, d=3, n=8, train="rand", seed=6)
    cand = torch.quasirandom.SobolEngine(3, scramble=True, seed=0).draw(512, dtype=torch.float64)
    model = crs.B200ModelListGP(desc, device=DEV)
    ref = gp_oracle.model_from_description(desc)
    got = crs.MinZeta(model)(cand.unsqueeze(1))
    want = ocba_oracle.min_zeta_ref(ref, cand.unsqueeze(1))
    assert got.shape == (512,) and got.device == cand.device
    torch.testing.assert_close(got, want, rtol=1e-9, atol=1e-12)
    assert int(got.argmax()) == int(want.argmax())
    for i in (int(want.argmax()), 0, 17):
        a = crs.find_next_arm_given_context(cand[i].view(1, -1), model, kernel_scale=1.0)
        b = ocba_oracle.find_next_arm_given_context_ref(cand[i].view(1, -1), ref, kernel_scale=1.0)
        assert a.dim() == 0 and int(a) == int(b)
    with pytest.raises(AssertionError):
        crs.MinZeta(model)(cand)


def test_philox_words(crs, golden):
    from contextual_rs_b200 import _lib
    lib = _lib.get()
    for v in golden("philox_kat.json")["vectors"]:
        ctr = torch.from_numpy(np.array([int(x, 16) for x in v["ctr"]], dtype=np.uint32).view(np.int32)).to(DEV)
        out = torch.zeros(4, dtype=torch.int32, device=DEV)
        _lib.check(lib.crs_philox_words(ctr.data_ptr(), 1, int(v["key"][0], 16), int(v["key"][1], 16), out.data_ptr(), None))
        assert [f"{int(x) & 0xffffffff:08x}" for x in out.cpu()] == v["out"]
    rng = np.random.default_rng(0)
    ctr = rng.integers(0, 2 ** 32, size=(1000, 4), dtype=np.uint64).astype(np.uint32)
    out = torch.zeros(4000, dtype=torch.int32, device=DEV)
    _lib.check(lib.crs_philox_words(torch.from_numpy(ctr.view(np.int32)).to(DEV).data_ptr(), 1000, 123, 456, out.data_ptr(), None))
    want = np.stack(philox_ref.philox4x32_10(ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3], 123, 456), axis=1)
    assert np.array_equal(out.cpu().numpy().view(np.uint32).reshape(1000, 4), want)


@pytest.mark.parametrize("dtype,S", [(torch.float64, 8192), (torch.float32, 5003)])
def test_pcs_counts_parity(crs, dtype, S):
    from contextual_rs_b200.generalized_pcs import pcs_counts
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=12, n_c=48, d=2, n=10, seed=5)
    model = crs.B200ModelListGP(desc, device=DEV, dtype=dtype)
    p = model.posterior(ctx.to(dtype))
    mean, var = p.mean.contiguous(), p.variance.contiguous()
    stats = torch.zeros(4, dtype=torch.int64, device=DEV)
    cnt = pcs_counts(model, mean, var, S, seed=77, offset=3, stats=stats, sampler="flat").cpu().numpy()
    want = philox_ref.pcs_counts_philox_ref(mean.cpu().numpy(), var.cpu().numpy(), S, seed=77, offset=3)
    # shared Philox stream: identical up to borderline draws (MUFU vs libm rounding of the normals)
    assert np.abs(cnt - want).max() <= 2 and (cnt != want).sum() <= 3
    exact = pcs_oracle.pcs_exact_quadrature(mean.double().cpu().numpy(), var.double().cpu().numpy())
    se = np.sqrt(np.maximum(exact * (1 - exact), 1e-4) / S)
    assert np.all(np.abs(cnt / S - exact) < 5 * se)
    drawn, skipped = stats.cpu().tolist()[:2]
    assert drawn + skipped == S * 12 * 48  # executed + skipped = nominal S * K * n_c
    # sharding invariance: context / arm offsets only re-key the stream
    half = pcs_counts(model, mean[24:], var[24:], S, seed=77, offset=3, ctx_offset=24, sampler="flat").cpu().numpy()
    assert np.array_equal(half, cnt[24:])
    other = pcs_counts(model, mean, var, S, seed=78, offset=3, sampler="flat").cpu().numpy()
    assert not np.array_equal(other, cnt)


def test_pcs_screening_is_exact(crs):
    """Arms that cannot win (|z| <= 5.7683) are skipped without changing any count."""
    from contextual_rs_b200.generalized_pcs import pcs_counts
    gen = torch.Generator().manual_seed(1)
    n_c, K, S = 32, 40, 4096
    mean = torch.randn(n_c, K, generator=gen, dtype=torch.float64)
    mean[:, 20:] -= 50.0  # hopeless arms
    var = 0.05 + torch.rand(n_c, K, generator=gen, dtype=torch.float64)
    stats = torch.zeros(4, dtype=torch.int64, device=DEV)
    cnt = pcs_counts(None, mean.to(DEV), var.to(DEV), S, seed=5, stats=stats, sampler="flat").cpu().numpy()
    want = philox_ref.pcs_counts_philox_ref(mean.numpy(), var.numpy(), S, seed=5)
    assert np.abs(cnt - want).max() <= 2
    drawn, skipped = stats.cpu().tolist()[:2]
    assert skipped >= S * n_c * 20 and drawn + skipped == S * n_c * K
    # degenerate rows: a dominant best arm gives PCS = 1 exactly, equal arms give ~1/K
    mean2 = torch.zeros(4, 5, dtype=torch.float64); mean2[:2, 3] = 100.0
    var2 = torch.ones(4, 5, dtype=torch.float64)
    for sampler in ("flat", "tree"):
        c2 = pcs_counts(None, mean2.to(DEV), var2.to(DEV), S, seed=1, sampler=sampler).cpu().numpy()
        assert c2[0] == S and c2[1] == S and abs(c2[2] / S - 0.2) < 0.04


@pytest.mark.parametrize("dtype,S", [(torch.float64, 8192), (torch.float32, 5003)])
def test_pcs_tree_parity(crs, dtype, S):
    """Max-tree stream: the kernel's lazy walk reproduces the brute-force restatement
    (oracle/pcs_tree_ref.py: every node and leaf generated) count for count, up to borderline
    MUFU-vs-libm cases, and both estimate the quadrature value."""
    from contextual_rs_b200.generalized_pcs import pcs_counts
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=12, n_c=48, d=2, n=10, seed=5)
    model = crs.B200ModelListGP(desc, device=DEV, dtype=dtype)
    p = model.posterior(ctx.to(dtype))
    mean, var = p.mean.contiguous(), p.variance.contiguous()
    stats = torch.zeros(4, dtype=torch.int64, device=DEV)
    cnt = pcs_counts(model, mean, var, S, seed=77, offset=3, stats=stats, sampler="tree").cpu().numpy()
    want = pcs_tree_ref.pcs_counts_tree_ref(mean.cpu().numpy(), var.cpu().numpy(), S, seed=77, offset=3)
    assert np.abs(cnt - want).max() <= 2 and (cnt != want).sum() <= 3
    exact = pcs_oracle.pcs_exact_quadrature(mean.double().cpu().numpy(), var.double().cpu().numpy())
    se = np.sqrt(np.maximum(exact * (1 - exact), 1e-4) / S)
    assert np.all(np.abs(cnt / S - exact) < 5 * se)
    roots, expansions, directs = stats.cpu().tolist()[:3]
    assert roots == S * 48 and 0 < directs < roots and 0 < expansions < 12 * roots
    half = pcs_counts(model, mean[24:], var[24:], S, seed=77, offset=3, ctx_offset=24, sampler="tree").cpu().numpy()
    assert np.array_equal(half, cnt[24:])  # sharding invariance: context offsets only re-key the stream
    other = pcs_counts(model, mean, var, S, seed=78, offset=3, sampler="tree").cpu().numpy()
    assert not np.array_equal(other, cnt)
    again = pcs_counts(model, mean, var, S, seed=77, offset=3, sampler="tree").cpu().numpy()  # deterministic
    assert np.array_equal(again, cnt)


@pytest.mark.parametrize("K", [2, 3, 5, 6, 7, 10, 21, 22, 69, 70, 261, 262, 1000, 1029, 1030])
def test_pcs_tree_shapes(crs, K):
    """Every tree depth (1..6), full and ragged last levels, fp64 and fp32 tables, ragged sample
    counts, rows with hopeless arms and exact mean ties."""
    from contextual_rs_b200.generalized_pcs import pcs_counts
    gen = torch.Generator().manual_seed(K)
    n_c = 6
    S = 3001 if K <= 300 else 700
    mean = 0.6 * torch.randn(n_c, K, generator=gen, dtype=torch.float64)
    var = 0.05 + 1.5 * torch.rand(n_c, K, generator=gen, dtype=torch.float64)
    if K > 4:
        mean[1, K // 2:] -= 40.0          # hopeless arms
        mean[2, 1] = mean[2, 0] = mean[2].max() + 0.1  # tied best arms: first index is the predicted best
        var[3] = 1e-10                    # variance floor: near-deterministic rows
    for dtype in (torch.float64, torch.float32):
        m, v = mean.to(dtype), var.to(dtype)
        cnt = pcs_counts(None, m.to(DEV), v.to(DEV), S, seed=31 + K, offset=1, ctx_offset=7, sampler="tree").cpu().numpy()
        want = pcs_tree_ref.pcs_counts_tree_ref(m.numpy(), v.numpy(), S, seed=31 + K, offset=1, ctx_offset=7)
        assert np.abs(cnt - want).max() <= 2 and (cnt != want).sum() <= 2, (K, dtype, cnt, want)
    exact = pcs_oracle.pcs_exact_quadrature(mean.numpy(), var.numpy())
    se = np.sqrt(np.maximum(exact * (1 - exact), 1e-3) / S)
    assert np.all(np.abs(want / S - exact) < 5 * se)


def test_pcs_tree_inverse_normal(crs):
    """The kernel's fp32 inverse normal zfun(t), Phi(z) = exp(-t): equal to the oracle's restatement
    of the same operation sequence up to the MUFU approximations (lg2 / ex2 / sqrt), within 3e-6 of
    scipy over the whole range the sampler produces, and non-increasing (exact pruning relies on it)."""
    from contextual_rs_b200 import _lib
    from scipy.special import ndtri_exp
    rng = np.random.default_rng(0)
    t = np.concatenate([10 ** rng.uniform(-11, -3, 200000), rng.uniform(1e-3, 0.7, 200000),
                        rng.uniform(0.6, 5, 200000), rng.uniform(5, 25, 200000),
                        np.sort(10 ** rng.uniform(-11, 1.8, 400000))]).astype(np.float32)
    td = torch.from_numpy(t).to(DEV)
    out = torch.empty_like(td)
    _lib.check(_lib.get().crs_debug_zfun(td.data_ptr(), td.numel(), out.data_ptr(), None))
    z = out.cpu().numpy()
    assert np.isfinite(z).all()
    assert np.abs(z - pcs_tree_ref.zfun32(t)).max() < 2e-6
    ok = t <= 25.0  # beyond w = 27 (z < -7.1) the function saturates by design
    assert np.abs(z - ndtri_exp(-t.astype(np.float64)))[ok].max() < 3.5e-6
    zs = z[800000:]
    assert (np.diff(zs[0::2]) <= 0).all() and (np.diff(zs[1::2]) <= 0).all()  # both variants, sorted t
    assert (np.diff(zs) <= 2e-7).all()  # and they agree where both apply


def test_pcs_tree_limits(crs):
    from contextual_rs_b200.generalized_pcs import pcs_counts
    mean = torch.randn(2, 4110, dtype=torch.float64, device=DEV)
    var = torch.ones(2, 4110, dtype=torch.float64, device=DEV)
    with pytest.raises(ValueError):
        pcs_counts(None, mean, var, 64, seed=1, sampler="tree")
    a = pcs_counts(None, mean, var, 256, seed=1)          # auto -> cdf stream (any K >= 2); flat also serves K > 4102
    b = pcs_counts(None, mean, var, 256, seed=1, sampler="cdf")
    assert torch.equal(a, b)
    f = pcs_counts(None, mean, var, 256, seed=1, sampler="flat")
    assert int(f.max()) <= 256 and abs(int(f.sum()) - int(a.sum())) <= 40  # two contexts, PCS ~ 1/4110: both near 0
    one = pcs_counts(None, mean[:, :1].contiguous(), var[:, :1].contiguous(), 100, seed=1)  # K = 1: no competitor
    assert one.tolist() == [100, 100]
    from contextual_rs_b200 import _lib
    lib = _lib.get()
    cnt = torch.zeros(2, dtype=torch.int32, device=DEV)
    assert lib.crs_pcs_counts_tree(mean.data_ptr(), var.data_ptr(), 2, 4110, 4110, 0, 64, 1, 0, 0, cnt.data_ptr(), None, None) == 2


def test_estimate_current_generalized_pcs_dropin(crs):
    from contextual_rs_b200.synthetic import make_problem
    f = lambda X: (X > 0).to(torch.float64)
    vals = []
    for n in (4, 60):
        desc, ctx = make_problem(K=3, n_c=10, d=2, n=n, seed=6)
        model = crs.B200ModelListGP(desc, device=DEV)
        ref = gp_oracle.model_from_description(desc)
        arm_set = torch.arange(3.0, dtype=torch.float64).view(-1, 1)
        per_ctx = []
        rho = lambda x: (per_ctx.append(x), x.mean(dim=-2))[1]
        v = crs.estimate_current_generalized_pcs(model, arm_set, ctx, 16384, None, f, rho, seed=4)
        assert v.shape == torch.Size([]) and 0.0 <= float(v) <= 1.0  # test_generalized_pcs.py:290-301
        assert per_ctx[0].shape == (10, 1)
        rp = ref.posterior(ctx)
        exact = pcs_oracle.pcs_exact_quadrature(rp.mean.numpy(), rp.variance.numpy())
        se = np.sqrt(np.maximum(exact * (1 - exact), 1e-4) / 16384)
        assert np.all(np.abs(per_ctx[0].squeeze(-1).cpu().numpy() - exact) < 5 * se)
        worst = crs.estimate_current_generalized_pcs(model, arm_set, ctx, 16384, None, f,
                                                     lambda x: x.min(dim=-2).values, seed=4)
        assert abs(float(worst) - float(per_ctx[0].min())) < 1e-12
        vals.append(float(v))
        with pytest.raises(NotImplementedError):
            crs.estimate_current_generalized_pcs(model, arm_set, ctx.repeat(3, 1, 1), 8, None, f, rho)
        with pytest.raises(NotImplementedError):
            crs.estimate_current_generalized_pcs(model, arm_set, ctx, 8, None, f, rho, use_approximation=True)
        with pytest.raises(NotImplementedError):
            crs.estimate_current_generalized_pcs(model, arm_set, ctx, 8, None, lambda X: X.clamp_min(0), rho)
        with pytest.raises(AssertionError):
            crs.estimate_current_generalized_pcs(model, arm_set.view(-1), ctx, 8, None, f, rho)
    assert vals[1] > vals[0]  # more data -> higher PCS (test_generalized_pcs.py:303-329)
    # default seeding consumes the global torch RNG like the reference's rsample does
    torch.manual_seed(0)
    a = crs.estimate_current_generalized_pcs(model, arm_set, ctx, 64, None, f, lambda x: x.mean(dim=-2))
    torch.manual_seed(0)
    b = crs.estimate_current_generalized_pcs(model, arm_set, ctx, 64, None, f, lambda x: x.mean(dim=-2))
    assert float(a) == float(b)


def test_full_size_properties(crs):
    """Size-independent properties at BASELINE config-3 scale (256 arms x 8,192 contexts):
    a random slice of the grid against the oracle, scan outputs against torch reductions of the
    same device tables, and shard-invariance of the PCS counts."""
    from contextual_rs_b200.generalized_pcs import pcs_counts
    from contextual_rs_b200.synthetic import make_config
    desc, ctx = make_config("c3")
    model = crs.B200ModelListGP(desc, device=DEV)
    p = model.posterior(ctx)
    sel = torch.randperm(8192, generator=torch.Generator().manual_seed(0))[:64]
    mu, var = gp_oracle.posterior_grid_batched(desc, ctx[sel])
    assert _scaled_err(p.mean.cpu()[sel], mu) < 1e-10 and _scaled_err(p.variance.cpu()[sel], var) < 1e-10
    arm, x = crs.gao_modellist(model, ctx)
    ws = model.workspace(8192)
    mean_d, var_d = ws["mean"], ws["var"]
    best = mean_d.argmax(dim=-1)
    assert torch.equal(ws["best_arm"].long(), best)
    mb = mean_d.gather(1, best[:, None]); vb = var_d.gather(1, best[:, None])
    zeta = (mb - mean_d).pow(2) / (vb + var_d)
    zeta.scatter_(1, best[:, None], float("inf"))
    assert torch.equal(ws["min_zeta"], zeta.min(dim=-1).values)
    c_star = int(zeta.min(dim=-1).values.argmin())
    assert torch.equal(x, ctx[c_star])
    S = 1024
    for sampler in ("flat", "tree"):
        full = pcs_counts(model, mean_d, var_d, S, seed=3, sampler=sampler)
        parts = torch.cat([pcs_counts(model, mean_d[i:i + 2048], var_d[i:i + 2048], S, seed=3, ctx_offset=i, sampler=sampler)
                           for i in range(0, 8192, 2048)])
        assert torch.equal(full, parts)
        assert int(full.min()) >= 0 and int(full.max()) <= S


def test_fp64_kernel_math(crs):
    """The grid kernel's table-based exp and Newton sqrt against libm (fp64)."""
    from contextual_rs_b200 import _lib
    lib = _lib.get()
    gen = torch.Generator().manual_seed(0)
    x = torch.cat([torch.zeros(1, dtype=torch.float64), 10.0 ** (torch.rand(20000, generator=gen, dtype=torch.float64) * 12 - 8),
                   torch.tensor([1e-300, 1e-30, 1.0, 4.0, 700.0, 1e4, 1e6], dtype=torch.float64)]).to(DEV)
    out = torch.empty_like(x)

    def run(kernel, variant, inp=x):
        _lib.check(lib.crs_debug_corr(inp.data_ptr(), inp.numel(), kernel, variant, out.data_ptr(), None))
        torch.cuda.synchronize()
        return out.cpu().clone()

    xc = x.cpu()
    for iters_variant in (2, 3):
        s = run(0, iters_variant)
        ref = (xc + 1e-290).sqrt()
        assert float(((s - ref).abs() / ref).max()) < (3e-16 if iters_variant == 3 else 1e-12)
    small = xc.clamp(max=690.0).to(DEV)  # the kernel's exp clamps below exp(-693) ~ 1e-301
    e = run(0, 4, small)
    ref = torch.exp(-small.cpu())
    assert float(((e - ref).abs() / ref).max()) < 5e-16
    for kernel in (0, 1):
        got, lib_dev = run(kernel, 0), run(kernel, 1)
        r = xc.sqrt()
        want = ((1 + 5 ** 0.5 * r + 5.0 / 3.0 * xc) * torch.exp(-(5 ** 0.5) * r)) if kernel == 0 else torch.exp(-0.5 * xc)
        assert float((got - want).abs().max()) < 1e-15  # absolute: correlations live in [0, 1]
        assert float((lib_dev - want).abs().max()) < 1e-15
        big = want > 1e-280
        # exp(-s) amplifies the rounding of its argument by s: bound relative to (2 + s)
        arg = (5 ** 0.5) * r if kernel == 0 else 0.5 * xc
        assert float(((got - want).abs() / (want * (2 + arg)))[big].max()) < 4e-16


def test_sequential_loop_matches_oracle_loop(crs):
    """The cached-grid / dirty-column loop (SequentialRS) against the reference's loop structure
    (experiments/wsc_experiments/main.py:345-470) restated with the oracle: identical decision
    sequence, identical best-arm readout, PCS within MC error of exact quadrature."""
    from contextual_rs_b200.rs_loop import SequentialRS
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=8, n_c=50, d=2, n=6, seed=11)
    model = crs.B200ModelListGP(desc, device=DEV)
    ref = gp_oracle.model_from_description(desc)
    gen = torch.Generator().manual_seed(3)
    noise = torch.randn(40, generator=gen, dtype=torch.float64)

    def truth(arm, x, it):
        return torch.sin(10.0 * torch.cat([torch.tensor([arm / 7.0], dtype=torch.float64), x])).sum() + 0.1 * noise[it]

    S = 8192
    loop = SequentialRS(model, ctx, kernel_scale=1.0, num_pcs_samples=S, pcs_seed=5)
    for it in range(20):
        rarm, rx = ocba_oracle.gao_modellist_ref(ref, ctx, kernel_scale=1.0)
        rp = ref.posterior(ctx)
        out = loop.step(lambda arm, x: truth(arm, x, it))
        assert out["next_arm"] == int(rarm) and torch.equal(out["next_context"], rx), it
        assert torch.equal(out["best_arms"].cpu(), rp.mean.argmax(dim=-1))
        exact = pcs_oracle.pcs_exact_quadrature(rp.mean.numpy(), rp.variance.numpy()).mean()
        assert abs(float(out["pcs"]) - exact) < 5 * np.sqrt(0.25 / (S * 50)) + 2e-3
        y = truth(int(rarm), rx, it)
        ref.models[int(rarm)] = ref.models[int(rarm)].condition_on_observations(rx.view(1, -1), y.view(1, 1))
    # the cached grid equals a from-scratch posterior of the updated model
    p = model.posterior(ctx)
    assert torch.equal(p.mean, loop.ws["mean"]) and torch.equal(p.variance, loop.ws["var"])
    rp = ref.posterior(ctx)
    assert _scaled_err(p.mean, rp.mean) < 1e-10 and _scaled_err(p.variance, rp.variance) < 1e-10


# ------------------------------------------------------------------------------------------ LEVI
@pytest.mark.parametrize("K,n_c,d,n,kernel", [(10, 10, 1, 20, "matern52"), (100, 1024, 4, 20, "matern52"),
                                              (37, 333, 3, 11, "rbf"), (2, 5, 2, 6, "matern52")])
def test_levi_matches_oracle_fp64(crs, K, n_c, d, n, kernel):
    """discrete_levi / PredictiveEI / the EI table (contextual_rs/levi.py:132-227): indices
    bit-exact, EI within 1e-9 of the table's scale (the posterior itself is within 1e-10)."""
    from contextual_rs_b200.levi import PredictiveEI, _compute_all_EI_reviewer, discrete_levi
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=K, n_c=n_c, d=d, n=n, kernel=kernel, seed=3)
    model = crs.B200ModelListGP(desc, device=DEV, dtype=torch.float64)
    ref = gp_oracle.model_from_description(desc)
    X = ctx.unsqueeze(-2)
    want = levi_oracle.compute_all_ei_reviewer(X, ref)
    got = _compute_all_EI_reviewer(X, model)
    assert got.shape == (n_c, K) and got.device == ctx.device
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= 1e-9 * scale
    torch.testing.assert_close(PredictiveEI(model)(X), want.max(dim=-1).values, rtol=0, atol=1e-9 * scale)
    arm, x = discrete_levi(model, ctx)
    rarm, rx = levi_oracle.discrete_levi_ref(ref, ctx)
    assert arm.dtype == torch.long and arm.dim() == 0
    assert int(arm) == int(rarm) and torch.equal(x, rx)
    g = torch.Generator().manual_seed(4)
    w = torch.rand(n_c, dtype=torch.float64, generator=g)
    arm, x = discrete_levi(model, ctx, w)
    rarm, rx = levi_oracle.discrete_levi_ref(ref, ctx, w)
    assert int(arm) == int(rarm) and torch.equal(x, rx)


def test_levi_tables_and_fp32(crs):
    """Kernel-level check on raw tables (ties at the top of a row, clamped variances) + fp32."""
    import ctypes as C
    from contextual_rs_b200 import _lib
    from contextual_rs_b200.levi import discrete_levi
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=6, n_c=40, d=2, n=9, seed=8)
    for dtype, tol in ((torch.float64, 1e-12), (torch.float32, 2e-5)):
        model = crs.B200ModelListGP(desc, device=DEV, dtype=dtype)
        g = torch.Generator().manual_seed(2)
        mean = torch.randn(40, 6, generator=g, dtype=torch.float64).to(dtype)
        mean[3, 4] = mean[3].max()          # exact tie with the row maximum
        mean[5, :] = 1.5                    # all arms tied
        var = (0.05 + torch.rand(40, 6, generator=g, dtype=torch.float64)).to(dtype)
        var[7, 2] = 1e-12                   # below the 1e-9 clamp
        st = model.state
        noise_orig = (model.description()["noise"].double() * model.description()["y_std"].double() ** 2)
        want = levi_oracle.ei_from_tables(mean.unsqueeze(-2), var.unsqueeze(-2), noise_orig.to(torch.float64))
        ei = torch.empty(40, 6, dtype=dtype, device=DEV)
        mx = torch.empty(40, dtype=dtype, device=DEV)
        am = torch.empty(40, dtype=torch.int32, device=DEV)
        dec = torch.zeros(24, dtype=torch.uint8, device=DEV)
        md, vd = mean.to(DEV), var.to(DEV)
        _lib.check(_lib.get().crs_levi_scan(C.byref(st), md.data_ptr(), vd.data_ptr(), 40, 6, None, ei.data_ptr(), 6,
                                            mx.data_ptr(), am.data_ptr(), dec.data_ptr(), None))
        torch.cuda.synchronize()
        scale = float(want.abs().max())
        assert float((ei.cpu().double() - want.double()).abs().max()) <= tol * scale
        if dtype == torch.float64:
            wmax = want.max(dim=-1)
            assert torch.equal(am.cpu().long(), wmax.indices)
            d = _lib.LeviDecision.from_buffer_copy(dec.cpu().numpy().tobytes())
            assert d.context == int(wmax.values.argmax()) and d.arm == int(wmax.indices[d.context])
    arm, x = discrete_levi(crs.B200ModelListGP(desc, device=DEV, dtype=torch.float32), ctx.float())
    assert 0 <= int(arm) < 6 and x.shape == (2,)


# ------------------------------------------------------------------------- MinZeta gradient
@pytest.mark.parametrize("K,d,n,kernel,dtype", [(6, 2, 9, "matern52", torch.float64), (11, 8, 18, "matern52", torch.float64),
                                                (5, 3, 40, "rbf", torch.float64), (6, 2, 9, "matern52", torch.float32)])
def test_min_zeta_gradient_matches_autograd_of_the_oracle(crs, K, d, n, kernel, dtype):
    """d(-min zeta)/d(context): the analytic backward kernel against torch autograd through the
    oracle's MinZeta (continuous_context.py:57-74), plus a finite-difference spot check and a few
    optimiser steps over the contexts (what optimize_acqf does, main.py:252-263)."""
    from contextual_rs_b200.synthetic import make_problem
    desc, _ = make_problem(K=K, n_c=16, d=d, n=n, kernel=kernel, train="rand", seed=9)
    model = crs.B200ModelListGP(desc, device=DEV, dtype=dtype)
    ref = gp_oracle.model_from_description(desc)
    g = torch.Generator().manual_seed(1)
    X0 = torch.rand(200, 1, d, generator=g, dtype=torch.float64)
    Xr = X0.clone().requires_grad_(True)
    want = ocba_oracle.min_zeta_ref(ref, Xr)
    want.sum().backward()
    Xg = X0.to(dtype).clone().requires_grad_(True)
    acq = crs.MinZeta(model)
    got = acq(Xg)
    assert got.requires_grad and got.shape == (200,)
    (got * torch.arange(1, 201, dtype=dtype)).sum().backward()  # non-trivial upstream gradient
    mine = Xg.grad.double() / torch.arange(1, 201, dtype=torch.float64).view(-1, 1, 1)
    scale = float(Xr.grad.abs().max())
    tol = 1e-8 if dtype == torch.float64 else 2e-3
    assert float((mine - Xr.grad).abs().max()) <= tol * scale
    if dtype == torch.float64:
        # central finite differences of the CUDA forward itself
        eps = 1e-6
        for k in range(min(d, 3)):
            dx = torch.zeros(1, 1, d, dtype=dtype); dx[..., k] = eps
            fd = (acq(X0 + dx) - acq(X0 - dx)) / (2 * eps)
            assert float((fd - mine[:, 0, k]).abs().max()) <= 1e-5 * scale
        # gradient ascent on the acquisition value moves every context uphill
        X = X0[:32].clone().requires_grad_(True)
        opt = torch.optim.SGD([X], lr=1e-3 / max(scale, 1.0))
        v0 = acq(X).detach().clone()
        for _ in range(5):
            opt.zero_grad()
            (-acq(X).sum()).backward()
            opt.step()
        assert bool((acq(X).detach() >= v0 - 1e-12).all())


# --------------------------------------------------------------------------------- lookahead PCS
def test_lookahead_pcs_modellist(crs):
    """estimate_lookahead_generalized_pcs_modellist (generalized_pcs.py:203-315): shape and [0, 1]
    like the reference's test (test/test_generalized_pcs.py:162-240), plus every value within MC
    error of exact quadrature on the oracle's fantasy tables, and the model left untouched."""
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=3, n_c=4, d=2, n=10, train="rand", seed=13)
    model = crs.B200ModelListGP(desc, device=DEV)
    ref = gp_oracle.model_from_description(desc)
    before = model.posterior(ctx)
    g = torch.Generator().manual_seed(0)
    cand = torch.cat([torch.randint(0, 3, (5, 1, 1), generator=g).double(), torch.rand(5, 1, 2, generator=g, dtype=torch.float64)], dim=-1)
    func_I = lambda X: (X > 0).to(torch.float64)
    rho = lambda X: X.mean(dim=-2)
    S = 1 << 14
    tol = 5 * np.sqrt(0.25 / (S * 4)) + 2e-3
    # certainty equivalent
    pcs = crs.estimate_lookahead_generalized_pcs_modellist(cand, model, None, ctx, S, None, func_I, rho, seed=3)
    assert pcs.shape == (5,) and torch.equal(pcs, pcs.clamp(min=0, max=1))
    want = torch.zeros(5, dtype=torch.float64)
    for i, f, mean, var in pcs_oracle.lookahead_pcs_tables_ref(ref, cand, ctx):
        want[i] += pcs_oracle.pcs_exact_quadrature(mean.numpy(), var.numpy()).mean()
    assert float((pcs - want).abs().max()) < tol
    # two fantasies per candidate with shared standard-normal draws
    zb = torch.randn(5, 2, generator=g, dtype=torch.float64)
    pcs2 = crs.estimate_lookahead_generalized_pcs_modellist(cand, model, 2, ctx, S, None, func_I, rho, seed=4, fantasy_base=zb)
    assert pcs2.shape == (5,) and torch.equal(pcs2, pcs2.clamp(min=0, max=1))
    want2 = torch.zeros(5, dtype=torch.float64)
    for i, f, mean, var in pcs_oracle.lookahead_pcs_tables_ref(ref, cand, ctx, zb):
        want2[i] += pcs_oracle.pcs_exact_quadrature(mean.numpy(), var.numpy()).mean() / 2
    assert float((pcs2 - want2).abs().max()) < tol
    # sampler object with a sample_shape, global RNG for the fantasies
    class _Sampler:
        sample_shape = torch.Size([2])
    pcs3 = crs.estimate_lookahead_generalized_pcs_modellist(cand, model, _Sampler(), ctx, 1024, None, func_I, rho)
    assert pcs3.shape == (5,) and torch.equal(pcs3, pcs3.clamp(min=0, max=1))
    with pytest.raises(NotImplementedError):
        crs.estimate_lookahead_generalized_pcs_modellist(cand, model, None, ctx, 16, None, func_I, rho, use_approximation=True)
    # the fantasies were rolled back
    after = model.posterior(ctx)
    assert torch.equal(before.mean, after.mean) and torch.equal(before.variance, after.variance)
    assert model.num_observations() == [10, 10, 10]


# ------------------------------------------------------------------------------ hyper-parameter fit
def _fit_problem(K=5, n=14, d=3, seed=2):
    g = torch.Generator().manual_seed(seed)
    rows, ys = [], []
    for k in range(K):
        nk = n + k  # ragged arms
        Xk = torch.rand(nk, d, generator=g, dtype=torch.float64)
        yk = torch.sin(6.0 * Xk + k).sum(-1) * (0.5 + k) + 0.05 * torch.randn(nk, generator=g, dtype=torch.float64) + 3.0 * k
        rows.append(torch.cat([torch.full((nk, 1), float(k), dtype=torch.float64), Xk], dim=-1))
        ys.append(yk)
    return torch.cat(rows), torch.cat(ys).reshape(-1, 1), K


@pytest.mark.parametrize("kernel", ["matern52", "rbf"])
def test_gp_mll_value_and_gradient_match_oracle_autograd(crs, kernel):
    """crs_gp_mll against the oracle's torch MLL + autograd at random hyper-parameters: value and
    gradient to 1e-9 relative (the objective IS pinned to rounding; the optimiser is not)."""
    import ctypes as C
    from contextual_rs_b200 import _lib
    from oracle import fit_oracle
    X, Y, K = _fit_problem()
    d = X.shape[1] - 1
    g = torch.Generator().manual_seed(5)
    tx = [X[X[:, 0] == k][:, 1:] for k in range(K)]
    ty = [Y[X[:, 0] == k] for k in range(K)]
    desc = {"train_X": tx, "train_Y": ty, "kernel": kernel, "normalize": True, "standardize": True,
            "lengthscale": 0.2 + torch.rand(K, d, generator=g, dtype=torch.float64),
            "outputscale": 0.5 + torch.rand(K, generator=g, dtype=torch.float64),
            "noise": 0.01 + 0.1 * torch.rand(K, generator=g, dtype=torch.float64),
            "mean_const": 0.3 * torch.randn(K, generator=g, dtype=torch.float64)}
    model = crs.B200ModelListGP(desc, device=DEV)
    val = torch.empty(K, dtype=torch.float64, device=DEV)
    grad = torch.empty(K, d + 3, dtype=torch.float64, device=DEV)
    _lib.check(_lib.get().crs_gp_mll(C.byref(model.state), 0, K, val.data_ptr(), grad.data_ptr(), None))
    torch.cuda.synchronize()
    for k in range(K):
        mins, ranges, ym, ys = fit_oracle.transforms(tx[k], ty[k].reshape(-1))
        Xn, yt = (tx[k] - mins) / ranges, (ty[k].reshape(-1) - ym) / ys
        p = [desc[n][k].clone().requires_grad_(True) for n in ("lengthscale", "outputscale", "noise", "mean_const")]
        v = fit_oracle.exact_mll(Xn, yt, *p, kernel)
        gs = torch.autograd.grad(v, p)
        want = torch.cat([gs[0], gs[1].reshape(1), gs[2].reshape(1), gs[3].reshape(1)])
        assert abs(float(val[k]) - float(v.detach())) <= 1e-10 * abs(float(v.detach()))
        assert float((grad[k].cpu() - want).abs().max()) <= 1e-9 * float(want.abs().max())
    # not positive definite -> -inf, zero gradient, no crash
    model.noise.fill_(-10.0)
    _lib.check(_lib.get().crs_gp_mll(C.byref(model.state), 0, K, val.data_ptr(), grad.data_ptr(), None))
    assert bool(torch.isinf(val).all()) and float(grad.abs().max()) == 0.0


def test_fit_modellist_reaches_the_scipy_optimum(crs):
    """fit_modellist / fit_modellist_with_reuse (experiment_utils.py:15-48, 68-121): the batched
    device L-BFGS and the oracle's scipy L-BFGS-B start from gpytorch's initial values and must end
    at the same optimum: objective within 1e-6, hyper-parameters within 2 %, posterior within 1e-3
    of scale.  (Optimiser tolerance, not arithmetic parity: see contextual_rs_b200/fit.py.)"""
    from contextual_rs_b200.fit import fit_modellist, fit_modellist_with_reuse
    from oracle import fit_oracle
    X, Y, K = _fit_problem()
    model = fit_modellist(X, Y, K, device=DEV)
    rdesc, robj = fit_oracle.fit_modellist_ref(X, Y, K)
    got = model.description()
    obj = model.fit_info["objective"]
    assert float((obj - robj).max()) <= 1e-6, (obj, robj)   # never worse than scipy
    assert float((obj - robj).abs().max()) <= 1e-5
    for name in ("lengthscale", "outputscale", "mean_const"):
        torch.testing.assert_close(got[name], rdesc[name], rtol=2e-2, atol=2e-3)
    torch.testing.assert_close(got["noise"], rdesc["noise"], rtol=5e-2, atol=2e-4)
    for name in ("norm_mins", "norm_ranges", "y_mean", "y_std"):
        torch.testing.assert_close(got[name], rdesc[name], rtol=1e-13, atol=1e-13)
    ctx = torch.rand(64, 3, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    p, rp = model.posterior(ctx), gp_oracle.model_from_description(rdesc).posterior(ctx)
    assert _scaled_err(p.mean, rp.mean) < 1e-3 and _scaled_err(p.variance, rp.variance) < 1e-2
    # reuse: add one observation to arm 2 only -> only arm 2 is re-fitted, the others are kept bit for bit
    X2 = torch.cat([X, torch.tensor([[2.0, 0.3, 0.6, 0.9]], dtype=torch.float64)])
    Y2 = torch.cat([Y, torch.tensor([[6.5]], dtype=torch.float64)])
    model2 = fit_modellist_with_reuse(X2, Y2, K, old_model=model, device=DEV)
    d2 = model2.description()
    keep = [0, 1, 3, 4]
    for name in ("lengthscale", "outputscale", "noise", "mean_const", "y_mean", "y_std"):
        assert torch.equal(d2[name][keep], got[name][keep])
    assert model2.num_observations()[2] == model.num_observations()[2] + 1
    assert int(model2.fit_info["iterations"][2]) > 0 and int(model2.fit_info["iterations"][[0, 1, 3, 4]].sum()) == 0
    arm, x = crs.gao_modellist(model2, ctx)
    assert 0 <= int(arm) < K


# ----------------------------------------------------------- every dispatch branch of the scan / PCS
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("K", [2, 33, 129, 256, 300, 600, 1000, 1024, 1500, 3000])
def test_scan_all_row_widths(crs, K, dtype):
    """crs_ocba_scan on raw tables for every kernel variant (register rows K <= 256, TMA-staged
    rows with 1 or 4 warps per row up to 1024, generic beyond): best arm, min zeta, zeta arm and
    tie count bit-identical to torch reductions of `squared_diffs / added_vars`
    (contextual_rs_strategies.py:139-149), including exact ties and a tied row maximum."""
    from contextual_rs_b200 import _lib
    lib = _lib.get()
    n_c = 257
    g = torch.Generator().manual_seed(K)
    mean = torch.randn(n_c, K, generator=g, dtype=torch.float64).to(dtype)
    var = (0.1 + torch.rand(n_c, K, generator=g, dtype=torch.float64)).to(dtype)
    if K >= 4:
        mean[3, K - 1] = mean[3, 1]; var[3, K - 1] = var[3, 1]       # two arms with identical zeta
        mean[5, K // 2] = mean[5].max()                               # tied row maximum (lowest id wins)
        mean[7] = mean[7, 0]                                          # all arms equal: zeta = 0 everywhere
    md, vd = mean.to(DEV), var.to(DEV)
    out = {k: torch.zeros(n_c, dtype=torch.int32, device=DEV) for k in ("best", "zarm", "ties")}
    minz = torch.zeros(n_c, dtype=dtype, device=DEV)
    dec = torch.zeros(64, dtype=torch.uint8, device=DEV)
    wsb = torch.zeros(int(lib.crs_scan_workspace_bytes()), dtype=torch.uint8, device=DEV)
    _lib.check(lib.crs_ocba_scan(md.data_ptr(), vd.data_ptr(), n_c, K, K, _lib.DTYPES[dtype], 0, None, None, None,
                                 out["best"].data_ptr(), minz.data_ptr(), out["zarm"].data_ptr(), out["ties"].data_ptr(),
                                 dec.data_ptr(), wsb.data_ptr(), None))
    torch.cuda.synchronize()
    best = mean.argmax(dim=-1)
    assert torch.equal(out["best"].cpu().long(), best)
    mb, vb = mean.gather(1, best[:, None]), var.gather(1, best[:, None])
    zeta = (mb - mean).pow(2) / (vb + var)
    zeta.scatter_(1, best[:, None], float("inf"))
    zmin = zeta.min(dim=-1).values
    assert torch.equal(minz.cpu(), zmin)
    assert torch.equal(out["ties"].cpu().long(), (zeta == zmin[:, None]).sum(dim=-1))
    za = out["zarm"].cpu().long()
    assert bool((zeta.gather(1, za[:, None]).squeeze(1) == zmin).all())
    d = _lib.Decision.from_buffer_copy(dec.cpu().numpy().tobytes())
    gmin = zmin.min()
    assert d.min_zeta == float(gmin) and d.context == int((zmin == gmin).nonzero()[0])
    assert d.tie_count == int((zeta == gmin).sum())


def test_wide_model_decision_and_pcs(crs):
    """K = 300 arms: the fused scan + select of the TMA-staged variant (4 warps per row) and a PCS
    plan of 75 arm blocks with K not a multiple of the block size beyond 1024 arms."""
    from contextual_rs_b200.generalized_pcs import pcs_counts
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=300, n_c=96, d=3, n=7, seed=17)
    model = crs.B200ModelListGP(desc, device=DEV)
    ref = gp_oracle.model_from_description(desc)
    arm, x = crs.gao_modellist(model, ctx)
    rarm, rx = ocba_oracle.gao_modellist_ref(ref, ctx)
    assert int(arm) == int(rarm) and torch.equal(x, rx)
    arm, x = crs.gao_modellist(model, ctx, use_kde=True, kernel_scale=1.0)
    rarm, rx = ocba_oracle.gao_modellist_ref(ref, ctx, use_kde=True, kernel_scale=1.0)
    assert int(arm) == int(rarm) and torch.equal(x, rx)
    g = torch.Generator().manual_seed(3)
    mean = torch.randn(5, 1301, generator=g, dtype=torch.float64)
    var = 0.05 + torch.rand(5, 1301, generator=g, dtype=torch.float64)
    S = 512
    cnt = pcs_counts(None, mean.to(DEV), var.to(DEV), S, seed=9, sampler="flat").cpu().numpy()
    want = philox_ref.pcs_counts_philox_ref(mean.numpy(), var.numpy(), S, seed=9)
    assert np.abs(cnt - want).max() <= 2 and (cnt != want).sum() <= 2
    cnt = pcs_counts(None, mean.to(DEV), var.to(DEV), S, seed=9, sampler="tree").cpu().numpy()  # depth 6
    want = pcs_tree_ref.pcs_counts_tree_ref(mean.numpy(), var.numpy(), S, seed=9)
    assert np.abs(cnt - want).max() <= 2 and (cnt != want).sum() <= 2
