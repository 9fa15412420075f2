"""CPU-side checks: the C-ABI library loads and exports every symbol of include/crs_b200.h, the
ctypes mirrors match the header, host-side argument handling, and the product never falls back
to the CPU.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session", autouse=True)
def built_library():
    from contextual_rs_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "contextual_rs_b200", "csrc"), "-j8"], check=True)
    return _lib.get()


def test_header_symbols_all_exported(built_library):
    from contextual_rs_b200 import _lib
    header = open(os.path.join(ROOT, "include", "crs_b200.h")).read()
    declared = set(re.findall(r"CRS_API\s+[\w\s\*]+?\b(crs_\w+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(built_library, name), name
    assert built_library.crs_abi_version() == 1
    assert built_library.crs_scan_workspace_bytes() > 0
    assert built_library.crs_status_string(2).startswith(b"unsupported")


def test_struct_layouts_match_header():
    from contextual_rs_b200 import _lib
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "crs_b200.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(crs_gp_state), offsetof(crs_gp_state, train_x),
             offsetof(crs_gp_state, status), sizeof(crs_decision), offsetof(crs_decision, context),
             offsetof(crs_decision, reserved), sizeof(crs_decide_ws));
      return 0;
    }'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "t.c"), "-o", os.path.join(td, "t")], check=True)
        out = subprocess.run([os.path.join(td, "t")], capture_output=True, text=True, check=True).stdout.split()
    got = [C.sizeof(_lib.GPState), _lib.GPState.train_x.offset, _lib.GPState.status.offset,
           C.sizeof(_lib.Decision), _lib.Decision.context.offset, _lib.Decision.reserved.offset,
           C.sizeof(_lib.DecideWS)]
    assert [int(x) for x in out] == got


def test_argument_validation_without_gpu(built_library):
    lib = built_library
    assert lib.crs_gp_prepare(None, 0, 1, None) == 1  # CRS_ERR_INVALID_ARG
    assert lib.crs_posterior_grid(None, None, 4, 0, 1, None, None, 1, 0, None) == 1
    assert lib.crs_ocba_scan(None, None, 4, 3, 3, 0, 0, None, None, None, None, None, None, None, None, None, None) == 1
    assert lib.crs_pcs_counts(None, None, 4, 3, 3, 0, 16, 0, 0, 0, 0, None, None, None) == 1
    assert lib.crs_philox_words(None, 1, 0, 0, None, None) == 1


def test_no_cpu_fallback():
    import contextual_rs_b200 as crs
    from contextual_rs_b200.synthetic import make_problem
    desc, ctx = make_problem(K=3, n_c=4, d=1, n=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crs.B200ModelListGP(desc, device="cpu")
    # the product package never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "contextual_rs_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from contextual_rs_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="is missing"):
        _lib.get()


def test_strict_indicator_probe():
    from contextual_rs_b200.generalized_pcs import _is_strict_indicator
    assert _is_strict_indicator(lambda X: (X > 0).to(torch.float64))
    assert _is_strict_indicator(lambda X: (X > 0).to(dtype=torch.float))
    assert not _is_strict_indicator(lambda X: (X >= 0).to(torch.float64))
    assert not _is_strict_indicator(lambda X: torch.sigmoid(X))
    assert not _is_strict_indicator(lambda X: X.sum())


def test_pcs_sampler_argument_handling(built_library):
    """Host-side validation of ``pcs_counts`` happens before any launch (no GPU needed): unknown
    sampler, tree sampler on an arm slice or beyond 4,102 arms, too small a ``stats`` tensor."""
    from contextual_rs_b200.generalized_pcs import TREE_MAX_ARMS, pcs_counts
    mean, var = torch.zeros(2, 6), torch.ones(2, 6)
    with pytest.raises(ValueError):
        pcs_counts(None, mean, var, 8, seed=0, sampler="bogus")
    with pytest.raises(ValueError):
        pcs_counts(None, mean, var, 8, seed=0, arm_offset=4, sampler="tree")
    wide = torch.zeros(1, TREE_MAX_ARMS + 1)
    with pytest.raises(ValueError):
        pcs_counts(None, wide, wide + 1, 8, seed=0, sampler="tree")
    with pytest.raises(ValueError):
        pcs_counts(None, mean, var, 8, seed=0, stats=torch.zeros(2, dtype=torch.int64))
    # the C entry validates too: NULL tables / too many arms / K = 1 (no competitor: use the flat stream)
    lib = built_library
    assert lib.crs_pcs_counts_tree(None, None, 4, 3, 3, 0, 16, 0, 0, 0, None, None, None) == 1
    assert lib.crs_debug_zfun(None, 4, None, None) == 1


def test_model_rejects_unknown_train_inputs_mode():
    from contextual_rs_b200.model import B200ModelListGP
    with pytest.raises((ValueError, RuntimeError)):
        B200ModelListGP({"train_X": [torch.zeros(2, 1)], "train_Y": [torch.zeros(2, 1)], "train_inputs": "bogus",
                         "lengthscale": torch.ones(1, 1), "outputscale": torch.ones(1), "noise": torch.ones(1),
                         "mean_const": torch.zeros(1)}, device="cuda")


class _FakeParam:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class MaternKernel:  # the extractor keys on the class name, as with gpytorch.kernels.MaternKernel
    def __init__(self, lengthscale):
        self.lengthscale, self.nu = lengthscale, 2.5


class _FakeNormalize:
    def __init__(self, mins, rng):
        self.bounds, self._mins, self._rng = torch.stack([mins, mins + rng]), mins, rng

    def untransform(self, X):
        return X * self._rng + self._mins


class _FakeSingleTaskGP:
    """Duck-typed stand-in for a fitted BoTorch SingleTaskGP(Normalize, Standardize) in eval mode: the
    attribute paths of SURVEY.md §8b.  ``transformed_in_eval`` mimics BoTorch >= 0.4 (eval mode holds
    the Normalize-d train inputs, ``_has_transformed_inputs``, raw copy in ``_original_train_inputs``);
    False mimics the behaviour of the reference's recorded runs.  ``conditioned`` mimics a model
    returned by ``condition_on_observations``: ``_original_train_inputs`` is the STALE pre-conditioning
    copy (one row short).  Switching modes is forbidden: the extractor must not touch the model."""

    def __init__(self, X, Y, ls, osc, noise, mean, transformed_in_eval, conditioned=False):
        self._X, self.training, self._tie = X, False, transformed_in_eval
        mins, rng = X.min(dim=0).values, X.max(dim=0).values - X.min(dim=0).values
        ym, ys = Y.mean(), Y.std()
        self.train_targets = ((Y - ym) / ys).reshape(-1)
        self.covar_module = _FakeParam(base_kernel=MaternKernel(ls.reshape(1, -1)), outputscale=osc)
        self.likelihood = _FakeParam(noise=noise.reshape(1))
        self.mean_module = _FakeParam(constant=mean)
        self.input_transform = _FakeNormalize(mins, rng)
        self.outcome_transform = _FakeParam(means=ym.reshape(1, 1), stdvs=ys.reshape(1, 1))
        self._mins, self._rng = mins, rng
        if transformed_in_eval:
            self._has_transformed_inputs = True
            self._set_transformed_inputs = lambda: None
            self._original_train_inputs = X[:-1] if conditioned else X

    def train(self):
        raise AssertionError("description_from_botorch must not switch the caller's model to train mode")

    def eval(self):
        raise AssertionError("description_from_botorch must not toggle the caller's model")

    @property
    def train_inputs(self):
        if self._tie:
            if not hasattr(self, "_Xt"):  # BoTorch stores the transformed copy; it is not rebuilt per access
                self._Xt = (self._X - self._mins) / self._rng
            return (self._Xt,)
        return (self._X,)


@pytest.mark.parametrize("transformed_in_eval,conditioned", [(False, False), (True, False), (True, True)])
def test_description_from_botorch_attribute_paths(transformed_in_eval, conditioned):
    """``description_from_botorch`` against a duck-typed BoTorch model: hyper-parameters, transform
    constants and raw data come out right WITHOUT touching the model (no train()/eval()), the
    eval-mode semantics of ``train_inputs`` are detected, a conditioned model's stale
    ``_original_train_inputs`` is not trusted, and the description drives the oracle to the same
    posterior as a directly built one."""
    from contextual_rs_b200.model import botorch_fingerprint, description_from_botorch
    from oracle import gp_oracle
    gen = torch.Generator().manual_seed(0)
    subs, want = [], []
    for k in range(3):
        X = torch.rand(7 + k, 2, generator=gen, dtype=torch.float64)
        Y = torch.randn(7 + k, 1, generator=gen, dtype=torch.float64) * 3 + 5
        ls = 0.2 + torch.rand(2, generator=gen, dtype=torch.float64)
        osc, noise, mean = torch.tensor(1.3 + k), torch.tensor(0.01 * (k + 1)), torch.tensor(0.1 * k)
        subs.append(_FakeSingleTaskGP(X, Y, ls, osc.double(), noise.double(), mean.double(), transformed_in_eval, conditioned))
        want.append(gp_oracle.OracleSingleTaskGP(X, Y, ls, osc.double(), noise.double(), mean.double()))
    fake = _FakeParam(models=subs)
    desc = description_from_botorch(fake)
    assert desc["train_inputs"] == ("transformed" if transformed_in_eval else "raw")
    assert desc["kernel"] == "matern52"
    for k in range(3):  # raw inputs recovered (exactly, or through untransform for the conditioned model)
        torch.testing.assert_close(desc["train_X"][k], subs[k]._X, rtol=0, atol=1e-15)
    got = gp_oracle.model_from_description(desc)
    ctx = torch.rand(11, 2, generator=gen, dtype=torch.float64)
    pw = gp_oracle.OracleModelListGP(*want).posterior(ctx)
    pg = got.posterior(ctx)
    torch.testing.assert_close(pg.mean, pw.mean, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(pg.variance, pw.variance, rtol=1e-10, atol=1e-12)
    for (a,), m in zip(got.train_inputs, subs):  # what OCBA step 2(b) will see == what the model shows in eval mode
        torch.testing.assert_close(a, m.train_inputs[0], rtol=0, atol=1e-15)
    # the cached-handle key changes when the caller mutates the model in place, and only then
    fp0 = botorch_fingerprint(fake)
    assert botorch_fingerprint(fake) == fp0
    subs[1].train_targets.mul_(1.0)  # in-place edit bumps the tensor's version counter
    assert botorch_fingerprint(fake) != fp0
    # inconsistent model (targets one short of inputs): refuse instead of packing garbage
    subs[0].train_targets = subs[0].train_targets[:-1]
    with pytest.raises(ValueError):
        description_from_botorch(fake)


def test_initialize_q_batch_keeps_the_best_and_samples_without_replacement():
    """contextual_rs_b200.optimize.initialize_q_batch (BoTorch's heuristic): the arg-max is always among the
    starts, starts are distinct raw candidates, equal values fall back to a random subset, n >= n_raw returns all."""
    from contextual_rs_b200.optimize import initialize_q_batch
    g = torch.Generator().manual_seed(0)
    X = torch.rand(64, 3, dtype=torch.float64, generator=g)
    Y = torch.randn(64, dtype=torch.float64, generator=g)
    for seed in range(5):
        x0 = initialize_q_batch(X, Y, 8, eta=2.0, generator=torch.Generator().manual_seed(seed))
        assert x0.shape == (8, 3)
        rows = [int((X == r).all(dim=-1).nonzero()[0]) for r in x0]
        assert len(set(rows)) == 8 and int(torch.argmax(Y)) in rows
    x0 = initialize_q_batch(X, torch.zeros(64, dtype=torch.float64), 8, generator=torch.Generator().manual_seed(1))
    assert x0.shape == (8, 3)
    assert initialize_q_batch(X[:5], Y[:5], 8).shape == (5, 3)


def test_batched_lbfgs_box_constraints_on_cpu():
    """fit.batched_lbfgs with the box projection of optimize.py on a CPU objective: independent quadratics whose
    unconstrained minima lie inside / outside the unit box end at the projected minima."""
    from contextual_rs_b200.fit import batched_lbfgs
    target = torch.tensor([[0.3, 0.7], [1.5, -0.4], [0.5, 2.0]], dtype=torch.float64)
    scale = torch.tensor([[1.0, 10.0], [3.0, 0.5], [2.0, 2.0]], dtype=torch.float64)

    class Obj:
        d = 2

        def __call__(self, x):
            r = x - target
            return 0.5 * (scale * r * r).sum(-1), scale * r

    lo, up = torch.zeros(2, dtype=torch.float64), torch.ones(2, dtype=torch.float64)
    project = lambda x, _d: torch.minimum(torch.maximum(x, lo), up)  # noqa: E731
    pg = lambda x, g, _d: torch.where(((x <= lo) & (g > 0)) | ((x >= up) & (g < 0)), torch.zeros_like(g), g)  # noqa: E731
    x, f, g, it = batched_lbfgs(Obj(), torch.full((3, 2), 0.5, dtype=torch.float64), project=project, projected_gradient=pg,
                                gtol=1e-10, ftol=1e-15)
    torch.testing.assert_close(x, target.clamp(0.0, 1.0), rtol=0, atol=1e-7)


def test_optimize_acqf_without_a_backward_uses_central_differences():
    """An acquisition function that returns values without a graph (like PredictiveEI) is refined with
    central differences: a smooth bump inside the box is found; one on the boundary ends on the boundary."""
    from contextual_rs_b200.optimize import optimize_acqf
    bounds = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]], dtype=torch.float64)
    peak = torch.tensor([0.3, 0.8, 0.55], dtype=torch.float64)

    def acq(X):  # batch x 1 x d -> batch, deliberately detached
        with torch.no_grad():
            return torch.exp(-4.0 * ((X.squeeze(-2) - peak) ** 2).sum(-1))

    cand, val, det = optimize_acqf(acq, bounds, q=1, num_restarts=6, raw_samples=64, seed=3, return_details=True)
    torch.testing.assert_close(cand.reshape(-1), peak, rtol=0, atol=1e-4)
    assert abs(float(val) - 1.0) < 1e-7 and float(val) >= det["raw_best"]
    peak2 = torch.tensor([1.4, 0.2, -0.3], dtype=torch.float64)  # outside the box

    def acq2(X):
        with torch.no_grad():
            return -((X.squeeze(-2) - peak2) ** 2).sum(-1)

    cand2, _ = optimize_acqf(acq2, bounds, q=1, num_restarts=4, raw_samples=32, seed=1)
    torch.testing.assert_close(cand2.reshape(-1), torch.tensor([1.0, 0.2, 0.0], dtype=torch.float64), rtol=0, atol=1e-5)
