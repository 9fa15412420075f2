"""Hyper-parameter fitting of the per-arm GPs on the B200: drop-ins for ``fit_modellist`` and
``fit_modellist_with_reuse`` of ``contextual_rs/experiment_utils.py:15-48, 68-121``.

The reference builds one BoTorch ``SingleTaskGP`` per arm (``Standardize(m=1)``, ``Normalize(d)``,
ScaleKernel(ARD Matern-5/2) with the library's default Gamma priors and noise constraint) and calls
``fit_gpytorch_mll(ExactMarginalLogLikelihood(...))``, i.e. scipy L-BFGS-B on

    -( log N(y~ | m 1, o C(l) + s2 I) + log Gamma(l; 3, 6) + log Gamma(o; 2, 0.15)
       + log Gamma(s2; 1.1, 0.05) ) / n

over (softplus^-1 l, softplus^-1 o, s2 >= 1e-4, m), one arm at a time, from the default initial
values (l = o = log 2, s2 = 2, m = 0).  Here the objective and its gradient come from ONE launch of
``crs_gp_mll`` for all arms, and the K independent problems are advanced in lockstep by a batched
L-BFGS (two-loop recursion and Armijo back-tracking as [K, P] tensor operations on the device).
The optimiser TRAJECTORY is not the reference's (scipy's More-Thuente line search is not
replicated); the objective is, to rounding, and the tests check that both reach the same optimum.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib
from .model import B200ModelListGP

# BoTorch SingleTaskGP defaults of the reference's era (mirrored at contextual_rs/models/lce_gp.py:101-129)
LENGTHSCALE_PRIOR = (3.0, 6.0)     # GammaPrior(concentration, rate)
OUTPUTSCALE_PRIOR = (2.0, 0.15)
NOISE_PRIOR = (1.1, 0.05)
MIN_NOISE = 1e-4                   # GreaterThan(MIN_INFERRED_NOISE_LEVEL, transform=None)
INIT_NOISE = (NOISE_PRIOR[0] - 1.0) / NOISE_PRIOR[1]  # initial_value = prior mode = 2.0
MIN_RANGE, MIN_STDV = 1e-8, 1e-8   # Normalize(min_range), Standardize(min_stdv)


def _softplus_inv(x: float) -> float:
    return math.log(math.expm1(x))


def default_raw_parameters(K: int, d: int, device) -> Tensor:
    """Raw parameters ``[K, d + 3]`` = (raw lengthscales, raw outputscale, noise, mean) at
    gpytorch's initial values: raw = 0 (softplus -> log 2), noise = prior mode, mean = 0."""
    raw = torch.zeros(K, d + 3, dtype=torch.float64, device=device)
    raw[:, d + 1] = INIT_NOISE
    return raw


class _Objective:
    """-(MLL + log priors) / n for every arm and its gradient in the raw parameters."""

    def __init__(self, model: B200ModelListGP):
        self.m = model
        K, d = model.num_arms, model.dim
        self.K, self.d = K, d
        self.val = torch.empty(K, dtype=torch.float64, device=model.device)
        self.grad = torch.zeros(K, d + 3, dtype=torch.float64, device=model.device)
        self.n = model.n_obs.to(torch.float64).clamp_min(1.0)
        self.evaluations = 0

    def natural(self, raw: Tensor):
        d = self.d
        sp = torch.nn.functional.softplus
        return sp(raw[:, :d]), sp(raw[:, d]), raw[:, d + 1], raw[:, d + 2]

    def __call__(self, raw: Tensor):
        m, d = self.m, self.d
        ls, osc, noise, mean = self.natural(raw)
        m.lengthscale.copy_(ls)
        m.outputscale.copy_(osc)
        m.noise.copy_(noise)
        m.mean_const.copy_(mean)
        with torch.cuda.device(m.device):
            rc = _lib.get().crs_gp_mll(C.byref(m.state), 0, self.K, self.val.data_ptr(), self.grad.data_ptr(),
                                       _lib.stream_ptr(m.device))
        _lib.check(rc, "crs_gp_mll")
        self.evaluations += 1
        (a_l, b_l), (a_o, b_o), (a_n, b_n) = LENGTHSCALE_PRIOR, OUTPUTSCALE_PRIOR, NOISE_PRIOR
        logp = ((a_l - 1) * ls.log() - b_l * ls).sum(-1) + (a_o - 1) * osc.log() - b_o * osc \
            + (a_n - 1) * noise.log() - b_n * noise
        g = self.grad.clone()
        g[:, :d] += (a_l - 1) / ls - b_l
        g[:, d] += (a_o - 1) / osc - b_o
        g[:, d + 1] += (a_n - 1) / noise - b_n
        g[:, :d] *= torch.sigmoid(raw[:, :d])       # d softplus(u) / du
        g[:, d] *= torch.sigmoid(raw[:, d])
        f = -(self.val + logp) / self.n
        g = -g / self.n.unsqueeze(-1)
        bad = ~torch.isfinite(f)
        f = torch.where(bad, torch.full_like(f, float("inf")), f)
        g = torch.where(bad.unsqueeze(-1), torch.zeros_like(g), g)
        return f, g


def _project(raw: Tensor, d: int) -> Tensor:
    raw[:, d + 1].clamp_(min=MIN_NOISE)
    return raw


def _projected_gradient(raw: Tensor, g: Tensor, d: int) -> Tensor:
    """Gradient with the component that pushes the noise below its bound removed."""
    pg = g.clone()
    at_bound = (raw[:, d + 1] <= MIN_NOISE) & (g[:, d + 1] > 0)
    pg[:, d + 1] = torch.where(at_bound, torch.zeros_like(pg[:, d + 1]), pg[:, d + 1])
    return pg


def batched_lbfgs(obj, raw0: Tensor, active: Optional[Tensor] = None, maxiter: int = 500,
                  history: int = 10, gtol: float = 1e-6, ftol: float = 1e-12, max_backtracks: int = 25,
                  project=None, projected_gradient=None):
    """K independent L-BFGS minimisations advanced in lockstep.  Stopping rules are scipy
    L-BFGS-B's (projected-gradient sup-norm, relative decrease) with 10x / 2000x tighter
    tolerances than its defaults (1e-5, 2.2e-9): an objective launch costs < 1 ms for all arms.
    ``obj(x) -> (f [K], g [K, P])``; ``project(x, d)`` / ``projected_gradient(x, g, d)`` encode the
    feasible set (default: the noise bound of the hyper-parameter fit; ``contextual_rs_b200.optimize``
    passes box bounds) — ``d`` is ``obj.d``."""
    d = obj.d
    _project = project if project is not None else globals()["_project"]
    _projected_gradient = projected_gradient if projected_gradient is not None else globals()["_projected_gradient"]
    x = _project(raw0.clone(), d)
    f, g = obj(x)
    K = x.shape[0]
    live = torch.ones(K, dtype=torch.bool, device=x.device) if active is None else active.clone()
    live &= torch.isfinite(f)
    hist: List = []
    iters = torch.zeros(K, dtype=torch.int32, device=x.device)
    retry = torch.zeros(K, dtype=torch.bool, device=x.device)  # line search failed: next step is steepest descent
    for _ in range(maxiter):
        pg = _projected_gradient(x, g, d)
        live &= pg.abs().amax(dim=-1) > gtol
        if not bool(live.any()):
            break
        # two-loop recursion, batched over arms (pairs skipped by an arm carry rho = 0)
        q = pg.clone()
        alphas = []
        for s, y, rho in reversed(hist):
            a = rho * (s * q).sum(-1)
            q -= a.unsqueeze(-1) * y
            alphas.append(a)
        if hist:
            s, y, _ = hist[-1]
            yy = (y * y).sum(-1)
            gamma = torch.where(yy > 0, (s * y).sum(-1) / yy.clamp_min(1e-300), torch.ones_like(yy))
        else:
            gamma = 1.0 / pg.abs().sum(-1).clamp_min(1.0)  # scipy's first step: 1 / ||g||_1
        r = gamma.unsqueeze(-1) * q
        for (s, y, rho), a in zip(hist, reversed(alphas)):
            b = rho * (y * r).sum(-1)
            r += s * (a - b).unsqueeze(-1)
        dirn = -r
        gd = (pg * dirn).sum(-1)
        reset = ~(gd < 0) | retry
        dirn = torch.where(reset.unsqueeze(-1), -pg, dirn)
        gd = torch.where(reset, -(pg * pg).sum(-1), gd)
        # Armijo back-tracking with a per-arm step
        t = torch.ones(K, dtype=torch.float64, device=x.device)
        x_new, f_new, g_new = x.clone(), f.clone(), g.clone()
        pending = live.clone()
        for _bt in range(max_backtracks):
            trial = _project(x + t.unsqueeze(-1) * dirn, d)
            ft, gt = obj(trial)
            ok = pending & (ft <= f + 1e-4 * t * gd)
            x_new = torch.where(ok.unsqueeze(-1), trial, x_new)
            f_new = torch.where(ok, ft, f_new)
            g_new = torch.where(ok.unsqueeze(-1), gt, g_new)
            pending &= ~ok
            if not bool(pending.any()):
                break
            t = torch.where(pending, t * 0.5, t)
        moved = live & ~pending
        # a failed quasi-Newton line search (kinks of a min over arms, stale curvature) gets ONE steepest-descent
        # retry with its curvature pairs dropped before the problem is declared finished
        failed = live & pending
        second = failed & retry
        retry = failed & ~retry
        if bool(retry.any()):
            keep = (~retry).unsqueeze(-1)
            hist = [(s_ * keep, y_ * keep, rho_ * (~retry)) for s_, y_, rho_ in hist]
        s, y = x_new - x, g_new - g
        sy = (s * y).sum(-1)
        good = moved & (sy > 1e-10 * (y * y).sum(-1).clamp_min(1e-300))
        rho = torch.where(good, 1.0 / sy.clamp_min(1e-300), torch.zeros_like(sy))
        hist.append((torch.where(good.unsqueeze(-1), s, torch.zeros_like(s)),
                     torch.where(good.unsqueeze(-1), y, torch.zeros_like(y)), rho))
        if len(hist) > history:
            hist.pop(0)
        small = (f - f_new) <= ftol * torch.maximum(torch.maximum(f.abs(), f_new.abs()), torch.ones_like(f))
        iters += moved.to(torch.int32)
        x, f, g = x_new, f_new, g_new
        live &= (moved & ~small) | retry  # failed twice in a row, or stopped improving: done
        live &= ~second
    return x, f, g, iters


def fit_hyperparameters(model: B200ModelListGP, arms: Optional[Sequence[int]] = None, raw0: Optional[Tensor] = None,
                        maxiter: int = 500, max_retries: int = 3) -> Dict:
    """Maximise every (selected) arm's exact MLL + priors in place and re-prepare the model.

    Returns ``{"objective": [K] final -(MLL + priors)/n, "iterations": [K], "evaluations": int,
    "raw": [K, d+3]}``.  Arms not in ``arms`` keep their hyper-parameters."""
    K, d = model.num_arms, model.dim
    obj = _Objective(model)
    keep = (model.lengthscale.clone(), model.outputscale.clone(), model.noise.clone(), model.mean_const.clone())
    raw = default_raw_parameters(K, d, model.device) if raw0 is None else raw0.to(model.device, torch.float64).clone()
    active = torch.ones(K, dtype=torch.bool, device=model.device)
    if arms is not None:
        active[:] = False
        active[torch.as_tensor(list(arms), dtype=torch.long, device=model.device)] = True
        # frozen arms: raw parameters that reproduce their current values (they are never moved)
        sp_inv = lambda v: torch.where(v > 20, v, torch.log(torch.expm1(v.clamp_max(20.0))))
        cur = torch.cat([sp_inv(keep[0]), sp_inv(keep[1]).unsqueeze(-1), keep[2].unsqueeze(-1), keep[3].unsqueeze(-1)], dim=-1)
        raw = torch.where(active.unsqueeze(-1), raw, cur)
    raw, f, g, iters = batched_lbfgs(obj, raw, active=active, maxiter=maxiter)
    # experiment_utils.py:107-113: a fit that fails (NotPSDError in the reference; here a non-finite
    # objective, i.e. the arm's Cholesky broke down at its starting point or along the way) is retried
    # from hyper-parameters sampled from the priors (`sample_all_priors`), a few times, arm by arm
    retried: List[int] = []
    for attempt in range(max_retries):
        bad = active & ~torch.isfinite(f)
        if not bool(bad.any()):
            break
        retried = sorted(set(retried) | set(torch.nonzero(bad).flatten().tolist()))
        gen = torch.Generator().manual_seed(1234 + attempt)
        sp_inv = lambda v: torch.where(v > 20, v, torch.log(torch.expm1(v.clamp_max(20.0))))  # noqa: E731
        gamma = lambda conc, rate, shape: torch.distributions.Gamma(conc, rate).sample(shape)  # noqa: E731
        with torch.random.fork_rng():
            torch.manual_seed(int(gen.initial_seed()))
            ls0 = gamma(LENGTHSCALE_PRIOR[0], LENGTHSCALE_PRIOR[1], (K, d)).to(torch.float64)
            os0 = gamma(OUTPUTSCALE_PRIOR[0], OUTPUTSCALE_PRIOR[1], (K,)).to(torch.float64)
            nz0 = gamma(NOISE_PRIOR[0], NOISE_PRIOR[1], (K,)).to(torch.float64).clamp_min(10 * MIN_NOISE)
        fresh = torch.cat([sp_inv(ls0), sp_inv(os0).unsqueeze(-1), nz0.unsqueeze(-1), torch.zeros(K, 1, dtype=torch.float64)],
                          dim=-1).to(model.device)
        start = torch.where(bad.unsqueeze(-1), fresh, raw)
        raw2, f2, g2, it2 = batched_lbfgs(obj, start, active=bad, maxiter=maxiter)
        better = bad & torch.isfinite(f2)
        raw = torch.where(better.unsqueeze(-1), raw2, raw)
        f = torch.where(better, f2, f)
        iters = torch.where(better, it2, iters)
    still = active & ~torch.isfinite(f)
    if bool(still.any()):
        raise RuntimeError("hyper-parameter fit failed (kernel matrix not positive definite at every start) for arms "
                           f"{torch.nonzero(still).flatten().tolist()}")
    ls, osc, noise, mean = obj.natural(raw)
    sel = active
    model.lengthscale.copy_(torch.where(sel.unsqueeze(-1), ls, keep[0]))
    model.outputscale.copy_(torch.where(sel, osc, keep[1]))
    model.noise.copy_(torch.where(sel, noise, keep[2]))
    model.mean_const.copy_(torch.where(sel, mean, keep[3]))
    model.prepare(relearn=False)
    model.check_status()
    return {"objective": f.cpu(), "iterations": iters.cpu(), "evaluations": obj.evaluations, "raw": raw.cpu(),
            "retried_arms": retried}


def _transforms(Xk: Tensor, Yk: Tensor):
    """Normalize(d) bounds and Standardize(m=1) constants learnt from one arm's data."""
    mins = Xk.min(dim=0).values
    ranges = Xk.max(dim=0).values - mins
    ranges = torch.where(ranges < MIN_RANGE, torch.ones_like(ranges), ranges)
    ym = Yk.mean()
    ys = Yk.std() if Yk.numel() > 1 else torch.ones((), dtype=Yk.dtype)
    if not bool(ys >= MIN_STDV):
        ys = torch.ones((), dtype=Yk.dtype)
    return mins, ranges, ym, ys


def fit_modellist_with_reuse(X: Tensor, Y: Tensor, num_arms: int, old_model: Optional[B200ModelListGP] = None,
                             device="cuda", dtype: torch.dtype = torch.float64, kernel: str = "matern52") -> B200ModelListGP:
    r"""Fit one GP per arm; arms whose number of observations did not change since ``old_model``
    keep its hyper-parameters and transforms (``contextual_rs/experiment_utils.py:68-121``).  As in
    the reference (``num_last_train_inputs``, ``:66, 93-97, 113-117``) "did not change" refers to the
    arm's last FIT: observations appended by ``condition_on_observations`` since then make the arm
    due for a re-fit (``model.n_at_fit`` carries the counts).

    Args:
        X: all evaluated arm-context pairs, first column = arm index.
        Y: the corresponding evaluations (``N x 1``).
        num_arms: number of arms.
        old_model: the previous :class:`B200ModelListGP` (or ``None``: fit everything).
    """
    X64, Y64 = X.detach().cpu().to(torch.float64), Y.detach().cpu().to(torch.float64).reshape(-1)
    d = X64.shape[-1] - 1
    tx, ty, mins, rngs, ym, ys = [], [], [], [], [], []
    ls = torch.full((num_arms, d), math.log(2.0), dtype=torch.float64)
    osc = torch.full((num_arms,), math.log(2.0), dtype=torch.float64)
    noise = torch.full((num_arms,), INIT_NOISE, dtype=torch.float64)
    mc = torch.zeros(num_arms, dtype=torch.float64)
    refit: List[int] = []
    old = old_model.description() if old_model is not None and old_model.num_arms == num_arms else None
    old_n_fit = getattr(old_model, "n_at_fit", None) if old is not None else None
    n_at_fit: List[int] = []
    for i in range(num_arms):
        mask = X64[..., 0] == i
        Xi, Yi = X64[mask][..., 1:], Y64[mask]
        if Yi.numel() == 0:
            raise RuntimeError("Missing data!")  # experiment_utils.py:98-100
        tx.append(Xi)
        ty.append(Yi.reshape(-1, 1))
        n_old = old_n_fit[i] if old_n_fit is not None else (old["train_X"][i].shape[0] if old is not None else -1)
        n_at_fit.append(int(Yi.numel()))
        if old is not None and n_old == Yi.numel():  # :93-97 reuse
            ls[i], osc[i], noise[i], mc[i] = old["lengthscale"][i], old["outputscale"][i], old["noise"][i], old["mean_const"][i]
            mins.append(old["norm_mins"][i]); rngs.append(old["norm_ranges"][i])
            ym.append(old["y_mean"][i]); ys.append(old["y_std"][i])
        else:
            a, b, c, e = _transforms(Xi, Yi)
            mins.append(a); rngs.append(b); ym.append(c); ys.append(e)
            refit.append(i)
    desc = {"train_X": tx, "train_Y": ty, "lengthscale": ls, "outputscale": osc, "noise": noise, "mean_const": mc,
            "kernel": kernel, "normalize": True, "standardize": True, "norm_mins": torch.stack(mins),
            "norm_ranges": torch.stack(rngs), "y_mean": torch.stack(ym), "y_std": torch.stack(ys)}
    if old_model is not None and old is not None:
        desc["train_inputs"] = old_model.train_inputs_mode
    model = B200ModelListGP(desc, device=device, dtype=dtype)
    model.n_at_fit = n_at_fit
    if refit:
        model.fit_info = fit_hyperparameters(model, arms=refit)
    print(f"Skipped model fitting for {num_arms - len(refit)} out of {num_arms}.")  # :120
    return model


def fit_modellist(X: Tensor, Y: Tensor, num_arms: int, device="cuda", dtype: torch.dtype = torch.float64,
                  kernel: str = "matern52") -> B200ModelListGP:
    r"""Fit one GP per arm from scratch (``contextual_rs/experiment_utils.py:15-48``)."""
    return fit_modellist_with_reuse(X, Y, num_arms, None, device=device, dtype=dtype, kernel=kernel)
