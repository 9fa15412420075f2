"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's hyper-parameter fit
(``contextual_rs/experiment_utils.py:15-48, 68-121``): one ``SingleTaskGP`` per arm with
``Standardize(m=1)`` / ``Normalize(d)`` and BoTorch's default priors and constraints (mirrored in
the reference at ``contextual_rs/models/lce_gp.py:101-129``), fitted by
``fit_gpytorch_mll(ExactMarginalLogLikelihood(...))`` = scipy L-BFGS-B on

    -( log N(y~ | m 1, o C(l) + s2 I) + log Gamma(l; 3, 6).sum() + log Gamma(o; 2, 0.15)
       + log Gamma(s2; 1.1, 0.05) ) / n

over (raw_lengthscale, raw_outputscale: softplus; noise >= 1e-4 as a box bound; mean constant),
from gpytorch's initial values (raw = 0, noise = 2.0, mean = 0).  The objective is written in
plain torch (autograd supplies the gradient) and minimised with ``scipy.optimize.minimize(method=
"L-BFGS-B")``.  BoTorch/GPyTorch are absent from this image, so the published algorithm is restated
(prior normalisation constants are dropped on both sides); what pins it is the replay of the
reference's recorded runs (``trace_replay.py``): with these fits 90 of the first 100 recorded
GP-C-OCBA decisions and 968 of 1,000 recorded best-arm readouts of ``0000_ML_Gao.pt`` are reproduced.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch
from torch import Tensor

from .gp_oracle import scaled_sq_dist

LENGTHSCALE_PRIOR, OUTPUTSCALE_PRIOR, NOISE_PRIOR = (3.0, 6.0), (2.0, 0.15), (1.1, 0.05)
MIN_NOISE = 1e-4


def transforms(X: Tensor, Y: Tensor):
    mins = X.min(dim=0).values
    ranges = X.max(dim=0).values - mins
    ranges = torch.where(ranges < 1e-8, torch.ones_like(ranges), ranges)
    ym = Y.mean()
    ys = Y.std() if Y.numel() > 1 else torch.ones((), dtype=Y.dtype)
    if not bool(ys >= 1e-8):
        ys = torch.ones((), dtype=Y.dtype)
    return mins, ranges, ym, ys


def correlation(r2: Tensor, kernel: str) -> Tensor:
    """Same correlations as gp_oracle.correlation, with gpytorch's differentiable distance
    (``dist = r2.clamp_min(1e-30).sqrt()``): the gradient at coincident points is 0, not NaN."""
    if kernel == "matern52":
        s5r = math.sqrt(5.0) * r2.clamp_min(1e-30).sqrt()
        return ((s5r + 1.0) + (5.0 / 3.0) * r2) * torch.exp(-s5r)
    if kernel == "rbf":
        return torch.exp(-0.5 * r2)
    raise ValueError(kernel)


def exact_mll(Xn: Tensor, yt: Tensor, ls: Tensor, osc: Tensor, noise: Tensor, mean: Tensor, kernel: str) -> Tensor:
    """log N(yt | mean 1, osc C(ls) + noise I) for normalised inputs / standardised targets."""
    n = Xn.shape[0]
    Xs = Xn / ls
    Khat = osc * correlation(scaled_sq_dist(Xs, Xs), kernel) + noise * torch.eye(n, dtype=Xn.dtype)
    L = torch.linalg.cholesky(Khat)
    r = (yt - mean).reshape(-1, 1)
    z = torch.linalg.solve_triangular(L, r, upper=False)
    return -0.5 * (z * z).sum() - L.diagonal().log().sum() - 0.5 * n * math.log(2.0 * math.pi)


def objective(raw: Tensor, Xn: Tensor, yt: Tensor, kernel: str) -> Tensor:
    """-(MLL + log priors) / n in the raw parameters (raw_ls[d], raw_os, noise, mean)."""
    d = Xn.shape[1]
    sp = torch.nn.functional.softplus
    ls, osc, noise, mean = sp(raw[:d]), sp(raw[d]), raw[d + 1], raw[d + 2]
    (a_l, b_l), (a_o, b_o), (a_n, b_n) = LENGTHSCALE_PRIOR, OUTPUTSCALE_PRIOR, NOISE_PRIOR
    logp = ((a_l - 1) * ls.log() - b_l * ls).sum() + (a_o - 1) * osc.log() - b_o * osc \
        + (a_n - 1) * noise.log() - b_n * noise
    return -(exact_mll(Xn, yt, ls, osc, noise, mean, kernel) + logp) / Xn.shape[0]


def fit_arm(X: Tensor, Y: Tensor, kernel: str = "matern52") -> Tuple[Dict, float]:
    """Fit one arm; returns (natural hyper-parameters + transforms, final objective)."""
    from scipy.optimize import minimize
    X, Y = X.to(torch.float64), Y.to(torch.float64).reshape(-1)
    d = X.shape[1]
    mins, ranges, ym, ys = transforms(X, Y)
    Xn, yt = (X - mins) / ranges, (Y - ym) / ys
    x0 = np.zeros(d + 3)
    x0[d + 1] = (NOISE_PRIOR[0] - 1.0) / NOISE_PRIOR[1]

    def fun(x):
        raw = torch.tensor(x, dtype=torch.float64, requires_grad=True)
        try:
            f = objective(raw, Xn, yt, kernel)
        except Exception:  # not positive definite
            return float("inf"), np.zeros_like(x)
        (g,) = torch.autograd.grad(f, raw)
        return float(f.detach()), g.numpy()

    bounds = [(None, None)] * (d + 1) + [(MIN_NOISE, None), (None, None)]
    res = minimize(fun, x0, jac=True, method="L-BFGS-B", bounds=bounds, options={"maxiter": 2000, "maxfun": 5000})
    raw = torch.tensor(res.x, dtype=torch.float64)
    sp = torch.nn.functional.softplus
    hp = {"lengthscale": sp(raw[:d]), "outputscale": sp(raw[d]), "noise": raw[d + 1], "mean_const": raw[d + 2],
          "norm_mins": mins, "norm_ranges": ranges, "y_mean": ym, "y_std": ys}
    return hp, float(res.fun)


def fit_modellist_ref(X: Tensor, Y: Tensor, num_arms: int, kernel: str = "matern52") -> Tuple[Dict, Tensor]:
    """``fit_modellist`` (experiment_utils.py:15-48) -> (fitted model description, objectives)."""
    X, Y = X.to(torch.float64), Y.to(torch.float64).reshape(-1)
    keys = ("lengthscale", "outputscale", "noise", "mean_const", "norm_mins", "norm_ranges", "y_mean", "y_std")
    cols = {k: [] for k in keys}
    tx, ty, objs = [], [], []
    for i in range(num_arms):
        mask = X[..., 0] == i
        hp, f = fit_arm(X[mask][..., 1:], Y[mask], kernel)
        tx.append(X[mask][..., 1:]); ty.append(Y[mask].reshape(-1, 1)); objs.append(f)
        for k in keys:
            cols[k].append(hp[k])
    desc = {"train_X": tx, "train_Y": ty, "kernel": kernel, "normalize": True, "standardize": True}
    desc.update({k: torch.stack(v) for k, v in cols.items()})
    return desc, torch.tensor(objs, dtype=torch.float64)
