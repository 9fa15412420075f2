"""Multi-start gradient refinement of an acquisition function over the context box on the B200:
the ``optimize_acqf(acq_function, bounds, q=1, num_restarts=20, raw_samples=1024)`` call that the
reference's continuous-context driver wraps around ``MinZeta`` / ``PredictiveEI``
(``experiments/continuous_experiments/main.py:252-263, 274-281``; BoTorch ``[3P]``).

What BoTorch does for ``q = 1`` and what is mirrored here:

1. ``raw_samples`` scrambled-Sobol points of the box are scored with the acquisition function
   (one grid + scan launch for all of them);
2. ``num_restarts`` starting points are drawn from them — the arg-max always, the others without
   replacement with probability ``exp(eta * standardised value)`` (``initialize_q_batch``, eta = 2);
3. the restarts are refined by L-BFGS under the box constraints (BoTorch: scipy L-BFGS-B on the
   stacked restarts; here the restarts are independent box-constrained L-BFGS problems advanced in
   lockstep by ``fit.batched_lbfgs`` — every iteration is ONE objective launch for all restarts, with
   the analytic gradient of ``crs_min_zeta_grad`` when the acquisition function is ``MinZeta``; an
   acquisition function without a backward, e.g. ``PredictiveEI``, gets central differences from one
   more launch);
4. the best refined point and its value are returned as ``(candidate [1, d], value)``.

The optimiser trajectory is not scipy's (its More-Thuente line search is not replicated); the
objective, the initialisation heuristic and the stopping rules are the same.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor

from .fit import batched_lbfgs


class _BoxObjective:
    """f = -acq(X) and its gradient for a batch of restarts ``[R, d]`` (autograd through the
    acquisition function: ``MinZeta`` carries the analytic backward of ``crs_min_zeta_grad``)."""

    def __init__(self, acq: Callable[[Tensor], Tensor], d: int, fd_step: Optional[Tensor] = None):
        self.acq, self.d = acq, d
        self.fd_step = fd_step if fd_step is not None else torch.full((d,), 1e-6, dtype=torch.float64)
        self.evaluations = 0

    def __call__(self, x: Tensor):
        X = x.detach().clone().unsqueeze(-2).requires_grad_(True)
        with torch.enable_grad():
            val = self.acq(X)
            differentiable = val.requires_grad
            if differentiable:
                (g,) = torch.autograd.grad(val.sum(), X)
        self.evaluations += 1
        if not differentiable:
            # an acquisition function without a backward (``PredictiveEI``: the LEVI branch of the continuous
            # driver, experiments/continuous_experiments/main.py:269-281): central differences, all 2 d
            # perturbations of all restarts in ONE further launch
            R, d = x.shape
            h = self.fd_step.to(device=x.device, dtype=x.dtype)
            eye = torch.eye(d, dtype=x.dtype, device=x.device) * h
            pts = torch.cat([x.unsqueeze(1) + eye, x.unsqueeze(1) - eye], dim=1).reshape(R * 2 * d, 1, d)
            with torch.no_grad():
                v = self.acq(pts).reshape(R, 2, d).to(torch.float64)
            g = ((v[:, 0] - v[:, 1]) / (2.0 * h.to(torch.float64))).unsqueeze(-2)
            self.evaluations += 1
        f = -val.detach().to(torch.float64)
        g = -g.squeeze(-2).to(torch.float64)
        bad = ~torch.isfinite(f)
        f = torch.where(bad, torch.full_like(f, float("inf")), f)
        g = torch.where(bad.unsqueeze(-1), torch.zeros_like(g), g)
        return f, g


def initialize_q_batch(X: Tensor, Y: Tensor, n: int, eta: float = 2.0, generator=None) -> Tensor:
    """BoTorch's heuristic for picking ``n`` of the raw candidates: the best one always, the rest
    sampled without replacement with weights ``exp(eta * (Y - mean) / std)``; a random subset when
    all values are equal."""
    n_raw = X.shape[0]
    if n >= n_raw:
        return X
    Yd = Y.detach().to(torch.float64).reshape(-1).cpu()
    std = Yd.std()
    if not torch.isfinite(std) or float(std) == 0.0:
        idx = torch.randperm(n_raw, generator=generator)[:n]
        return X[idx.to(X.device)]
    Z = (Yd - Yd.mean()) / std
    w = torch.exp(eta * Z)
    w = torch.where(torch.isfinite(w), w, torch.zeros_like(w))
    top = int(torch.argmax(Yd))
    idx = torch.multinomial(w, n, replacement=False, generator=generator) if int((w > 0).sum()) >= n else \
        torch.randperm(n_raw, generator=generator)[:n]
    if top not in idx.tolist():
        idx[-1] = top
    return X[idx.to(X.device)]


def optimize_acqf(acq_function: Callable[[Tensor], Tensor], bounds: Tensor, q: int = 1, num_restarts: int = 20,
                  raw_samples: int = 1024, options: Optional[Dict] = None, *, seed: Optional[int] = None,
                  return_details: bool = False):
    r"""Maximise ``acq_function`` over the box ``bounds`` (``2 x d``) for ``q = 1``; signature and return
    value follow BoTorch's ``optimize_acqf``: ``(candidate [1, d], acq_value 0-dim)`` on ``bounds.device``.

    ``options``: ``maxiter`` (200, BoTorch's default), ``eta`` (2.0), ``gtol`` / ``ftol``.  ``seed`` pins the
    Sobol scrambling and the restart sampling (default: drawn from the global torch RNG, which BoTorch also
    consumes here)."""
    if q != 1:
        raise NotImplementedError("only q = 1 (the reference's call) is supported")
    options = dict(options or {})
    d = bounds.shape[-1]
    lower, upper = bounds[0].to(torch.float64), bounds[1].to(torch.float64)
    if seed is None:
        seed = int(torch.randint(2 ** 31 - 1, (1,)))
    gen = torch.Generator().manual_seed(seed)
    eng = torch.quasirandom.SobolEngine(dimension=d, scramble=True, seed=seed)
    raw = eng.draw(raw_samples, dtype=torch.float64).to(bounds.device)
    raw = lower + (upper - lower) * raw
    with torch.no_grad():
        y_raw = acq_function(raw.unsqueeze(-2).to(bounds.dtype))
    x0 = initialize_q_batch(raw, y_raw, num_restarts, eta=float(options.get("eta", 2.0)), generator=gen)

    def project(x: Tensor, _d: int) -> Tensor:
        return torch.minimum(torch.maximum(x, lower.to(x.device)), upper.to(x.device))

    def projected_gradient(x: Tensor, g: Tensor, _d: int) -> Tensor:
        lo, up = lower.to(x.device), upper.to(x.device)
        blocked = ((x <= lo) & (g > 0)) | ((x >= up) & (g < 0))  # descent direction -g leaves the box
        return torch.where(blocked, torch.zeros_like(g), g)

    obj = _BoxObjective(lambda X: acq_function(X.to(bounds.dtype)), d, fd_step=1e-6 * (upper - lower).clamp_min(1e-12))
    x, f, g, iters = batched_lbfgs(obj, x0.to(torch.float64), maxiter=int(options.get("maxiter", 200)),
                                   gtol=float(options.get("gtol", 1e-5)), ftol=float(options.get("ftol", 2.2e-9)),
                                   project=project, projected_gradient=projected_gradient)
    best = int(torch.argmin(f))
    cand = x[best].reshape(1, d).to(device=bounds.device, dtype=bounds.dtype)
    value = (-f[best]).to(device=bounds.device, dtype=bounds.dtype)
    if return_details:
        return cand, value, {"restarts": x, "values": -f, "iterations": iters, "evaluations": obj.evaluations,
                             "raw_best": float(y_raw.max()), "x0": x0}
    return cand, value
