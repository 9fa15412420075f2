"""Generalized-PCS oracle (CPU).  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).

  * ``estimate_current_generalized_pcs_ref`` — the ModelListGP branch of
    ``contextual_rs/generalized_pcs.py:318-479`` restated with the reference's *joint* sampler
    (``posterior.rsample``: block-diagonal over arms, dense over contexts; O(n_c^3), small n_c only);
  * ``pcs_marginal_ref`` — the same estimator with per-(arm, context) marginal draws, chunked; this
    is what scales and what the CUDA kernel computes.  PCS(c) depends only on the marginals across
    arms at c (arms are independent GPs; the reference argues the same at
    ``contextual_rs/generalized_pcs.py:382-395``), so both estimate the same per-context quantity;
  * ``pcs_exact_quadrature`` — noise-free known answer
    PCS(c) = int phi(z) prod_{i != b} Phi((mu_b + s_b z - mu_i) / s_i) dz.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
from torch import Tensor


def strict_indicator(dtype=torch.float64) -> Callable[[Tensor], Tensor]:
    """The func_I every reference caller passes (``experiments/wsc_experiments/main.py:453``)."""
    return lambda X: (X > 0).to(dtype=dtype)


def estimate_current_generalized_pcs_ref(model, arm_set: Tensor, context_set: Tensor,
                                         num_samples: int, base_samples: Optional[Tensor],
                                         func_I: Callable[[Tensor], Tensor],
                                         rho: Callable[[Tensor], Tensor],
                                         use_approximation: bool = False) -> Tensor:
    """generalized_pcs.py:318-479, ModelListGP branch (:441-479)."""
    assert arm_set.dim() == 2 and context_set.dim() in [2, 3]  # :360
    if context_set.dim() == 3:  # :361-362
        raise NotImplementedError
    if use_approximation:  # :396-399
        raise NotImplementedError
    K = arm_set.shape[0]
    post = model.posterior(context_set)  # :443 (context_set[0] of the K-fold repeat)
    mu = post.mean.t().unsqueeze(-1)  # K x n_c x 1, :444/:455
    draws = post.rsample(sample_shape=torch.Size([num_samples]), base_samples=base_samples)  # :449
    draws = draws.transpose(-1, -2).unsqueeze(-1)  # S x K x n_c x 1, :454
    _, order = mu.sort(dim=-3, descending=True)  # :464
    draws = draws.gather(dim=-3, index=order.expand(num_samples, *order.shape))  # :465-467
    runner_up, _ = draws[..., 1:, :, :].max(dim=-3)  # :470
    delta = draws[..., 0, :, :] - runner_up  # :471
    pcs_c = func_I(delta).mean(dim=0)  # :473-475, n_c x 1
    return rho(pcs_c).squeeze(-1)  # :477-479


def pcs_marginal_ref(mean: Tensor, variance: Tensor, num_samples: int,
                     generator: Optional[torch.Generator] = None, chunk: int = 1024,
                     context_chunk: int = 1024) -> Tensor:
    """Per-context PCS estimate ``[n_c]`` from independent N(mean, variance) draws at every
    (context, arm); same sort/gather/max/compare as generalized_pcs.py:464-475 written with an
    arg-max instead of the full sort (identical under tie-free means).  mean/variance: ``[n_c, K]``."""
    n_c, K = mean.shape
    out = torch.zeros(n_c, dtype=mean.dtype)
    for c0 in range(0, n_c, context_chunk):
        mu = mean[c0:c0 + context_chunk]
        sd = variance[c0:c0 + context_chunk].sqrt()
        best = mu.argmax(dim=-1)  # [m]
        hits = torch.zeros(mu.shape[0], dtype=mean.dtype)
        mask = torch.zeros_like(mu, dtype=torch.bool)
        mask[torch.arange(mu.shape[0]), best] = True
        for s0 in range(0, num_samples, chunk):
            s = min(chunk, num_samples - s0)
            z = torch.randn(s, *mu.shape, dtype=mean.dtype, generator=generator)
            y = mu + sd * z
            yb = y.gather(-1, best.expand(s, -1).unsqueeze(-1)).squeeze(-1)
            rest = y.masked_fill(mask, float("-inf")).max(dim=-1).values
            hits += (yb - rest > 0).to(mean.dtype).sum(dim=0)
        out[c0:c0 + context_chunk] = hits / num_samples
    return out


def pcs_exact_quadrature(mean: np.ndarray, variance: np.ndarray, num_nodes: int = 40001,
                         z_lim: float = 10.0) -> np.ndarray:
    """Trapezoid quadrature of PCS(c) on a dense z grid (robust when some s_i are tiny and the
    integrand has near-steps).  mean/variance ``[n_c, K]`` float64 -> ``[n_c]``."""
    from scipy.special import log_ndtr

    mean = np.asarray(mean, dtype=np.float64)
    sd = np.sqrt(np.asarray(variance, dtype=np.float64))
    n_c, K = mean.shape
    z = np.linspace(-z_lim, z_lim, num_nodes)
    w = np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi)
    h = z[1] - z[0]
    out = np.zeros(n_c)
    for c in range(n_c):
        b = int(np.argmax(mean[c]))
        yb = mean[c, b] + sd[c, b] * z  # [Z]
        others = np.delete(np.arange(K), b)
        t = (yb[None, :] - mean[c, others, None]) / sd[c, others, None]
        logp = log_ndtr(t).sum(axis=0)
        f = w * np.exp(logp)
        out[c] = h * (f.sum() - 0.5 * (f[0] + f[-1]))
    return out


def lookahead_pcs_tables_ref(model, candidate: Tensor, context_set: Tensor, fantasy_base: Optional[Tensor] = None):
    """Fantasy posterior tables of ``estimate_lookahead_generalized_pcs_modellist``
    (contextual_rs/generalized_pcs.py:246-276): for every candidate ``(arm, x)`` the arm's oracle GP
    is conditioned on Y = posterior mean at x (``fantasy_base is None``: certainty equivalent,
    :262-265) or on ``mean + sqrt(var + noise * stdv^2) * z`` for each z in ``fantasy_base[i]``
    (``fantasize``, :259-260).  Yields ``(i, f, mean [n_c, K], variance [n_c, K])``."""
    base = model.posterior(context_set)
    for i in range(candidate.shape[0]):
        arm = int(candidate[i, 0, 0])
        x = candidate[i, :, 1:]
        sub = model.models[arm]
        px = sub.posterior(x)
        mu_x, var_x = px.mean.reshape(()), px.variance.reshape(())
        if fantasy_base is None:
            ys = [mu_x]
        else:
            sd = (var_x + sub.noise * sub.y_std * sub.y_std).sqrt()
            ys = [mu_x + sd * z for z in fantasy_base[i]]
        for f, y in enumerate(ys):
            fm = sub.condition_on_observations(x, y.reshape(1, 1))
            pf = fm.posterior(context_set)
            mean, var = base.mean.clone(), base.variance.clone()
            mean[:, arm], var[:, arm] = pf.mean.squeeze(-1), pf.variance.squeeze(-1)
            yield i, f, mean, var
