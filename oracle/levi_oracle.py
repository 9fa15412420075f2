"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's LEVI acquisition
(``contextual_rs/levi.py``), operation for operation in PyTorch:

- ``compute_all_ei_reviewer``  <- ``_compute_all_EI_reviewer`` (levi.py:132-172)
- ``predictive_ei``            <- ``PredictiveEI.forward``      (levi.py:183-196)
- ``discrete_levi_ref``        <- ``discrete_levi``             (levi.py:199-227)

The model only has to expose what the reference reads: ``model.posterior(X)`` ->
``.mean/.variance`` (``batch x 1 x K``), ``model.models[i].noise`` (``model.likelihood.
likelihoods[i].noise``) and ``model.models[i].y_std`` (``outcome_transform._stdvs_sq`` = its
square).  The reference has no test of LEVI, so the known answers in tests/test_oracle_golden.py
are hand-derived from the formulas with scalar ``math.erf`` arithmetic; on top of that the
reference's recorded LEVI run (``0000_LEVI_new.pt``) is replayed by ``trace_replay.py``: 77 of its
first 100 decisions are reproduced (``tests/golden/reference_trace_levi_new.json``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.distributions import Normal


def ei_from_tables(mean: Tensor, variance: Tensor, noise_orig: Tensor) -> Tensor:
    """levi.py:147-168 on ``batch x 1 x K`` posterior tensors; ``noise_orig[i]`` =
    ``raw_noise_i * stdvs_sq_i`` (levi.py:158-161)."""
    expanded_mean = mean.expand(-1, mean.shape[-1], -1)
    expanded_mean = expanded_mean - torch.full(mean.shape[-1:], float("inf"), dtype=mean.dtype).diag()
    delta = -(mean.squeeze(-2) - expanded_mean.max(dim=-1).values).abs()
    variance = variance.clamp_min(1e-9).squeeze(-2)
    variance_w_noise = variance.clone()
    for i in range(variance.shape[-1]):
        variance_w_noise[..., i] = variance_w_noise[..., i] + float(noise_orig[i])
    sigma = variance / variance_w_noise.sqrt()
    u = delta / sigma
    normal = Normal(torch.zeros_like(u), torch.ones_like(u))
    ucdf = normal.cdf(u)
    updf = torch.exp(normal.log_prob(u))
    return sigma * (updf + u * ucdf)


def _noise_orig(model) -> Tensor:
    return torch.stack([(m.noise * m.y_std * m.y_std).to(torch.float64).reshape(()) for m in model.models])


def compute_all_ei_reviewer(X: Tensor, model) -> Tensor:
    posterior = model.posterior(X)
    return ei_from_tables(posterior.mean, posterior.variance, _noise_orig(model))


def predictive_ei(X: Tensor, model) -> Tensor:
    return compute_all_ei_reviewer(X, model).max(dim=-1).values


def discrete_levi_ref(model, context_set: Tensor, weights: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    all_ei = compute_all_ei_reviewer(context_set.unsqueeze(-2), model)
    max_ei, max_idcs = all_ei.max(dim=-1)
    if weights is not None:
        max_ei = max_ei * weights
    context_idx = max_ei.argmax()
    return max_idcs[context_idx], context_set[context_idx]
