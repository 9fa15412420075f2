"""GP-C-OCBA decision oracle (CPU, PyTorch).  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).

Restates, operation for operation (same torch ops in the same order, so that given the same
posterior tables the outputs are bit-identical to the reference's), the allocation logic of

  * ``gao_modellist``                — ``contextual_rs/contextual_rs_strategies.py:132-202``
  * ``_get_hat_Z`` / ``MinZeta``     — ``contextual_rs/continuous_context.py:21-36, 57-74``
  * ``find_next_arm_given_context``  — ``contextual_rs/continuous_context.py:96-132``
  * ``gao_sampling_strategy``        — ``contextual_rs/contextual_rs_strategies.py:75-97``
    (kept only as a logic cross-check: it shares the zeta table and has a pinned known answer
    in ``test/test_contextual_rs_strategies.py:55-59``)

``model`` is any object with ``.models``, ``.posterior(X) -> (.mean, .variance)``,
``.train_inputs`` and per-sub-model ``.forward(x).variance`` — the de-facto boundary the
reference's own MockModel tests establish (``test/test_contextual_rs_strategies.py:169-203``).
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor


def zeta_table(means: Tensor, variances: Tensor) -> Tuple[Tensor, Tensor]:
    """``[n_c, K]`` posterior tables -> (zeta ``[n_c, K-1]`` in descending-mean order with the
    best arm dropped, sort permutation ``[n_c, K]``).
    contextual_rs_strategies.py:137-143 == continuous_context.py:29-36."""
    mu_desc, order = means.sort(dim=-1, descending=True)
    var_desc = variances.gather(dim=-1, index=order)
    rest = means.shape[-1] - 1
    var_sum = var_desc[:, :1].expand(-1, rest) + var_desc[:, 1:]
    gap = mu_desc[:, :1].expand(-1, rest) - mu_desc[:, 1:]
    return gap.pow(2) / var_sum, order


def pick_minimiser(flat: Tensor, randomize_ties: bool) -> Tensor:
    """First minimum of the flattened zeta; if ``randomize_ties`` and the minimum is attained
    more than once, a uniform draw among the tied positions using the global torch RNG, which is
    consumed ONLY in that case.  contextual_rs_strategies.py:145-154, continuous_context.py:99-108."""
    z_min, where = torch.min(flat, dim=0)
    if randomize_ties:
        tied = flat == z_min
        n_tied = tied.sum()
        if n_tied > 1:
            positions = torch.arange(0, flat.shape[0], device=flat.device)[tied]
            where = positions[torch.randint(n_tied, (1,), device=flat.device)].squeeze()
    return where


def ocba_arm_choice(y_vals: Tensor, order_row: Tensor, zeta_rank: Tensor) -> Tensor:
    """Step 2(b): compare psi of the best arm with the summed psi of the others, in the sorted
    order (``y_s_1 < y_s_2`` -> best arm, else the zeta-minimising arm, rank shifted by the
    dropped best).  contextual_rs_strategies.py:194-201, continuous_context.py:124-131."""
    y_sorted = y_vals[order_row]
    psi_best = y_sorted[0]
    psi_rest = y_sorted[1:].sum()
    rank = 0 if psi_best < psi_rest else zeta_rank + 1
    return order_row[rank]


def gao_modellist_ref(model, context_set: Tensor, randomize_ties: bool = True,
                      infer_p: bool = False, use_kde: bool = False,
                      kernel_scale: float = 2.0) -> Tuple[Tensor, Tensor]:
    """CPU restatement of ``gao_modellist`` (contextual_rs_strategies.py:100-202)."""
    post = model.posterior(context_set)  # :133
    means, variances = post.mean, post.variance  # n_c x K, :134-136
    K = len(model.models)  # :139
    hat_Z, order = zeta_table(means, variances)
    where = pick_minimiser(hat_Z.view(-1), randomize_ties)  # :144-154
    c_star = torch.div(where, K - 1, rounding_mode="floor")  # :155
    x_star = context_set[c_star]  # :156
    zeta_rank = where % (K - 1)  # :157
    if infer_p:  # :158-167
        prior_var = [m.forward(context_set[c_star].view(1, -1)).variance for m in model.models]
        ratio = torch.cat(prior_var) / variances[c_star]
        y_vals = ratio / variances[c_star]
    elif use_kde:  # :168-186
        weights = torch.zeros(K).to(context_set)
        for i, inputs in enumerate(model.train_inputs):
            dist = torch.cdist(inputs[0], x_star.unsqueeze(0))
            weights[i] = torch.exp(-kernel_scale * dist).sum(dim=0)
        y_vals = weights / variances[c_star]
    else:  # :187-193
        counts = torch.zeros(K).to(context_set)
        for i, inputs in enumerate(model.train_inputs):
            counts[i] = (inputs[0] == x_star).all(dim=-1).sum()
        y_vals = counts / variances[c_star]
    next_arm = ocba_arm_choice(y_vals, order[c_star], zeta_rank)  # :194-201
    return next_arm, x_star


def min_zeta_ref(model, X: Tensor) -> Tensor:
    """``MinZeta.forward``: ``X`` is ``batch x 1 x d``; returns ``-min_j zeta_{c,j}`` per
    context.  continuous_context.py:57-74."""
    assert X.shape[-2] == 1 and X.ndim == 3
    post = model.posterior(X)
    hat_Z, _ = zeta_table(post.mean.squeeze(-2), post.variance.squeeze(-2))
    return -hat_Z.min(dim=-1).values


def find_next_arm_given_context_ref(next_context: Tensor, model, kernel_scale: float = 2.0,
                                    randomize_ties: bool = True) -> Tensor:
    """continuous_context.py:77-132: steps 4-7 for one given context, always with KDE counts."""
    post = model.posterior(next_context.view(1, 1, -1))  # :96
    hat_Z, order = zeta_table(post.mean.squeeze(-2), post.variance.squeeze(-2))
    zeta_rank = pick_minimiser(hat_Z.view(-1), randomize_ties)  # :99-109
    weights = torch.zeros(len(model.models)).to(hat_Z)  # :112-122
    for i, inputs in enumerate(model.train_inputs):
        dist = torch.cdist(inputs[0], next_context)
        weights[i] = torch.exp(-kernel_scale * dist).sum(dim=0)
    y_vals = weights / post.variance.squeeze()  # :123
    return ocba_arm_choice(y_vals, order.view(-1), zeta_rank)  # :124-132


def gao_sampling_strategy_ref(means: Tensor, vars_: Tensor, num_obs: Tensor) -> Tuple[Tensor, Tensor]:
    """C-OCBA on the IID-normal tabular model (tables are ``K x n_c`` here, arm-major as in the
    reference's ContextualIndependentModel).  contextual_rs_strategies.py:75-97."""
    K, n_c = means.shape
    p = num_obs / num_obs.sum()
    mu_desc, order = means.sort(dim=0, descending=True)
    var_desc = vars_.gather(dim=0, index=order)
    p_desc = p.gather(dim=0, index=order)
    scaled = var_desc / p_desc
    hat_Z = (mu_desc[:1].expand(K - 1, -1) - mu_desc[1:]).pow(2) / (scaled[:1].expand(K - 1, -1) + scaled[1:])
    where = torch.argmin(hat_Z)
    zeta_rank = where // n_c
    ctx = where % n_c
    y = p_desc[:, ctx].pow(2) / var_desc[:, ctx]
    rank = 0 if y[0] < y[1:].sum() else zeta_rank + 1
    return order[rank, ctx], ctx


def best_arms_ref(model, context_set: Tensor) -> Tensor:
    """Empirical best-arm readout ``posterior.mean.t().argmax(dim=0)``
    (``experiments/wsc_experiments/main.py:459-468``)."""
    return model.posterior(context_set).mean.t().argmax(dim=0)
