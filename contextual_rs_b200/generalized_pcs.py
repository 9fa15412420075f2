"""Monte-Carlo generalized PCS on the B200: drop-in for the ModelListGP branch of
``estimate_current_generalized_pcs`` (``contextual_rs/generalized_pcs.py:318-479``).

The reference draws ``S x n_c x K`` joint posterior samples, sorts arms by posterior mean,
compares the best arm's sample with the max of the others, applies ``func_I`` and averages.
PCS(c) depends only on the marginals across arms at context c (arms are independent GPs; the
reference argues the same at generalized_pcs.py:382-395), so ``crs_pcs_counts`` draws per-(arm,
context) marginals from a Philox stream inside the kernel and returns the success counts; ``rho``
(a Python callable) is applied to the resulting ``n_c x 1`` tensor on the device.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .contextual_rs_strategies import _as_b200
from .model import B200ModelListGP


def _is_strict_indicator(func_I: Callable[[Tensor], Tensor]) -> bool:
    probe = torch.tensor([-1.0, -1e-300, 0.0, 1e-300, 1.0], dtype=torch.float64)
    try:
        out = func_I(probe)
    except Exception:
        return False
    return out.shape == probe.shape and out.to(torch.float64).tolist() == [0.0, 0.0, 0.0, 1.0, 1.0]


def pcs_counts(model: B200ModelListGP, mean: Tensor, var: Tensor, num_samples: int, seed: int,
               offset: int = 0, arm_offset: int = 0, ctx_offset: int = 0,
               counts: Optional[Tensor] = None, stats: Optional[Tensor] = None) -> Tensor:
    """Launch ``crs_pcs_counts`` on device tables ``[n_c, K]``; returns int32 counts ``[n_c]``."""
    n_c, K = mean.shape
    if counts is None:
        counts = torch.empty(n_c, dtype=torch.int32, device=mean.device)
    with torch.cuda.device(mean.device):
        rc = _lib.get().crs_pcs_counts(mean.data_ptr(), var.data_ptr(), n_c, K, mean.stride(0),
                                       _lib.DTYPES[mean.dtype], int(num_samples), int(seed) & (2 ** 64 - 1),
                                       int(offset) & 0xFFFFFFFF, int(arm_offset), int(ctx_offset),
                                       counts.data_ptr(), _lib.ptr(stats), _lib.stream_ptr(mean.device))
    _lib.check(rc, "crs_pcs_counts")
    return counts


def estimate_current_generalized_pcs(
    model,
    arm_set: Tensor,
    context_set: Tensor,
    num_samples: int,
    base_samples: Optional[Tensor],
    func_I: Callable[[Tensor], Tensor],
    rho: Callable[[Tensor], Tensor],
    use_approximation: bool = False,
    *,
    seed: Optional[int] = None,
    offset: int = 0,
    reuse_grid: bool = False,
) -> Tensor:
    r"""Estimate of the generalized PCS implied by the current model
    (``contextual_rs/generalized_pcs.py:318-479``; argument meaning as in the reference).

    Extra keyword-only arguments (not in the reference): ``seed``/``offset`` pin the Philox
    stream (by default a seed is drawn from the global torch RNG, which the reference also
    consumes through ``posterior.rsample``); ``reuse_grid`` skips the posterior launch when the
    model's workspace already holds the grid of this ``context_set`` (same iteration's decision).

    Returns:
        0-dim tensor on ``context_set.device``.
    """
    assert arm_set.dim() == 2 and context_set.dim() in [2, 3]  # generalized_pcs.py:360
    m = _as_b200(model)
    if context_set.dim() == 3:  # :361-362
        raise NotImplementedError
    if use_approximation:  # :396-399
        raise NotImplementedError
    if base_samples is not None:
        raise NotImplementedError(
            "base_samples cannot be honoured: samples are generated inside the kernel from a "
            "Philox stream and never materialised (pass seed=/offset= to pin the stream)")
    if not _is_strict_indicator(func_I):
        raise NotImplementedError(
            "only the strict indicator func_I = (X > 0) is fused into the PCS kernel "
            "(the only func_I the reference's callers use); there is no CPU fallback")
    if arm_set.shape[0] != m.num_arms:
        raise ValueError("arm_set must list one row per arm of the model")
    n_c = context_set.shape[0]
    ws = m.workspace(n_c)
    if not (reuse_grid and ws.get("valid_for") == (context_set.data_ptr(), n_c)):
        ws["ctx"].copy_(context_set.to(dtype=m.dtype), non_blocking=True)
        m.posterior_tables(ws["ctx"], ws["mean"], ws["var"], arms=(0, m.num_arms))
        ws["valid_for"] = (context_set.data_ptr(), n_c)
    if seed is None:
        seed = int(torch.randint(2 ** 62, (1,)))
    counts = pcs_counts(m, ws["mean"], ws["var"], num_samples, seed, offset,
                        counts=ws["counts"], stats=ws["stats"])
    pcs_c_est = (counts.to(m.dtype) / float(num_samples)).unsqueeze(-1)  # n_c x 1, :475
    rho_vals = rho(pcs_c_est)  # :477
    return rho_vals.squeeze(-1).to(context_set.device)  # :479


def empirical_best_arms(model, context_set: Tensor, reuse_grid: bool = False) -> Tensor:
    """``model.posterior(context_map).mean.t().argmax(dim=0)`` of
    ``experiments/wsc_experiments/main.py:459-468`` — a by-product of the scan kernel."""
    import ctypes as C
    m = _as_b200(model)
    n_c = context_set.shape[0]
    ws = m.workspace(n_c)
    lib = _lib.get()
    K = m.num_arms
    with torch.cuda.device(m.device):
        stream = _lib.stream_ptr(m.device)
        if not (reuse_grid and ws.get("valid_for") == (context_set.data_ptr(), n_c)):
            ws["ctx"].copy_(context_set.to(dtype=m.dtype), non_blocking=True)
            m.posterior_tables(ws["ctx"], ws["mean"], ws["var"], arms=(0, K))
            ws["valid_for"] = (context_set.data_ptr(), n_c)
        _lib.check(lib.crs_ocba_scan(ws["mean"].data_ptr(), ws["var"].data_ptr(), n_c, K, K,
                                     _lib.DTYPES[m.dtype], 0, None, None, None,
                                     ws["best_arm"].data_ptr(), None, None, None, None, None, stream),
                   "crs_ocba_scan")
    return ws["best_arm"].to(device=context_set.device, dtype=torch.long)
