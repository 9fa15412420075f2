"""LEVI (Pearce & Branke 2017) on the B200: drop-ins for ``discrete_levi``, ``PredictiveEI`` and
``_compute_all_EI_reviewer`` of ``contextual_rs/levi.py:132-227`` — the acquisition the
reference's drivers run as "LEVI-new" (``experiments/wsc_experiments/main.py:406-410``,
``experiments/continuous_experiments/main.py:269-284``).

The grid posterior comes from ``crs_posterior_grid``; ``crs_levi_scan`` evaluates the expected
improvement of every (context, arm) pair, the row maxima and the first arg-max over contexts.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .contextual_rs_strategies import _as_b200
from .model import B200ModelListGP


def _levi_launch(m: B200ModelListGP, ctx2: Tensor, weights: Optional[Tensor], want_table: bool,
                 reuse_grid: bool = False):
    """Grid posterior + ``crs_levi_scan`` for ``ctx2`` (``[n_c, d]``); returns the workspace."""
    lib = _lib.get()
    n_c = ctx2.shape[0]
    ws = m.workspace(n_c)
    K = m.num_arms
    if "levi_max" not in ws:
        ws["levi_max"] = torch.empty(n_c, dtype=m.dtype, device=m.device)
        ws["levi_arm"] = torch.empty(n_c, dtype=torch.int32, device=m.device)
        ws["levi_dec"] = torch.zeros(24, dtype=torch.uint8, device=m.device)
    if want_table and "levi_ei" not in ws:
        ws["levi_ei"] = torch.empty(n_c, K, dtype=m.dtype, device=m.device)
    w_dev = None
    if weights is not None:
        w_dev = weights.reshape(-1).to(device=m.device, dtype=m.dtype).contiguous()
        if w_dev.shape[0] != n_c:
            raise ValueError("weights must hold one entry per context")
    with torch.cuda.device(m.device):
        stream = _lib.stream_ptr(m.device)
        if not (reuse_grid and ws.get("valid_for") == m.grid_key(ctx2)):
            ws["ctx"].copy_(ctx2.to(dtype=m.dtype), non_blocking=True)
            m.posterior_tables(ws["ctx"], ws["mean"], ws["var"], arms=(0, K))
            ws["valid_for"] = m.grid_key(ctx2)
        _lib.check(lib.crs_levi_scan(C.byref(m.state), ws["mean"].data_ptr(), ws["var"].data_ptr(), n_c, K,
                                     _lib.ptr(w_dev), ws["levi_ei"].data_ptr() if want_table else None, K,
                                     ws["levi_max"].data_ptr(), ws["levi_arm"].data_ptr(),
                                     ws["levi_dec"].data_ptr(), stream), "crs_levi_scan")
    return ws


def _compute_all_EI_reviewer(X: Tensor, model) -> Tensor:
    r"""EI of every arm at every context (``contextual_rs/levi.py:132-172``).

    Args:
        X: ``batch x 1 x d_c`` contexts.
    Returns:
        ``batch x num_arms`` tensor of EI values on ``X.device``.
    """
    m = _as_b200(model)
    ws = _levi_launch(m, X.reshape(X.shape[0], -1), None, want_table=True)
    # a NEW tensor, like the reference: `.to` alone would alias the workspace buffer when X already
    # lives on the model's GPU in the model's dtype, and the next call would overwrite the result
    return ws["levi_ei"].to(device=X.device, dtype=X.dtype, copy=True)


class PredictiveEI:
    r"""Max over arms of the LEVI expected improvement (``contextual_rs/levi.py:175-196``); the
    acquisition ``optimize_acqf`` maximises over contexts in the continuous-context driver."""

    def __init__(self, model) -> None:
        self.model = _as_b200(model)

    def forward(self, X: Tensor) -> Tensor:
        r"""``X``: ``batch x 1 x d_c`` -> ``batch`` max-EI values (on ``X.device``)."""
        assert X.shape[-2] == 1 and X.ndim == 3
        ws = _levi_launch(self.model, X.reshape(X.shape[0], -1), None, want_table=False)
        return ws["levi_max"].to(device=X.device, dtype=X.dtype, copy=True)  # never an alias of the workspace

    __call__ = forward


def discrete_levi(model, context_set: Tensor, weights: Optional[Tensor] = None,
                  *, reuse_grid: bool = False) -> Tuple[Tensor, Tensor]:
    r"""Next alternative / context by maximising LEVI over a finite context set
    (``contextual_rs/levi.py:199-227``).

    Args:
        model: a :class:`B200ModelListGP` (or a BoTorch ``ModelListGP``).
        context_set: ``num_contexts x d_c``.
        weights: optional ``num_contexts`` weights applied to the per-context max EI.
    Returns:
        ``(next_alternative, next_context)``: 0-dim int64 tensor and the selected row of
        ``context_set``, both on ``context_set.device``.
    """
    m = _as_b200(model)
    if m.num_arms < 2:
        raise ValueError("discrete_levi needs at least two arms")
    ws = _levi_launch(m, context_set, weights, want_table=False, reuse_grid=reuse_grid)
    dec = _lib.LeviDecision.from_buffer_copy(ws["levi_dec"].cpu().numpy().tobytes())
    next_alternative = torch.tensor(dec.arm, dtype=torch.long, device=context_set.device)
    return next_alternative, context_set[dec.context]
