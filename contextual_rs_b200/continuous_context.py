"""Continuous-context GP-C-OCBA on the B200: drop-ins for ``MinZeta`` and
``find_next_arm_given_context`` of ``contextual_rs/continuous_context.py:39-74, 77-132``.

``MinZeta.forward`` sweeps a batch of candidate contexts (config 4: a 65,536-point Sobol set)
through the grid kernel and the scan kernel and returns ``-min_j zeta_{c,j}`` per context; the
L-BFGS refinement the reference wraps around it (``optimize_acqf``,
``experiments/continuous_experiments/main.py:252-263``) needs gradients with respect to the
contexts: ``MinZeta`` is a differentiable function with an analytic backward kernel
(``crs_min_zeta_grad``); the optimiser itself is host-side and stays the caller's.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from . import _lib
from .contextual_rs_strategies import _as_b200, decide


def _min_zeta_launch(m, X2: Tensor):
    """Grid posterior + scan for the ``[n_c, d]`` contexts ``X2``; returns the workspace."""
    lib = _lib.get()
    n_c = X2.shape[0]
    ws = m.workspace(n_c)
    K = m.num_arms
    with torch.cuda.device(m.device):
        stream = _lib.stream_ptr(m.device)
        ws["ctx"].copy_(X2.to(dtype=m.dtype), non_blocking=True)
        _lib.check(lib.crs_posterior_grid(C.byref(m.state), ws["ctx"].data_ptr(), n_c, 0, K,
                                          ws["mean"].data_ptr(), ws["var"].data_ptr(), K, 0, stream),
                   "crs_posterior_grid")
        _lib.check(lib.crs_ocba_scan(ws["mean"].data_ptr(), ws["var"].data_ptr(), n_c, K, K,
                                     _lib.DTYPES[m.dtype], 0, None, None, None,
                                     ws["best_arm"].data_ptr(), ws["min_zeta"].data_ptr(),
                                     ws["min_arm"].data_ptr(), ws["tie_cnt"].data_ptr(),
                                     ws["decision"].data_ptr(), ws["scan_ws"].data_ptr(), stream),
                   "crs_ocba_scan")
    ws["valid_for"] = None  # the workspace no longer holds the grid of a gao_modellist context set
    return ws


class _MinZetaFunction(torch.autograd.Function):
    """``-min_j zeta`` with the analytic backward of ``crs_min_zeta_grad`` — what the reference
    gets from autograd through ``model.posterior`` when ``optimize_acqf`` refines the context
    (``experiments/continuous_experiments/main.py:252-263``)."""

    @staticmethod
    def forward(ctx, X: Tensor, holder) -> Tensor:
        m = holder.model
        n_c = X.shape[0]
        ws = _min_zeta_launch(m, X.detach().reshape(n_c, -1))
        lib = _lib.get()
        grad = torch.empty(n_c, m.dim, dtype=m.dtype, device=m.device)
        with torch.cuda.device(m.device):
            _lib.check(lib.crs_min_zeta_grad(C.byref(m.state), ws["ctx"].data_ptr(), n_c, ws["mean"].data_ptr(),
                                             ws["var"].data_ptr(), m.num_arms, ws["best_arm"].data_ptr(),
                                             ws["min_arm"].data_ptr(), grad.data_ptr(),
                                             _lib.stream_ptr(m.device)), "crs_min_zeta_grad")
        ctx.save_for_backward(grad)
        ctx.x_shape, ctx.x_device, ctx.x_dtype = X.shape, X.device, X.dtype
        return (-ws["min_zeta"]).to(device=X.device, dtype=X.dtype)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        (grad,) = ctx.saved_tensors
        g = grad.to(device=ctx.x_device, dtype=ctx.x_dtype) * grad_out.reshape(-1, 1)
        return g.reshape(ctx.x_shape), None


class MinZeta:
    r"""Acquisition value ``-min_{j != best} zeta(c, j)`` (continuous_context.py:39-74).

    Differentiable: when ``X.requires_grad`` the value carries an analytic backward
    (``crs_min_zeta_grad``), so it can be handed to a gradient-based optimiser over contexts the
    way the reference hands it to ``optimize_acqf``.

    Args:
        model: a :class:`B200ModelListGP` (or BoTorch ``ModelListGP``) with one outcome per arm.
    """

    def __init__(self, model) -> None:
        self.model = _as_b200(model)

    def forward(self, X: Tensor) -> Tensor:
        r"""``X``: ``batch x 1 x d_c`` contexts -> ``batch`` values (on ``X.device``)."""
        assert X.shape[-2] == 1 and X.ndim == 3
        if torch.is_grad_enabled() and X.requires_grad:
            return _MinZetaFunction.apply(X, self)
        ws = _min_zeta_launch(self.model, X.reshape(X.shape[0], -1))
        return (-ws["min_zeta"]).to(device=X.device, dtype=X.dtype)

    __call__ = forward


def find_next_arm_given_context(
    next_context: Tensor,
    model,
    kernel_scale: float = 2.0,
    randomize_ties: bool = True,
) -> Tensor:
    r"""The arm to sample at a given context (continuous_context.py:77-132): zeta over the
    K-1 non-best arms at that single context, then OCBA step 2(b) with KDE counts.

    Args:
        next_context: ``1 x d_c`` tensor.
    Returns:
        0-dim int64 tensor (on ``next_context.device``).
    """
    m = _as_b200(model)
    dec = decide(m, next_context.reshape(1, -1), randomize_ties, _lib.CRS_P_KDE, kernel_scale)
    return torch.tensor(dec.next_arm, dtype=torch.long, device=next_context.device)
