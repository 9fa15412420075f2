"""GP-C-OCBA allocation on the B200: drop-in for ``gao_modellist`` of
``contextual_rs/contextual_rs_strategies.py:100-202`` (same name, arguments, defaults and return
value).  All arithmetic runs in ``libcrs_b200.so``; this module only marshals pointers, keeps the
reference's RNG contract for tie randomisation and shapes the outputs.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from .model import B200ModelListGP, botorch_fingerprint, from_botorch


def _as_b200(model) -> B200ModelListGP:
    """A :class:`B200ModelListGP` is used as is; a BoTorch ``ModelListGP`` is converted once and the
    handle cached on the object, keyed on :func:`botorch_fingerprint` so that in-place changes of the
    model (new data, re-fit, ``load_state_dict``) are seen."""
    if isinstance(model, B200ModelListGP):
        return model
    fp = botorch_fingerprint(model)
    cached = getattr(model, "_crs_b200_handle", None)
    if cached is None or cached[0] != fp:
        cached = (fp, from_botorch(model))
        try:
            model._crs_b200_handle = cached
        except Exception:  # pragma: no cover
            pass
    return cached[1]


def _p_mode(infer_p: bool, use_kde: bool) -> int:
    # the reference tests infer_p first, then use_kde (strategies.py:158,168,187)
    if infer_p:
        return _lib.CRS_P_INFER
    if use_kde:
        return _lib.CRS_P_KDE
    return _lib.CRS_P_COUNT


def decide(model: B200ModelListGP, context_set: Tensor, randomize_ties: bool, p_mode: int,
           kernel_scale: float, prepare_arms: Optional[Tuple[int, int]] = None) -> _lib.Decision:
    """One GP-C-OCBA decision: grid posterior -> scan -> step 2(b).  Returns the host copy of
    ``crs_decision``; the posterior tables stay in ``model.workspace(n_c)`` for re-use by the PCS
    estimate and the best-arm readout of the same iteration."""
    lib = _lib.get()
    dev = model.device
    ctx2 = context_set.reshape(-1, context_set.shape[-1])
    n_c = ctx2.shape[0]
    ws = model.workspace(n_c)
    dec = ws["decision_view"]
    st = C.byref(model.state)
    pb, pe = prepare_arms if prepare_arms is not None else (0, 0)
    with torch.cuda.device(dev):
        stream = _lib.stream_ptr(dev)
        if ctx2.device.type == "cpu":
            host = ctx2.to(dtype=model.dtype).contiguous()
            rc = lib.crs_gao_decide_host(st, host.data_ptr(), n_c, pb, pe, p_mode, float(kernel_scale),
                                         int(bool(randomize_ties)), C.byref(ws["struct"]),
                                         C.byref(dec), stream)
            _lib.check(rc, "crs_gao_decide_host")
        else:
            ws["ctx"].copy_(ctx2.to(dtype=model.dtype), non_blocking=True)
            if pe > pb:
                model.prepare(pb, pe, relearn=False)
            K = model.num_arms
            _lib.check(lib.crs_posterior_grid(st, ws["ctx"].data_ptr(), n_c, 0, K, ws["mean"].data_ptr(),
                                              ws["var"].data_ptr(), K, 0, stream), "crs_posterior_grid")
            _scan_select(lib, model, ws, n_c, p_mode, kernel_scale, stream)
            ws["decision_host"].copy_(ws["decision"], non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            if dec.reserved[0] != 0:
                raise RuntimeError(f"arm {dec.reserved[0] - 1}: kernel matrix not positive definite")
            rc = _lib.CRS_TIES_PENDING if (randomize_ties and dec.tie_count > 1) else _lib.CRS_OK
        if rc == _lib.CRS_TIES_PENDING:
            # strategies.py:147-154 — the global RNG is consumed only when ties exist, with the
            # reference's own call so that seeded runs stay stream-compatible.
            r = int(torch.randint(int(dec.tie_count), (1,), device=context_set.device))
            K = model.num_arms
            _lib.check(lib.crs_ocba_resolve_tie(ws["mean"].data_ptr(), ws["var"].data_ptr(), n_c, K, K,
                                                _lib.DTYPES[model.dtype], ws["best_arm"].data_ptr(),
                                                ws["min_zeta"].data_ptr(), ws["tie_cnt"].data_ptr(), r,
                                                ws["decision"].data_ptr(), stream), "crs_ocba_resolve_tie")
            _lib.check(lib.crs_ocba_select(st, ws["ctx"].data_ptr(), ws["mean"].data_ptr(), ws["var"].data_ptr(),
                                           K, 0, p_mode, float(kernel_scale), ws["decision"].data_ptr(), stream),
                       "crs_ocba_select")
            ws["decision_host"].copy_(ws["decision"], non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
    ws["valid_for"] = model.grid_key(ctx2)
    return dec


def _scan_select(lib, model: B200ModelListGP, ws: dict, n_c: int, p_mode: int, kernel_scale: float,
                 stream: int, ctx_ptr: Optional[int] = None) -> None:
    """zeta scan + global argmin + OCBA step 2(b) in one launch (``crs_ocba_scan_select``)."""
    K = model.num_arms
    _lib.check(lib.crs_ocba_scan_select(C.byref(model.state), ctx_ptr or ws["ctx"].data_ptr(),
                                        ws["mean"].data_ptr(), ws["var"].data_ptr(), n_c, K, 0, p_mode,
                                        float(kernel_scale), ws["best_arm"].data_ptr(),
                                        ws["min_zeta"].data_ptr(), ws["min_arm"].data_ptr(),
                                        ws["tie_cnt"].data_ptr(), ws["decision"].data_ptr(),
                                        ws["scan_ws"].data_ptr(), stream), "crs_ocba_scan_select")


def _scan(lib, model: B200ModelListGP, ws: dict, n_c: int, stream: int) -> None:
    K = model.num_arms
    _lib.check(lib.crs_ocba_scan(ws["mean"].data_ptr(), ws["var"].data_ptr(), n_c, K, K,
                                 _lib.DTYPES[model.dtype], 0, None, None, None,
                                 ws["best_arm"].data_ptr(), ws["min_zeta"].data_ptr(),
                                 ws["min_arm"].data_ptr(), ws["tie_cnt"].data_ptr(),
                                 ws["decision"].data_ptr(), ws["scan_ws"].data_ptr(), stream),
               "crs_ocba_scan")


def gao_modellist(
    model,
    context_set: Tensor,
    randomize_ties: bool = True,
    infer_p: bool = False,
    use_kde: bool = False,
    kernel_scale: float = 2.0,
) -> Tuple[Tensor, Tensor]:
    r"""GP-C-OCBA (``contextual_rs/contextual_rs_strategies.py:100-202``).

    Args:
        model: a :class:`B200ModelListGP` (or a fitted BoTorch ``ModelListGP``, converted once
            with :func:`contextual_rs_b200.model.from_botorch`), one sub-model per arm.
        context_set: ``num_contexts x d_c`` tensor, on the CPU (as in the reference's driver) or
            on the model's GPU.
        randomize_ties, infer_p, use_kde, kernel_scale: as in the reference.

    Returns:
        ``(next_arm, next_context)``: a 0-dim int64 tensor and the selected row of
        ``context_set`` (a view, like ``context_set[min_context]`` at strategies.py:156), both on
        ``context_set.device``.
    """
    m = _as_b200(model)
    if m.num_arms < 2:
        raise ValueError("gao_modellist needs at least two arms")
    dec = decide(m, context_set, randomize_ties, _p_mode(infer_p, use_kde), kernel_scale)
    next_arm = torch.tensor(dec.next_arm, dtype=torch.long, device=context_set.device)
    return next_arm, context_set[dec.context]
