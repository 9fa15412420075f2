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


TREE_MAX_ARMS = 4102  # crs_pcs_counts_tree: 5 direct arms + 4^6 leaves + the best arm


def pcs_counts(model: B200ModelListGP, mean: Tensor, var: Tensor, num_samples: int, seed: int,
               offset: int = 0, arm_offset: int = 0, ctx_offset: int = 0,
               counts: Optional[Tensor] = None, stats: Optional[Tensor] = None,
               sampler: str = "auto", workspace: Optional[Tensor] = None) -> Tensor:
    """Correct-selection counts of device tables ``[n_c, K]``; returns int32 counts ``[n_c]``.

    ``sampler``: ``"cdf"`` = ``crs_pcs_counts_cdf`` (default: one normal + one uniform per sample
    against the tabulated conditional success probability F_c(z); any K >= 2, full arm row),
    ``"tree"`` = ``crs_pcs_counts_tree`` (max-tree stream: K - 1 competitor normals sampled lazily;
    2 <= K <= 4102), ``"flat"`` = ``crs_pcs_counts`` (one draw per arm, keyed by arm id — also serves
    arm slices via ``arm_offset``), ``"auto"`` = cdf whenever the full row is given.  All three are
    Monte-Carlo estimates of the same quantity with the same per-sample law; the streams differ, so
    counts agree within MC error, not draw for draw.
    ``stats`` (int64, optional): cdf (>= 6) -> {Philox calls, table terms, contexts with exact sharp
    competitors, contexts over the sharp limit, total successes, 0}; flat (>= 3) -> {draws executed,
    nominal - executed, -}; tree (>= 3) -> {root tests, node expansions, direct-arm tests}.
    ``workspace`` (cdf): uint8 tensor of ``crs_pcs_cdf_workspace_bytes(n_c)`` bytes; allocated per
    call (torch's caching allocator) when omitted."""
    n_c, K = mean.shape
    if sampler not in ("auto", "cdf", "tree", "flat"):
        raise ValueError(f"unknown sampler {sampler!r}")
    full_row = arm_offset == 0
    tree_ok = 2 <= K <= TREE_MAX_ARMS and full_row
    if sampler == "tree" and not tree_ok:
        raise ValueError("the tree sampler needs the full arm row with 2 <= K <= 4102 and arm_offset = 0")
    if sampler == "cdf" and not (full_row and K >= 2):
        raise ValueError("the cdf sampler needs the full arm row (arm_offset = 0) and K >= 2")
    if sampler == "auto":
        sampler = "cdf" if (full_row and K >= 2) else "flat"
    if counts is None:
        counts = torch.empty(n_c, dtype=torch.int32, device=mean.device)
    need = 6 if sampler == "cdf" else 3
    if stats is not None and stats.numel() < need:
        raise ValueError(f"stats needs room for {need} int64 values")
    lib = _lib.get()
    args = (mean.data_ptr(), var.data_ptr(), n_c, K, mean.stride(0), _lib.DTYPES[mean.dtype], int(num_samples),
            int(seed) & (2 ** 64 - 1), int(offset) & 0xFFFFFFFF)
    with torch.cuda.device(mean.device):
        stream = _lib.stream_ptr(mean.device)
        if sampler == "cdf":
            nbytes = int(lib.crs_pcs_cdf_workspace_bytes(n_c))
            if workspace is None:
                workspace = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=mean.device)
            elif workspace.numel() * workspace.element_size() < nbytes:
                raise ValueError(f"cdf workspace needs {nbytes} bytes")
            rc = lib.crs_pcs_counts_cdf(*args, int(ctx_offset), counts.data_ptr(), _lib.ptr(stats),
                                        workspace.data_ptr(), stream)
        elif sampler == "tree":
            rc = lib.crs_pcs_counts_tree(*args, int(ctx_offset), counts.data_ptr(), _lib.ptr(stats), stream)
        else:
            rc = lib.crs_pcs_counts(*args, int(arm_offset), int(ctx_offset), counts.data_ptr(), _lib.ptr(stats), stream)
    _lib.check(rc, {"cdf": "crs_pcs_counts_cdf", "tree": "crs_pcs_counts_tree", "flat": "crs_pcs_counts"}[sampler])
    return counts


def pcs_counts_base(mean: Tensor, var: Tensor, base_samples: Tensor, counts: Optional[Tensor] = None) -> Tensor:
    """Correct-selection counts from caller-supplied standard normals (``crs_pcs_counts_base``):
    ``base_samples`` is ``[S, n_c, K]`` (the shape the reference hands to ``posterior.rsample``,
    ``generalized_pcs.py:449-451``); entry (s, c, k) is the marginal draw of arm k at context c."""
    n_c, K = mean.shape
    S = int(base_samples.shape[0])
    if tuple(base_samples.shape[-2:]) != (n_c, K) or base_samples.numel() != S * n_c * K:
        raise ValueError(f"base_samples must have shape [num_samples, {n_c}, {K}], got {tuple(base_samples.shape)}")
    z = base_samples.reshape(S, n_c, K).to(device=mean.device, dtype=mean.dtype).contiguous()
    if counts is None:
        counts = torch.empty(n_c, dtype=torch.int32, device=mean.device)
    with torch.cuda.device(mean.device):
        rc = _lib.get().crs_pcs_counts_base(mean.data_ptr(), var.data_ptr(), n_c, K, mean.stride(0),
                                            _lib.DTYPES[mean.dtype], z.data_ptr(), S, counts.data_ptr(),
                                            _lib.stream_ptr(mean.device))
    _lib.check(rc, "crs_pcs_counts_base")
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

    ``base_samples`` (``num_samples x n_c x K`` standard normals, the tensor the reference passes to
    ``posterior.rsample``) are honoured as the MARGINAL draws of their (sample, context, arm) —
    ``crs_pcs_counts_base``; with ``None`` the draws come from the in-kernel Philox stream and nothing
    sample-sized exists.  Extra keyword-only arguments (not in the reference): ``seed``/``offset`` pin the Philox
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
    if not _is_strict_indicator(func_I):
        raise NotImplementedError(
            "only the strict indicator func_I = (X > 0) is fused into the PCS kernel "
            "(the only func_I the reference's callers use); there is no CPU fallback")
    if arm_set.shape[0] != m.num_arms:
        raise ValueError("arm_set must list one row per arm of the model")
    n_c = context_set.shape[0]
    ws = m.workspace(n_c)
    if not (reuse_grid and ws.get("valid_for") == m.grid_key(context_set)):
        ws["ctx"].copy_(context_set.to(dtype=m.dtype), non_blocking=True)
        m.posterior_tables(ws["ctx"], ws["mean"], ws["var"], arms=(0, m.num_arms))
        ws["valid_for"] = m.grid_key(context_set)
    if base_samples is not None:
        # generalized_pcs.py:449-451: caller-supplied standard normals, sample_shape x n_c x K, used as
        # the marginal draws of their (sample, context, arm); num_samples must agree with their count
        if base_samples.dim() < 3 or base_samples.numel() != num_samples * n_c * m.num_arms:
            raise ValueError(f"base_samples must hold num_samples x {n_c} x {m.num_arms} standard normals")
        counts = pcs_counts_base(ws["mean"], ws["var"], base_samples, counts=ws["counts"])
    else:
        if seed is None:
            seed = int(torch.randint(2 ** 62, (1,)))
        counts = pcs_counts(m, ws["mean"], ws["var"], num_samples, seed, offset,
                            counts=ws["counts"], stats=ws["stats"])
    pcs_c_est = (counts.to(m.dtype) / float(num_samples)).unsqueeze(-1)  # n_c x 1, :475
    rho_vals = rho(pcs_c_est)  # :477
    return rho_vals.squeeze(-1).to(context_set.device)  # :479


def estimate_lookahead_generalized_pcs_modellist(
    candidate: Tensor,
    model,
    model_sampler,
    context_set: Tensor,
    num_samples: int,
    base_samples: Optional[Tensor],
    func_I: Callable[[Tensor], Tensor],
    rho: Callable[[Tensor], Tensor],
    use_approximation: bool = False,
    *,
    seed: Optional[int] = None,
    fantasy_base: Optional[Tensor] = None,
) -> Tensor:
    r"""One-step look-ahead generalized PCS for a ModelListGP
    (``contextual_rs/generalized_pcs.py:203-315``): for every candidate ``(arm, context)`` the
    arm's GP is conditioned on a fantasy observation at that context — the posterior mean when
    ``model_sampler is None`` (certainty equivalent, ``:262-265``), else ``num_fantasies`` draws of
    the noisy posterior (``arm_model.fantasize``, ``:259-260``) — and the PCS of the fantasy model
    is estimated like :func:`estimate_current_generalized_pcs`, averaged over fantasies (``:311``).

    Only the fantasised arm's grid column changes, so a candidate costs one ``crs_gp_prepare`` of
    one arm, one column of ``crs_posterior_grid`` and one ``crs_pcs_counts`` launch.

    Args:
        candidate: ``n x 1 x (d_c + 1)``; column 0 is the arm index (as a float).
        model_sampler: ``None``, an ``int`` number of fantasies, or any object with a
            ``sample_shape`` (BoTorch ``MCSampler``); its base samples are NOT replicated — the
            standard-normal fantasy draws come from ``fantasy_base`` (``n x num_fantasies``) or
            the global torch RNG.
        seed: Philox seed of the PCS estimates (default: drawn from the global torch RNG).
    Returns:
        ``n``-dim tensor on ``candidate.device``.
    """
    assert context_set.dim() == 2  # :224
    assert context_set.shape[-1] + 1 == candidate.shape[-1]  # :225
    assert candidate.dim() == 3 and candidate.shape[-2] == 1  # :226
    m = _as_b200(model)
    if use_approximation:  # :228-229
        raise NotImplementedError
    if base_samples is not None:
        raise NotImplementedError(
            "base_samples cannot be honoured: samples are generated inside the kernel from a "
            "Philox stream and never materialised (pass seed= to pin the stream)")
    if not _is_strict_indicator(func_I):
        raise NotImplementedError("only the strict indicator func_I = (X > 0) is fused into the PCS kernel")
    if model_sampler is None:
        num_fantasies = 1
    elif isinstance(model_sampler, int):
        num_fantasies = model_sampler
    else:
        num_fantasies = int(tuple(model_sampler.sample_shape)[0])
    n, n_c, K = candidate.shape[0], context_set.shape[0], m.num_arms
    m.reserve(1)  # a fantasy append must not re-create the workspace
    ws = m.workspace(n_c)
    ws["ctx"].copy_(context_set.to(dtype=m.dtype), non_blocking=True)
    m.posterior_tables(ws["ctx"], ws["mean"], ws["var"], arms=(0, K))
    ws["valid_for"] = None
    if seed is None:
        seed = int(torch.randint(2 ** 62, (1,)))
    if model_sampler is not None and fantasy_base is None:
        fantasy_base = torch.randn(n, num_fantasies, dtype=torch.float64)
    # Everything below is enqueued without a host round-trip per candidate: the candidates' own posterior
    # (one [n, K] launch), the fantasy values (device arithmetic), then per (candidate, fantasy) one arm's
    # prepare + one grid column + the PCS kernels + rho — the host reads the n results once at the end.
    arms = [int(a) for a in candidate[:, 0, 0].tolist()]
    for i, arm in enumerate(arms):
        if not 0 <= arm < K:
            raise ValueError(f"candidate {i}: arm index {arm} out of range")
    xs = candidate[:, 0, 1:].to(device=m.device, dtype=torch.float64)
    at_mu, at_var = m.posterior_tables(xs.to(m.dtype).contiguous())
    pick = torch.tensor(arms, dtype=torch.long, device=m.device).unsqueeze(-1)
    mu_x = at_mu.gather(1, pick).squeeze(-1).to(torch.float64)
    var_x = at_var.gather(1, pick).squeeze(-1).to(torch.float64)
    if model_sampler is None:
        ys = mu_x.unsqueeze(-1)  # certainty equivalent, :262-265
    else:  # fantasize(): noisy posterior draws, in the outcome's original units, :259-260
        noise_orig = (m.noise * m.y_std * m.y_std).to(torch.float64)
        sd = (var_x + noise_orig[pick.squeeze(-1)]).sqrt()
        ys = mu_x.unsqueeze(-1) + sd.unsqueeze(-1) * fantasy_base.to(device=m.device, dtype=torch.float64)
    keep_m, keep_v = ws["mean"].clone(), ws["var"].clone()
    vals = []
    for i, arm in enumerate(arms):
        x = xs[i:i + 1]
        for f in range(num_fantasies):
            m.condition_on_observations(arm, x, ys[i, f].reshape(1, 1))
            m.posterior_tables(ws["ctx"], ws["mean"], ws["var"], arms=(arm, arm + 1))
            counts = pcs_counts(m, ws["mean"], ws["var"], num_samples, seed, offset=i * num_fantasies + f,
                                counts=ws["counts"], stats=ws["stats"])
            pcs_c_est = (counts.to(m.dtype) / float(num_samples)).unsqueeze(-1)
            vals.append(rho(pcs_c_est).squeeze(-1).double().reshape(()))  # :309, stays on the device
            m.pop_observations(arm, 1, prepare=False)  # the next append (or the final restore) re-prepares the arm
        ws["mean"][:, arm] = keep_m[:, arm]
        ws["var"][:, arm] = keep_v[:, arm]
    for arm in sorted(set(arms)):
        m.prepare(arm, arm + 1, relearn=False)  # back to the un-fantasised state
    out = torch.stack(vals).reshape(n, num_fantasies).mean(dim=-1)  # :311
    return out.to(device=candidate.device, dtype=candidate.dtype)


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
        if not (reuse_grid and ws.get("valid_for") == m.grid_key(context_set)):
            ws["ctx"].copy_(context_set.to(dtype=m.dtype), non_blocking=True)
            m.posterior_tables(ws["ctx"], ws["mean"], ws["var"], arms=(0, K))
            ws["valid_for"] = m.grid_key(context_set)
        _lib.check(lib.crs_ocba_scan(ws["mean"].data_ptr(), ws["var"].data_ptr(), n_c, K, K,
                                     _lib.DTYPES[m.dtype], 0, None, None, None,
                                     ws["best_arm"].data_ptr(), None, None, None, None, None, stream),
                   "crs_ocba_scan")
    return ws["best_arm"].to(device=context_set.device, dtype=torch.long)
