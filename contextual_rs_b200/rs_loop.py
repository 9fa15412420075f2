"""Sequential ranking-and-selection loop on the B200 — the caller of the hot path
(``experiments/wsc_experiments/main.py:345-470``) restated around a device-resident posterior grid.

Per iteration the reference (i) conditions ONE arm's GP on the new observation
(``main.py:356-362``), (ii) calls ``gao_modellist`` — a full grid posterior (``:392-400``),
(iii) calls ``estimate_current_generalized_pcs`` — a second full grid posterior plus joint sampling
(``:447-455``) and (iv) takes a third grid posterior for the empirical best arms (``:459-468``).
Here the ``n_c x K`` grid stays in HBM between iterations and only the *dirty arm's column* is
recomputed (``crs_gp_prepare`` on one arm + ``crs_posterior_grid`` on one column, SURVEY.md §8f
rank 1); the decision, the PCS counts and the best-arm readout all consume that one cached grid.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional

import torch
from torch import Tensor

from . import _lib
from .contextual_rs_strategies import _p_mode, _scan_select
from .generalized_pcs import pcs_counts
from .model import B200ModelListGP


class SequentialRS:
    """GP-C-OCBA loop state: model + cached grid.  ``context_set`` is ``[n_c, d]`` (any device)."""

    def __init__(self, model: B200ModelListGP, context_set: Tensor, randomize_ties: bool = True,
                 infer_p: bool = False, use_kde: bool = False, kernel_scale: float = 1.0,
                 num_pcs_samples: int = 64, pcs_seed: int = 0,
                 rho: Optional[Callable[[Tensor], Tensor]] = None):
        self.model = model
        self.context_host = context_set.detach().cpu()
        self.ctx = context_set.to(device=model.device, dtype=model.dtype).contiguous()
        self.n_c = self.ctx.shape[0]
        self.randomize_ties = randomize_ties
        self.p_mode = _p_mode(infer_p, use_kde)
        self.kernel_scale = kernel_scale
        self.num_pcs_samples = num_pcs_samples
        self.pcs_seed = pcs_seed
        self.rho = rho if rho is not None else (lambda x: x.mean(dim=-2))
        self.iteration = 0
        self.ws = model.private_workspace(self.n_c)  # owned: not shared with the stateless entry points
        self.ws["ctx"].copy_(self.ctx)
        self._dirty: Optional[int] = None
        if getattr(model, "n_at_fit", None) is None:  # observation counts at the last fit: refit() re-fits arms that grew since
            model.n_at_fit = model.num_observations()
        # full grid once; afterwards only dirty columns
        model.posterior_tables(self.ws["ctx"], self.ws["mean"], self.ws["var"], arms=(0, model.num_arms))

    # -- model update ----------------------------------------------------------------------------
    def observe(self, arm: int, x: Tensor, y: Tensor) -> None:
        """Append one observation (``condition_on_observations``) and refresh that arm's column."""
        m = self.model
        cap_before = m.n_cap
        m.condition_on_observations(int(arm), x, y)
        if m.n_cap != cap_before:  # storage was regrown: every arm was re-prepared
            m.posterior_tables(self.ws["ctx"], self.ws["mean"], self.ws["var"], arms=(0, m.num_arms))
        else:
            m.posterior_tables(self.ws["ctx"], self.ws["mean"], self.ws["var"], arms=(int(arm), int(arm) + 1))

    def refit(self) -> None:
        """Re-fit the hyper-parameters of every arm that received observations since its last fit and rebuild
        the grid — what the reference's loop does every ``fit_frequency`` iterations with
        ``fit_modellist_with_reuse(X, Y, num_arms, old_model)`` (``experiments/wsc_experiments/main.py:345-370``,
        ``contextual_rs/experiment_utils.py:68-121``): a re-fitted arm gets fresh Normalize / Standardize constants
        and L-BFGS-fitted hyper-parameters, the others keep theirs."""
        from .fit import fit_modellist_with_reuse
        old = self.model
        desc = old.description()
        K = old.num_arms
        X = torch.cat([torch.cat([torch.full((x.shape[0], 1), float(k), dtype=torch.float64), x.to(torch.float64)], dim=-1)
                       for k, x in enumerate(desc["train_X"])])
        Y = torch.cat([y.to(torch.float64).reshape(-1, 1) for y in desc["train_Y"]])
        new = fit_modellist_with_reuse(X, Y, K, old_model=old, device=old.device, dtype=old.dtype, kernel=old.kernel)
        self.model = new
        self.ws = new.private_workspace(self.n_c)
        self.ws["ctx"].copy_(self.ctx)
        new.posterior_tables(self.ws["ctx"], self.ws["mean"], self.ws["var"], arms=(0, K))

    # -- one iteration -----------------------------------------------------------------------------
    def decide(self) -> Dict:
        """GP-C-OCBA decision on the cached grid (scan + step 2(b) in one launch)."""
        m, ws, lib = self.model, self.ws, _lib.get()
        dec = ws["decision_view"]
        with torch.cuda.device(m.device):
            stream = _lib.stream_ptr(m.device)
            _scan_select(lib, m, ws, self.n_c, self.p_mode, self.kernel_scale, stream)
            ws["decision_host"].copy_(ws["decision"], non_blocking=True)
            torch.cuda.current_stream(m.device).synchronize()
            if dec.reserved[0] != 0:
                raise RuntimeError(f"arm {dec.reserved[0] - 1}: kernel matrix not positive definite")
            if self.randomize_ties and dec.tie_count > 1:
                r = int(torch.randint(int(dec.tie_count), (1,)))
                K = m.num_arms
                _lib.check(lib.crs_ocba_resolve_tie(ws["mean"].data_ptr(), ws["var"].data_ptr(), self.n_c, K, K,
                                                    _lib.DTYPES[m.dtype], ws["best_arm"].data_ptr(),
                                                    ws["min_zeta"].data_ptr(), ws["tie_cnt"].data_ptr(), r,
                                                    ws["decision"].data_ptr(), stream), "crs_ocba_resolve_tie")
                _lib.check(lib.crs_ocba_select(C.byref(m.state), ws["ctx"].data_ptr(), ws["mean"].data_ptr(),
                                               ws["var"].data_ptr(), K, 0, self.p_mode, float(self.kernel_scale),
                                               ws["decision"].data_ptr(), stream), "crs_ocba_select")
                ws["decision_host"].copy_(ws["decision"], non_blocking=True)
                torch.cuda.current_stream(m.device).synchronize()
        return {"next_arm": int(dec.next_arm), "context_index": int(dec.context),
                "next_context": self.context_host[dec.context], "min_zeta": float(dec.min_zeta)}

    def pcs(self) -> Tensor:
        """Generalized-PCS estimate from the cached grid (Philox offset = iteration index)."""
        ws = self.ws
        counts = pcs_counts(self.model, ws["mean"], ws["var"], self.num_pcs_samples, self.pcs_seed,
                            offset=self.iteration, counts=ws["counts"], stats=ws["stats"])
        pcs_c = (counts.to(self.model.dtype) / float(self.num_pcs_samples)).unsqueeze(-1)
        return self.rho(pcs_c).squeeze(-1)

    def best_arms(self) -> Tensor:
        """Predicted best arm per context: a by-product of the scan (``main.py:459-468``)."""
        return self.ws["best_arm"].to(torch.long)

    def step(self, evaluate: Callable[[int, Tensor], Tensor], fit_frequency: Optional[int] = None) -> Dict:
        """decide -> simulate -> update -> PCS, the order of ``main.py:392-468``; with ``fit_frequency`` the
        model is re-fitted (:meth:`refit`) at the top of every ``fit_frequency``-th iteration instead of being
        conditioned only (``main.py:352-370``)."""
        if fit_frequency and self.iteration > 0 and self.iteration % int(fit_frequency) == 0:
            self.refit()
        out = self.decide()
        y = evaluate(out["next_arm"], out["next_context"])
        pcs = self.pcs()               # PCS of the model that made the decision (:447-455)
        out["best_arms"] = self.best_arms().clone()
        self.observe(out["next_arm"], out["next_context"].view(1, -1), torch.as_tensor(y).reshape(1, 1))
        out["pcs"] = pcs
        self.iteration += 1
        return out
