"""TEST INFRASTRUCTURE ONLY — replay of a recorded run of the reference's sequential loop
(``experiments/wsc_experiments/main.py:345-402`` with ``label="ML_Gao"``) through the CPU oracle.

The reference tree ships outputs of real runs, e.g. ``experiments/wsc_experiments/config_b_3_mean/
0000_ML_Gao.pt`` (10 arms x 10 contexts, Branin, two initial samples per pair, ``fit_frequency = 10``,
``gao_modellist(model, context_map, randomize_ties, kernel_scale=1.0)``): ``X[200 + i]`` is the
(arm, context) the reference chose at iteration i given ``X[:200 + i], Y[:200 + i]``.  The fitted
hyper-parameters were not stored, so the replay re-fits (``fit_oracle.fit_arm``: the restatement of
``fit_gpytorch_mll`` on BoTorch's default SingleTaskGP) following ``fit_modellist_with_reuse``
(``contextual_rs/experiment_utils.py:68-121``: an arm is re-fitted at a fit iteration iff it received
observations since its last fit) and conditions in between (``main.py:352-362``).  Agreement with
the recorded decisions pins the whole CPU restatement — transforms, fit, posterior, zeta table and
OCBA step 2(b) — on outputs of the reference itself; disagreements come from the re-fitted optimum
(scipy L-BFGS-B here vs the BoTorch version of the day) and from tie randomisation.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch import Tensor

from . import fit_oracle, gp_oracle, levi_oracle, ocba_oracle

HP_KEYS = ("lengthscale", "outputscale", "noise", "mean_const", "norm_mins", "norm_ranges", "y_mean", "y_std")


def replay_decisions(X: Tensor, Y: Tensor, context_map: Tensor, num_arms: int, num_init: int, iterations: int,
                     fit_frequency: int = 10, kernel_scale: float = 1.0,
                     modes: Tuple[str, ...] = ("raw",), policy: str = "gao",
                     weights: Tensor = None) -> Dict[str, List[Tuple[int, float]]]:
    """Decisions of the oracle at iterations 0..iterations-1 of the recorded run, for each
    ``train_inputs`` mode (the fits are shared).  ``policy``: "gao" = ``gao_modellist`` (label
    ML_Gao, main.py:391-402), "levi" = ``discrete_levi(model, context_map, weights)`` (label LEVI,
    main.py:406-413).  Returns ``{mode: [(arm, context value), ...]}`` plus, under the keys
    ``"best_arms"`` and ``"pcs_exact"``, the per-context arg-max of the posterior mean (what the
    reference stores in ``all_maximizers``, main.py:459-466) and the exact per-context PCS of the
    same posterior (quadrature; the reference stores a 64-sample MC estimate of ``rho`` of it in
    ``pcs_estimates``, main.py:447-455)."""
    from . import pcs_oracle
    X, Y = X.to(torch.float64), Y.to(torch.float64).reshape(-1, 1)
    hp = [None] * num_arms
    n_at_fit = [0] * num_arms          # experiment_utils.py:66 ``num_last_train_inputs``
    out: Dict[str, List] = {m: [] for m in modes}
    out["best_arms"], out["pcs_exact"] = [], []
    for i in range(iterations):
        Xi, Yi = X[: num_init + i], Y[: num_init + i]
        masks = [Xi[:, 0] == k for k in range(num_arms)]
        if i % fit_frequency == 0:     # main.py:346-364
            for k in range(num_arms):
                n_k = int(masks[k].sum())
                if n_k != n_at_fit[k]:  # experiment_utils.py:93-97: same inputs -> reuse
                    hp[k], _ = fit_oracle.fit_arm(Xi[masks[k]][:, 1:], Yi[masks[k]].reshape(-1))
                    n_at_fit[k] = n_k
        desc = {"train_X": [Xi[m][:, 1:] for m in masks], "train_Y": [Yi[m] for m in masks],
                "kernel": "matern52", "normalize": True, "standardize": True}
        desc.update({key: torch.stack([hp[k][key] for k in range(num_arms)]) for key in HP_KEYS})
        post = gp_oracle.model_from_description(desc).posterior(context_map)
        out["best_arms"].append(post.mean.argmax(dim=-1).tolist())
        out["pcs_exact"].append(pcs_oracle.pcs_exact_quadrature(post.mean.numpy(), post.variance.numpy()).tolist())
        for mode in modes:
            desc["train_inputs"] = mode
            model = gp_oracle.model_from_description(desc)  # conditioning keeps hyper-parameters and transforms
            if policy == "levi":
                arm, ctx = levi_oracle.discrete_levi_ref(model, context_map, weights)
            else:
                arm, ctx = ocba_oracle.gao_modellist_ref(model, context_map, True, kernel_scale=kernel_scale)
            out[mode].append((int(arm), float(ctx.reshape(-1)[0])))
    return out


def match_counts(decisions: List[Tuple[int, float]], X: Tensor, num_init: int) -> Dict[str, int]:
    """How many decisions equal the recorded ones ``X[num_init + i] = (arm, context)``."""
    arm = ctx = both = 0
    for i, (a, c) in enumerate(decisions):
        ra, rc = int(X[num_init + i, 0]), float(X[num_init + i, 1])
        ea, ec = a == ra, abs(c - rc) < 1e-12
        arm += ea
        ctx += ec
        both += ea and ec
    return {"arm": arm, "context": ctx, "both": both, "iterations": len(decisions)}
