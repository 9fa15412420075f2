"""contextual_rs_b200 — B200 (sm_100a) implementation of contextual_rs's per-iteration decision
hot path: ModelListGP grid posterior, GP-C-OCBA allocation, Monte-Carlo generalized PCS and the
LEVI grid acquisition,
behind the reference's own function signatures.  All compute is in ``libcrs_b200.so``
(hand-written CUDA, C ABI in ``include/crs_b200.h``); there is no CPU fallback.
"""
from .model import B200ModelListGP, B200Posterior, from_botorch
from .contextual_rs_strategies import gao_modellist
from .continuous_context import MinZeta, find_next_arm_given_context
from .generalized_pcs import (estimate_current_generalized_pcs, estimate_lookahead_generalized_pcs_modellist,
                              empirical_best_arms)
from .levi import PredictiveEI, discrete_levi
from .optimize import optimize_acqf

__all__ = [
    "B200ModelListGP", "B200Posterior", "from_botorch", "gao_modellist", "MinZeta",
    "find_next_arm_given_context", "estimate_current_generalized_pcs", "empirical_best_arms",
    "PredictiveEI", "discrete_levi", "estimate_lookahead_generalized_pcs_modellist", "optimize_acqf",
]
