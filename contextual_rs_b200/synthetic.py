"""Seeded synthetic workloads for the hot path: context sets, per-arm training
data and stand-in 'fitted' hyper-parameters, returned as a plain *fitted model description*
dict of CPU tensors — the interchange format both the CUDA model handle
(:class:`contextual_rs_b200.model.B200ModelListGP`) and the test oracle are built from.

Shapes follow the reference's experiment driver: contexts ``torch.rand(n_c, d)`` under seed 0
(``experiments/wsc_experiments/main.py:237-240``) or a scrambled Sobol set
(``experiments/continuous_experiments/main.py:176-178``); training inputs are every context
``num_full_train`` times (``main.py:267-296``) or ``n`` random rows of the context set per arm
(``main.py:301-311``); targets are the reference tests' ``sine_test``
(``test/test_generalized_pcs.py:243-244``) with the arm mapped to an extra coordinate.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
from torch import Tensor

# name -> (K arms, n_c contexts, d, n obs/arm, context generator, training design)
CONFIGS = {
    "c1": dict(K=10, n_c=10, d=1, n=20, contexts="rand", train="full"),
    "c2": dict(K=100, n_c=1024, d=4, n=20, contexts="rand", train="rows"),
    "c3": dict(K=256, n_c=8192, d=4, n=20, contexts="rand", train="rows"),
    "c4": dict(K=1000, n_c=65536, d=8, n=18, contexts="sobol", train="rand"),
    "c5": dict(K=1000, n_c=16384, d=4, n=20, contexts="rand", train="rows"),
}


def make_contexts(n_c: int, d: int, kind: str = "rand", dtype=torch.float64) -> Tensor:
    if kind == "rand":
        g = torch.Generator().manual_seed(0)
        return torch.rand(n_c, d, dtype=torch.float64, generator=g).to(dtype)
    if kind == "sobol":
        eng = torch.quasirandom.SobolEngine(dimension=d, scramble=True, seed=0)
        return eng.draw(n_c, dtype=torch.float64).to(dtype)
    raise ValueError(kind)


def make_problem(K: int, n_c: int, d: int, n: int, contexts: str = "rand", train: str = "rows",
                 seed: int = 0, kernel: str = "matern52", dtype=torch.float64,
                 target_scale: float = 1.0) -> Tuple[Dict, Tensor]:
    """Returns ``(description, context_set)``; everything is generated in fp64 on the CPU and cast."""
    ctx = make_contexts(n_c, d, contexts, torch.float64)
    g = torch.Generator().manual_seed(1000 + seed)
    if train == "full":
        reps = max(1, n // n_c)
        TX = ctx.repeat(reps, 1).unsqueeze(0).expand(K, -1, -1).clone()
    elif train == "rows":
        idx = torch.randint(n_c, (K, n), generator=g)
        TX = ctx[idx]
    elif train == "rand":
        TX = torch.rand(K, n, d, dtype=torch.float64, generator=g)
    else:
        raise ValueError(train)
    n_eff = TX.shape[1]
    arm_coord = torch.arange(K, dtype=torch.float64) / max(K - 1, 1)
    full = torch.cat([arm_coord.view(K, 1, 1).expand(-1, n_eff, -1), TX], dim=-1)
    TY = torch.sin(10.0 * full).sum(dim=-1, keepdim=True) * target_scale
    TY = TY + 0.1 * torch.randn(K, n_eff, 1, dtype=torch.float64, generator=g)
    ls = 0.2 + 0.8 * torch.rand(K, d, dtype=torch.float64, generator=g)
    osc = 0.5 + 1.5 * torch.rand(K, dtype=torch.float64, generator=g)
    noise = torch.exp(math.log(1e-3) + (math.log(1e-1) - math.log(1e-3)) * torch.rand(K, dtype=torch.float64, generator=g))
    mconst = 0.1 * torch.randn(K, dtype=torch.float64, generator=g)
    desc = {
        "train_X": [TX[k].to(dtype) for k in range(K)],
        "train_Y": [TY[k].to(dtype) for k in range(K)],
        "lengthscale": ls.to(dtype),
        "outputscale": osc.to(dtype),
        "noise": noise.to(dtype),
        "mean_const": mconst.to(dtype),
        "kernel": kernel,
        "normalize": True,
        "standardize": True,
    }
    return desc, ctx.to(dtype)


def make_config(name: str, **overrides) -> Tuple[Dict, Tensor]:
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    return make_problem(**cfg)
