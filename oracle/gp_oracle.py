"""Exact-GP posterior oracle: a CPU (pure PyTorch) stand-in for the BoTorch ``ModelListGP`` of
``SingleTaskGP(outcome_transform=Standardize(m=1), input_transform=Normalize(d))`` sub-models
that the reference builds at ``contextual_rs/experiment_utils.py:29-40`` and queries at
``contextual_rs/contextual_rs_strategies.py:133-136``, ``contextual_rs/generalized_pcs.py:441-451``,
``contextual_rs/continuous_context.py:68,96``.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__``).

BoTorch/GPyTorch are absent from this image, so the arithmetic below restates their published
exact-GP prediction equations (posterior VALUES are unpinned at the bit level; the decisions they
lead to are pinned on the reference's recorded run, see ``trace_replay.py``):

    x~  = (x - mins) / ranges                          Normalize, bounds learnt from train X
    y~  = (y - ybar) / s                               Standardize(m=1), unbiased std
    K^  = o * c(X~, X~) + noise * I,  L = chol(K^),  alpha = K^-1 (y~ - m)
    mu~(x*)  = m + k* . alpha,        k* = o * c(x~*, X~)
    var~(x*) = o - || L^-1 k* ||^2                     (latent f, observation_noise=False)
    mu = ybar + s * mu~,  var = max(s^2 * var~, floor) floor = 1e-10 (fp64) / 1e-6 (fp32)
                                                       [clamped AFTER un-standardising: the floor
                                                        lives in gpytorch's MultivariateNormal.variance,
                                                        which BoTorch reads on the posterior that
                                                        Standardize.untransform_posterior returns]

with c = Matern-5/2 ARD (reference-era SingleTaskGP default, mirrored in-tree at
``contextual_rs/models/lce_gp.py:117-129``) or RBF ARD.  Scaled distances are computed by
direct differences (GPyTorch expands ||a||^2+||b||^2-2ab; the two agree to O(eps*||x||^2)).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

MIN_RANGE = 1e-8  # botorch Normalize(min_range)
MIN_STDV = 1e-8  # botorch Standardize(min_stdv)


def variance_floor(dtype: torch.dtype) -> float:
    """gpytorch.settings.min_variance: 1e-10 for double, 1e-6 for float."""
    return 1e-10 if dtype == torch.float64 else 1e-6


def scaled_sq_dist(a: Tensor, b: Tensor) -> Tensor:
    """Pairwise squared distance by direct differences. a: [..., p, d], b: [..., q, d]."""
    diff = a.unsqueeze(-2) - b.unsqueeze(-3)
    return (diff * diff).sum(dim=-1)


def correlation(r2: Tensor, kernel: str) -> Tensor:
    """Stationary correlation c(r) for squared scaled distance ``r2``."""
    if kernel == "matern52":
        r = r2.clamp_min(0.0).sqrt()
        s5r = math.sqrt(5.0) * r
        poly = (s5r + 1.0) + (5.0 / 3.0) * r2
        return poly * torch.exp(-s5r)
    if kernel == "rbf":
        return torch.exp(-0.5 * r2)
    raise ValueError(f"unknown kernel {kernel!r}")


class OraclePosterior:
    """Duck-type of ``GPyTorchPosterior``: ``.mean``, ``.variance``, ``.rsample``."""

    def __init__(self, mean: Tensor, variance: Tensor, joint=None):
        self.mean = mean
        self.variance = variance
        self._joint = joint  # callable returning per-arm [n_c, n_c] covariances

    def rsample(self, sample_shape=torch.Size([1]), base_samples: Optional[Tensor] = None) -> Tensor:
        """Joint draw, block-diagonal over arms and dense over contexts within an arm — what
        ``posterior.rsample`` does for the reference at ``contextual_rs/generalized_pcs.py:449-451``.
        Returns ``sample_shape x n_c x K``.  Only feasible for small n_c (O(n_c^3) per arm)."""
        if self._joint is None:
            raise RuntimeError("posterior was built without joint covariance")
        covs = self._joint()  # [K, n_c, n_c]
        K, n_c, _ = covs.shape
        S = int(torch.Size(sample_shape).numel())
        if base_samples is None:
            base_samples = torch.randn(S, n_c, K, dtype=covs.dtype)
        base_samples = base_samples.reshape(S, n_c, K).to(covs.dtype)
        jitter = 0.0
        eye = torch.eye(n_c, dtype=covs.dtype)
        for _ in range(8):
            try:
                root = torch.linalg.cholesky(covs + jitter * eye)
                break
            except RuntimeError:
                jitter = 1e-8 if jitter == 0.0 else jitter * 10.0
        else:  # pragma: no cover
            raise RuntimeError("joint covariance not PSD even with jitter")
        # y[s, c, k] = mean[c, k] + sum_c' root[k, c, c'] z[s, c', k]
        y = torch.einsum("kcd,sdk->sck", root, base_samples)
        y = y + self.mean.reshape(1, n_c, K)
        return y.reshape(*torch.Size(sample_shape), n_c, K)


class _Prior:
    def __init__(self, variance: Tensor):
        self.variance = variance


class OracleSingleTaskGP:
    """One arm.  ``spec`` keys: train_X [n,d], train_Y [n,1] (raw), lengthscale [d],
    outputscale, noise, mean_const, kernel, normalize, standardize and optional frozen transform
    constants norm_mins/norm_ranges [d], y_mean/y_std (used after ``condition_on_observations``,
    which re-uses the fitted transforms instead of re-learning them)."""

    def __init__(self, train_X: Tensor, train_Y: Tensor, lengthscale: Tensor, outputscale,
                 noise, mean_const, kernel: str = "matern52", normalize: bool = True,
                 standardize: bool = True, norm_mins: Optional[Tensor] = None,
                 norm_ranges: Optional[Tensor] = None, y_mean=None, y_std=None,
                 train_inputs_mode: str = "raw"):
        assert train_inputs_mode in ("raw", "transformed")
        self.train_inputs_mode = train_inputs_mode
        dt = train_X.dtype
        self.dtype = dt
        self.kernel = kernel
        self.raw_X = train_X
        self.raw_Y = train_Y.reshape(-1, 1)
        self.lengthscale = torch.as_tensor(lengthscale, dtype=dt).reshape(-1)
        self.outputscale = torch.as_tensor(outputscale, dtype=dt).reshape(())
        self.noise = torch.as_tensor(noise, dtype=dt).reshape(())
        self.mean_const = torch.as_tensor(mean_const, dtype=dt).reshape(())
        d = train_X.shape[-1]
        if normalize:
            if norm_mins is None:
                norm_mins = train_X.min(dim=0).values
                rng = train_X.max(dim=0).values - norm_mins
                norm_ranges = torch.where(rng < MIN_RANGE, torch.ones_like(rng), rng)
            self.norm_mins = torch.as_tensor(norm_mins, dtype=dt).reshape(-1)
            self.norm_ranges = torch.as_tensor(norm_ranges, dtype=dt).reshape(-1)
        else:
            self.norm_mins = torch.zeros(d, dtype=dt)
            self.norm_ranges = torch.ones(d, dtype=dt)
        if standardize:
            if y_mean is None:
                y_mean = self.raw_Y.mean()
                sd = self.raw_Y.std() if self.raw_Y.shape[0] > 1 else torch.ones((), dtype=dt)
                y_std = sd if float(sd) >= MIN_STDV else torch.ones((), dtype=dt)
            self.y_mean = torch.as_tensor(y_mean, dtype=dt).reshape(())
            self.y_std = torch.as_tensor(y_std, dtype=dt).reshape(())
        else:
            self.y_mean = torch.zeros((), dtype=dt)
            self.y_std = torch.ones((), dtype=dt)
        self.normalize = normalize
        self.standardize = standardize
        self._cache = None

    # ---- transforms -------------------------------------------------------------------
    def transform_inputs(self, X: Tensor) -> Tensor:
        return (X - self.norm_mins) / self.norm_ranges

    @property
    def train_inputs(self) -> Tuple[Tensor]:
        """What ``gao_modellist`` reads after ``model.posterior`` (``contextual_rs/
        contextual_rs_strategies.py:133`` then ``:176/:188``).  ``train_inputs_mode == "raw"``
        (default): the raw inputs — the behaviour of the reference's recorded runs (replaying
        ``experiments/wsc_experiments/config_b_3_mean/0000_ML_Gao.pt`` reproduces 90 of its first
        100 decisions this way and 15 with transformed inputs); ``"transformed"``: the Normalize-d
        inputs of a BoTorch >= 0.4 model in eval mode."""
        if self.train_inputs_mode == "raw":
            return (self.raw_X,)
        return (self.transform_inputs(self.raw_X),)

    # ---- exact GP ---------------------------------------------------------------------
    def _prepare(self):
        if self._cache is None:
            Xs = self.transform_inputs(self.raw_X) / self.lengthscale
            yt = (self.raw_Y - self.y_mean) / self.y_std
            n = Xs.shape[0]
            Khat = self.outputscale * correlation(scaled_sq_dist(Xs, Xs), self.kernel)
            Khat = Khat + self.noise * torch.eye(n, dtype=self.dtype)
            L = torch.linalg.cholesky(Khat)
            alpha = torch.cholesky_solve(yt - self.mean_const, L)
            self._cache = (Xs, L, alpha)
        return self._cache

    def _latent(self, X: Tensor, full_cov: bool = False):
        Xs, L, alpha = self._prepare()
        xs = self.transform_inputs(X) / self.lengthscale  # [q, d]
        kstar = self.outputscale * correlation(scaled_sq_dist(xs, Xs), self.kernel)  # [q, n]
        mu = self.mean_const + (kstar @ alpha).squeeze(-1)
        V = torch.linalg.solve_triangular(L, kstar.t(), upper=False)  # [n, q]
        if full_cov:
            kss = self.outputscale * correlation(scaled_sq_dist(xs, xs), self.kernel)
            return mu, kss - V.t() @ V
        return mu, self.outputscale - (V * V).sum(dim=0)

    def posterior(self, X: Tensor) -> OraclePosterior:
        shape = X.shape[:-1]
        mu, var = self._latent(X.reshape(-1, X.shape[-1]))
        mean = self.y_mean + self.y_std * mu
        # [3P] the floor is applied by MultivariateNormal.variance, i.e. on the un-standardised
        # posterior (Standardize.untransform_posterior rescales the covariance first)
        variance = (self.y_std * self.y_std * var).clamp_min(variance_floor(self.dtype))
        return OraclePosterior(mean.reshape(*shape, 1), variance.reshape(*shape, 1))

    def joint_covariance(self, X: Tensor) -> Tensor:
        _, cov = self._latent(X, full_cov=True)
        return self.y_std * self.y_std * cov

    def forward(self, x: Tensor) -> _Prior:
        """Prior at x in the model's own (standardised) scale: variance = k(x,x) = outputscale
        (used by ``infer_p``, ``contextual_rs/contextual_rs_strategies.py:160-167``)."""
        return _Prior(self.outputscale.expand(x.shape[:-1]).clone())

    def condition_on_observations(self, X: Tensor, Y: Tensor) -> "OracleSingleTaskGP":
        """Append observations keeping hyper-parameters and the *fitted* transforms
        (``experiments/wsc_experiments/main.py:356-362``)."""
        return OracleSingleTaskGP(
            torch.cat([self.raw_X, X.reshape(-1, self.raw_X.shape[-1])], dim=0),
            torch.cat([self.raw_Y, Y.reshape(-1, 1)], dim=0),
            self.lengthscale, self.outputscale, self.noise, self.mean_const, self.kernel,
            self.normalize, self.standardize, self.norm_mins, self.norm_ranges,
            self.y_mean, self.y_std, train_inputs_mode=self.train_inputs_mode)


class OracleModelListGP:
    """K independent arms.  ``posterior(X)`` stitches the K marginal posteriors into
    ``mean``/``variance`` of shape ``[..., K]`` exactly like ``ModelListGP.posterior``."""

    def __init__(self, *models: OracleSingleTaskGP):
        self.models = list(models)

    @property
    def train_inputs(self) -> List[Tuple[Tensor]]:
        return [m.train_inputs for m in self.models]

    def posterior(self, X: Tensor) -> OraclePosterior:
        posts = [m.posterior(X) for m in self.models]
        mean = torch.cat([p.mean for p in posts], dim=-1)
        variance = torch.cat([p.variance for p in posts], dim=-1)
        Xf = X.reshape(-1, X.shape[-1])
        joint = lambda: torch.stack([m.joint_covariance(Xf) for m in self.models], dim=0)
        return OraclePosterior(mean, variance, joint)


def model_from_description(desc: Dict, dtype: torch.dtype = torch.float64) -> OracleModelListGP:
    """Build the oracle model list from the plain 'fitted model description' dict that the
    synthetic generator and the tests share with the CUDA product (no product code imported)."""
    K = len(desc["train_X"])
    models = []
    for k in range(K):
        kw = {}
        if "norm_mins" in desc and desc["norm_mins"] is not None:
            kw.update(norm_mins=desc["norm_mins"][k].to(dtype), norm_ranges=desc["norm_ranges"][k].to(dtype))
        if "y_mean" in desc and desc["y_mean"] is not None:
            kw.update(y_mean=desc["y_mean"][k].to(dtype), y_std=desc["y_std"][k].to(dtype))
        models.append(OracleSingleTaskGP(
            desc["train_X"][k].to(dtype), desc["train_Y"][k].to(dtype),
            desc["lengthscale"][k].to(dtype), desc["outputscale"][k].to(dtype),
            desc["noise"][k].to(dtype), desc["mean_const"][k].to(dtype),
            kernel=desc.get("kernel", "matern52"), normalize=desc.get("normalize", True),
            standardize=desc.get("standardize", True), train_inputs_mode=desc.get("train_inputs", "raw"), **kw))
    return OracleModelListGP(*models)


def posterior_grid_batched(desc: Dict, X: Tensor, dtype: torch.dtype = torch.float64,
                           chunk: int = 16384) -> Tuple[Tensor, Tensor]:
    """Same arithmetic as ``OracleModelListGP.posterior`` but batched over arms with equal n
    (bmm / batched Cholesky) — the multi-threaded CPU baseline for the large configs.
    Returns (mean, variance), each ``[n_c, K]``."""
    K = len(desc["train_X"])
    n = desc["train_X"][0].shape[0]
    assert all(x.shape[0] == n for x in desc["train_X"]), "batched oracle needs equal n per arm"
    kernel = desc.get("kernel", "matern52")
    TX = torch.stack([x.to(dtype) for x in desc["train_X"]])  # [K, n, d]
    TY = torch.stack([y.to(dtype).reshape(-1, 1) for y in desc["train_Y"]])  # [K, n, 1]
    ls = desc["lengthscale"].to(dtype).reshape(K, 1, -1)
    o = desc["outputscale"].to(dtype).reshape(K, 1, 1)
    noise = desc["noise"].to(dtype).reshape(K, 1, 1)
    m = desc["mean_const"].to(dtype).reshape(K, 1, 1)
    if desc.get("normalize", True):
        if desc.get("norm_mins") is not None:
            mins = desc["norm_mins"].to(dtype).reshape(K, 1, -1)
            rng = desc["norm_ranges"].to(dtype).reshape(K, 1, -1)
        else:
            mins = TX.min(dim=1, keepdim=True).values
            rng = TX.max(dim=1, keepdim=True).values - mins
            rng = torch.where(rng < MIN_RANGE, torch.ones_like(rng), rng)
    else:
        mins, rng = torch.zeros_like(ls), torch.ones_like(ls)
    if desc.get("standardize", True):
        if desc.get("y_mean") is not None:
            ybar = desc["y_mean"].to(dtype).reshape(K, 1, 1)
            sd = desc["y_std"].to(dtype).reshape(K, 1, 1)
        else:
            ybar = TY.mean(dim=1, keepdim=True)
            sd = TY.std(dim=1, keepdim=True)
            sd = torch.where(sd >= MIN_STDV, sd, torch.ones_like(sd))
    else:
        ybar, sd = torch.zeros_like(m), torch.ones_like(m)
    Xs = (TX - mins) / rng / ls
    yt = (TY - ybar) / sd
    Khat = o * correlation(scaled_sq_dist(Xs, Xs), kernel) + noise * torch.eye(n, dtype=dtype)
    L = torch.linalg.cholesky(Khat)
    alpha = torch.cholesky_solve(yt - m, L)  # [K, n, 1]
    means, variances = [], []
    floor = variance_floor(dtype)
    for s in range(0, X.shape[0], chunk):
        xc = X[s:s + chunk].to(dtype)
        xs = (xc.unsqueeze(0) - mins) / rng / ls  # [K, q, d]
        kstar = o * correlation(scaled_sq_dist(xs, Xs), kernel)  # [K, q, n]
        mu = m.reshape(K, 1) + torch.bmm(kstar, alpha).squeeze(-1)  # [K, q]
        V = torch.linalg.solve_triangular(L, kstar.transpose(1, 2), upper=False)  # [K, n, q]
        var = o.reshape(K, 1) - (V * V).sum(dim=1)
        means.append((ybar.reshape(K, 1) + sd.reshape(K, 1) * mu).t())
        variances.append((sd.reshape(K, 1) ** 2 * var).clamp_min(floor).t())
    return torch.cat(means, dim=0).contiguous(), torch.cat(variances, dim=0).contiguous()
