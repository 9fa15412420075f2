"""Device-resident stand-in for the reference's ``ModelListGP`` of per-arm ``SingleTaskGP``s
(``contextual_rs/experiment_utils.py:29-40``).

The reference's strategies only touch a small duck-typed surface of the BoTorch model
(SURVEY.md §8b): ``len(model.models)``, ``model.posterior(X) -> (.mean, .variance)``,
``model.train_inputs`` and ``m.forward(x).variance`` per sub-model.  :class:`B200ModelListGP`
offers exactly that surface on top of ``libcrs_b200.so``: the fitted description (training data
and hyper-parameters) lives in HBM, ``crs_gp_prepare`` builds the per-arm Cholesky state and
``crs_posterior_grid`` evaluates the whole arm x context grid in one launch.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _dim_pad(d: int) -> int:
    p = 1
    while p < d:
        p <<= 1
    return p


class B200Posterior:
    """``.mean`` / ``.variance`` of shape ``[..., K]`` like ``ModelListGP.posterior(X)``
    (``contextual_rs/contextual_rs_strategies.py:133-136``)."""

    def __init__(self, mean: Tensor, variance: Tensor):
        self.mean = mean
        self.variance = variance

    MAX_RSAMPLE_ELEMS = 1 << 28

    def rsample(self, sample_shape=torch.Size([1]), base_samples: Optional[Tensor] = None) -> Tensor:
        """``sample_shape x [..., K]`` draws like ``posterior.rsample`` (``generalized_pcs.py:449-451``),
        from the MARGINALS: ``mean + sqrt(variance) * z`` with ``z = base_samples`` or fresh standard
        normals.  Materialising samples is only meant for the reference's small call shapes — the PCS
        kernels draw in registers — so anything above 2^28 elements is refused."""
        shape = torch.Size(sample_shape) + self.mean.shape
        if shape.numel() > self.MAX_RSAMPLE_ELEMS:
            raise NotImplementedError(
                "a sample tensor of this size is what the PCS kernels exist to avoid; "
                "use estimate_current_generalized_pcs")
        if base_samples is None:
            base_samples = torch.randn(shape, dtype=self.mean.dtype, device=self.mean.device)
        z = base_samples.reshape(shape).to(device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.variance.sqrt() * z


class _Prior:
    def __init__(self, variance: Tensor):
        self.variance = variance


class B200ArmView:
    """``model.models[i]``: the slice of the packed state that belongs to one arm."""

    def __init__(self, parent: "B200ModelListGP", index: int):
        self._parent = parent
        self.index = index

    @property
    def train_inputs(self) -> Tuple[Tensor]:
        return self._parent.train_inputs[self.index]

    def forward(self, x: Tensor) -> _Prior:
        """Prior at ``x`` in the arm's standardised scale: variance = k(x, x) = outputscale
        (what ``m.forward(x).variance`` yields for ``infer_p``, strategies.py:160-167)."""
        o = self._parent.outputscale[self.index].to(self._parent.dtype)
        return _Prior(o.expand(x.shape[:-1]).clone())

    def posterior(self, X: Tensor) -> B200Posterior:
        return self._parent.posterior(X, arms=(self.index, self.index + 1))


class B200ModelListGP:
    """K independent exact GPs (one per arm) packed for the B200 kernels.

    ``desc`` is the *fitted model description* dict (see ``synthetic.make_problem``): per-arm raw
    ``train_X [n_k, d]`` / ``train_Y [n_k, 1]``, ``lengthscale [K, d]``, ``outputscale``,
    ``noise``, ``mean_const [K]``, ``kernel`` ("matern52" | "rbf"), ``normalize``,
    ``standardize`` and optionally frozen transform constants ``norm_mins/norm_ranges [K, d]``,
    ``y_mean/y_std [K]``.  ``dtype`` is the compute dtype of every table (fp64 as the
    reference's default, ``experiments/wsc_experiments/main.py:184``, or fp32).

    ``train_inputs`` (or ``desc["train_inputs"]``) selects what ``model.train_inputs`` holds — the
    tensors OCBA step 2(b) compares with the chosen context (exact-match counts / KDE weights,
    ``contextual_rs_strategies.py:168-193``): ``"raw"`` (default) = the raw inputs, which is what the
    reference's own recorded runs show (replaying ``experiments/wsc_experiments/config_b_3_mean/
    0000_ML_Gao.pt`` reproduces 90 of its first 100 decisions with raw inputs, 15 with transformed
    ones — ``tests/golden/reference_trace_ml_gao.json``); ``"transformed"`` = the Normalize-d inputs
    a BoTorch >= 0.4 model exposes in eval mode.
    """

    def __init__(self, desc: Dict, device="cuda", dtype: torch.dtype = torch.float64,
                 n_cap: Optional[int] = None, headroom: int = 8, train_inputs: Optional[str] = None):
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _lib.require_cuda(self.device)
        if dtype not in _lib.DTYPES:
            raise ValueError(f"unsupported dtype {dtype}")
        self.dtype = dtype
        self.kernel = desc.get("kernel", "matern52")
        K = len(desc["train_X"])
        d = int(desc["train_X"][0].shape[-1])
        if d > _lib.CRS_MAX_DIM:
            raise ValueError(f"context dimension {d} > {_lib.CRS_MAX_DIM}")
        self.num_arms, self.dim = K, d
        self._n_host = [int(x.shape[0]) for x in desc["train_X"]]
        need = max(self._n_host) if self._n_host else 0
        self.n_cap = n_cap if n_cap is not None else max(8, _round_up(need + headroom, 8))
        if self.n_cap < need:
            raise ValueError("n_cap smaller than the largest arm")
        self.normalize = bool(desc.get("normalize", True))
        self.standardize = bool(desc.get("standardize", True))
        self.train_inputs_mode = train_inputs if train_inputs is not None else desc.get("train_inputs", "raw")
        if self.train_inputs_mode not in ("raw", "transformed"):
            raise ValueError(f"train_inputs must be 'raw' or 'transformed', got {self.train_inputs_mode!r}")
        self._alloc(desc)
        self.models: List[B200ArmView] = [B200ArmView(self, k) for k in range(K)]
        self._ws_cache: Dict[int, dict] = {}
        self._update_count = 0
        self.prepare()

    # ------------------------------------------------------------------ storage
    def _alloc(self, desc: Dict) -> None:
        K, d, ncap = self.num_arms, self.dim, self.n_cap
        dev, f64, T = self.device, torch.float64, self.dtype
        tx = torch.zeros(K, ncap, d, dtype=f64)
        ty = torch.zeros(K, ncap, dtype=f64)
        for k in range(K):
            n = self._n_host[k]
            tx[k, :n] = desc["train_X"][k].to(f64)
            ty[k, :n] = desc["train_Y"][k].to(f64).reshape(-1)
        self.train_x = tx.to(dev)
        self.train_y = ty.to(dev)
        self.n_obs = torch.tensor(self._n_host, dtype=torch.int32, device=dev)
        self.lengthscale = torch.as_tensor(desc["lengthscale"], dtype=f64).reshape(K, d).contiguous().to(dev)
        self.outputscale = torch.as_tensor(desc["outputscale"], dtype=f64).reshape(K).contiguous().to(dev)
        self.noise = torch.as_tensor(desc["noise"], dtype=f64).reshape(K).contiguous().to(dev)
        self.mean_const = torch.as_tensor(desc["mean_const"], dtype=f64).reshape(K).contiguous().to(dev)
        flags = 0
        if self.normalize and desc.get("norm_mins") is not None:
            self.norm_mins = torch.as_tensor(desc["norm_mins"], dtype=f64).reshape(K, d).contiguous().to(dev)
            self.norm_ranges = torch.as_tensor(desc["norm_ranges"], dtype=f64).reshape(K, d).contiguous().to(dev)
        else:
            self.norm_mins = torch.zeros(K, d, dtype=f64, device=dev)
            self.norm_ranges = torch.ones(K, d, dtype=f64, device=dev)
            if self.normalize:
                flags |= _lib.CRS_LEARN_NORMALIZE
        if self.standardize and desc.get("y_mean") is not None:
            self.y_mean = torch.as_tensor(desc["y_mean"], dtype=f64).reshape(K).contiguous().to(dev)
            self.y_std = torch.as_tensor(desc["y_std"], dtype=f64).reshape(K).contiguous().to(dev)
        else:
            self.y_mean = torch.zeros(K, dtype=f64, device=dev)
            self.y_std = torch.ones(K, dtype=f64, device=dev)
            if self.standardize:
                flags |= _lib.CRS_LEARN_STANDARDIZE
        self._learn_flags = flags
        dp = _dim_pad(d)
        self.dim_pad = dp
        self.xn = torch.zeros(K, ncap, d, dtype=T, device=dev)
        self.xs = torch.zeros(K, ncap, dp, dtype=T, device=dev)
        self.alpha = torch.zeros(K, ncap, dtype=T, device=dev)
        self.linv_t = torch.zeros(K, int(_lib.get().crs_linv_stride(ncap, _lib.DTYPES[T])), dtype=T, device=dev)
        self.arm_head = torch.zeros(K, _lib.CRS_HEAD_ELEMS, dtype=T, device=dev)
        self.status = torch.zeros(K, dtype=torch.int32, device=dev)
        raw = _lib.CRS_TRAIN_INPUTS_RAW if self.train_inputs_mode == "raw" else 0
        self._state = self._make_state(flags | raw)
        self._state_frozen = self._make_state(raw)

    def _make_state(self, flags: int) -> _lib.GPState:
        p = _lib.ptr
        return _lib.GPState(
            self.num_arms, self.n_cap, self.dim, self.dim_pad, _lib.KERNELS[self.kernel],
            _lib.DTYPES[self.dtype], flags, max(self._n_host) if self._n_host else 0,
            p(self.train_x), p(self.train_y), p(self.n_obs), p(self.lengthscale),
            p(self.outputscale), p(self.noise), p(self.mean_const),
            p(self.norm_mins), p(self.norm_ranges), p(self.y_mean), p(self.y_std),
            p(self.xn), p(self.xs), p(self.alpha), p(self.linv_t),
            p(self.arm_head), p(self.status))

    @property
    def state(self) -> _lib.GPState:
        """The ``crs_gp_state`` (transforms frozen) to pass to grid / select entry points."""
        return self._state_frozen

    # ------------------------------------------------------------------ prepare / update
    def prepare(self, arm_begin: int = 0, arm_end: Optional[int] = None, relearn: Optional[bool] = None) -> None:
        """``crs_gp_prepare`` on arms [arm_begin, arm_end).  The very first call learns the
        Normalize / Standardize constants (a fresh fit); later calls keep them frozen, which is
        the ``condition_on_observations`` semantics of ``experiments/wsc_experiments/main.py:356-362``."""
        arm_end = self.num_arms if arm_end is None else arm_end
        first = not getattr(self, "_prepared", False)
        st = self._state if (relearn if relearn is not None else first) else self._state_frozen
        with torch.cuda.device(self.device):
            rc = _lib.get().crs_gp_prepare(C.byref(st), arm_begin, arm_end, _lib.stream_ptr(self.device))
        _lib.check(rc, "crs_gp_prepare")
        self._prepared = True
        self._update_count += 1  # cached grids (reuse_grid=True, grid_key) describe the old state
        for ws in self._ws_cache.values():
            ws["valid_for"] = None

    def check_status(self) -> None:
        """Synchronising check of the per-arm Cholesky status."""
        bad = torch.nonzero(self.status).flatten()
        if bad.numel():
            k = int(bad[0])
            raise RuntimeError(f"arm {k}: kernel matrix not positive definite "
                               f"(pivot {int(self.status[k]) - 1})")

    def condition_on_observations(self, arm: int, X: Tensor, Y: Tensor) -> "B200ModelListGP":
        """Append observations to one arm and rebuild only that arm's state (hyper-parameters
        and transforms kept) — ``models[i].condition_on_observations(X, Y)`` +
        ``ModelListGP(*models)`` of ``experiments/wsc_experiments/main.py:356-362``, in place."""
        X = X.reshape(-1, self.dim).to(device=self.device, dtype=torch.float64)
        Y = Y.reshape(-1).to(device=self.device, dtype=torch.float64)
        n, add = self._n_host[arm], X.shape[0]
        if n + add > self.n_cap:
            self._grow(_round_up(n + add + 8, 8))
        self.train_x[arm, n:n + add] = X
        self.train_y[arm, n:n + add] = Y
        self._n_host[arm] = n + add
        self.n_obs[arm] = n + add
        self._state.n_max = self._state_frozen.n_max = max(self._n_host)
        self.prepare(arm, arm + 1, relearn=False)
        return self

    def reserve(self, extra: int = 1) -> None:
        """Make room for ``extra`` more observations on any arm WITHOUT touching the data (grows
        the row capacity now, so that later appends do not re-create the workspaces)."""
        need = max(self._n_host) + int(extra)
        if need > self.n_cap:
            self._grow(_round_up(need + 8, 8))

    def pop_observations(self, arm: int, count: int = 1, prepare: bool = True) -> "B200ModelListGP":
        """Undo the last ``count`` appends to one arm (fantasy roll-back) and rebuild its state
        (``prepare=False``: the caller re-prepares the arm itself, e.g. with the next fantasy append)."""
        n = self._n_host[arm]
        if count < 0 or count > n:
            raise ValueError("cannot remove more observations than the arm holds")
        self._n_host[arm] = n - count
        self.n_obs[arm] = n - count
        self._state.n_max = self._state_frozen.n_max = max(self._n_host)
        if prepare:
            self.prepare(arm, arm + 1, relearn=False)
        return self

    def _grow(self, new_cap: int) -> None:
        if new_cap > _lib.CRS_MAX_OBS:
            raise RuntimeError(f"more than {_lib.CRS_MAX_OBS} observations on one arm is not supported yet")
        desc = self.description()
        self.n_cap = new_cap
        self._alloc(desc)
        self._prepared = True  # transforms come frozen from description()
        self.prepare(relearn=False)

    def description(self) -> Dict:
        """Round-trip back to a (CPU) fitted model description with the transforms frozen."""
        return {
            "train_X": [self.train_x[k, :n].cpu() for k, n in enumerate(self._n_host)],
            "train_Y": [self.train_y[k, :n].cpu().reshape(-1, 1) for k, n in enumerate(self._n_host)],
            "lengthscale": self.lengthscale.cpu(), "outputscale": self.outputscale.cpu(),
            "noise": self.noise.cpu(), "mean_const": self.mean_const.cpu(),
            "kernel": self.kernel, "normalize": self.normalize, "standardize": self.standardize,
            "norm_mins": self.norm_mins.cpu() if self.normalize else None,
            "norm_ranges": self.norm_ranges.cpu() if self.normalize else None,
            "y_mean": self.y_mean.cpu() if self.standardize else None,
            "y_std": self.y_std.cpu() if self.standardize else None,
            "train_inputs": self.train_inputs_mode,
        }

    # ------------------------------------------------------------------ duck-typed surface
    @property
    def train_inputs(self) -> List[Tuple[Tensor]]:
        """One 1-tuple per arm — what ``gao_modellist`` reads at strategies.py:176/188 after
        ``model.posterior`` ran: the raw inputs (``train_inputs="raw"``, the behaviour of the
        reference's recorded runs) or the normalised ones (``"transformed"``, BoTorch >= 0.4 eval mode)."""
        return [(self.xn[k, :n],) for k, n in enumerate(self._n_host)]

    def num_observations(self) -> List[int]:
        return list(self._n_host)

    def _as_ctx(self, X: Tensor) -> Tensor:
        X2 = X.reshape(-1, X.shape[-1])
        if X2.shape[-1] != self.dim:
            raise ValueError(f"expected contexts of dimension {self.dim}, got {X2.shape[-1]}")
        return X2.to(device=self.device, dtype=self.dtype, non_blocking=True).contiguous()

    def posterior_tables(self, ctx_dev: Tensor, mean: Optional[Tensor] = None, var: Optional[Tensor] = None,
                         arms: Optional[Tuple[int, int]] = None) -> Tuple[Tensor, Tensor]:
        """Launch ``crs_posterior_grid`` for a device-resident ``[n_c, d]`` context tensor.  With
        ``mean``/``var`` given (``[n_c, K]``) only the columns of ``arms`` are rewritten — the
        dirty-arm update of the sequential loop."""
        n_c = ctx_dev.shape[0]
        a0, a1 = arms if arms is not None else (0, self.num_arms)
        if mean is None:
            width = a1 - a0 if arms is not None else self.num_arms
            mean = torch.empty(n_c, width, dtype=self.dtype, device=self.device)
            var = torch.empty(n_c, width, dtype=self.dtype, device=self.device)
            col = 0
        else:
            col = a0
        with torch.cuda.device(self.device):
            rc = _lib.get().crs_posterior_grid(C.byref(self._state_frozen), ctx_dev.data_ptr(), n_c, a0, a1,
                                               mean.data_ptr(), var.data_ptr(), mean.shape[1], col,
                                               _lib.stream_ptr(self.device))
        _lib.check(rc, "crs_posterior_grid")
        return mean, var

    def posterior(self, X: Tensor, arms: Optional[Tuple[int, int]] = None) -> B200Posterior:
        """``model.posterior(X)``: ``X`` is ``[n_c, d]`` or ``[b, 1, d]``; the result has the
        reference's shapes ``[n_c, K]`` / ``[b, 1, K]`` and lives on this model's device."""
        mean, var = self.posterior_tables(self._as_ctx(X), arms=arms)
        shape = tuple(X.shape[:-1]) + (mean.shape[-1],)
        return B200Posterior(mean.reshape(shape), var.reshape(shape))

    # ------------------------------------------------------------------ decision workspace
    WS_CACHE_SIZE = 4

    def _make_workspace(self, n_c: int) -> dict:
        dev, T, K = self.device, self.dtype, self.num_arms
        ws = {
            "ctx": torch.empty(n_c, self.dim, dtype=T, device=dev),
            "mean": torch.empty(n_c, K, dtype=T, device=dev),
            "var": torch.empty(n_c, K, dtype=T, device=dev),
            "best_arm": torch.empty(n_c, dtype=torch.int32, device=dev),
            "min_zeta": torch.empty(n_c, dtype=T, device=dev),
            "min_arm": torch.empty(n_c, dtype=torch.int32, device=dev),
            "tie_cnt": torch.empty(n_c, dtype=torch.int32, device=dev),
            "decision": torch.zeros(64, dtype=torch.uint8, device=dev),
            "scan_ws": torch.zeros(int(_lib.get().crs_scan_workspace_bytes()), dtype=torch.uint8, device=dev),
            "decision_host": torch.zeros(64, dtype=torch.uint8).pin_memory(),
            "counts": torch.empty(n_c, dtype=torch.int32, device=dev),
            "stats": torch.zeros(8, dtype=torch.int64, device=dev),
            "valid_for": None,  # a fresh workspace holds no grid yet
        }
        p = _lib.ptr
        ws["struct"] = _lib.DecideWS(p(ws["ctx"]), p(ws["mean"]), p(ws["var"]), p(ws["best_arm"]),
                                     p(ws["min_zeta"]), p(ws["min_arm"]), p(ws["tie_cnt"]),
                                     p(ws["decision"]), p(ws["scan_ws"]))
        ws["decision_view"] = _lib.Decision.from_address(ws["decision_host"].data_ptr())
        return ws

    def workspace(self, n_c: int) -> dict:
        """Shared per-``n_c`` device buffers of a stateless call (context copy, ``[n_c, K]`` tables,
        scan outputs, counts): a true LRU over the last ``WS_CACHE_SIZE`` batch sizes.  Whether its
        tables hold the grid of a given context set is recorded in ``ws["valid_for"]``
        (:meth:`grid_key`); consumers that keep a grid ACROSS calls must not use this cache — see
        :meth:`private_workspace`."""
        ws = self._ws_cache.pop(n_c, None)
        if ws is None:
            ws = self._make_workspace(n_c)
            while len(self._ws_cache) >= self.WS_CACHE_SIZE:
                self._ws_cache.pop(next(iter(self._ws_cache)))  # least recently used first
        self._ws_cache[n_c] = ws  # most recently used last
        return ws

    def private_workspace(self, n_c: int) -> dict:
        """Buffers owned by ONE long-lived consumer (sequential loops, sharded deciders): never
        shared with, evicted by or overwritten through the stateless entry points, whatever batch
        sizes those are called with on the same model."""
        return self._make_workspace(n_c)

    def grid_key(self, context_set: Tensor) -> tuple:
        """Identity of "the grid of this context set under the current model state": storage address,
        shape, the tensor's in-place version counter and the model's own update counter — an in-place
        edit of the contexts or any re-prepare invalidates a cached grid (``reuse_grid=True``)."""
        return (context_set.data_ptr(), tuple(context_set.shape), int(context_set._version), self._update_count)


def description_from_botorch(model) -> Dict:
    """The *fitted model description* dict of a BoTorch ``ModelListGP`` of ``SingleTaskGP`` sub-models,
    read by attribute path (pure host code).  The caller's model is NOT touched: no
    ``train()`` / ``eval()`` toggling (that would drop GPyTorch's prediction caches and, on a model
    produced by ``condition_on_observations``, revert ``train_inputs`` to the stale pre-conditioning
    copy).  ``train_inputs[0]`` is read as the model shows it now; if BoTorch holds the
    Normalize-transformed copy (``_has_transformed_inputs``, eval mode of BoTorch >= 0.4) the raw inputs
    are recovered from ``_original_train_inputs`` when its length still matches ``train_targets``
    (not the case after conditioning) and otherwise by ``input_transform.untransform``.
    ``desc["train_inputs"]`` records what ``model.train_inputs`` holds in eval mode under the
    installed BoTorch (raw, or transformed), i.e. what the reference's ``gao_modellist`` sees.

    BoTorch is not installed in this image: the function is written against the documented attribute
    names and has only been exercised with a duck-typed stand-in (``tests/test_host_cpu.py``)."""
    tx, ty, ls, osc, noise, mc, mins, rngs, ym, ys = [], [], [], [], [], [], [], [], [], []
    kernel = None
    detected = "raw"
    for k, m in enumerate(model.models):
        X_seen = m.train_inputs[0].detach()
        Yt = m.train_targets.detach()
        it = getattr(m, "input_transform", None)
        holds_transformed = it is not None and bool(getattr(m, "_has_transformed_inputs", False))
        if it is not None and (holds_transformed or hasattr(m, "_set_transformed_inputs")):
            detected = "transformed"  # eval mode swaps in the transformed inputs under this BoTorch
        if holds_transformed:
            orig = getattr(m, "_original_train_inputs", None)
            if orig is not None and orig.shape == X_seen.shape:
                X = orig.detach()
            else:
                X = it.untransform(X_seen).detach()
        else:
            X = X_seen
        if X.shape[0] != Yt.reshape(-1).shape[0]:
            raise ValueError(f"arm {k}: {X.shape[0]} training inputs but {Yt.reshape(-1).shape[0]} targets")
        cov = m.covar_module
        base = getattr(cov, "base_kernel", cov)
        name = type(base).__name__.lower()
        kern = "rbf" if "rbf" in name else "matern52"
        if kern == "matern52" and abs(float(getattr(base, "nu", 2.5)) - 2.5) > 1e-12:
            raise NotImplementedError("only Matern nu=2.5 and RBF kernels are supported")
        if kernel not in (None, kern):
            raise NotImplementedError("all arms must share one kernel family")
        kernel = kern
        d = X.shape[-1]
        ls.append(base.lengthscale.detach().reshape(-1).expand(d).cpu())
        osc.append(cov.outputscale.detach().reshape(()).cpu() if hasattr(cov, "outputscale") else torch.ones(()))
        noise.append(m.likelihood.noise.detach().reshape(-1)[0].cpu())
        mc.append(m.mean_module.constant.detach().reshape(()).cpu())
        if it is not None:
            b = it.bounds.detach() if hasattr(it, "bounds") else torch.stack([it.mins, it.mins + it.ranges]).squeeze(1)
            mins.append(b[0].reshape(-1).cpu())
            rngs.append((b[1] - b[0]).reshape(-1).cpu())
        else:
            mins.append(torch.zeros(d))
            rngs.append(torch.ones(d))
        ot = getattr(m, "outcome_transform", None)
        if ot is not None:
            ym.append(ot.means.detach().reshape(()).cpu())
            ys.append(ot.stdvs.detach().reshape(()).cpu())
            Y = Yt.cpu() * ys[-1] + ym[-1]  # train_targets are stored standardised
        else:
            ym.append(torch.zeros(()))
            ys.append(torch.ones(()))
            Y = Yt.cpu()
        tx.append(X.cpu())
        ty.append(Y.reshape(-1, 1))
    desc = {"train_X": tx, "train_Y": ty, "lengthscale": torch.stack(ls), "outputscale": torch.stack(osc),
            "noise": torch.stack(noise), "mean_const": torch.stack(mc), "kernel": kernel,
            "normalize": True, "standardize": True, "norm_mins": torch.stack(mins),
            "norm_ranges": torch.stack(rngs), "y_mean": torch.stack(ym), "y_std": torch.stack(ys),
            "train_inputs": detected}
    return desc


def botorch_fingerprint(model) -> tuple:
    """Cheap identity of a BoTorch model's data and hyper-parameters: object ids, shapes and the
    in-place version counters of the training tensors and of every parameter / buffer.  Used to
    invalidate the cached B200 handle when the caller mutates the model in place
    (``load_state_dict``, ``set_train_data``, a re-fit)."""
    parts = []
    for m in model.models:
        X, Y = m.train_inputs[0], m.train_targets
        item = [id(m), tuple(X.shape), int(getattr(X, "_version", 0)), X.data_ptr(),
                tuple(Y.shape), int(getattr(Y, "_version", 0)), Y.data_ptr(), bool(getattr(m, "training", False))]
        params = getattr(m, "parameters", None)
        if callable(params):
            for prm in params():
                item.append((prm.data_ptr(), int(prm._version)))
        buffers = getattr(m, "buffers", None)
        if callable(buffers):
            for b in buffers():
                item.append((b.data_ptr(), int(b._version)))
        parts.append(tuple(item))
    return tuple(parts)


def from_botorch(model, device="cuda", dtype: Optional[torch.dtype] = None,
                 train_inputs: Optional[str] = None) -> B200ModelListGP:
    """Pack a fitted BoTorch ``ModelListGP`` for the B200 kernels (:func:`description_from_botorch`).
    ``train_inputs=None`` keeps the detected eval-mode semantics; pass "raw" / "transformed" to force."""
    desc = description_from_botorch(model)
    return B200ModelListGP(desc, device=device, dtype=dtype or desc["train_X"][0].dtype,
                           train_inputs=train_inputs or desc["train_inputs"])
