#!/usr/bin/env python
"""bench.py — the GP-C-OCBA decision hot path on BASELINE.json's headline configuration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--workload headline|c2|c3|c4|c5|loop] [--exchange mailbox|nccl] [--partition contexts|arms]

Default workload `headline` (BASELINE.json north_star target, config 4 shapes), the same at every N:
a *step* is one whole GP-C-OCBA decision on a model whose caches are cold — per-arm Cholesky state
(crs_gp_prepare, all arms), ModelListGP posterior over the 1,000-arm x 65,536-context grid
(crs_posterior_grid), zeta scan + global argmin + OCBA step 2(b) (crs_ocba_scan_select) — followed by
one 2^16-sample generalized-PCS estimate on the same grid (crs_pcs_counts_cdf): what one iteration of the
reference's loop does with `gao_modellist(model, context_set)` and
`estimate_current_generalized_pcs(model, arm_set, context_set, 2**16, None, func_I, rho)`
(contextual_rs/contextual_rs_strategies.py:100-202, contextual_rs/generalized_pcs.py:318-479,
experiments/wsc_experiments/main.py:392-456).

  N > 1 : ONE step is sharded over the ranks by contexts ("strong" scaling): every rank runs prepare +
          its slice of the grid + scan/select + PCS; the ranks' local minimisers are folded and the PCS
          totals summed ON THE DEVICE through peer-mapped mailboxes over NVLink (one 64-byte store per
          peer each; --exchange nccl uses one all-gather + one all-reduce instead).
          --partition arms measures the partition the north_star describes (arms sharded, all-gather of
          per-context bests, all-to-all re-shard for the PCS) on the same step.
  value : decisions/s (each with its PCS estimate), every input resident in HBM, device timed.
  e2e   : the same step through the public API with the context set in pinned HOST memory:
          host->device copy of the contexts, device->host read of the decision and of the PCS value
          inside every timed step.
  The L2 is flushed (256 MiB write) between timed steps; each step is bracketed by its own pair of
  CUDA events on the launching stream; per-step durations are summed and the max over ranks is taken.
`--impl reference` times the CPU restatement of the reference (oracle/, PyTorch on the host cores;
BoTorch/GPyTorch are not installable here) on the same workload, every step a bounded
sample of it scaled linearly.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner
# to fd 1 when NCCL_DEBUG is set), so fd 1 is pointed at stderr for the whole run and the result line
# goes to a private duplicate of the original stdout.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_RESULT_FD, (json.dumps(obj) + "\n").encode())


import torch  # noqa: E402

PCS_SAMPLES = 1 << 16
WORKLOADS = {
    "headline": "c4 + PCS: one GP-C-OCBA decision (cold caches) + one 2^16-sample generalized-PCS estimate, "
                "1,000 arms x 65,536 Sobol contexts, 8-D, 18 obs/arm, Matern-5/2",
    "c2": "c2: GP-C-OCBA decision, 100 arms x 1,024 contexts, 4-D, 20 obs/arm, Matern-5/2",
    "c3": "c3: GP-C-OCBA decision, 256 arms x 8,192 contexts, 4-D, 20 obs/arm, Matern-5/2",
    "c4": "c4: GP-C-OCBA decision, 1,000 arms x 65,536 Sobol contexts, 8-D, 18 obs/arm, Matern-5/2",
    "c5": "c5: GP-C-OCBA decision, 1,000 arms x 16,384 contexts, 4-D, 20 obs/arm, Matern-5/2",
    "loop": "c5 loop: sequential R&S, 1,000 arms x 16,384 contexts, 4-D, 20 obs/arm growing by one per iteration; "
            "every iteration = GP-C-OCBA decision on the cached grid + 2^16-sample PCS + dirty-arm update",
}
SHAPES = {"headline": "c4", "loop": "c5"}
METRIC = {"loop": ("rs_loop_iterations_per_sec", "iterations/s")}


def workload_config(name, world, args):
    """The `config` object — identical for the GPU arm and `--impl reference` (same flags -> same dict)."""
    from contextual_rs_b200.synthetic import CONFIGS
    c = CONFIGS[SHAPES.get(name, name)]
    with_pcs = name in ("headline", "loop")
    if world == 1:
        sharding = "none"
    elif name == "c2" and not args.shard:
        sharding = f"replicas x{world}: one independent decision stream per GPU, no collective on the data path"
    else:
        sharding = f"{args.partition}/{world}: one step sharded over the ranks"
    cfg = {"workload": WORKLOADS[name], "arms": c["K"], "contexts": c["n_c"], "dim": c["d"], "obs_per_arm": c["n"],
           "pcs_samples": PCS_SAMPLES if with_pcs else 0, "sharding": sharding,
           "l2": "flushed between timed steps (256 MiB write)" if not args.no_flush else "not flushed"}
    return cfg


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(workload, kernel, dtype):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full
    capture of the same workload at N = 1 (profiles/r02_traffic.json, else r01d), else None."""
    for fn in ("r02_traffic.json", "r01d_traffic.json"):
        path = os.path.join(ROOT, "profiles", fn)
        if os.path.exists(path):
            with open(path) as f:
                v = json.load(f).get(workload, {}).get(f"{kernel}:{dtype}")
            if v is not None:
                return v
    if dtype == "f64":
        path = os.path.join(ROOT, "profiles", "r01d_traffic.json")
        if os.path.exists(path):
            with open(path) as f:
                return json.load(f).get(workload, {}).get(kernel)
    return None


def algorithmic_flops(K, n_c, n, d):
    """distance + kernel eval + mean dot + triangular solve + squared norm."""
    return K * n_c * (n * (3 * d + 10) + 2 * n + n * n)


# ------------------------------------------------------------------------------------------- CPU
class _TableModel:
    def __init__(self, mean, var, train_inputs, K):
        self._p = type("P", (), {"mean": mean, "variance": var})()
        self.models = [None] * K
        self.train_inputs = train_inputs

    def posterior(self, X):
        return self._p


def cpu_decision_step(desc, ctx, variant):
    """One decision with the oracle on the host cores (cold model caches, like the GPU step)."""
    from oracle import gp_oracle, ocba_oracle
    if variant == "batched":
        mean, var = gp_oracle.posterior_grid_batched(desc, ctx)
        K = len(desc["train_X"])
        TX = torch.stack(desc["train_X"])
        mins = TX.min(dim=1, keepdim=True).values
        rng = TX.max(dim=1, keepdim=True).values - mins
        rng = torch.where(rng < 1e-8, torch.ones_like(rng), rng)
        xn = (TX - mins) / rng
        model = _TableModel(mean, var, [(xn[k],) for k in range(K)], K)
    else:
        model = gp_oracle.model_from_description(desc)
    arm, x = ocba_oracle.gao_modellist_ref(model, ctx)
    return int(arm)


def pick_cpu_variant(desc, ctx):
    """Two CPU restatements: "modellist" walks the K sub-models like the reference's ModelListGP does,
    "batched" stacks them (bmm / batched Cholesky); the faster one (after a warm-up call) is the baseline."""
    best = None
    for variant in ("batched", "modellist"):
        cpu_decision_step(desc, ctx, variant)
        dt = float("inf")
        for _ in range(2):
            t0 = time.perf_counter()
            cpu_decision_step(desc, ctx, variant)
            dt = min(dt, time.perf_counter() - t0)
        if best is None or dt < best[1]:
            best = (variant, dt)
    return best


class CpuReference:
    """The reference's CPU path on a bounded sample of a workload: decision on the first `sub_c` contexts
    (all arms), PCS with the marginal sampler on `pcs_c` contexts x `pcs_s` samples, both scaled linearly to
    the full workload.  `step()` returns the scaled seconds of one full step."""

    def __init__(self, name, budget_per_step_s=2.0):
        from contextual_rs_b200.synthetic import make_config
        from oracle import gp_oracle
        torch.set_num_threads(os.cpu_count() or 1)
        self.name = name
        self.cores = torch.get_num_threads()
        self.desc, self.ctx = make_config(SHAPES.get(name, name))
        self.K, self.n_c = len(self.desc["train_X"]), self.ctx.shape[0]
        self.with_pcs = name in ("headline", "loop")
        self.sub_c = self.n_c if name == "c2" else max(256, self.n_c // 128)
        self.variant, est = pick_cpu_variant(self.desc, self.ctx[: self.sub_c])
        while est > budget_per_step_s and self.sub_c > 128:
            self.sub_c //= 2
            est /= 2
        self.ctx_s = self.ctx[: self.sub_c]
        if self.with_pcs:
            self.pcs_c, self.pcs_s = 32, 512
            self.mu, self.vr = gp_oracle.posterior_grid_batched(self.desc, self.ctx[: self.pcs_c])
            self.gen = torch.Generator().manual_seed(0)

    def step(self):
        from oracle import pcs_oracle
        t0 = time.perf_counter()
        cpu_decision_step(self.desc, self.ctx_s, self.variant)
        t1 = time.perf_counter()
        dec = (t1 - t0) * (self.n_c / self.sub_c)
        pcs = 0.0
        if self.with_pcs:
            pcs_oracle.pcs_marginal_ref(self.mu, self.vr, self.pcs_s, generator=self.gen)
            pcs = (time.perf_counter() - t1) * (self.n_c / self.pcs_c) * (PCS_SAMPLES / self.pcs_s)
        return dec, pcs

    def sample(self, steps):
        s = (f"{steps} steps; decision: oracle '{self.variant}' variant on {self.sub_c}/{self.n_c} contexts x {self.K} arms, "
             f"time scaled x{self.n_c / self.sub_c:.0f}")
        if self.with_pcs:
            s += (f"; PCS: marginal sampler (oracle.pcs_marginal_ref, fp64) on {self.pcs_c} contexts x {self.K} arms x "
                  f"{self.pcs_s} samples, time scaled x{(self.n_c / self.pcs_c) * (PCS_SAMPLES / self.pcs_s):.0f} — the "
                  f"reference's joint sampler is O(n_c^3) and cannot run at this size")
        return s + f"; PyTorch CPU, {self.cores} threads"

    def run(self, warmup, steps):
        for _ in range(warmup):
            self.step()
        dec = pcs = 0.0
        for _ in range(steps):
            d, p = self.step()
            dec += d
            pcs += p
        return dec / steps, pcs / steps


def run_reference(args, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    ref = CpuReference(name)
    W, steps = max(args.warmup, 0), max(args.steps, 1)
    dec, pcs = ref.run(W, steps)
    dt = dec + pcs
    metric, unit = METRIC.get(name, ("gp_c_ocba_decisions_per_sec", "decisions/s"))
    val = 1.0 / dt
    sharded = world > 1 and not (name == "c2" and not args.shard)
    emit(({
        "impl": "reference", "metric": metric, "value": val, "unit": unit,
        "n_gpus": args.gpus, "steps": steps, "warmup": W, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong" if (sharded or world == 1) and name != "c2" else "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(name, world, args),
        "breakdown_ms": {"decision": dec * 1e3, "pcs": pcs * 1e3},
        "cpu_baseline": {"value": val, "unit": unit, "cores": ref.cores, "kind": "port", "sample": ref.sample(steps)},
        "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------- GPU
class Env:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: contextual_rs_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush_buf = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)

    def flush(self):
        if self.flush_buf is not None:
            self.flush_buf.fill_(1)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """Per-step CUDA-event pairs on the launching stream, L2 flushed in between; list of ms."""
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            self.flush()
            starts[i].record()
            fn()
            ends[i].record()
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in zip(starts, ends)]

    def max_over_ranks(self, values):
        t = torch.tensor(values, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(self, values):
        t = torch.tensor(values, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t)
        return t.tolist()

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def expected_headline_decision():
    """(next_arm, next_context) of the fp64 oracle on config 4, committed by
    tests/golden/make_quoted_configs_golden.py (tests/golden/quoted_configs_expected.json)."""
    with open(os.path.join(ROOT, "tests", "golden", "quoted_configs_expected.json")) as f:
        exp = json.load(f)["c4"]
    return exp["default"]["next_arm"], exp["default"]["next_context"]


def run_headline(args, env):
    """The default: config 4 decision + 2^16-sample PCS, contexts sharded over the ranks."""
    import contextual_rs_b200 as crs
    from contextual_rs_b200 import _lib
    from contextual_rs_b200.contextual_rs_strategies import _scan, decide
    from contextual_rs_b200.generalized_pcs import pcs_counts
    from contextual_rs_b200.peaks import measure_fma_peak, measure_pcs_cdf_rates, pcs_cdf_roofline
    from contextual_rs_b200.sharded import ContextShardedDecider, REC_CONTEXT, REC_NEXT_ARM
    from contextual_rs_b200.synthetic import make_config
    dev, world, rank = env.dev, env.world, env.rank
    name = "headline"
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    w = 8 if dtype == torch.float64 else 4
    desc, ctx = make_config("c4")
    K, n_c, d = len(desc["train_X"]), ctx.shape[0], ctx.shape[1]
    n_obs = desc["train_X"][0].shape[0]
    S = PCS_SAMPLES
    model = crs.B200ModelListGP(desc, device=dev, dtype=dtype)
    model.check_status()
    lib = _lib.get()
    peaks, peak_src = load_peaks()
    decider = ContextShardedDecider(model, ctx, exchange=args.exchange)
    begin, n_loc, ws = decider.begin, decider.n_loc, decider.ws
    ctx_dev = decider.local_ctx
    stream = _lib.stream_ptr(dev)
    st = C.byref(model.state)
    stats = ws["stats"]
    W, K_steps = max(args.warmup, 3), max(args.steps, 1)
    it = [0]

    def pcs_launch(num_samples=S):
        if n_loc:
            decider.pcs_counts(num_samples, seed=0, offset=it[0])
        else:
            stats.zero_()

    def device_step():
        """prepare -> grid -> scan/select (+ peer fold) -> PCS build + sample (+ peer sum): no host round-trip."""
        it[0] += 1
        if decider.exchange == "mailbox":
            decider._launch_local(_lib.CRS_P_COUNT, 2.0, (0, K), True, decider.mailbox.next_exchange(begin))
            pcs_launch()
            px = decider.mailbox.next_exchange(begin)
            _lib.check(lib.crs_peer_sum_u64(stats.data_ptr() + 32, 1, decider._sum_dev.data_ptr(), None, C.byref(px), stream))
        else:
            _lib.check(lib.crs_gp_prepare(st, 0, K, stream))
            if n_loc:
                _lib.check(lib.crs_posterior_grid(st, ctx_dev.data_ptr(), n_loc, 0, K, ws["mean"].data_ptr(),
                                                  ws["var"].data_ptr(), K, 0, stream))
                from contextual_rs_b200.contextual_rs_strategies import _scan_select
                _scan_select(lib, model, ws, n_loc, 0, 2.0, stream, ctx_ptr=ctx_dev.data_ptr())
            pcs_launch()
            if world > 1:
                env.dist.all_gather_into_tensor(decider._gather, ws["decision"] if n_loc else decider._empty)
                env.dist.all_reduce(stats[4:5])

    ctx_host = ctx.to(dtype).pin_memory()
    arm_set = torch.arange(K, dtype=dtype).view(-1, 1)
    func_I = lambda X: (X > 0).to(dtype)  # noqa: E731  (experiments/wsc_experiments/main.py:453)
    rho = lambda P: P.mean(dim=-2)  # noqa: E731  (main.py:248-251 "mean")
    last = {}

    ctx_pageable = ctx.to(dtype).clone()  # ordinary host memory, as the reference's caller holds the context set

    def e2e_step(host=None):
        it[0] += 1
        host = ctx_host if host is None else host
        if world == 1:  # the reference-facing drop-in calls, context set on the host
            dec = decide(model, host, True, 0, 2.0, prepare_arms=(0, K))
            pcs = crs.estimate_current_generalized_pcs(model, arm_set, host, S, None, func_I, rho, seed=0,
                                                       offset=it[0], reuse_grid=True)
            last.update(next_arm=dec.next_arm, next_context=dec.context, pcs=float(pcs))
        else:  # every rank uploads its slice of the pinned host context set; results come back to the host
            decider.local_ctx.copy_(ctx_host[begin:begin + n_loc], non_blocking=True)
            rec = decider.decide(prepare_arms=(0, K))
            pcs = decider.pcs_mean(S, seed=0, offset=it[0])
            last.update(next_arm=int(rec[REC_NEXT_ARM]), next_context=int(rec[REC_CONTEXT]), pcs=float(pcs))

    # nvidia-smi needs ~0.1 s before its first sample: it is started ahead of the warm-up steps (same kernels, same
    # load) so that it is already sampling, every 20 ms, when the timed region begins
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        env.flush()
        device_step()
    for _ in range(W):
        env.flush()
        e2e_step()
    env.barrier()
    dev_ms = env.timed(device_step, K_steps)
    env.barrier()
    e2e_ms = env.timed(e2e_step, K_steps)
    env.barrier()
    page_ms = None
    if world == 1:
        e2e_step(ctx_pageable)
        page_ms = env.timed(lambda: e2e_step(ctx_pageable), min(K_steps, 10))
    clocks = sampler.stop() if rank == 0 else None
    # ---- per-kernel durations, same region style (flushed, own event pair, launching stream) ---------
    ksteps = min(K_steps, 20)

    def grid_only():
        if n_loc:
            lib.crs_posterior_grid(st, ctx_dev.data_ptr(), n_loc, 0, K, ws["mean"].data_ptr(), ws["var"].data_ptr(), K, 0, stream)

    def scan_only():
        if n_loc:
            _scan(lib, model, ws, n_loc, stream)

    grid_ms = env.timed(grid_only, ksteps)
    scan_ms = env.timed(scan_only, ksteps)
    prep_ms = env.timed(lambda: lib.crs_gp_prepare(st, 0, K, stream), ksteps)
    build_ms = env.timed(lambda: pcs_launch(0), ksteps)
    pcs_ms = env.timed(pcs_launch, ksteps)
    work = stats[:2].to(torch.float64).tolist()  # {Philox calls, table terms} of this rank's last estimate
    env.barrier()
    avg = lambda v: sum(v) / len(v)  # noqa: E731
    t_dev, t_e2e, grid_avg, scan_avg, prep_avg, build_avg, pcs_avg = env.max_over_ranks(
        [sum(dev_ms) / 1e3, sum(e2e_ms) / 1e3, avg(grid_ms), avg(scan_ms), avg(prep_ms), avg(build_ms), avg(pcs_ms)])
    calls_all, terms_all = env.sum_over_ranks(work)
    sample_avg = max(pcs_avg - build_avg, 1e-6)
    out = None
    if rank == 0:
        exp_arm, exp_ctx = expected_headline_decision()
        got = (last["next_arm"], last["next_context"])
        if dtype == torch.float64 and got != (exp_arm, exp_ctx):
            raise SystemExit(f"headline decision {got} != fp64 oracle's {(exp_arm, exp_ctx)} (tests/golden/quoted_configs_expected.json)")
        value = K_steps / t_dev
        flops = algorithmic_flops(K, n_c, n_obs, d)
        fma_peak = measure_fma_peak(0 if dtype == torch.float64 else 1, dev)
        cdf_rates = measure_pcs_cdf_rates(dev)
        roof_grid = {"kernel": "posterior_grid_kernel", "bound": "fp64_fma" if dtype == torch.float64 else "fp32_fma",
                     "achieved": flops / world / (grid_avg * 1e-3) / 1e12, "peak": fma_peak, "unit": "TFLOP/s",
                     "frac": flops / world / (grid_avg * 1e-3) / 1e12 / fma_peak,
                     "traffic": ncu_traffic(name, "posterior_grid_kernel", args.dtype) if world == 1 else None,
                     "avg_launch_ms": grid_avg, "algorithmic_flops_per_launch": flops / world,
                     "peak_source": "measured in this run by crs_peak_probe (pure FMA micro-kernel, 16 chains/thread, CUDA "
                                    "events); MEASURED_PEAKS.json has no fp64/fp32-FMA figure (nominal 37.2 / 74.5)"}
        if dtype == torch.float64:
            # what the kernel actually has to issue on the FP64 pipe (SASS count; ncu profiles/r02c):
            # per observation and (context, arm) pair 2d distance + 7 sqrt + 10 exp + 2 polynomial + 3 product /
            # accumulation DFMA-class instructions, plus the DMMA tiles of the triangular contraction at 8
            # DFMA-equivalents each (config 4: 44 tiles per warp and arm) -- software exp / sqrt count as 17 here,
            # as 2 in F1
            n4 = (n_obs + 3) // 4 * 4
            dmma = None
            if n4 <= 32:  # one 32-column panel: k-steps 2 sg, 2 sg + 1 reach the row tiles mt >= sg (posterior_grid.cu)
                mt_tiles, ksteps = (8 if n_obs <= 8 else 16 if n_obs <= 16 else 24 if n_obs <= 24 else 32) // 8, n4 // 4
                dmma = 4 * sum((mt_tiles - sg) * sum(1 for ks in (2 * sg, 2 * sg + 1) if ks < ksteps) for sg in range(mt_tiles))
            if dmma is not None:
                exec_flops = (n4 * (2 * d + 22) + dmma * 8) * 2.0 * K * n_c / world
                roof_grid["instruction_floor"] = {"fp64_flops_issued_per_launch": exec_flops,
                                                  "frac": exec_flops / (grid_avg * 1e-3) / 1e12 / fma_peak,
                                                  "note": "issued FP64 work (DFMA-class instructions + DMMA at 8 DFMA-equivalents) over the "
                                                          "measured FMA peak: the FP64 pipe's busy fraction (ncu: 75 %)"}
        scan_bytes = (2 * K * n_c * w + n_c * (w + 12)) / world
        roof_scan = {"kernel": "ocba_scan_kernel", "bound": "hbm", "achieved": scan_bytes / (scan_avg * 1e-3) / 1e9,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": scan_bytes / (scan_avg * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "traffic": ncu_traffic(name, "ocba_scan_kernel", args.dtype) if world == 1 else None,
                     "avg_launch_ms": scan_avg, "peak_source": peak_src, "algorithmic_bytes_per_launch": scan_bytes}
        roof_sample = pcs_cdf_roofline("pcs_cdf_sample_kernel", calls_all, sample_avg * 1e-3, world, cdf_rates)
        roof_build = pcs_cdf_roofline("pcs_cdf_build_kernel", terms_all, build_avg * 1e-3, world, cdf_rates)
        if world == 1:
            roof_sample["traffic"] = ncu_traffic(name, "pcs_cdf_sample_kernel", args.dtype)
            roof_build["traffic"] = ncu_traffic(name, "pcs_cdf_build_kernel", args.dtype)
        build_bytes = 2 * K * n_c * w / world
        roof_build["hbm"] = {"algorithmic_bytes_per_launch": build_bytes, "achieved_gbs": build_bytes / (build_avg * 1e-3) / 1e9,
                             "frac_of_hbm_peak": build_bytes / (build_avg * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        roofs = {"posterior_grid": (grid_avg, roof_grid), "pcs_cdf_sample": (sample_avg, roof_sample),
                 "pcs_cdf_build": (build_avg, roof_build), "ocba_scan": (scan_avg, roof_scan)}
        dominant = max(roofs, key=lambda k: roofs[k][0])
        cfg = workload_config(name, world, args)
        out = {
            "metric": "gp_c_ocba_decisions_per_sec", "value": value, "unit": "decisions/s",
            "n_gpus": world, "steps": K_steps, "warmup": W, "ms_per_step": t_dev / K_steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": cfg,
            "step": "prepare(all arms) + posterior grid + fused zeta scan / OCBA select (+ peer fold) + PCS table build + PCS "
                    "sampling (+ peer sum): 5 launches per rank at N = 1, 6 at N > 1; exchange = " + decider.exchange,
            "pcs_mc_samples_per_sec": {"value": float(S) * n_c * K / (pcs_avg * 1e-3), "unit": "equivalent N(0,1) draws/s",
                                       "ms_per_estimate": pcs_avg,
                                       "note": "nominal S x n_c x K draws of the reference's estimator per PCS time; the cdf stream "
                                               "executes one normal + one uniform per (sample, context): %.3g Philox calls" % calls_all},
            "e2e": {"value": K_steps / t_e2e, "unit": "decisions/s", "ms_per_step": t_e2e / K_steps * 1e3,
                    "h2d_bytes_per_step": int(n_c * d * w), "d2h_bytes_per_step": (64 + 80) * world if world > 1 else 64 + w,
                    "api": ("gao_modellist's decide(model, host context_set) + estimate_current_generalized_pcs(..., reuse_grid=True)"
                            if world == 1 else "ContextShardedDecider.decide + .pcs_mean, every rank uploading its context slice"),
                    "pageable_contexts": ({"ms_per_step": sum(page_ms) / len(page_ms), "value": len(page_ms) / (sum(page_ms) * 1e-3),
                                           "note": "same calls with the context set in ordinary (pageable) host memory"}
                                          if page_ms else None)},
            "gpu_launches": (5 + (1 if world > 1 else 0)) * 2 * K_steps * world + 6 * ksteps * world,
            "roofline": dict(roofs[dominant][1], dominant_of="kernel_ms"),
            "kernel_ms": {"gp_prepare": prep_avg, "posterior_grid": grid_avg, "ocba_scan": scan_avg,
                          "pcs_cdf_build": build_avg, "pcs_cdf_sample": sample_avg},
            "decision": {"next_arm": got[0], "next_context": got[1], "pcs_mean": last["pcs"],
                         "matches_fp64_oracle": got == (exp_arm, exp_ctx),
                         "parity": "partial: decision index-exact vs the CPU oracle (a line-by-line restatement of the reference); the GP "
                                   "arithmetic of the absent GPyTorch/BoTorch is restated, not pinned bit for bit (DESIGN.md)"},
            "clocks": clocks,
        }
        for k, (_, r) in roofs.items():
            if k != dominant:
                out["roofline_" + k] = r
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        ref = CpuReference(name)
        dec_s, pcs_s = ref.run(1, 6)
        out["cpu_baseline"] = {"value": 1.0 / (dec_s + pcs_s), "unit": "decisions/s", "cores": ref.cores, "kind": "port",
                               "sample": ref.sample(6), "breakdown_ms": {"decision": dec_s * 1e3, "pcs": pcs_s * 1e3}}
    if world == 1 and not args.no_sub:
        env.flush_buf = None
        torch.cuda.empty_cache()
        env.flush_buf = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        sub_c2 = run_decision(args, env, "c2", steps=300, sub=True)
        sub_pcs = run_pcs(env, cpu=not args.no_cpu_baseline)
        sub_cont = run_continuous(env, model, ctx)
        if rank == 0:
            out["c2"] = sub_c2
            out["pcs"] = sub_pcs
            out["continuous"] = sub_cont
            out["gpu_launches"] += sub_c2["gpu_launches"] + sub_pcs["gpu_launches"] + sub_cont["gpu_launches"]
    if rank == 0:
        emit(out)
    if decider.mailbox is not None:
        decider.mailbox.close()
    env.finish()


def run_headline_arms(args, env):
    """The headline step under the partition BASELINE.json's north_star describes: ARMS sharded over the
    ranks (each owns K/N independent GPs for all 65,536 contexts), all-gather of the per-context bests,
    all-gather of the ranks' 64-byte minimisers, all-reduce of the psi partials; the PCS needs every arm
    of a context, so the grid is re-sharded by context with one all-to-all first.  Device timed like the
    default; e2e adds the host->device copy of the whole context set on every rank and the read-back."""
    import contextual_rs_b200 as crs  # noqa: F401
    from contextual_rs_b200.sharded import ArmShardedDecider
    from contextual_rs_b200.synthetic import make_config
    dev, world, rank = env.dev, env.world, env.rank
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    w = 8 if dtype == torch.float64 else 4
    desc, ctx = make_config("c4")
    K, n_c, d = len(desc["train_X"]), ctx.shape[0], ctx.shape[1]
    S = PCS_SAMPLES
    dec = ArmShardedDecider(desc, ctx, dev, dtype=dtype)
    W, K_steps = max(args.warmup, 3), max(args.steps, 1)
    it = [0]
    last = {}
    ctx_host = ctx.to(dtype).pin_memory()

    def device_step():
        it[0] += 1
        dec.launch(prepare=True)
        dec.pcs_total_launch(S, seed=0, offset=it[0])

    def e2e_step():
        it[0] += 1
        dec.ws["ctx"].copy_(ctx_host, non_blocking=True)
        d_ = dec.decide_struct(prepare=True)
        pcs = dec.pcs_mean(S, seed=0, offset=it[0])
        last.update(next_arm=int(d_.next_arm), next_context=int(d_.context), pcs=float(pcs))

    def decision_only():
        dec.launch(prepare=True)

    def reshard_only():
        dec.reshard_by_context()

    for _ in range(W):
        env.flush()
        device_step()
        e2e_step()
    env.barrier()
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    env.barrier()
    dev_ms = env.timed(device_step, K_steps)
    env.barrier()
    e2e_ms = env.timed(e2e_step, K_steps)
    env.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ksteps = min(K_steps, 20)
    for _ in range(2):
        decision_only()
    env.barrier()
    dec_ms = env.timed(decision_only, ksteps)
    env.barrier()
    for _ in range(2):
        reshard_only()
    env.barrier()
    rs_ms = env.timed(reshard_only, ksteps)
    env.barrier()
    avg = lambda v: sum(v) / len(v)  # noqa: E731
    t_dev, t_e2e, dec_avg, rs_avg = env.max_over_ranks([sum(dev_ms) / 1e3, sum(e2e_ms) / 1e3, avg(dec_ms), avg(rs_ms)])
    if rank == 0:
        exp_arm, exp_ctx = expected_headline_decision()
        got = (last["next_arm"], last["next_context"])
        if dtype == torch.float64 and got != (exp_arm, exp_ctx):
            raise SystemExit(f"headline decision {got} != fp64 oracle's {(exp_arm, exp_ctx)}")
        sent_decision = (world - 1) * (3 * 8 * n_c + 64 + 24) // 1
        sent_reshard = int(2 * w * n_c * (K / world) * (world - 1) / world)
        emit({
            "metric": "gp_c_ocba_decisions_per_sec", "value": K_steps / t_dev, "unit": "decisions/s",
            "n_gpus": world, "steps": K_steps, "warmup": W, "ms_per_step": t_dev / K_steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config("headline", world, args),
            "step": "arms sharded: prepare + local grid + scan + pack | all-gather bests | combine + scan vs global best | "
                    "all-gather records | fold + select | all-reduce psi | finish | all-to-all re-shard | PCS build + sample | all-reduce",
            "breakdown_ms": {"decision": dec_avg, "context_reshard_all_to_all": rs_avg,
                             "pcs_and_rest": t_dev / K_steps * 1e3 - dec_avg - rs_avg},
            "bytes_sent_per_rank_per_step": {"decision_collectives": sent_decision, "context_reshard": sent_reshard},
            "e2e": {"value": K_steps / t_e2e, "unit": "decisions/s", "ms_per_step": t_e2e / K_steps * 1e3,
                    "h2d_bytes_per_step": int(n_c * d * w) * world, "d2h_bytes_per_step": (64 + 8) * world,
                    "api": "ArmShardedDecider.decide_struct + .pcs_mean, every rank uploading the whole context set"},
            "gpu_launches": 14 * 2 * K_steps * world,
            "decision": {"next_arm": got[0], "next_context": got[1], "pcs_mean": last["pcs"], "matches_fp64_oracle": True},
            "roofline": None, "clocks": clocks})
    env.finish()


def run_continuous(env, model, ctx):
    """BASELINE config 4 as the reference's continuous-context driver runs it
    (experiments/continuous_experiments/main.py:252-268, 304-306) on the headline's model (1,000 arms, 8-D):
    the MinZeta sweep over the 65,536-point Sobol set, the 20-restart / 1,024-raw-sample optimize_acqf
    refinement of MinZeta, find_next_arm_given_context at the refined context, and a 10-candidate
    look-ahead PCS (generalized_pcs.py:203-315) on the first 4,096 contexts.  Wall clock of the host
    API calls (they end in a device->host read), median of 3."""
    import contextual_rs_b200 as crs
    dev = env.dev
    K = model.num_arms
    acq = crs.MinZeta(model)
    X = ctx.to(device=dev, dtype=model.dtype).unsqueeze(-2)
    bounds = torch.stack([torch.zeros(ctx.shape[1], dtype=model.dtype), torch.ones(ctx.shape[1], dtype=model.dtype)]).to(dev)

    def med(fn, n=3):
        ts = []
        res = None
        for _ in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return sorted(ts)[len(ts) // 2] * 1e3, res

    acq(X)
    sweep_ms, vals = med(lambda: acq(X))
    crs.optimize_acqf(acq, bounds, q=1, num_restarts=20, raw_samples=1024, seed=1)
    opt_ms, (cand, val, det) = med(lambda: crs.optimize_acqf(acq, bounds, q=1, num_restarts=20, raw_samples=1024, seed=1,
                                                            return_details=True))
    arm_ms, arm = med(lambda: crs.find_next_arm_given_context(cand, model, kernel_scale=2.0, randomize_ties=False))
    sub_ctx = ctx[:4096].to(device=dev, dtype=model.dtype)
    cand_la = torch.cat([torch.arange(10, dtype=model.dtype, device=dev).view(-1, 1, 1) * (K // 10),
                         sub_ctx[:10].unsqueeze(-2)], dim=-1)
    f_I = lambda Xd: (Xd > 0).to(model.dtype)  # noqa: E731
    rho = lambda P: P.mean(dim=-2)  # noqa: E731
    crs.estimate_lookahead_generalized_pcs_modellist(cand_la, model, None, sub_ctx, 4096, None, f_I, rho, seed=2)
    la_ms, la = med(lambda: crs.estimate_lookahead_generalized_pcs_modellist(cand_la, model, None, sub_ctx, 4096, None, f_I, rho, seed=2))
    return {"workload": "c4 continuous: MinZeta sweep over 65,536 Sobol contexts, optimize_acqf(MinZeta, q=1, num_restarts=20, "
                        "raw_samples=1024), find_next_arm_given_context; look-ahead PCS of 10 candidates on 4,096 contexts x 4,096 samples",
            "min_zeta_sweep_ms": sweep_ms, "sweep_best": float(vals.max()),
            "optimize_acqf_ms": opt_ms, "optimize_acqf_value": float(val), "raw_best": det["raw_best"],
            "objective_launches": int(det["evaluations"]), "lbfgs_iterations_max": int(det["iterations"].max()),
            "find_next_arm_ms": arm_ms, "next_arm": int(arm),
            "lookahead_pcs_ms": la_ms, "lookahead_pcs_ms_per_candidate": la_ms / 10, "lookahead_values": [round(float(v), 5) for v in la],
            "gpu_launches": 4 * 2 + 8 * (2 + int(det["evaluations"])) + 8 + 4 * 10 * 5}


# ---------------------------------------------------------------------------------- decision only
def run_decision(args, env, name, steps=None, sub=False):
    """A decision-only workload (configs 2-5): prepare + grid + fused scan/select.  `sub=True` returns a
    compact object (config 2 inside the headline line) instead of emitting."""
    import contextual_rs_b200 as crs
    from contextual_rs_b200 import _lib
    from contextual_rs_b200.contextual_rs_strategies import _scan, _scan_select, decide
    from contextual_rs_b200.peaks import measure_fma_peak
    from contextual_rs_b200.sharded import ContextShardedDecider, REC_CONTEXT, REC_NEXT_ARM
    from contextual_rs_b200.synthetic import make_config
    dev, world, rank = env.dev, env.world, env.rank
    W, K_steps = max(args.warmup, 3), steps or max(args.steps, 1)
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    w = 8 if dtype == torch.float64 else 4
    # N > 1: c2 runs one independent replication per rank (weak scaling) unless --shard
    sharded = world > 1 and (name != "c2" or args.shard)
    desc, ctx = make_config(name) if (sharded or world == 1) else make_config(name, seed=rank)
    K, n_c, d = len(desc["train_X"]), ctx.shape[0], ctx.shape[1]
    n_obs = desc["train_X"][0].shape[0]
    model = crs.B200ModelListGP(desc, device=dev, dtype=dtype)
    model.check_status()
    lib = _lib.get()
    peaks, peak_src = load_peaks()
    if sharded:
        decider = ContextShardedDecider(model, ctx, exchange=args.exchange)
        begin, n_loc, ctx_dev, ws = decider.begin, decider.n_loc, decider.local_ctx, decider.ws
    else:
        decider, begin, n_loc = None, 0, n_c
        ctx_dev = ctx.to(device=dev, dtype=dtype).contiguous()
        ws = model.workspace(n_c)
    stream = _lib.stream_ptr(dev)
    st = C.byref(model.state)

    def device_step():
        if sharded and decider.exchange == "mailbox":
            decider._launch_local(_lib.CRS_P_COUNT, 2.0, (0, K), True, decider.mailbox.next_exchange(begin))
            return
        _lib.check(lib.crs_gp_prepare(st, 0, K, stream))
        if n_loc:
            _lib.check(lib.crs_posterior_grid(st, ctx_dev.data_ptr(), n_loc, 0, K, ws["mean"].data_ptr(),
                                              ws["var"].data_ptr(), K, 0, stream))
            _scan_select(lib, model, ws, n_loc, 0, 2.0, stream, ctx_ptr=ctx_dev.data_ptr())
        if sharded:
            env.dist.all_gather_into_tensor(decider._gather, ws["decision"] if n_loc else decider._empty)

    ctx_host = ctx.to(dtype).pin_memory()
    ctx_pageable = ctx.to(dtype).clone()

    def e2e_step(host=ctx_host):
        if not sharded:
            dec = decide(model, host, True, 0, 2.0, prepare_arms=(0, K))
            return dec.next_arm, dec.context
        decider.local_ctx.copy_(host[begin:begin + n_loc], non_blocking=True)
        rec = decider.decide(prepare_arms=(0, K))
        return int(rec[REC_NEXT_ARM]), int(rec[REC_CONTEXT])

    for _ in range(W):
        env.flush()
        device_step()
    for _ in range(W):
        env.flush()
        e2e_step()
        e2e_step(ctx_pageable)
    env.barrier()
    sampler = ClockSampler(env.local_rank)
    if rank == 0 and not sub:
        sampler.start()
    env.barrier()
    dev_ms = env.timed(device_step, K_steps)
    env.barrier()
    ksteps = min(K_steps, 500)

    def grid_only():
        if n_loc:
            lib.crs_posterior_grid(st, ctx_dev.data_ptr(), n_loc, 0, K, ws["mean"].data_ptr(), ws["var"].data_ptr(), K, 0, stream)

    def scan_only():
        if n_loc:
            _scan(lib, model, ws, n_loc, stream)

    grid_ms = env.timed(grid_only, ksteps)
    scan_ms = env.timed(scan_only, ksteps)
    prep_ms = env.timed(lambda: lib.crs_gp_prepare(st, 0, K, stream), ksteps)
    env.barrier()
    e2e_ms = env.timed(e2e_step, K_steps)
    page_ms = env.timed(lambda: e2e_step(ctx_pageable), min(K_steps, 200)) if not sharded else [0.0]
    env.barrier()
    clocks = sampler.stop() if (rank == 0 and not sub) else None
    avg = lambda v: sum(v) / len(v)  # noqa: E731
    t_dev, t_e2e, grid_avg, scan_avg, prep_avg = env.max_over_ranks(
        [sum(dev_ms) / 1e3, sum(e2e_ms) / 1e3, avg(grid_ms), avg(scan_ms), avg(prep_ms)])
    out = None
    if rank == 0:
        mult = 1 if (sharded or world == 1) else world  # replicas: every rank made K_steps decisions
        value = mult * K_steps / t_dev
        flops = algorithmic_flops(K, n_loc, n_obs, d)
        fma_peak = measure_fma_peak(0 if dtype == torch.float64 else 1, dev)
        achieved = flops / (grid_avg * 1e-3) / 1e12
        scan_bytes = 2 * K * n_loc * w + n_loc * (w + 12)
        # gp_prepare is a latency chain, not a throughput kernel: one CTA per arm walks the n x n Cholesky
        # column by column; its model is n dependent column steps (sqrt + divide + rank-1 update + two
        # substitution steps), reported as ns per dependent column step
        out = {
            "metric": "gp_c_ocba_decisions_per_sec", "value": value, "unit": "decisions/s",
            "n_gpus": world, "steps": K_steps, "warmup": W, "ms_per_step": t_dev / K_steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": workload_config(name, world, args),
            "step": "3 kernels: prepare(all arms) + posterior grid + fused zeta scan / OCBA select",
            "e2e": {"value": mult * K_steps / t_e2e, "unit": "decisions/s", "ms_per_step": t_e2e / K_steps * 1e3,
                    "h2d_bytes_per_step": int(n_c * d * w) * mult, "d2h_bytes_per_step": 64 * world,
                    "pageable_contexts": ({"value": mult / (avg(page_ms) * 1e-3), "ms_per_step": avg(page_ms),
                                           "note": "same call with the context set in ordinary (pageable) host memory, as the "
                                                   "reference's caller holds it: plain cudaMemcpyAsync staging instead of the "
                                                   "zero-copy read of pinned memory"} if not sharded else None)},
            "gpu_launches": (3 * K_steps + 3 * K_steps + 3 * ksteps) * mult,
            "roofline": {"kernel": "posterior_grid_kernel", "bound": "fp64_fma" if dtype == torch.float64 else "fp32_fma",
                         "achieved": achieved, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved / fma_peak,
                         "traffic": ncu_traffic(name, "posterior_grid_kernel", args.dtype) if world == 1 else None,
                         "avg_launch_ms": grid_avg,
                         "peak_source": "measured in this run by crs_peak_probe (pure FMA micro-kernel, 16 chains/thread, "
                                        "CUDA events); MEASURED_PEAKS.json has no fp64/fp32-FMA figure (nominal 37.2 / 74.5)",
                         "algorithmic_flops_per_launch": flops},
            "roofline_scan": {"kernel": "ocba_scan_kernel", "bound": "hbm", "achieved": scan_bytes / (scan_avg * 1e-3) / 1e9,
                              "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": scan_bytes / (scan_avg * 1e-3) / 1e9 / peaks["hbm_gbs"],
                              "traffic": ncu_traffic(name, "ocba_scan_kernel", args.dtype) if world == 1 else None,
                              "avg_launch_ms": scan_avg, "peak_source": peak_src,
                              "algorithmic_bytes_per_launch": scan_bytes},
            "latency_prepare": {"kernel": "gp_prepare_kernel", "bound": "latency", "avg_launch_ms": prep_avg,
                                "dependent_column_steps": n_obs, "ns_per_column_step": prep_avg * 1e6 / n_obs,
                                "note": "one CTA per arm; the n x n Cholesky + two substitutions are a chain of n dependent "
                                        "column steps (fp64 sqrt / divide ~ 100+ cycles each); K <= 148 x resident CTAs arms run "
                                        "in one wave, so the launch time is the chain, not throughput"},
            "kernel_ms": {"gp_prepare": prep_avg, "posterior_grid": grid_avg, "ocba_scan": scan_avg},
            "clocks": clocks,
        }
        if sub:
            for k in ("higher_is_better", "vs_baseline", "data", "clocks", "n_gpus"):
                out.pop(k, None)
    if sub:
        return out
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        ref = CpuReference(name)
        dec_s, _ = ref.run(1, 20 if name == "c2" else 6)
        out["cpu_baseline"] = {"value": 1.0 / dec_s, "unit": "decisions/s", "cores": ref.cores, "kind": "port",
                               "sample": ref.sample(20 if name == "c2" else 6)}
    if rank == 0:
        emit(out)
    if decider is not None and decider.mailbox is not None:
        decider.mailbox.close()
    env.finish()


# ------------------------------------------------------------------------------------------- PCS
PCS_CONFIG = "c3: generalized-PCS MC estimate, 2^16 samples, 256 arms x 8,192 contexts, fp32"


def run_pcs(env, reps=5, num_samples=PCS_SAMPLES, cpu=True):
    """PCS MC samples/s on config 3 (fp32, one GPU): the default cdf stream (crs_pcs_counts_cdf), with the
    max-tree stream and the flat per-arm stream on the same tables beside it.  'value' counts the S*n_c*K
    nominal draws of the reference's estimator ("equivalent draws"); the executed work is reported beside it."""
    import contextual_rs_b200 as crs
    from contextual_rs_b200.generalized_pcs import pcs_counts
    from contextual_rs_b200.peaks import measure_pcs_cdf_rates, pcs_cdf_roofline
    from contextual_rs_b200.synthetic import make_config
    dev = env.dev
    desc, ctx = make_config("c3")
    K, n_c = len(desc["train_X"]), ctx.shape[0]
    model = crs.B200ModelListGP(desc, device=dev, dtype=torch.float32)
    post = model.posterior(ctx.float().to(dev))
    mean, var = post.mean.contiguous(), post.variance.contiguous()
    stats = torch.zeros(8, dtype=torch.int64, device=dev)
    wsb = torch.empty(int(crs._lib.get().crs_pcs_cdf_workspace_bytes(n_c)), dtype=torch.uint8, device=dev)
    res = {}
    launches = 0
    for sampler in ("cdf", "cdf_build", "tree", "flat"):
        def step(i, sampler=sampler):
            if sampler == "cdf_build":
                return pcs_counts(model, mean, var, 0, seed=0, offset=i, stats=stats, sampler="cdf", workspace=wsb)
            return pcs_counts(model, mean, var, num_samples, seed=0, offset=i, stats=stats, sampler=sampler,
                              workspace=wsb if sampler == "cdf" else None)
        step(0)
        torch.cuda.synchronize()
        n = reps if sampler.startswith("cdf") else 2
        ms = env.timed(lambda: step(1), n)
        c = step(1)
        launches += (n + 2) * (1 if sampler in ("tree", "flat", "cdf_build") else 2)
        res[sampler] = {"ms": sum(ms) / len(ms), "stats": stats.cpu().tolist(),
                        "pcs_mean": float(c.sum(dtype=torch.int64)) / (num_samples * n_c) if sampler != "cdf_build" else None}
    nominal = float(num_samples) * n_c * K
    t_cdf, t_build = res["cdf"]["ms"] * 1e-3, res["cdf_build"]["ms"] * 1e-3
    rates = measure_pcs_cdf_rates(dev)
    out = {"metric": "pcs_mc_samples_per_sec", "value": nominal / t_cdf, "unit": "equivalent N(0,1) draws/s",
           "config": {"workload": PCS_CONFIG, "samples": num_samples, "arms": K, "contexts": n_c,
                      "sampler": "cdf stream (crs_pcs_counts_cdf): table build + sampling kernels"},
           "ms_per_estimate": t_cdf * 1e3, "sample_context_pairs_per_sec": float(num_samples) * n_c / t_cdf,
           "philox_calls_per_sample": res["cdf"]["stats"][0] / (float(num_samples) * n_c),
           "pcs_mean": res["cdf"]["pcs_mean"], "gpu_launches": launches,
           "kernel_ms": {"pcs_cdf_build": t_build * 1e3, "pcs_cdf_sample": (t_cdf - t_build) * 1e3},
           "roofline": pcs_cdf_roofline("pcs_cdf_sample_kernel", res["cdf"]["stats"][0], max(t_cdf - t_build, 1e-9), 1, rates),
           "roofline_build": pcs_cdf_roofline("pcs_cdf_build_kernel", res["cdf"]["stats"][1], t_build, 1, rates),
           "tree_stream": {"ms_per_estimate": res["tree"]["ms"], "pcs_mean": res["tree"]["pcs_mean"],
                           "philox_calls_per_sample": sum(res["tree"]["stats"][:3]) / (float(num_samples) * n_c)},
           "flat_stream": {"ms_per_estimate": res["flat"]["ms"], "pcs_mean": res["flat"]["pcs_mean"],
                           "executed_draws_per_sec": res["flat"]["stats"][0] / (res["flat"]["ms"] * 1e-3)}}
    if cpu:
        from oracle import pcs_oracle
        torch.set_num_threads(os.cpu_count() or 1)
        sub_c, sub_s = 64, 2048
        g = torch.Generator().manual_seed(0)
        mu, vr = mean[:sub_c].double().cpu(), var[:sub_c].double().cpu()
        pcs_oracle.pcs_marginal_ref(mu, vr, 256, generator=g)
        t0 = time.perf_counter()
        n_rep = 0
        while time.perf_counter() - t0 < 6.0:
            pcs_oracle.pcs_marginal_ref(mu, vr, sub_s, generator=g)
            n_rep += 1
        dt = (time.perf_counter() - t0) / n_rep
        out["cpu_baseline"] = {"value": sub_s * sub_c * K / dt, "unit": "N(0,1) draws/s", "cores": torch.get_num_threads(),
                               "kind": "port", "sample": f"{n_rep} x marginal sampler (oracle.pcs_marginal_ref) on {sub_c} contexts x {K} arms x "
                               f"{sub_s} samples, fp64, PyTorch CPU; the reference's joint sampler is O(n_c^3) and cannot run at this size"}
        # the reference-faithful JOINT sampler at config 1 (10 arms x 10 contexts), the case of the
        # reference's notebooks/pcs_sampling_overhead.ipynb, next to the kernel on the same posterior
        from oracle import gp_oracle
        d1, c1 = make_config("c1")
        ref1 = gp_oracle.model_from_description(d1)
        arm_set = torch.arange(10, dtype=torch.float64).view(-1, 1)
        f_I, rho = pcs_oracle.strict_indicator(), (lambda P: P.mean(dim=-2))
        for S1 in (64, 1024):
            pcs_oracle.estimate_current_generalized_pcs_ref(ref1, arm_set, c1, S1, None, f_I, rho)
        t0, n_rep = time.perf_counter(), 0
        while time.perf_counter() - t0 < 2.0:
            v_ref = pcs_oracle.estimate_current_generalized_pcs_ref(ref1, arm_set, c1, 1024, None, f_I, rho)
            n_rep += 1
        dt_ref = (time.perf_counter() - t0) / n_rep
        m1 = crs.B200ModelListGP(d1, device=dev)
        c1d = c1.to(dev)
        crs.estimate_current_generalized_pcs(m1, arm_set, c1d, 1024, None, f_I, rho, seed=1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(50):
            v_gpu = crs.estimate_current_generalized_pcs(m1, arm_set, c1d, 1024, None, f_I, rho, seed=2 + i)
        torch.cuda.synchronize()
        dt_gpu = (time.perf_counter() - t0) / 50
        out["c1_joint_sampler"] = {"workload": "c1: estimate_current_generalized_pcs, 10 arms x 10 contexts, 1,024 samples, posterior included",
                                   "reference_port_ms": dt_ref * 1e3, "b200_ms": dt_gpu * 1e3, "cores": torch.get_num_threads(),
                                   "pcs_reference_port": float(v_ref), "pcs_b200": float(v_gpu),
                                   "note": "oracle.estimate_current_generalized_pcs_ref = the reference's joint MVN sampler "
                                           "(generalized_pcs.py:441-479) restated; the B200 number is the host API call (wall clock)"}
    return out


# ------------------------------------------------------------------------------------------ loop
def run_loop(args, env):
    """BASELINE config 5: the end-to-end sequential loop (experiments/wsc_experiments/main.py:345-470)
    with the context set sharded over the ranks.  One step = decide (scan + step 2(b) on the cached
    grid, peer fold, decision read back by the host) -> simulate -> PCS (2^16 samples, peer sum) ->
    condition the sampled arm on the new observation and refresh its grid column."""
    import contextual_rs_b200 as crs
    from contextual_rs_b200.peaks import measure_pcs_cdf_rates, pcs_cdf_roofline
    from contextual_rs_b200.sharded import ShardedSequentialRS
    from contextual_rs_b200.synthetic import make_config
    dev, world, rank, dist = env.dev, env.world, env.rank, env.dist
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    desc, ctx = make_config("c5")
    K, n_c, d = len(desc["train_X"]), ctx.shape[0], ctx.shape[1]
    S = PCS_SAMPLES
    model = crs.B200ModelListGP(desc, device=dev, dtype=dtype)
    loop = ShardedSequentialRS(model, ctx.to(dtype), kernel_scale=1.0, num_pcs_samples=S, pcs_seed=0, exchange=args.exchange)
    gen = torch.Generator().manual_seed(7)
    W, steps = max(args.warmup, 3), args.steps
    noise = 0.1 * torch.randn(W + steps, generator=gen, dtype=torch.float64)
    it_box = [0]

    def truth(arm, x):  # the synthetic sine surface of SURVEY.md §8d; identical on every rank
        full = torch.cat([torch.tensor([arm / (K - 1.0)], dtype=torch.float64), x.double()])
        return torch.sin(10.0 * full).sum() + noise[it_box[0]]

    stats = loop.dec.ws["stats"]
    for _ in range(W):
        loop.step(truth)
        it_box[0] += 1
    env.barrier()
    sampler = ClockSampler(env.local_rank)
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    work = torch.zeros(2, dtype=torch.float64)
    env.barrier()
    t_wall = time.perf_counter()
    for i in range(steps):
        ev[i][0].record()
        out = loop.decide()
        ev[i][1].record()
        y = truth(out["next_arm"], out["next_context"])
        pcs = loop.pcs()
        ev[i][2].record()
        work += stats[:2].cpu().to(torch.float64)  # {Philox calls, table terms} of this estimate
        loop.observe(out["next_arm"], out["next_context"].view(1, -1), y.reshape(1, 1))
        loop.iteration += 1
        it_box[0] += 1
        ev[i][3].record()
    env.barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    parts = torch.tensor([[e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in ev], dtype=torch.float64).sum(dim=0)
    dec_ms, pcs_ms, obs_ms, tot_ms, wall_s = env.max_over_ranks(parts.tolist() + [float(parts.sum()), t_wall])
    calls_all, terms_all = env.sum_over_ranks(work.tolist())
    if rank == 0:
        w = 8 if dtype == torch.float64 else 4
        val = steps / (tot_ms * 1e-3)
        rates = measure_pcs_cdf_rates(dev)
        emit(({
            "metric": "rs_loop_iterations_per_sec", "value": val, "unit": "iterations/s",
            "n_gpus": world, "steps": steps, "warmup": W, "ms_per_step": tot_ms / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": dict(workload_config("loop", world, args),
                           l2="no flush needed: the grid tables a step re-reads (2 x %d MB per rank) exceed the 126 MB L2" % (K * n_c * w // world // 1000000)),
            "breakdown_ms": {"decide": dec_ms / steps, "pcs": pcs_ms / steps, "observe": obs_ms / steps},
            "e2e": {"value": steps / wall_s, "unit": "iterations/s", "h2d_bytes_per_step": (d + 1) * 8 * world, "d2h_bytes_per_step": (64 + 80) * world,
                    "note": "the loop is host-driven: every iteration reads the decision and the PCS value back and uploads the new "
                            "observation; e2e = wall clock of the same steps"},
            "gpu_launches": 6 * steps * world, "exchange": loop.dec.exchange,
            "roofline": dict(pcs_cdf_roofline("pcs_cdf_sample_kernel", calls_all / steps, pcs_ms / steps * 1e-3, world, rates),
                             note="time = both cdf kernels + the peer sum of one iteration, work = the sampling calls of one iteration over "
                                  "the bare sampling-call rate: a lower bound of the sampling kernel's fraction",
                             table_terms_per_iteration=terms_all / steps),
            "pcs_last": float(pcs), "clocks": clocks}))
    if loop.dec.mailbox is not None:
        loop.dec.mailbox.close()
    env.finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--exchange", default="mailbox", choices=["mailbox", "nccl"])
    ap.add_argument("--partition", default="contexts", choices=["contexts", "arms"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush between steps (profiling only)")
    ap.add_argument("--no-sub", action="store_true", help="headline: skip the config-2 and config-3 sub-objects")
    ap.add_argument("--shard", action="store_true", help="N > 1: shard ONE decision over the ranks even for c2")
    args = ap.parse_args()
    name = args.workload
    if args.impl == "reference":
        return run_reference(args, name)
    env = Env(args)
    if name == "loop":
        return run_loop(args, env)
    if name == "headline":
        if args.partition == "arms":
            return run_headline_arms(args, env)
        return run_headline(args, env)
    return run_decision(args, env, name)


if __name__ == "__main__":
    main()
