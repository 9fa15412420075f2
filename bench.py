#!/usr/bin/env python
"""bench.py — the GP-C-OCBA decision benchmark on BASELINE.json config 2 (default workload).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl reference]

A *step* is one whole GP-C-OCBA decision on a model whose caches are cold: per-arm Cholesky
state (crs_gp_prepare, all arms), ModelListGP posterior over the arm x context grid
(crs_posterior_grid), zeta scan + global argmin + OCBA step 2(b) fused in one launch
(crs_ocba_scan_select) — what `gao_modellist(model, context_set)` does in the reference
(contextual_rs/contextual_rs_strategies.py:100-202) on a freshly fitted model.

  value : decisions/s with every input already resident in HBM (device timed, CUDA events).
  e2e   : decisions/s through the public drop-in `contextual_rs_b200.gao_modellist(model,
          context_set)` with the context set in pinned HOST memory: host->device copy of the
          contexts and device->host read of the decision inside every timed step.
  N > 1 : default workload (c2, a 65-us decision): every rank runs its own independent replication
          of the decision stream (own seeded problem of the same shape, no collective on the data
          path — the way the reference's experiments use many workers) -> "weak" scaling, value =
          decisions of all ranks / max-over-ranks time.  With --workload c3|c4|c5 (or --shard) ONE
          decision is sharded over the ranks by contexts (NCCL: one 64-byte-per-rank all-gather per
          decision) -> "strong" scaling.  The `pcs` (config 3) and `headline` (config 4: decision +
          2^16-sample PCS over 1,000 arms x 65,536 contexts, target < 10 ms on 8 GPUs) sub-objects
          are always context-sharded over all ranks.

The L2 is flushed (256 MiB write) between timed steps; each step is bracketed by its own pair of
CUDA events on the launching stream and the per-step durations are summed.
`--impl reference` times the CPU restatement of the reference (oracle/, PyTorch on the host
cores; BoTorch/GPyTorch are not installable here) on the same workload.
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

WORKLOADS = {
    "c2": "c2: GP-C-OCBA decision, 100 arms x 1,024 contexts, 4-D, 20 obs/arm, Matern-5/2",
    "c3": "c3: GP-C-OCBA decision, 256 arms x 8,192 contexts, 4-D, 20 obs/arm, Matern-5/2",
    "c4": "c4: GP-C-OCBA decision, 1,000 arms x 65,536 Sobol contexts, 8-D, 18 obs/arm, Matern-5/2",
    "c5": "c5: GP-C-OCBA decision, 1,000 arms x 16,384 contexts, 4-D, 20 obs/arm, Matern-5/2",
    "loop": "c5 loop: sequential R&S, 1,000 arms x 16,384 contexts, 4-D, 20 obs/arm growing by one per iteration; "
            "every iteration = GP-C-OCBA decision on the cached grid + 2^16-sample PCS + dirty-arm update",
}


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
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
    capture of the same workload (profiles/r01d_traffic.json; fp64 captures only), else None."""
    path = os.path.join(ROOT, "profiles", "r01d_traffic.json")
    if dtype != "f64" or not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get(workload, {}).get(kernel)


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


def time_cpu(desc, ctx, budget_s, max_steps):
    torch.set_num_threads(os.cpu_count() or 1)
    best = None
    # two CPU restatements: "modellist" walks the K sub-models like the reference's ModelListGP does,
    # "batched" stacks them (bmm / batched Cholesky); the faster one (after a warm-up call) is the baseline
    for variant in ("batched", "modellist"):
        cpu_decision_step(desc, ctx, variant)
        dt = float("inf")
        for _ in range(2):
            t0 = time.perf_counter()
            cpu_decision_step(desc, ctx, variant)
            dt = min(dt, time.perf_counter() - t0)
        if best is None or dt < best[1]:
            best = (variant, dt)
    variant, est = best
    steps = int(max(2, min(max_steps, budget_s / max(est, 1e-6))))
    cpu_decision_step(desc, ctx, variant)  # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_decision_step(desc, ctx, variant)
    dt = (time.perf_counter() - t0) / steps
    return variant, steps, dt


def run_reference(args, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from contextual_rs_b200.synthetic import make_config
    desc, ctx = make_config(name)
    K, n_c = len(desc["train_X"]), ctx.shape[0]
    full = name == "c2"
    if not full:  # bounded sample: a slice of the contexts, scaled linearly
        sub = max(256, n_c // 64)
        ctx_s = ctx[:sub]
    else:
        sub, ctx_s = n_c, ctx
    variant, steps, dt = time_cpu(desc, ctx_s, budget_s=min(60.0, 4.0 * max(args.steps, 1)), max_steps=max(args.steps, 2))
    dt_full = dt * (n_c / sub)
    cores = torch.get_num_threads()
    sample = (f"{steps} decisions on {sub}/{n_c} contexts x {K} arms, oracle '{variant}' variant "
              f"(PyTorch CPU, {cores} threads), time scaled x{n_c / sub:.0f}")
    val = 1.0 / dt_full
    emit(({
        "impl": "reference", "metric": "gp_c_ocba_decisions_per_sec", "value": val, "unit": "decisions/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": dt_full * 1e3,
        "higher_is_better": True, "scaling": "weak" if name == "c2" else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOADS[name]},
        "cpu_baseline": {"value": val, "unit": "decisions/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "decisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))



# ------------------------------------------------------------------------------------------- PCS
PCS_CONFIG = "c3: generalized-PCS MC estimate, 2^16 samples, 256 arms x 8,192 contexts, fp32"


def run_pcs(dev, world, rank, reps=3, num_samples=1 << 16, cpu=True):
    """PCS MC samples/s on config 3 (fp32).  Contexts are sharded over ranks; counts are combined
    with one all-reduce.  'value' counts the S*n_c*K nominal draws of the estimator ("equivalent
    samples").  The max-tree kernel decides a sample with a root test plus a few node expansions
    (one Philox call each) instead of K draws; the executed steps are reported beside the value
    and are what the roofline fraction is computed from.  The flat per-arm stream is timed too."""
    import torch.distributed as dist
    import contextual_rs_b200 as crs
    from contextual_rs_b200.generalized_pcs import pcs_counts
    from contextual_rs_b200.sharded import shard_bounds
    from contextual_rs_b200.synthetic import make_config
    desc, ctx = make_config("c3")
    K, n_c = len(desc["train_X"]), ctx.shape[0]
    model = crs.B200ModelListGP(desc, device=dev, dtype=torch.float32)
    b, e = shard_bounds(n_c, world, rank)
    post = model.posterior(ctx[b:e].float())
    mean, var = post.mean.contiguous(), post.variance.contiguous()
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.float64, device=dev)

    def step(i, sampler="tree"):
        c = pcs_counts(model, mean, var, num_samples, seed=0, offset=i, ctx_offset=b, stats=stats, sampler=sampler)
        total[0] = c.sum(dtype=torch.int64).to(torch.float64)
        if world > 1:
            dist.all_reduce(total)

    step(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = []
    for i in range(reps):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step(1 + i)
        t.record()
        t.synchronize()
        ms.append(s.elapsed_time(t))
    t_step = sum(ms) / len(ms) * 1e-3
    tt = torch.tensor([t_step], dtype=torch.float64, device=dev)
    st = stats.to(torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(st)
    t_step = float(tt)
    roots, expansions, directs = st.tolist()[:3]
    pcs_tree_mean = float(total) / (num_samples * n_c)
    # the flat per-arm stream (crs_pcs_counts) on the same tables, for comparison
    step(0, "flat")
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    step(1, "flat")
    t.record()
    t.synchronize()
    tf = torch.tensor([s.elapsed_time(t) * 1e-3], dtype=torch.float64, device=dev)
    sf = stats.to(torch.float64)
    if world > 1:
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        dist.all_reduce(sf)
    nominal = float(num_samples) * n_c * K
    out = {"metric": "pcs_mc_samples_per_sec", "value": nominal / t_step, "unit": "equivalent N(0,1) draws/s",
           "config": {"workload": PCS_CONFIG, "samples": num_samples, "arms": K, "contexts": n_c,
                      "sharding": f"contexts/{world}" if world > 1 else "none", "sampler": "max-tree stream (crs_pcs_counts_tree)"},
           "ms_per_estimate": t_step * 1e3, "sample_context_pairs_per_sec": float(num_samples) * n_c / t_step,
           "philox_calls_per_sample": (roots + expansions + directs) / (float(num_samples) * n_c),
           "pcs_mean": pcs_tree_mean, "gpu_launches": reps + 3,
           "flat_stream": {"ms_per_estimate": float(tf) * 1e3, "executed_draws_per_sec": float(sf[0]) / float(tf),
                           "executed_fraction": float(sf[0]) / float(sf[0] + sf[1]),
                           "pcs_mean": float(total) / (num_samples * n_c)}}
    from contextual_rs_b200.peaks import measure_pcs_tree_rates, pcs_tree_roofline
    out["roofline"] = pcs_tree_roofline(roots, expansions, t_step, world, measure_pcs_tree_rates(dev), directs)
    if cpu and world == 1 and rank == 0:
        from oracle import pcs_oracle
        torch.set_num_threads(os.cpu_count() or 1)
        sub_c, sub_s = 64, 2048
        g = torch.Generator().manual_seed(0)
        mu, vr = mean[:sub_c].double().cpu(), var[:sub_c].double().cpu()
        pcs_oracle.pcs_marginal_ref(mu, vr, 256, generator=g)
        t0 = time.perf_counter()
        n_rep = 0
        while time.perf_counter() - t0 < 10.0:
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


# -------------------------------------------------------------------------------------- headline
HEADLINE_CONFIG = ("c4: one GP-C-OCBA decision (fp64, cold caches) + one 2^16-sample generalized-PCS estimate over "
                   "1,000 arms x 65,536 Sobol contexts, 8-D, 18 obs/arm")


def run_headline(dev, world, rank, reps=3, num_samples=1 << 16):
    """BASELINE.json's target: decision + 2^16-sample PCS over 1,000 arms x 65,536 contexts in
    under 10 ms on 8 GPUs.  Contexts are sharded over the ranks; a step is prepare(all arms) +
    grid + scan/select + 64-byte all-gather, then PCS counts + one scalar all-reduce."""
    import torch.distributed as dist
    import contextual_rs_b200 as crs
    from contextual_rs_b200.sharded import ContextShardedDecider, REC_CONTEXT, REC_NEXT_ARM
    from contextual_rs_b200.synthetic import make_config
    desc, ctx = make_config("c4")
    K, n_c = len(desc["train_X"]), ctx.shape[0]
    model = crs.B200ModelListGP(desc, device=dev, dtype=torch.float64)
    decider = ContextShardedDecider(model, ctx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rec = decider.decide(prepare_arms=(0, K))
    pcs = decider.pcs_mean(num_samples, seed=0, offset=0)
    barrier()
    t_dec, t_pcs = [], []
    for i in range(reps):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        rec = decider.decide(prepare_arms=(0, K))
        e1.record()
        pcs = decider.pcs_mean(num_samples, seed=0, offset=1 + i)
        e2.record()
        e2.synchronize()
        t_dec.append(e0.elapsed_time(e1))
        t_pcs.append(e1.elapsed_time(e2))
        barrier()
    t = torch.tensor([sum(t_dec) / reps, sum(t_pcs) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    d_ms, p_ms = t.tolist()
    return {"workload": HEADLINE_CONFIG, "n_gpus": world, "sharding": f"contexts/{world}" if world > 1 else "none",
            "ms_per_step": d_ms + p_ms, "decision_ms": d_ms, "pcs_ms": p_ms, "target_ms": 10.0, "target_gpus": 8,
            "steps": reps, "next_arm": int(rec[REC_NEXT_ARM]), "next_context": int(rec[REC_CONTEXT]),
            "pcs_mean": float(pcs), "pcs_equivalent_draws_per_sec": float(num_samples) * n_c * K / (p_ms * 1e-3)}


# ------------------------------------------------------------------------------------------ loop
def run_loop(args, dev, world, rank):
    """BASELINE config 5: the end-to-end sequential loop (experiments/wsc_experiments/main.py:345-470)
    with the context set sharded over the ranks.  One step = decide (scan + step 2(b) on the cached
    grid, 64-byte all-gather, decision read back by the host) -> simulate -> PCS (2^16 samples, one
    all-reduce) -> condition the sampled arm on the new observation and refresh its grid column."""
    import torch.distributed as dist
    import contextual_rs_b200 as crs
    from contextual_rs_b200.peaks import measure_pcs_tree_rates, pcs_tree_roofline
    from contextual_rs_b200.sharded import ShardedSequentialRS
    from contextual_rs_b200.synthetic import make_config
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    desc, ctx = make_config("c5")
    K, n_c, d = len(desc["train_X"]), ctx.shape[0], ctx.shape[1]
    S = 1 << 16
    model = crs.B200ModelListGP(desc, device=dev, dtype=dtype)
    loop = ShardedSequentialRS(model, ctx.to(dtype), kernel_scale=1.0, num_pcs_samples=S, pcs_seed=0)
    gen = torch.Generator().manual_seed(7)
    W, steps = max(args.warmup, 3), args.steps
    noise = 0.1 * torch.randn(W + steps, generator=gen, dtype=torch.float64)
    it_box = [0]

    def truth(arm, x):  # the synthetic sine surface of SURVEY.md §8d; identical on every rank
        full = torch.cat([torch.tensor([arm / (K - 1.0)], dtype=torch.float64), x.double()])
        return torch.sin(10.0 * full).sum() + noise[it_box[0]]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stats = loop.dec.ws["stats"]
    for _ in range(W):
        loop.step(truth)
        it_box[0] += 1
    barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    drawn = torch.zeros(3, dtype=torch.float64)
    barrier()
    for i in range(steps):
        ev[i][0].record()
        out = loop.decide()
        ev[i][1].record()
        y = truth(out["next_arm"], out["next_context"])
        pcs = loop.pcs()
        ev[i][2].record()
        drawn += stats[:3].cpu().to(torch.float64)  # {root tests, node expansions, direct-arm tests} of this estimate
        loop.observe(out["next_arm"], out["next_context"].view(1, -1), y.reshape(1, 1))
        loop.iteration += 1
        it_box[0] += 1
        ev[i][3].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    parts = torch.tensor([[e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in ev], dtype=torch.float64).sum(dim=0)
    t = torch.cat([parts, parts.sum().reshape(1), drawn]).to(dev)
    if world > 1:
        tmax = t[:4].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dsum = t[4:].clone()
        dist.all_reduce(dsum)
        t = torch.cat([tmax, dsum])
    dec_ms, pcs_ms, obs_ms, tot_ms, roots_all, exp_all, dir_all = t.tolist()
    if rank == 0:
        w = 8 if dtype == torch.float64 else 4
        val = steps / (tot_ms * 1e-3)
        emit(({
            "metric": "rs_loop_iterations_per_sec", "value": val, "unit": "iterations/s (decision + 2^16-sample PCS + update)",
            "n_gpus": world, "steps": steps, "warmup": W, "ms_per_step": tot_ms / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOADS["loop"], "arms": K, "contexts": n_c, "dim": d, "pcs_samples": S,
                       "sharding": f"contexts/{world}" if world > 1 else "none",
                       "l2": "no flush needed: the grid tables a step re-reads (2 x %d MB per rank) exceed the 126 MB L2" % (K * n_c * w // world // 1000000)},
            "breakdown_ms": {"decide": dec_ms / steps, "pcs": pcs_ms / steps, "observe": obs_ms / steps},
            "e2e": {"value": val, "unit": "iterations/s", "h2d_bytes_per_step": (d + 1) * 8 * world, "d2h_bytes_per_step": 64 * world,
                    "note": "the loop is host-driven: every iteration reads the decision back and uploads the new observation"},
            "gpu_launches": 5 * steps * world,
            "roofline": pcs_tree_roofline(roots_all, exp_all, pcs_ms * 1e-3, world, measure_pcs_tree_rates(dev), dir_all),
            "pcs_last": float(pcs), "clocks": clocks}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- GPU
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush between steps (profiling only)")
    ap.add_argument("--no-pcs", action="store_true", help="skip the PCS (config 3) and headline (config 4) sub-benchmarks")
    ap.add_argument("--shard", action="store_true", help="N > 1: shard ONE decision over the ranks even for c2")
    args = ap.parse_args()
    name = args.workload
    if args.impl == "reference":
        return run_reference(args, name)

    import torch.distributed as dist
    import contextual_rs_b200 as crs
    from contextual_rs_b200 import _lib
    from contextual_rs_b200.contextual_rs_strategies import _scan, _scan_select
    from contextual_rs_b200.sharded import ContextShardedDecider, REC_CONTEXT, REC_NEXT_ARM, shard_bounds
    from contextual_rs_b200.synthetic import make_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: contextual_rs_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if name == "loop":
        if args.steps == 3000:
            args.steps = 20
        return run_loop(args, dev, world, rank)
    W, K_steps = max(args.warmup, 3), args.steps
    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    w = 8 if dtype == torch.float64 else 4

    # N > 1: c2 runs one independent replication per rank (weak scaling); the larger workloads
    # shard one decision over the ranks by contexts (strong scaling)
    sharded = world > 1 and (name != "c2" or args.shard)
    desc, ctx = make_config(name) if (sharded or world == 1) else make_config(name, seed=rank)
    K, n_c, d = len(desc["train_X"]), ctx.shape[0], ctx.shape[1]
    n_obs = desc["train_X"][0].shape[0]
    model = crs.B200ModelListGP(desc, device=dev, dtype=dtype)
    model.check_status()
    lib = _lib.get()
    peaks, peak_src = load_peaks()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush():
        if not args.no_flush:
            flush_buf.fill_(1)

    if sharded:
        begin, end = shard_bounds(n_c, world, rank)
        decider = ContextShardedDecider(model, ctx)
        ctx_dev = decider.local_ctx
    else:
        begin, end = 0, n_c
        decider = None
        ctx_dev = ctx.to(device=dev, dtype=dtype).contiguous()
    n_loc = end - begin
    ws = model.workspace(max(n_loc, 1))
    stream = _lib.stream_ptr(dev)
    st = C.byref(model.state)

    def device_step():
        """prepare -> grid -> fused scan + select, all inputs resident, no host round-trip."""
        _lib.check(lib.crs_gp_prepare(st, 0, K, stream))
        if n_loc:
            _lib.check(lib.crs_posterior_grid(st, ctx_dev.data_ptr(), n_loc, 0, K, ws["mean"].data_ptr(),
                                              ws["var"].data_ptr(), K, 0, stream))
            _scan_select(lib, model, ws, n_loc, 0, 2.0, stream, ctx_ptr=ctx_dev.data_ptr())
        if sharded:
            dist.all_gather_into_tensor(decider._gather, ws["decision"] if n_loc else decider._empty)

    ctx_host = ctx.to(dtype).pin_memory()

    def e2e_step():
        if not sharded:
            from contextual_rs_b200.contextual_rs_strategies import decide
            dec = decide(model, ctx_host, True, 0, 2.0, prepare_arms=(0, K))
            return dec.next_arm, dec.context
        # sharded: every rank uploads its slice of the pinned host context set, then one all-gather
        decider.local_ctx.copy_(ctx_host[begin:end], non_blocking=True)
        rec = decider.decide(prepare_arms=(0, K))
        return int(rec[REC_NEXT_ARM]), int(rec[REC_CONTEXT])

    def timed_loop(fn, steps, per_kernel=False):
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            flush()
            starts[i].record()
            fn()
            ends[i].record()
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in zip(starts, ends)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -------------------------------------------------------------------------------
    for _ in range(W):
        flush()
        device_step()
    for _ in range(W):
        flush()
        e2e_step()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident value ---------------------------------------------------------------------
    barrier()
    dev_ms = timed_loop(device_step, K_steps)
    barrier()
    # ---- dominant-kernel duration (posterior grid), same region style -------------------------------
    def grid_only():
        if n_loc:
            lib.crs_posterior_grid(st, ctx_dev.data_ptr(), n_loc, 0, K, ws["mean"].data_ptr(),
                                   ws["var"].data_ptr(), K, 0, stream)
    def scan_only():
        if n_loc:
            _scan(lib, model, ws, n_loc, stream)
    def prep_only():
        lib.crs_gp_prepare(st, 0, K, stream)
    ksteps = min(K_steps, 500)
    grid_ms = timed_loop(grid_only, ksteps)
    scan_ms = timed_loop(scan_only, ksteps)
    prep_ms = timed_loop(prep_only, ksteps)
    barrier()
    # ---- end to end through the public API -------------------------------------------------------
    e2e_ms = timed_loop(e2e_step, K_steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    t_dev = sum(dev_ms) / 1e3
    t_e2e = sum(e2e_ms) / 1e3
    if world > 1:
        t = torch.tensor([t_dev, t_e2e, sum(grid_ms) / len(grid_ms), sum(scan_ms) / len(scan_ms)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, grid_avg, scan_avg = t.tolist()
    else:
        grid_avg, scan_avg = sum(grid_ms) / len(grid_ms), sum(scan_ms) / len(scan_ms)
    prep_avg = sum(prep_ms) / len(prep_ms)

    if rank == 0:
        mult = 1 if (sharded or world == 1) else world  # replicas: every rank made K_steps decisions
        value = mult * K_steps / t_dev
        flops = algorithmic_flops(K, n_loc, n_obs, d)
        from contextual_rs_b200.peaks import measure_fma_peak
        fp64_peak = measure_fma_peak(0 if dtype == torch.float64 else 1, dev)  # live FMA micro-kernel
        achieved = flops / (grid_avg * 1e-3) / 1e12
        scan_bytes = 2 * K * n_loc * w + n_loc * (w + 12)
        out = {
            "metric": "gp_c_ocba_decisions_per_sec", "value": value, "unit": "decisions/s",
            "n_gpus": world, "steps": K_steps, "warmup": W, "ms_per_step": t_dev / K_steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOADS[name], "arms": K, "contexts": n_c, "dim": d, "obs_per_arm": n_obs,
                       "sharding": (f"contexts/{world}: one decision sharded over the ranks, 64-byte all-gather per decision"
                                    if sharded else (f"replicas x{world}: one independent decision stream per GPU, no "
                                                     "collective on the data path" if world > 1 else "none")),
                       "l2": "flushed between timed steps (256 MiB write)" if not args.no_flush else "not flushed",
                       "step": "3 kernels: prepare(all arms) + posterior grid + fused zeta scan / OCBA select"},
            "e2e": {"value": mult * K_steps / t_e2e, "unit": "decisions/s", "ms_per_step": t_e2e / K_steps * 1e3,
                    "h2d_bytes_per_step": int(n_c * d * w) * mult, "d2h_bytes_per_step": 64 * world},
            "gpu_launches": (3 * K_steps + 3 * K_steps + 3 * ksteps) * mult,
            "roofline": {"kernel": "posterior_grid_kernel", "bound": "fp64_fma" if dtype == torch.float64 else "fp32_fma",
                         "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
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
            "kernel_ms": {"gp_prepare": prep_avg, "posterior_grid": grid_avg, "ocba_scan": scan_avg},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            if name == "c2":
                variant, steps, dt = time_cpu(desc, ctx, budget_s=15.0, max_steps=200)
                scale = 1.0
                sub = n_c
            else:
                sub = max(256, n_c // 64)
                variant, steps, dt = time_cpu(desc, ctx[:sub], budget_s=15.0, max_steps=50)
                scale = n_c / sub
            out["cpu_baseline"] = {
                "value": 1.0 / (dt * scale), "unit": "decisions/s", "cores": torch.get_num_threads(), "kind": "port",
                "sample": f"{steps} decisions, oracle '{variant}' variant on {sub}/{n_c} contexts x {K} arms, "
                          f"time scaled x{scale:.0f}; PyTorch CPU, {torch.get_num_threads()} threads"}
    pcs = headline = None
    if not args.no_pcs and name == "c2":
        del flush_buf
        pcs = run_pcs(dev, world, rank, cpu=not args.no_cpu_baseline)
        headline = run_headline(dev, world, rank)
    if rank == 0:
        if pcs is not None:
            out["pcs"] = pcs
            out["headline"] = headline
            out["gpu_launches"] += pcs["gpu_launches"] + 4 * 4 * world  # headline: 4 kernels x (1 + 3) steps per rank
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
