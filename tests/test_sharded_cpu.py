"""world_size-2 (and 3) gloo tests of the multi-rank combination logic on the CPU.  The local
compute of each rank is done by the oracle on its shard (tests may use the oracle); what is under
test is shard_bounds / the gathered-record combination rules / the collectives wiring of
contextual_rs_b200.sharded, against the unsharded oracle decision."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _tables(seed=0):
    from contextual_rs_b200.synthetic import make_problem
    from oracle import gp_oracle
    desc, ctx = make_problem(K=9, n_c=37, d=2, n=7, seed=seed)
    ref = gp_oracle.model_from_description(desc)
    post = ref.posterior(ctx)
    return ref, ctx, post.mean, post.variance


def _local_record(mean, var, counts_fn, ctx_offset, arm_offset=0, ext=None):
    """Oracle restatement of scan + select on a shard -> float64 record."""
    from contextual_rs_b200 import sharded as sh
    n_c, K = mean.shape
    if n_c == 0:
        return sh.decision_record(float("inf"), 0, -1, -1, -1, -1)
    cols = torch.arange(K) + arm_offset
    if ext is None:
        best = mean.argmax(dim=-1)
        bm = mean.gather(1, best[:, None]); bv = var.gather(1, best[:, None])
        best_g = best + arm_offset
    else:
        bm, bv, best_g = ext[0][:, None], ext[1][:, None], ext[2]
    zeta = (bm - mean).pow(2) / (bv + var)
    zeta = torch.where(cols[None, :] == best_g[:, None], torch.full_like(zeta, float("inf")), zeta)
    zmin = zeta.min()
    ties = int((zeta == zmin).sum())
    flat = int(torch.argmin(zeta.view(-1)))
    c, a = divmod(flat, K)
    y = counts_fn(c + ctx_offset)[arm_offset:arm_offset + K] / var[c]
    is_best = cols == best_g[c]
    psi_best = float(y[is_best].sum())
    psi_rest = float(y[~is_best].sum())
    nxt = int(best_g[c]) if psi_best < psi_rest else a + arm_offset
    return sh.decision_record(float(zmin), ties, c + ctx_offset, a + arm_offset, int(best_g[c]), nxt, psi_best, psi_rest)


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from contextual_rs_b200 import sharded as sh
        ref, ctx, mean, var = _tables()
        n_c, K = mean.shape
        counts_all = torch.stack([torch.stack([(ti[0] == ctx[c]).all(dim=-1).sum() for ti in ref.train_inputs]).double()
                                  for c in range(n_c)])
        counts_all = counts_all + 1.0  # make psi non-trivial (normalised inputs rarely match exactly)
        counts_fn = lambda c: counts_all[c]
        if mode == "contexts":
            b, e = sh.shard_bounds(n_c, world, rank)
            rec = _local_record(mean[b:e], var[b:e], counts_fn, b)
            out = sh.combine_context_sharded(sh.all_gather_rows(rec))
        else:
            a0, a1 = sh.shard_bounds(K, world, rank)
            lm, lv = mean[:, a0:a1], var[:, a0:a1]
            lb = lm.argmax(dim=-1)
            g = sh.combine_local_best(sh.all_gather_rows(lm.gather(1, lb[:, None]).squeeze(1)),
                                      sh.all_gather_rows(lv.gather(1, lb[:, None]).squeeze(1)),
                                      sh.all_gather_rows(lb + a0))
            rec = _local_record(lm, lv, counts_fn, 0, arm_offset=a0, ext=g)
            out = sh.combine_arm_sharded(sh.all_gather_rows(rec))
            # psi partial sums at the winning context: one all-reduce
            c = int(out[sh.REC_CONTEXT])
            cols = torch.arange(a0, a1)
            y = counts_all[c, a0:a1] / lv[c]
            part = torch.stack([y[cols == int(g[2][c])].sum(), y[cols != int(g[2][c])].sum()])
            dist.all_reduce(part)
            out[sh.REC_PSI_BEST], out[sh.REC_PSI_REST] = part[0], part[1]
            out[sh.REC_BEST_ARM] = float(g[2][c])
            out[sh.REC_NEXT_ARM] = out[sh.REC_BEST_ARM] if part[0] < part[1] else out[sh.REC_ZETA_ARM]
        if rank == 0:
            q.put(out.tolist())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["contexts", "arms"])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_decision_equals_unsharded(mode, world):
    from contextual_rs_b200 import sharded as sh
    ctxq = mp.get_context("spawn")
    q = ctxq.Queue()
    port = _free_port()
    procs = [ctxq.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, ctx, mean, var = _tables()
    n_c, K = mean.shape
    counts_all = torch.stack([torch.stack([(ti[0] == ctx[c]).all(dim=-1).sum() for ti in ref.train_inputs]).double()
                              for c in range(n_c)]) + 1.0
    want = _local_record(mean, var, lambda c: counts_all[c], 0)
    assert got[sh.REC_MIN_ZETA] == float(want[sh.REC_MIN_ZETA])
    for col in (sh.REC_TIES, sh.REC_CONTEXT, sh.REC_ZETA_ARM, sh.REC_BEST_ARM, sh.REC_NEXT_ARM):
        assert int(got[col]) == int(want[col]), col
    assert abs(got[sh.REC_PSI_BEST] - float(want[sh.REC_PSI_BEST])) < 1e-12
    assert abs(got[sh.REC_PSI_REST] - float(want[sh.REC_PSI_REST])) < 1e-9
    # and the unsharded record agrees with the oracle's gao_modellist zeta argmin
    from oracle import ocba_oracle
    zeta, order = ocba_oracle.zeta_table(mean, var)
    flat = int(torch.argmin(zeta.view(-1)))
    assert int(want[sh.REC_CONTEXT]) == flat // (K - 1)
    assert int(want[sh.REC_ZETA_ARM]) == int(order[flat // (K - 1), flat % (K - 1) + 1])


def test_shard_bounds_cover_everything():
    from contextual_rs_b200.sharded import shard_bounds
    for total in (0, 1, 7, 1024, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_combine_rules_ties_and_empty_shards():
    from contextual_rs_b200 import sharded as sh
    recs = torch.stack([
        sh.decision_record(0.5, 2, 40, 3, 1, 3),
        sh.decision_record(0.5, 1, 12, 2, 0, 2),
        sh.decision_record(float("inf"), 0, -1, -1, -1, -1),
        sh.decision_record(0.7, 9, 3, 1, 0, 1),
    ])
    out = sh.combine_context_sharded(recs)
    assert (int(out[sh.REC_CONTEXT]), int(out[sh.REC_ZETA_ARM]), int(out[sh.REC_TIES])) == (12, 2, 3)


def test_arm_sharded_cross_shard_tie_rule():
    """Equal zeta at the SAME context on two shards: the zeta arm with the larger mean wins, then the lower
    arm id (the reference's descending-mean sort order = the single-GPU kernel's rule = crs_arm_fold)."""
    from contextual_rs_b200 import sharded as sh
    recs = torch.stack([sh.decision_record(0.25, 1, 4, 2, 7, -1),    # shard 0: arm 2 at context 4
                        sh.decision_record(0.25, 2, 4, 5, 7, -1),    # shard 1: arm 5 at context 4, same zeta
                        sh.decision_record(0.30, 1, 1, 8, 6, -1)])
    out = sh.combine_arm_sharded(recs, mean_at=torch.tensor([1.0, 3.0, 9.0], dtype=torch.float64))
    assert int(out[sh.REC_ZETA_ARM]) == 5 and int(out[sh.REC_TIES]) == 3 and int(out[sh.REC_CONTEXT]) == 4
    out = sh.combine_arm_sharded(recs, mean_at=torch.tensor([3.0, 3.0, 9.0], dtype=torch.float64))
    assert int(out[sh.REC_ZETA_ARM]) == 2                            # equal means: lower arm id
    out = sh.combine_arm_sharded(recs)                                # without means: lower arm id
    assert int(out[sh.REC_ZETA_ARM]) == 2
    recs[2, sh.REC_MIN_ZETA] = 0.25                                   # a third record at the minimum, lower context
    out = sh.combine_arm_sharded(recs, mean_at=torch.tensor([1.0, 3.0, 9.0], dtype=torch.float64))
    assert int(out[sh.REC_CONTEXT]) == 1 and int(out[sh.REC_ZETA_ARM]) == 8 and int(out[sh.REC_TIES]) == 4
