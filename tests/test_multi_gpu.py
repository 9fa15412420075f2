"""N-rank NCCL tests on real GPUs (run by `torchrun` under `gpurun --gpus N`; skipped otherwise):
context-sharded and arm-sharded decisions + PCS must equal the single-GPU result."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

WORLD = int(os.environ.get("WORLD_SIZE", "1"))


@pytest.mark.skipif(WORLD < 2, reason="needs torchrun with >= 2 ranks")
def test_sharded_deciders_equal_single_gpu():
    import torch.distributed as dist
    import contextual_rs_b200 as crs
    from contextual_rs_b200 import sharded as sh
    from contextual_rs_b200.generalized_pcs import pcs_counts
    from contextual_rs_b200.synthetic import make_problem
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    desc, ctx = make_problem(K=37, n_c=1000, d=3, n=11, seed=21)
    model = crs.B200ModelListGP(desc, device=dev)
    arm, x = crs.gao_modellist(model, ctx, use_kde=True, kernel_scale=1.0)
    ws = model.workspace(1000)
    single = pcs_counts(model, ws["mean"], ws["var"], 2048, seed=7).sum().item() / (2048 * 1000)
    c_star = int((ctx == x).all(dim=-1).nonzero()[0])

    dview = ws["decision_view"]
    want = (dview.min_zeta, dview.tie_count, dview.context, dview.zeta_arm, dview.best_arm, dview.next_arm,
            dview.psi_best, dview.psi_rest)
    for exchange in ("mailbox", "nccl"):
        cdec = sh.ContextShardedDecider(model, ctx, exchange=exchange)
        for rep in range(3 if exchange == "nccl" else 300):  # many exchanges: the double-buffered mailbox slots
            rec = cdec.decide(p_mode=1, kernel_scale=1.0, grid=rep == 0)
            assert (int(rec[sh.REC_NEXT_ARM]), int(rec[sh.REC_CONTEXT])) == (int(arm), c_star)
        assert tuple(rec.tolist()) == tuple(float(v) for v in want)  # every field of the combined record
        assert abs(float(cdec.pcs_mean(2048, seed=7)) - single) < 1e-12
        if cdec.mailbox is not None:
            cdec.mailbox.close()
    # a rank without contexts (n_c < world) takes part with an empty record
    tiny = sh.ContextShardedDecider(model, ctx[:1])
    a1, x1 = crs.gao_modellist(model, ctx[:1], use_kde=True, kernel_scale=1.0)
    rec = tiny.decide(p_mode=1, kernel_scale=1.0)
    assert (int(rec[sh.REC_NEXT_ARM]), int(rec[sh.REC_CONTEXT])) == (int(a1), 0)
    s1 = pcs_counts(model, model.workspace(1)["mean"], model.workspace(1)["var"], 512, seed=9).sum().item() / 512
    assert abs(float(tiny.pcs_mean(512, seed=9)) - s1) < 1e-12
    tiny.mailbox.close()

    adec = sh.ArmShardedDecider(desc, ctx, dev)
    rec = adec.decide(p_mode=1, kernel_scale=1.0)
    assert (int(rec[sh.REC_NEXT_ARM]), int(rec[sh.REC_CONTEXT])) == (int(arm), c_star)
    assert abs(float(adec.pcs_mean(2048, seed=7)) - single) < 1e-12
    dist.barrier()


@pytest.mark.skipif(WORLD < 2, reason="needs torchrun with >= 2 ranks")
def test_sharded_loop_equals_single_gpu_loop():
    """Config-5-style sequential loop: the context-sharded loop makes the same decisions and PCS
    estimates as the single-GPU SequentialRS."""
    import torch.distributed as dist
    import contextual_rs_b200 as crs
    from contextual_rs_b200 import sharded as sh
    from contextual_rs_b200.rs_loop import SequentialRS
    from contextual_rs_b200.synthetic import make_problem
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    desc, ctx = make_problem(K=12, n_c=101, d=2, n=6, seed=4)

    def truth(arm, x):
        return torch.sin(10.0 * torch.cat([torch.tensor([arm / 11.0], dtype=torch.float64), x])).sum()

    single = SequentialRS(crs.B200ModelListGP(desc, device=dev), ctx, randomize_ties=False, kernel_scale=1.0,
                          num_pcs_samples=1024, pcs_seed=3)
    shard = sh.ShardedSequentialRS(crs.B200ModelListGP(desc, device=dev), ctx, kernel_scale=1.0,
                                   num_pcs_samples=1024, pcs_seed=3)
    shard_nccl = sh.ShardedSequentialRS(crs.B200ModelListGP(desc, device=dev), ctx, kernel_scale=1.0,
                                        num_pcs_samples=1024, pcs_seed=3, exchange="nccl")
    shard_min = sh.ShardedSequentialRS(crs.B200ModelListGP(desc, device=dev), ctx, kernel_scale=1.0,
                                       num_pcs_samples=1024, pcs_seed=3, rho=lambda x: x.min(dim=-2).values)
    for it in range(12):
        a, b, c, e = single.step(truth), shard.step(truth), shard_min.step(truth), shard_nccl.step(truth)
        assert (a["next_arm"], a["context_index"]) == (b["next_arm"], b["context_index"]) == (c["next_arm"], c["context_index"]), it
        assert (a["next_arm"], a["context_index"]) == (e["next_arm"], e["context_index"]), it
        assert abs(float(a["pcs"]) - float(e["pcs"])) < 1e-12, it
        assert abs(float(a["pcs"]) - float(b["pcs"])) < 1e-12, it
        cnt = single.ws["counts"].double() / 1024
        assert abs(float(c["pcs"]) - float(cnt.min())) < 1e-12, it
    dist.barrier()


@pytest.mark.skipif(WORLD < 2, reason="needs torchrun with >= 2 ranks")
def test_mailbox_failure_on_one_rank_falls_back_to_nccl_on_all():
    """Peer mailboxes need CUDA IPC between the ranks' processes; where ONE rank cannot set them up, every
    rank must agree to use the NCCL exchange instead (no rank left waiting in a barrier), with equal results."""
    import torch.distributed as dist
    import contextual_rs_b200 as crs
    from contextual_rs_b200 import sharded as sh
    from contextual_rs_b200.synthetic import make_problem
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    desc, ctx = make_problem(K=9, n_c=130, d=2, n=8, seed=2)
    model = crs.B200ModelListGP(desc, device=dev)
    arm, x = crs.gao_modellist(model, ctx)
    c_star = int((ctx == x).all(dim=-1).nonzero()[0])
    if rank == 1:
        os.environ["CRS_DISABLE_IPC"] = "1"
    try:
        dec = sh.ContextShardedDecider(model, ctx)
    finally:
        os.environ.pop("CRS_DISABLE_IPC", None)
    assert dec.exchange == "nccl" and dec.mailbox is None
    rec = dec.decide()
    assert (int(rec[sh.REC_NEXT_ARM]), int(rec[sh.REC_CONTEXT])) == (int(arm), c_star)
    assert 0.0 <= float(dec.pcs_mean(1024, seed=3)) <= 1.0
    dist.barrier()
