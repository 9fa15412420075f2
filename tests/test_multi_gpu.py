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

    cdec = sh.ContextShardedDecider(model, ctx)
    rec = cdec.decide(p_mode=1, kernel_scale=1.0)
    assert (int(rec[sh.REC_NEXT_ARM]), int(rec[sh.REC_CONTEXT])) == (int(arm), c_star)
    assert abs(float(cdec.pcs_mean(2048, seed=7)) - single) < 1e-12

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
    shard_min = sh.ShardedSequentialRS(crs.B200ModelListGP(desc, device=dev), ctx, kernel_scale=1.0,
                                       num_pcs_samples=1024, pcs_seed=3, rho=lambda x: x.min(dim=-2).values)
    for it in range(12):
        a, b, c = single.step(truth), shard.step(truth), shard_min.step(truth)
        assert (a["next_arm"], a["context_index"]) == (b["next_arm"], b["context_index"]) == (c["next_arm"], c["context_index"]), it
        assert abs(float(a["pcs"]) - float(b["pcs"])) < 1e-12, it
        cnt = single.ws["counts"].double() / 1024
        assert abs(float(c["pcs"]) - float(cnt.min())) < 1e-12, it
    dist.barrier()
