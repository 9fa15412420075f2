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
