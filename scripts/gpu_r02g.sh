#!/bin/bash
# r02g: arm-sharded partition at N = 2 (tests + bench), context partition beside it.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 400 $TR -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r02g_multi2.log 2>&1; tail -4 gpurun_out/r02g_multi2.log
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --partition arms > gpurun_out/r02g_arms_n2.json 2> gpurun_out/r02g_arms_n2.err; echo "arms n2 rc=$?"; tail -3 gpurun_out/r02g_arms_n2.err
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02g_ctx_n2.json 2> gpurun_out/r02g_ctx_n2.err; echo "ctx n2 rc=$?"
python bench.py --gpus 1 --steps 5 --warmup 3 --partition arms > gpurun_out/r02g_arms_n1.json 2> gpurun_out/r02g_arms_n1.err; echo "arms n1 rc=$?"; tail -3 gpurun_out/r02g_arms_n1.err
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tree_limits" 2>&1 | tail -2
