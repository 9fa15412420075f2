#!/bin/bash
# r02a: new default bench (headline) at N = 1 and N = 2 (mailbox and nccl exchanges), reference arm.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_n1.json 2> gpurun_out/r02a_n1.err; echo "n1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02a_n2.json 2> gpurun_out/r02a_n2.err; echo "n2 rc=$?"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --exchange nccl > gpurun_out/r02a_n2_nccl.json 2> gpurun_out/r02a_n2_nccl.err; echo "n2 nccl rc=$?"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --workload loop > gpurun_out/r02a_loop_n2.json 2> gpurun_out/r02a_loop_n2.err; echo "loop n2 rc=$?"
python bench.py --steps 20 --warmup 5 --workload loop > gpurun_out/r02a_loop_n1.json 2> gpurun_out/r02a_loop_n1.err; echo "loop n1 rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err; echo "ref rc=$?"
tail -3 gpurun_out/r02a_n1.err gpurun_out/r02a_n2.err gpurun_out/r02a_n2_nccl.err gpurun_out/r02a_loop_n2.err gpurun_out/r02a_loop_n1.err
