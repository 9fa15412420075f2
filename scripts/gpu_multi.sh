#!/bin/bash
# usage: gpu_multi.sh N  — N-rank NCCL tests + bench on one box
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "multi pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/pytest_multi_$N.log | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1000 --warmup 5 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo "bench c2 n=$N rc=$?"
tail -1 gpurun_out/bench_c2_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c2 n=$N', d['scaling'], 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), 'pcs %.3e ms %.2f' % (d['pcs']['value'], d['pcs']['ms_per_estimate']), 'headline', d['headline']['ms_per_step'], d['headline']['decision_ms'], d['headline']['pcs_ms'], d['headline']['next_arm'], d['headline']['next_context'], d['headline']['pcs_mean'])" || tail -5 gpurun_out/bench_c2_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 50 --warmup 3 --workload c5 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "bench c5 n=$N rc=$?"
tail -1 gpurun_out/bench_c5_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c5 n=$N', d['scaling'], 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), d['kernel_ms'])" || tail -5 gpurun_out/bench_c5_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 3 --warmup 1 | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --workload loop --steps 10 --warmup 3 > gpurun_out/bench_loop_n$N.json 2> gpurun_out/bench_loop_n$N.err; echo "bench loop n=$N rc=$?"; tail -1 gpurun_out/bench_loop_n$N.json | cut -c1-1500; tail -3 gpurun_out/bench_loop_n$N.err
