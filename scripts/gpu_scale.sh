#!/bin/bash
# usage: gpu_scale.sh N — the driver's scaling run for one N: default bench (c2 replicas + sharded pcs + headline) and the loop
N=${1:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1000 --warmup 5 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo "bench c2 n=$N rc=$?"
tail -1 gpurun_out/bench_c2_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c2 n=$N', d['scaling'], 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), 'pcs %.3e ms %.2f' % (d['pcs']['value'], d['pcs']['ms_per_estimate']), 'headline ms %.2f dec %.2f pcs %.2f' % (d['headline']['ms_per_step'], d['headline']['decision_ms'], d['headline']['pcs_ms']), d['headline']['next_arm'], d['headline']['next_context'], d['headline']['pcs_mean'])" || tail -5 gpurun_out/bench_c2_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --workload loop --steps 10 --warmup 3 > gpurun_out/bench_loop_n$N.json 2> gpurun_out/bench_loop_n$N.err; echo "bench loop n=$N rc=$?"
tail -1 gpurun_out/bench_loop_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('loop n=$N value %.2f it/s ms %.2f' % (d['value'], d['ms_per_step']), d['breakdown_ms'], d['pcs_last'])" || tail -5 gpurun_out/bench_loop_n$N.err
