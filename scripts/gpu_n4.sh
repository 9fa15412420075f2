#!/bin/bash
N=4
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1000 --warmup 5 > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo "bench c2 n=$N rc=$?"
tail -1 gpurun_out/bench_c2_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c2 n=$N', d['scaling'], 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), 'pcs %.3e ms %.2f frac %.3f' % (d['pcs']['value'], d['pcs']['ms_per_estimate'], d['pcs']['roofline']['frac']), 'headline ms %.2f dec %.2f pcs %.2f' % (d['headline']['ms_per_step'], d['headline']['decision_ms'], d['headline']['pcs_ms']), d['headline']['next_arm'], d['headline']['next_context'], d['headline']['pcs_mean'])" || tail -5 gpurun_out/bench_c2_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 3 --warmup 1 | cut -c1-200
