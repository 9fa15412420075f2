#!/bin/bash
# usage: gpu_multi.sh N  — N-rank NCCL tests + bench on one box
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/pytest_multi_$N.log 2>&1; echo "multi pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/pytest_multi_$N.log | tail -4
for wl in c2 c5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 --workload $wl > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err; echo "bench $wl n=$N rc=$?"
  tail -1 gpurun_out/bench_${wl}_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl n=$N value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), d['kernel_ms'], ('pcs %.3e' % d['pcs']['value']) if 'pcs' in d else '')" || tail -5 gpurun_out/bench_${wl}_n$N.err
done
