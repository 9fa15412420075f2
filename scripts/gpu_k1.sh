#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "posterior or kernel_math or wsc_fixture or recorded or levi or min_zeta or full_size" 2>&1 | tail -3
for wl in c2 c4 c5; do
  python bench.py --workload $wl --steps 200 --warmup 5 --no-cpu-baseline --no-pcs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl value %.1f' % d['value'], {k: round(v*1e3,1) for k,v in d['kernel_ms'].items()}, 'grid frac %.3f' % d['roofline']['frac'])"
done
