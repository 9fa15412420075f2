#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "not pcs and not fit and not mll" 2>&1 | tail -3
for wl in c2 c3 c5; do
  timeout 300 python bench.py --workload $wl --steps 1000 --warmup 10 --no-pcs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl value %.1f e2e %.1f (%.1f us) h2d %d' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']*1e3, d['e2e']['h2d_bytes_per_step']), {k: round(v*1e3,2) for k,v in d['kernel_ms'].items()})"
done
