#!/bin/bash
for m in 592 400 296 200 104; do
  CRS_GRID_MIN_CTAS=$m python bench.py --steps 500 --warmup 10 --no-pcs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('min_ctas $m value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), {k: round(v*1e3,2) for k,v in d['kernel_ms'].items()})"
done
