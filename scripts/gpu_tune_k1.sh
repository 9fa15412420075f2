#!/bin/bash
for v in 592 370 296 148 74; do
  echo "CRS_GRID_MIN_CTAS=$v: $(CRS_GRID_MIN_CTAS=$v python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-pcs | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step %.1f us' % (d['ms_per_step']*1e3), {k: round(v*1e3,1) for k,v in d['kernel_ms'].items()})")"
done
for v in 592 296; do
  echo "c3 CRS_GRID_MIN_CTAS=$v: $(CRS_GRID_MIN_CTAS=$v python bench.py --workload c3 --steps 50 --warmup 5 --no-cpu-baseline --no-pcs | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step %.1f us' % (d['ms_per_step']*1e3), {k: round(v*1e3,1) for k,v in d['kernel_ms'].items()})")"
done
