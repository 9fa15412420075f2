#!/bin/bash
python - <<'PY'
import torch
from contextual_rs_b200.peaks import measure_fma_peak
for kind,name in ((0,"fp64 fma"),(2,"fp64 dmma"),(1,"fp32 fma")):
    for blocks in (148*4, 148*8):
        print(name, blocks, "%.2f TFLOP/s" % measure_fma_peak(kind, "cuda:0", blocks=blocks, iters=2048))
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
