#!/bin/bash
python - <<'PY'
import torch
from contextual_rs_b200.peaks import measure_fma_peak
for kind,name in ((0,"fp64 fma"),(3,"fp64 add (as 2 flop/op equiv)"),(4,"fp64 mul (as 2 flop/op equiv)"),(2,"fp64 dmma")):
    print(name, "%.2f TFLOP/s-equivalent" % measure_fma_peak(kind, "cuda:0", blocks=1184, iters=2048))
PY
