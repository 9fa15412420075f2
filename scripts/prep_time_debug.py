"""Phase timing of gp_prepare_small_kernel with clock64 (CTA 0): build the instrumented library first —
  cd contextual_rs_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++20 -Xcompiler -fPIC \
     --expt-relaxed-constexpr -DCRS_PREP_TIMING -c gp_prepare.cu -o /tmp/gp_prepare_timing.o && \
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libcrs_b200_timing.so $(ls *.o | grep -v gp_prepare.o) \
     /tmp/gp_prepare_timing.o -cudart static
then run this script on the GPU box (CRS_B200_LIB selects the library).  Round-2 numbers (config 2, n = 20, cycles):
load 1.3k | bounds 0.56k | transforms 2.1k | K^ 4.4k | Cholesky 11.1k | alpha + L^-1 12.3k -> 3.7k | write 1.4k."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CRS_B200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "contextual_rs_b200", "libcrs_b200_timing.so"))
import torch
import contextual_rs_b200 as crs
from contextual_rs_b200.synthetic import make_config
desc, ctx = make_config("c2")
model = crs.B200ModelListGP(desc, device="cuda:0")
for i in range(3):
    model.prepare(0, model.num_arms, relearn=False)
    torch.cuda.synchronize()
