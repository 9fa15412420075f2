"""One warm-up + a few headline steps (config 4 decision + 2^16-sample PCS) on cuda:0 — the command the
ncu captures of round 2 profile.  usage: headline_run.py [steps] [config] [dtype]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contextual_rs_b200 as crs
from contextual_rs_b200.sharded import ContextShardedDecider
from contextual_rs_b200.synthetic import make_config
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
name = sys.argv[2] if len(sys.argv) > 2 else "c4"
dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.float64
desc, ctx = make_config(name)
K = len(desc["train_X"])
model = crs.B200ModelListGP(desc, device="cuda:0", dtype=dt)
dec = ContextShardedDecider(model, ctx)
for i in range(1 + steps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rec = dec.decide(prepare_arms=(0, K))
    pcs = dec.pcs_mean(1 << 16, seed=0, offset=i)
    e.record(); torch.cuda.synchronize()
    print(f"step {i}: {s.elapsed_time(e):.3f} ms  next_arm {int(rec[5])} ctx {int(rec[2])} pcs {float(pcs):.6f}")
