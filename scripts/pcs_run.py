"""Run the PCS kernel alone on config 3 (256 arms x 8,192 contexts, fp32) — profiling target."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import contextual_rs_b200 as crs
from contextual_rs_b200.generalized_pcs import pcs_counts
from contextual_rs_b200.synthetic import make_config
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
desc, ctx = make_config("c3")
model = crs.B200ModelListGP(desc, device="cuda:0", dtype=torch.float32)
p = model.posterior(ctx.float())
mean, var = p.mean.contiguous(), p.variance.contiguous()
st = torch.zeros(2, dtype=torch.int64, device="cuda:0")
for _ in range(3):
    c = pcs_counts(model, mean, var, S, seed=1, stats=st)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); c = pcs_counts(model, mean, var, S, seed=1, stats=st); e.record(); torch.cuda.synchronize()
d, k = st.cpu().tolist()
print(f"S={S}: {s.elapsed_time(e):.3f} ms, executed {d:.3e} draws ({d/(d+k):.3f}), {d/s.elapsed_time(e)/1e6:.1f} G draws/s, pcs mean {float(c.float().mean())/S:.4f}")
