"""Time the hyper-parameter fit (fit_modellist) on config-2-shaped data: 100 arms x 20 obs, 4-D."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from contextual_rs_b200.fit import fit_modellist
from contextual_rs_b200.synthetic import make_config
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
desc, ctx = make_config(name)
K = len(desc["train_X"])
X = torch.cat([torch.cat([torch.full((x.shape[0], 1), float(k), dtype=torch.float64), x], dim=-1) for k, x in enumerate(desc["train_X"])])
Y = torch.cat(desc["train_Y"])
fit_modellist(X, Y, K)  # warm-up (library load, allocator)
torch.cuda.synchronize(); t0 = time.perf_counter()
m = fit_modellist(X, Y, K)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
info = m.fit_info
print(f"{name}: fit of {K} arms x {desc['train_X'][0].shape[0]} obs: {dt*1e3:.1f} ms wall, {info['evaluations']} objective launches, "
      f"iterations median {int(info['iterations'].median())} max {int(info['iterations'].max())}, objective mean {float(info['objective'].mean()):.4f}")
if len(sys.argv) > 2:  # CPU oracle on a few arms
    from oracle import fit_oracle
    n_ref = int(sys.argv[2])
    t0 = time.perf_counter()
    objs = [fit_oracle.fit_arm(desc["train_X"][k], desc["train_Y"][k])[1] for k in range(n_ref)]
    dt = time.perf_counter() - t0
    print(f"oracle (torch autograd + scipy L-BFGS-B, 1 arm at a time): {dt/n_ref*1e3:.1f} ms per arm; objective diff vs GPU "
          f"max {max(float(info['objective'][k]) - objs[k] for k in range(n_ref)):.2e}")
