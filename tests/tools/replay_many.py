"""Replay the first iterations of several recorded ML_Gao runs of the reference through the oracle."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import trace_replay
ROOT = "/root/reference/experiments/wsc_experiments"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
runs = [("config_b_3_mean", 1), ("config_b_3_mean", 2), ("config_b_3_worst", 0), ("config_g_4_mean", 0), ("config_h_4_mean", 0), ("config_c_5_mean", 0)]
out = []
for cfg, seed in runs:
    c = json.load(open(f"{ROOT}/{cfg}/config.json"))
    d = torch.load(f"{ROOT}/{cfg}/{seed:04d}_ML_Gao.pt", weights_only=False)
    X, Y = d["X"].double(), d["Y"].double()
    K, n_c, dim = c["num_arms"], c["num_contexts"], c.get("context_dim", 1)
    n0 = K * n_c * 2 if c["use_full_train"] else K * (2 * dim + 2)
    torch.manual_seed(0)
    cm = torch.rand(n_c, dim, dtype=torch.float64)
    t0 = time.time()
    torch.manual_seed(seed)
    dec = trace_replay.replay_decisions(X, Y, cm, K, n0, N, fit_frequency=c["fit_frequency"])
    # decisions compare (arm, first context coordinate)
    m = trace_replay.match_counts(dec["raw"], X, n0)
    print(cfg, seed, f"K={K} n_c={n_c} d={dim} init={n0}", m, f"{time.time() - t0:.0f} s", flush=True)
    out.append({"config": cfg, "seed": seed, "arms": K, "contexts": n_c, "dim": dim, **m})
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
