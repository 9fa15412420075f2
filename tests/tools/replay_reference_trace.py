"""Replay the reference's own recorded GP-C-OCBA run through the CPU oracle.

`experiments/wsc_experiments/config_b_3_mean/0000_ML_Gao.pt` in the reference tree holds the
training set and the decision sequence of one real run of `main(label="ML_Gao")`
(`experiments/wsc_experiments/main.py:165-505`; 10 arms x 10 contexts, Branin, 2 initial samples per
pair, re-fit every 10 iterations, `gao_modellist(..., kernel_scale=1.0)`): row 200 + i of X is the
(arm, context) the reference chose at iteration i.  The fitted hyper-parameters were not stored, so
the replay re-fits with the oracle's restatement of BoTorch's fit (`oracle/fit_oracle.py`) on the
recorded data and asks the oracle's `gao_modellist_ref` for the decision.
usage: replay_reference_trace.py [iterations] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import fit_oracle, gp_oracle, ocba_oracle

PATH = "/root/reference/experiments/wsc_experiments/config_b_3_mean/0000_ML_Gao.pt"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
out_path = sys.argv[2] if len(sys.argv) > 2 else None
d = torch.load(PATH, weights_only=False)
X, Y = d["X"].double(), d["Y"].double()
K, n_c, n0, fit_frequency = 10, 10, 200, 10
torch.manual_seed(0)
context_map = torch.rand(n_c, 1, dtype=torch.float64)  # main.py:237-240
hp = [None] * K          # fitted hyper-parameters + frozen transforms per arm
n_at_fit = [0] * K       # experiment_utils.py:66: number of inputs at the arm's last fit
match_arm = match_ctx = match_both = 0
rows = []
t0 = time.time()
torch.manual_seed(0)
for i in range(N):
    Xi, Yi = X[: n0 + i], Y[: n0 + i]
    if i % fit_frequency == 0:  # main.py:346-364 -> fit_modellist_with_reuse (experiment_utils.py:68-121)
        for k in range(K):
            mask = Xi[:, 0] == k
            if int(mask.sum()) != n_at_fit[k]:
                hp[k], _ = fit_oracle.fit_arm(Xi[mask][:, 1:], Yi[mask].reshape(-1))
                n_at_fit[k] = int(mask.sum())
    keys = ("lengthscale", "outputscale", "noise", "mean_const", "norm_mins", "norm_ranges", "y_mean", "y_std")
    desc = {"train_X": [Xi[Xi[:, 0] == k][:, 1:] for k in range(K)], "train_Y": [Yi[Xi[:, 0] == k] for k in range(K)],
            "kernel": "matern52", "normalize": True, "standardize": True}
    desc.update({key: torch.stack([hp[k][key] for k in range(K)]) for key in keys})
    model = gp_oracle.model_from_description(desc)  # conditioning = same hyper-parameters, frozen transforms
    if os.environ.get("RAW_INPUTS", "0") == "1":  # model.train_inputs = raw inputs (older BoTorch: transforms applied in forward only)
        class Raw:
            models = model.models
            posterior = staticmethod(model.posterior)
            train_inputs = [(m.raw_X,) for m in model.models]
        arm, ctx = ocba_oracle.gao_modellist_ref(Raw, context_map, True, kernel_scale=1.0)
    else:
        arm, ctx = ocba_oracle.gao_modellist_ref(model, context_map, True, kernel_scale=1.0)
    ref_arm, ref_ctx = int(X[n0 + i, 0]), float(X[n0 + i, 1])
    a, c = int(arm) == ref_arm, abs(float(ctx) - ref_ctx) < 1e-12
    match_arm += a; match_ctx += c; match_both += a and c
    rows.append({"iteration": i, "reference": [ref_arm, ref_ctx], "oracle": [int(arm), float(ctx)]})
    if (i + 1) % 20 == 0:
        print(f"iteration {i + 1}: arm {match_arm}/{i + 1} context {match_ctx}/{i + 1} both {match_both}/{i + 1} ({time.time() - t0:.0f} s)", flush=True)
print(f"reproduced {match_both}/{N} decisions (arm {match_arm}, context {match_ctx})")
if out_path:
    json.dump({"source": PATH, "iterations": N, "match_both": match_both, "match_arm": match_arm, "match_context": match_ctx, "rows": rows}, open(out_path, "w"))
