"""Generates the committed fixtures under tests/golden/.  Run in the build container only
(reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Fixtures
  mock_tables.json        posterior tables of the reference's MockModel block
                          (test/test_contextual_rs_strategies.py:169-176) with the zeta matrix,
                          sort permutation and decisions hand-derived from the reference formulas
                          (SURVEY.md §8c item 1).
  gao_iid_seed5.json      inputs of the reference's pinned C-OCBA known answer
                          (test/test_contextual_rs_strategies.py:22-62): torch.manual_seed(5),
                          3 arms x 2 contexts x 4 obs, last four Y overwritten; expected (arm 2, ctx 1).
  wsc_branin_seed0.json   real training data written by the reference itself: X/Y of
                          experiments/wsc_experiments/config_b_3_mean/0000_ML_Gao.pt (first 200 rows =
                          initial design, next 200 = the reference's own decision sequence).
  philox_kat.json         Random123 Philox4x32-10 known-answer vectors.
  reference_trace_ml_gao.json  the first 100 decisions the reference recorded in 0000_ML_Gao.pt next to
                          the decisions of the CPU oracle replaying the same run (oracle/trace_replay.py)
                          with raw and with transformed ``model.train_inputs`` — the measured pinning of
                          the oracle on outputs of the reference itself.
  reference_trace_levi_new.json  the same for the recorded LEVI run (0000_LEVI_new.pt, discrete_levi):
                          training data, recorded decisions and the oracle's replay.
"""
import json
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1)
    print("wrote", name)


def reference_trace(iterations=100):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import trace_replay
    d = torch.load(os.path.join(REF, "experiments/wsc_experiments/config_b_3_mean/0000_ML_Gao.pt"), weights_only=False)
    X, Y = d["X"].double(), d["Y"].double()
    torch.manual_seed(0)
    context_map = torch.rand(10, 1, dtype=torch.float64)  # experiments/wsc_experiments/main.py:237-240
    torch.manual_seed(0)
    dec = trace_replay.replay_decisions(X, Y, context_map, 10, 200, iterations, modes=("raw", "transformed"))
    dump("reference_trace_ml_gao.json", {
        "source": "experiments/wsc_experiments/config_b_3_mean/0000_ML_Gao.pt (X[200 + i] = decision of iteration i)",
        "generator": "tests/golden/make_golden.py::reference_trace -> oracle/trace_replay.py (re-fit with oracle/fit_oracle.py)",
        "iterations": iterations,
        "reference": [[int(X[200 + i, 0]), float(X[200 + i, 1])] for i in range(iterations)],
        "oracle_raw_train_inputs": dec["raw"],
        "oracle_transformed_train_inputs": dec["transformed"],
        "reference_best_arms": d["all_maximizers"][:iterations].long().tolist(),   # main.py:459-466
        "reference_pcs_estimates": d["pcs_estimates"][:iterations].tolist(),       # main.py:447-455, 64 samples
        "pcs_weights": [0.03, 0.07, 0.2, 0.1, 0.15, 0.2, 0.02, 0.08, 0.1, 0.05],     # config_b_3_mean/config.json
        "oracle_best_arms": dec["best_arms"],
        "oracle_pcs_exact": dec["pcs_exact"],
        "match_raw": trace_replay.match_counts(dec["raw"], X, 200),
        "match_transformed": trace_replay.match_counts(dec["transformed"], X, 200),
    })


def reference_trace_levi(iterations=100):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import trace_replay
    d = torch.load(os.path.join(REF, "experiments/wsc_experiments/config_b_3_mean/0000_LEVI_new.pt"), weights_only=False)
    X, Y = d["X"].double(), d["Y"].double()
    weights = [0.03, 0.07, 0.2, 0.1, 0.15, 0.2, 0.02, 0.08, 0.1, 0.05]  # config_b_3_mean/config.json
    torch.manual_seed(0)
    context_map = torch.rand(10, 1, dtype=torch.float64)
    dec = trace_replay.replay_decisions(X, Y, context_map, 10, 200, iterations, policy="levi",
                                        weights=torch.tensor(weights, dtype=torch.float64))
    dump("reference_trace_levi_new.json", {
        "source": "experiments/wsc_experiments/config_b_3_mean/0000_LEVI_new.pt (label LEVI_new = discrete_levi with "
                  "_compute_all_EI_reviewer, contextual_rs/levi.py:132-227; X[200 + i] = decision of iteration i)",
        "generator": "tests/golden/make_golden.py::reference_trace_levi -> oracle/trace_replay.py",
        "iterations": iterations, "num_arms": 10, "num_init": 200, "weights": weights,
        "context_map": context_map.reshape(-1).tolist(),
        "X": X[: 200 + iterations].tolist(), "Y": Y[: 200 + iterations].reshape(-1).tolist(),
        "oracle": dec["raw"],
        "match": trace_replay.match_counts(dec["raw"], X, 200),
    })


def mock_tables():
    dump("mock_tables.json", {
        "source": "test/test_contextual_rs_strategies.py:169-176",
        "mean": [[1, 2, 3], [2, 1, 4], [3, 6, 2]],  # rows = contexts, cols = arms
        "variance": [[2, 5, 10], [4, 10, 20], [10, 5, 4]],
        "order": [[2, 1, 0], [2, 0, 1], [1, 0, 2]],
        "zeta": [[1 / 15, 1 / 3], [1 / 6, 0.3], [0.6, 16 / 9]],
        "argmin_flat": 0,
        "best_arms": [2, 2, 1],
        "decisions": [  # train counts at the selected context -> next arm
            {"counts": [2, 2, 2], "next_arm": 2, "next_context": 0},
            {"counts": [1, 1, 10], "next_arm": 1, "next_context": 0},
            {"counts": [0, 0, 0], "next_arm": 1, "next_context": 0},
        ],
    })


def gao_iid():
    out = {"source": "test/test_contextual_rs_strategies.py:22-62", "expected": [2, 1], "cases": {}}
    for name, dtype in (("float32", torch.float32), ("float64", torch.float64)):
        torch.manual_seed(5)
        n_arms, n_ctx, n_obs = 3, 2, 4
        train_Y = torch.rand(n_arms * n_ctx * n_obs, dtype=dtype)
        train_Y[-4:] = torch.tensor([0, 1000, -1000, 1], dtype=dtype)
        obs = train_Y.reshape(n_arms, n_ctx, n_obs)  # rows ordered arm-major, then context, then repeat
        out["cases"][name] = {
            "means": obs.mean(dim=-1).tolist(),
            "vars": obs.std(dim=-1).pow(2).tolist(),
            "num_obs": [[float(n_obs)] * n_ctx] * n_arms,
        }
    dump("gao_iid_seed5.json", out)


def wsc_branin():
    d = torch.load(os.path.join(REF, "experiments/wsc_experiments/config_b_3_mean/0000_ML_Gao.pt"),
                   weights_only=False)
    X, Y = d["X"][:400], d["Y"][:400]
    dump("wsc_branin_seed0.json", {
        "source": "experiments/wsc_experiments/config_b_3_mean/0000_ML_Gao.pt (keys X, Y, all_maximizers)",
        "num_arms": 10, "num_contexts": 10, "num_init": 200,
        "context_map": X[:10, 1:].tolist(),
        "X": X.tolist(), "Y": Y.reshape(-1).tolist(),
        "all_maximizers_first200": d["all_maximizers"][:200].long().tolist(),
    })


def philox():
    dump("philox_kat.json", {
        "source": "Random123 kat_vectors, philox4x32 10 rounds",
        "vectors": [
            {"ctr": ["00000000"] * 4, "key": ["00000000"] * 2,
             "out": ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]},
            {"ctr": ["ffffffff"] * 4, "key": ["ffffffff"] * 2,
             "out": ["408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]},
            {"ctr": ["243f6a88", "85a308d3", "13198a2e", "03707344"], "key": ["a4093822", "299f31d0"],
             "out": ["d16cfe09", "94fdcceb", "5001e420", "24126ea1"]},
        ],
    })


if __name__ == "__main__":
    mock_tables()
    gao_iid()
    wsc_branin()
    philox()
    reference_trace()
    reference_trace_levi()
