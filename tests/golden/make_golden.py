"""Regenerates the golden fixtures from the reference checkout (/root/reference) -- run in the
build container only; the GPU box never reads /root/reference.

ref_40_DQN_GCN_seed0.npz : first 130 rows of data/test_result/40_DQN_GCN.csv (episode 0 = seed 0:
    landmark error, map entropy, max localisation uncertainty) + the shipped DQN+GCN weights
    (data/torch_weights/DQN_GCN/MyModel.pt) that produced them.
"""
import csv
import os

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

rows = [r for r in csv.DictReader(open(os.path.join(REF, "data/test_result/40_DQN_GCN.csv"))) if r["Step"]][:130]
sd = torch.load(os.path.join(REF, "data/torch_weights/DQN_GCN/MyModel.pt"), map_location="cpu", weights_only=False)
arrs = {"sd_" + k: v.numpy() for k, v in sd.items()}
arrs["landmark_error"] = np.array([float(r["Landmarks error"]) for r in rows])
arrs["entropy"] = np.array([float(r["Map entropy"]) for r in rows])
arrs["max_unc"] = np.array([float(r["Max localization uncertainty"]) for r in rows])
np.savez_compressed(os.path.join(OUT, "ref_40_DQN_GCN_seed0.npz"), **arrs)
print("wrote", os.path.join(OUT, "ref_40_DQN_GCN_seed0.npz"))

# ref_DQN_GCN_multi.npz : first 60 rows (landmark error, map entropy, max localisation uncertainty) of ALL 200 episodes of
# data/test_result/{40,60,80,100}_DQN_GCN.csv (test.py runs seeds lo = 0..49 per map size, each padded to plot_max_step rows).
EPISODES = {ms: list(range(50)) for ms in (40, 60, 80, 100)}      # every episode of the four result files
multi = {}
for ms, seeds in EPISODES.items():
    rows = [r for r in csv.DictReader(open(os.path.join(REF, f"data/test_result/{ms}_DQN_GCN.csv"))) if r["Step"]]
    eps = []
    for r in rows:
        if float(r["Step"]) == 1.0:
            eps.append([])
        eps[-1].append((float(r["Landmarks error"]), float(r["Map entropy"]), float(r["Max localization uncertainty"])))
    assert len(eps) == 50
    for s in seeds:
        multi[f"g_{ms}_{s}"] = np.array(eps[s][:60])
np.savez_compressed(os.path.join(OUT, "ref_DQN_GCN_multi.npz"), **multi)
print("wrote", os.path.join(OUT, "ref_DQN_GCN_multi.npz"))

# ref_state_dict_layouts.json : parameter names and shapes of the six shipped checkpoints (data/torch_weights/*/MyModel.pt):
# the drop-in contract of Networks.py is that `model.load_state_dict(torch.load('MyModel.pt'))` works unchanged.
import json
layouts = {}
for case in ("DQN_GCN", "DQN_GG-NN", "DQN_g-U-Net", "A2C_GCN", "A2C_GG-NN", "A2C_g-U-Net"):
    sdc = torch.load(os.path.join(REF, "data/torch_weights", case, "MyModel.pt"), map_location="cpu", weights_only=False)
    layouts[case] = {k: list(v.shape) for k, v in sdc.items()}
json.dump(layouts, open(os.path.join(OUT, "ref_state_dict_layouts.json"), "w"), indent=1, sort_keys=True)
print("wrote", os.path.join(OUT, "ref_state_dict_layouts.json"))

# ref_other_policies.npz : first 60 rows of ALL 1000 episodes the reference ships for its OTHER policies on the same simulator --
# A2C+GG-NN (data/test_result/{ms}_A2C_GG-NN.csv) and Supervised+GCN / Nearest Frontier / Random / EM ({ms}_Others.csv), 50 seeds x 4
# map sizes each.  Their decisions cannot be recomputed here (stochastic policy, unshipped weights, hand-written rules), but each
# decision of a frontier-driven policy is one of the env's frontier goals: tests/golden/scan_guided.py follows these episodes by
# trying every frontier and keeping the one whose rows match -- a policy-independent pin of simulator + SLAM + map + frontier +
# line-plan arithmetic on trajectories the DQN policy never takes.
OTHERS = {"A2C_GG-NN": ("{ms}_A2C_GG-NN.csv", "A2C+GG-NN"), "Supervised_GCN": ("{ms}_Others.csv", "Supervised+GCN"),
          "Nearest_Frontier": ("{ms}_Others.csv", "Nearest Frontier"), "Random": ("{ms}_Others.csv", "Random"), "EM": ("{ms}_Others.csv", "EM")}
others = {}
for key, (pattern, cat) in OTHERS.items():
    for ms in (40, 60, 80, 100):
        rows = [r for r in csv.DictReader(open(os.path.join(REF, "data/test_result", pattern.format(ms=ms)))) if r["Step"] and r["Category"] == cat]
        eps = []
        for r in rows:
            if float(r["Step"]) == 1.0:
                eps.append([])
            eps[-1].append((float(r["Landmarks error"]), float(r["Map entropy"]), float(r["Max localization uncertainty"])))
        assert len(eps) == 50, (key, ms, len(eps))
        for s in range(50):
            others[f"g_{key}_{ms}_{s}"] = np.array(eps[s][:60])
np.savez_compressed(os.path.join(OUT, "ref_other_policies.npz"), **others)
print("wrote", os.path.join(OUT, "ref_other_policies.npz"))
