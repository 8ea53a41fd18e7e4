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
