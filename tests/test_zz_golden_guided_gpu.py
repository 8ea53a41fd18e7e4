"""The CUDA path against the reference's OTHER result files (A2C+GG-NN, Supervised+GCN, Nearest Frontier, Random, EM): the
decisions of those policies cannot be recomputed, so the frontier taken at every decision is the one the CPU scan found to
reproduce the file (tests/golden/oracle_guided_scan.json, ``choices``); the GPU engine then runs the episode -- device-side
read-out of a one-hot "Q", line plan, steps -- and its per-step metrics are compared DIRECTLY with the file's rows, like in
test_golden_gpu.py.  Policy-independent, and on trajectories the DQN policy never takes (a random policy wanders).
(Written after the last GPU session of round 1, hence sorted behind the suites that have been seen green on a B200.)"""
import json
import os

import numpy as np
import pytest

from test_golden_gpu import GOLD, follow_episode

pytestmark = pytest.mark.gpu


def _cases():
    scan = json.load(open(os.path.join(GOLD, "oracle_guided_scan.json")))["episodes"]
    out = []
    for cat, ms in (("Random", 40), ("Random", 100), ("A2C_GG-NN", 60), ("Nearest_Frontier", 80), ("Supervised_GCN", 40), ("EM", 60)):
        # (only episodes without a nearest-frontier tie at any decision: at a tie the engine's and the oracle's 1e-13 rounding noise
        #  may pick different cells, and the recorded indices refer to the oracle's list)
        key = max((k for k in scan if k.startswith(f"{cat}/{ms}_") and scan[k]["tie_decisions"] == 0), key=lambda k: (scan[k]["rows"], -int(k.split("_")[-1])))
        out.append((cat, ms, int(key.split("_")[-1]), scan[key]["rows"], scan[key]["choices"]))
    return out


@pytest.mark.parametrize("cat,map_size,seed,rows,choices", _cases())
def test_cuda_path_follows_other_policies_episodes(cat, map_size, seed, rows, choices):
    gold = np.load(os.path.join(GOLD, "ref_other_policies.npz"))[f"g_{cat}_{map_size}_{seed}"]
    # the oracle followed `rows` (>= 40) rows; the CUDA path must follow at least half of them (knife-edge cells may end it earlier)
    # (entropy gate of the CPU scan: 0.5 nat -- whole occupancy quanta on knife-edge cells)
    follow_episode(map_size, seed, n_steps=rows, n_min=rows // 2, gold=gold, choices=[int(c) for c in choices.split()], ent_tol=0.5)
