"""How far does the CPU oracle follow each of the 200 episodes of the reference's DQN+GCN result files?
Reads only the committed fixtures (ref_DQN_GCN_multi.npz, weights in ref_40_DQN_GCN_seed0.npz); writes
oracle_golden_scan.json: per episode the number of consecutive rows (of the first 60) reproduced to 1e-5 relative in
landmark error and max localisation uncertainty and 0.5 nat in map entropy.  ~5 min on 8 cores (single-threaded).

    python tests/golden/scan_golden.py            # full scan
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from drl_graph_exploration_b200.config import EnvConfig   # noqa: E402
from helpers import RESET_ODOM                            # noqa: E402
from oracle import gnn_ref                                # noqa: E402
from oracle.oracle import OracleEnv                       # noqa: E402

DIFF = {40: 1200, 60: 1600, 80: 2000, 100: 2400}          # test.py:61-70


def load_model():
    g0 = np.load(os.path.join(HERE, "ref_40_DQN_GCN_seed0.npz"))
    model = gnn_ref.GCN()
    model.load_state_dict({k[3:]: torch.tensor(g0[k]) for k in g0.files if k.startswith("sd_")})
    return model.eval()


def follow(model, map_size, seed, gold, max_rows=60, tol=1e-5):
    """test.py:78-150 on the oracle (incl. the 'regenerate a environment' rule).  Returns (rows followed, worst rel err, why it stopped)."""
    cfg = EnvConfig(map_size=map_size)
    while True:
        e = OracleEnv(cfg, seed)
        for _ in range(4):
            e.step(RESET_ODOM)
        if int(np.sum(e.landmarks()["observed"])) >= 1:
            break
        seed += 50                                          # exploration_env.py:416-419
    row, worst = 0, 0.0
    with torch.no_grad():
        while row < max_rows:
            g = e.graph()
            if g["fro_size"] == 0:
                return row, worst, "no frontier"
            data = gnn_ref.Graph(torch.tensor(g["features"], dtype=torch.float32), torch.tensor(g["edge_index"]),
                                 torch.tensor(g["edge_attr"], dtype=torch.float32))
            q = model(data, 0.0).view(-1).numpy()
            a = int(np.argmax(q[-g["fro_size"]:]))
            for act in e.line_plan(*g["frontier_xy"][a]):
                e.step(act)
                m = e.metrics()
                p = e.vmap()["prob"]
                ent = -(p * np.log(p)).sum() + 0.5 * np.log(0.5) * DIFF[map_size]
                gl, ge, gm = gold[row]
                dl, dm = abs(m["landmark_error"] - gl) / gl, abs(m["max_traj_uncertainty"] - gm) / gm
                if dl > tol or dm > tol or abs(ent - ge) > 0.5:
                    return row, worst, f"row {row}: landmark error {dl:.1e}, uncertainty {dm:.1e}, entropy {ent - ge:+.3f}"
                worst = max(worst, dl, dm)
                row += 1
                if row >= max_rows or row >= len(gold):
                    return row, worst, "followed to the end of the fixture"
    return row, worst, "followed to the end of the fixture"


if __name__ == "__main__":
    model = load_model()
    multi = np.load(os.path.join(HERE, "ref_DQN_GCN_multi.npz"))
    out = {}
    for ms in (40, 60, 80, 100):
        for s in range(50):
            rows, worst, why = follow(model, ms, s, multi[f"g_{ms}_{s}"])
            out[f"{ms}_{s}"] = {"rows": rows, "worst_rel_err": worst, "stopped": why}
            print(ms, s, out[f"{ms}_{s}"], flush=True)
    summary = {str(ms): {"rows": sum(out[f"{ms}_{s}"]["rows"] for s in range(50)),
                         "episodes_ge_18_rows": sum(out[f"{ms}_{s}"]["rows"] >= 18 for s in range(50)),
                         "episodes_lt_5_rows": sum(out[f"{ms}_{s}"]["rows"] < 5 for s in range(50))} for ms in (40, 60, 80, 100)}
    json.dump({"summary": summary, "episodes": out}, open(os.path.join(HERE, "oracle_golden_scan.json"), "w"), indent=1)
    print(summary)
