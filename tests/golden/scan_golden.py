"""How far does the CPU oracle follow each of the 200 episodes of the reference's DQN+GCN result files?
Reads only the committed fixtures (ref_DQN_GCN_multi.npz, weights in ref_40_DQN_GCN_seed0.npz); writes
oracle_golden_scan.json: per episode the number of consecutive rows (of the first 60) reproduced to 1e-5 relative in
landmark error and max localisation uncertainty and 0.5 nat in map entropy.  ~40 s on 8 cores.

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


_TIES = ((1e-9, 1e-9), (1e-9, -1e-9), (-1e-9, 1e-9), (-1e-9, -1e-9))


def _policy_goal(model, g):
    data = gnn_ref.Graph(torch.tensor(g["features"], dtype=torch.float32), torch.tensor(g["edge_index"]),
                         torch.tensor(g["edge_attr"], dtype=torch.float32))
    q = model(data, 0.0).view(-1).numpy()
    return tuple(g["frontier_xy"][int(np.argmax(q[-g["fro_size"]:]))])


def _policy_goals(model, e, g):
    """The policy's goal on the oracle's own graph, then its goals on the graphs of the other resolutions of a nearest-frontier
    tie.  (The first decision of an episode is taken at the integer start pose -- the four forced reset steps walk a closed
    square --, where several frontier cell centres are exactly equidistant and the reference's `dist < min_dist`,
    exploration_env.py:350-358, is decided by the 1e-13 rounding noise of its pose estimate: which cell is the vehicle's
    frontier node, hence the graph the policy sees, is noise in the reference.  The ties are enumerated by shifting the vehicle
    position of that one query by (+-1e-9, +-1e-9), an analysis knob of the oracle that is 0 everywhere else.)"""
    import ctypes
    from oracle.oracle import lib
    goals = [_policy_goal(model, g)]
    L = lib()
    L.orc_set_knife.argtypes = [ctypes.c_double, ctypes.c_double]
    try:
        for dx, dy in _TIES:
            L.orc_set_knife(dx, dy)
            gt = e.graph()
            if gt["fro_size"] == g["fro_size"] and np.array_equal(gt["frontier_xy"][:gt["fro_size"]], g["frontier_xy"][:g["fro_size"]]):
                continue                                    # no tie at this decision
            goal = _policy_goal(model, gt)
            if goal not in goals:
                goals.append(goal)
    finally:
        L.orc_set_knife(0.0, 0.0)
    return goals


def follow(model, map_size, seed, gold, max_rows=60, tol=1e-5, ties=True):
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
    n_gold = min(max_rows, len(gold))

    def run_plan(env, goal, row, worst):
        """-> (rows after the plan, worst, None) or (rows reached, worst, reason) when a row disagrees."""
        for act in env.line_plan(*goal):
            env.step(act)
            m = env.metrics()
            p = env.vmap()["prob"]
            ent = -(p * np.log(p)).sum() + 0.5 * np.log(0.5) * DIFF[map_size]
            gl, ge, gm = gold[row]
            dl, dm = abs(m["landmark_error"] - gl) / gl, abs(m["max_traj_uncertainty"] - gm) / gm
            if dl > tol or dm > tol or abs(ent - ge) > 0.5:
                return row, worst, f"row {row}: landmark error {dl:.1e}, uncertainty {dm:.1e}, entropy {ent - ge:+.3f}"
            worst = max(worst, dl, dm)
            row += 1
            if row >= n_gold:
                break
        return row, worst, None

    with torch.no_grad():
        while row < n_gold:
            g = e.graph()
            if g["fro_size"] == 0:
                return row, worst, "no frontier"
            goals = _policy_goals(model, e, g) if ties else [_policy_goal(model, g)]
            first = None
            for goal in goals:
                c = e.clone()
                r, w, why = run_plan(c, goal, row, worst)
                if first is None:
                    first = (r, w, why)
                if why is None:
                    e, row, worst = c, r, w
                    break
            else:                                           # no resolution reproduces the whole plan: the oracle's own one counts
                return first
    return row, worst, "followed to the end of the fixture"


def _one(job):
    ms, s = job
    torch.set_num_threads(1)
    gold = np.load(os.path.join(HERE, "ref_DQN_GCN_multi.npz"))[f"g_{ms}_{s}"]
    rows, worst, why = follow(_one.model, ms, s, gold)
    rows0, _, _ = follow(_one.model, ms, s, gold, ties=False)
    return f"{ms}_{s}", {"rows": rows, "rows_without_tie_resolution": rows0, "worst_rel_err": worst, "stopped": why}


def _init():
    _one.model = load_model()


if __name__ == "__main__":
    from concurrent.futures import ProcessPoolExecutor
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    jobs = [(ms, s) for ms in (40, 60, 80, 100) for s in range(50)]
    with ProcessPoolExecutor(workers, initializer=_init) as ex:
        out = dict(ex.map(_one, jobs, chunksize=2))
    summary = {str(ms): {"rows": sum(out[f"{ms}_{s}"]["rows"] for s in range(50)),
                         "rows_without_tie_resolution": sum(out[f"{ms}_{s}"]["rows_without_tie_resolution"] for s in range(50)),
                         "episodes_ge_18_rows": sum(out[f"{ms}_{s}"]["rows"] >= 18 for s in range(50)),
                         "episodes_lt_5_rows": sum(out[f"{ms}_{s}"]["rows"] < 5 for s in range(50))} for ms in (40, 60, 80, 100)}
    json.dump({"summary": summary, "episodes": out}, open(os.path.join(HERE, "oracle_golden_scan.json"), "w"), indent=1)
    print(summary, "total", sum(v["rows"] for v in summary.values()), "without tie resolution", sum(v["rows_without_tie_resolution"] for v in summary.values()))
