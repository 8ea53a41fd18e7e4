"""Policy-independent pin of the simulator / SLAM / virtual-map / frontier / line-plan path on the reference's OTHER result files.

The reference ships, besides DQN+GCN, 50 test episodes per map size for A2C+GG-NN, Supervised+GCN, Nearest Frontier, Random and
EM (fixture ref_other_policies.npz, first 60 rows each).  Their decisions cannot be recomputed here -- a stochastic policy,
weights that are not shipped, hand-written rules -- but a decision of a frontier-driven policy is always one of the frontier
goals of ``ExplorationEnv.actions_all_goals`` followed along its line plan (test.py:100-143).  So the CPU oracle follows such an
episode WITHOUT the policy: at every decision it tries each frontier (and the cells tied with the vehicle's nearest one, see
``_candidates``) on a clone (RNG streams included) and keeps the one whose
rows reproduce the file to 1e-5 relative (landmark error, max localisation uncertainty) and 0.5 nat (map entropy).  A row that
no frontier reproduces ends the episode; only whole decisions are counted (the rows of a decision that fails half-way are not credited).  EM is not frontier-driven (its planner samples its own goals): it is scanned for the
record and expected to stop early.

    python tests/golden/scan_guided.py [workers]      # writes oracle_guided_scan.json (~15 s on 8 cores)
"""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from drl_graph_exploration_b200.config import EnvConfig   # noqa: E402
from helpers import RESET_ODOM                            # noqa: E402
from oracle.oracle import OracleEnv                       # noqa: E402

DIFF = {40: 1200, 60: 1600, 80: 2000, 100: 2400}          # test.py:61-70
CATEGORIES = ("A2C_GG-NN", "Supervised_GCN", "Nearest_Frontier", "Random", "EM")


def _row(e, map_size):
    m, p = e.metrics(), e.vmap()["prob"]
    return m["landmark_error"], -(p * np.log(p)).sum() + 0.5 * np.log(0.5) * DIFF[map_size], m["max_traj_uncertainty"]


_TIES = ((1e-9, 1e-9), (1e-9, -1e-9), (-1e-9, 1e-9), (-1e-9, -1e-9))


def _candidates(e, g):
    """(index, goal) of every frontier goal of the decision: the oracle's own list (index = position in it), then the cells TIED
    with the vehicle's nearest frontier (index -1).  The first decision of an episode is taken at the integer start pose (the
    four forced reset steps walk a closed square), where several frontier cell centres are exactly equidistant and the
    reference's `dist < min_dist` (exploration_env.py:350-358) is decided by the 1e-13 rounding noise of its pose estimate; the
    ties are enumerated by shifting the vehicle position of that query by (+-1e-9, +-1e-9) (analysis knob of the oracle)."""
    import ctypes
    from oracle.oracle import lib
    out = [(f, tuple(g["frontier_xy"][f])) for f in range(g["fro_size"])]
    L = lib()
    L.orc_set_knife.argtypes = [ctypes.c_double, ctypes.c_double]
    try:
        for dx, dy in _TIES:
            L.orc_set_knife(dx, dy)
            gt = e.graph()
            for f in range(gt["fro_size"]):
                goal = tuple(gt["frontier_xy"][f])
                if all(goal != o[1] for o in out):
                    out.append((-1, goal))
    finally:
        L.orc_set_knife(0.0, 0.0)
    return out


def follow_guided(map_size, seed, gold, max_rows=60, tol=1e-5):
    """Returns (rows followed, worst relative error, the frontier index chosen at every decision, why it stopped).
    ``follow_guided.tie_decisions`` = at how many of the accepted decisions a tied alternative existed at all."""
    cfg = EnvConfig(map_size=map_size)
    while True:
        e = OracleEnv(cfg, seed, record_noise=False)
        for _ in range(4):
            e.step(RESET_ODOM, record_noise=False)
        if int(np.sum(e.landmarks()["observed"])) >= 1:
            break
        seed += 50                                          # exploration_env.py:416-419
    row, worst, decisions = 0, 0.0, []
    follow_guided.tie_decisions = 0
    n_gold = min(max_rows, len(gold))
    while row < n_gold:
        g = e.graph()
        if g["fro_size"] == 0:
            return row, worst, decisions, "no frontier"
        best = None
        cands = _candidates(e, g)
        for f, goal in cands:
            c = e.clone()
            r, w, ok = row, 0.0, True
            for act in c.line_plan(*goal):
                c.step(act, record_noise=False)
                le, ent, mu = _row(c, map_size)
                gl, ge, gm = gold[r]
                dl, dm = abs(le - gl) / gl, abs(mu - gm) / gm
                if dl > tol or dm > tol or abs(ent - ge) > 0.5:
                    ok = False
                    break
                w = max(w, dl, dm)
                r += 1
                if r >= n_gold or c.metrics()["done"]:
                    break
            if ok and r > row and (best is None or w < best[2]):
                best = (c, r, w, f)
        if best is None:
            return row, worst, decisions, f"row {row}: none of the {g['fro_size']} frontiers (nor a tied one) reproduces it"
        e, row, worst = best[0], best[1], max(worst, best[2])
        decisions.append(best[3])
        follow_guided.tie_decisions += any(f < 0 for f, _ in cands)
        if e.metrics()["done"]:
            return row, worst, decisions, "episode done"
    return row, worst, decisions, "followed to the end of the fixture"


def _one(job):
    cat, ms, s = job
    gold = np.load(os.path.join(HERE, "ref_other_policies.npz"))[f"g_{cat}_{ms}_{s}"]
    rows, worst, dec, why = follow_guided(ms, s, gold)
    return f"{cat}/{ms}_{s}", {"rows": rows, "worst_rel_err": worst, "choices": " ".join(str(c) for c in dec), "tie_decisions": follow_guided.tie_decisions,
                                "stopped": why}


if __name__ == "__main__":
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    jobs = [(c, ms, s) for c in CATEGORIES for ms in (40, 60, 80, 100) for s in range(50)]
    with ProcessPoolExecutor(workers) as ex:
        out = dict(ex.map(_one, jobs, chunksize=4))
    summary = {}
    for c in CATEGORIES:
        rs = [v["rows"] for k, v in out.items() if k.startswith(c + "/")]
        summary[c] = {"episodes": len(rs), "rows": int(sum(rs)), "episodes_ge_18_rows": int(sum(r >= 18 for r in rs)),
                      "episodes_lt_5_rows": int(sum(r < 5 for r in rs)), "episodes_to_the_end": int(sum(r >= 60 for r in rs))}
    json.dump({"summary": summary, "episodes": out}, open(os.path.join(HERE, "oracle_guided_scan.json"), "w"), indent=1)
    print(json.dumps(summary, indent=1))
