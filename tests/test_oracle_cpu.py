"""CPU tests of the oracle itself (no GPU): structured vs dense SLAM solve, golden CSV tracking,
basic invariants.  These pin the checker before it is used to check the CUDA path."""
import csv
import json
import math
import os
import sys

import numpy as np
import pytest
import torch

from drl_graph_exploration_b200.config import EnvConfig, start_pose_for_seed
from helpers import RESET_ODOM, choose_actions
from oracle import gnn_ref
from oracle.oracle import OracleEnv, virtual_map_rebuild

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_start_pose_matches_reference_rng():
    # pyss2d.py:88-95 with lo = 0, S = 40 (legacy NumPy MT19937); values recorded from this container
    assert start_pose_for_seed(0, 40) == (17.0, -5.0, math.radians(152.0))


def test_structured_solver_matches_dense():
    cfg = EnvConfig(map_size=20, num_landmarks=30)
    a, b = OracleEnv(cfg, 3), OracleEnv(cfg, 3, dense=True)
    rng = np.random.default_rng(1)
    acts = [RESET_ODOM] * 4
    for _ in range(6):
        acts += choose_actions(a, rng)
        for od in acts:
            a.step(od); b.step(od)
        acts = []
        pa, pb = a.poses(), b.poses()
        assert np.allclose(pa["est"], pb["est"], rtol=1e-9, atol=1e-10)
        assert np.allclose(pa["cov"], pb["cov"], rtol=1e-7, atol=1e-13)
        la, lb = a.landmarks(), b.landmarks()
        assert np.array_equal(la["observed"], lb["observed"])
        assert np.allclose(la["est"], lb["est"], rtol=1e-9, atol=1e-10)
        assert np.allclose(la["cov"], lb["cov"], rtol=1e-7, atol=1e-13)
        assert np.array_equal(a.vmap()["seen"], b.vmap()["seen"])
    assert a.sizes()["update_count"] >= 20


def test_marginals_are_spd_and_map_values_are_in_closed_set():
    cfg = EnvConfig(map_size=40)
    e = OracleEnv(cfg, 7)
    for _ in range(4):
        e.step(RESET_ODOM)
    P = e.poses()
    for c in P["cov"]:
        assert np.allclose(c, c.T) and np.all(np.linalg.eigvalsh(c) > 0)
    vm = e.vmap()
    o = 3.0 / 7.0
    lmax = math.exp(0.95) / (1 + math.exp(0.95))   # MAX_LOGODDS quirk q7
    allowed = np.array([0.5, 0.3, o ** 2 / (1 + o ** 2), o ** 3 / (1 + o ** 3), 0.05, math.exp(lmax) / (1 + math.exp(lmax))])
    assert np.all(np.min(np.abs(vm["prob"][..., None] - allowed), axis=-1) < 1e-12)   # q8
    ev = np.linalg.eigvalsh(vm["info"])
    assert np.all(ev > 0)


def test_standalone_virtual_map_equals_env_path():
    cfg = EnvConfig(map_size=40)
    e = OracleEnv(cfg, 2)
    for _ in range(4):
        e.step(RESET_ODOM)
    for od in choose_actions(e, np.random.default_rng(0)):
        e.step(od)
    P, L = e.poses(), e.landmarks()
    prob, info, seen = virtual_map_rebuild(cfg, P["est"], P["info"].reshape(-1, 9), L["est"][L["observed"].astype(bool)])
    vm = e.vmap()
    assert np.array_equal(prob, vm["prob"]) and np.array_equal(seen, vm["seen"]) and np.array_equal(info, vm["info"])


def test_tracks_reference_golden_csv():
    """The only known-answer data of the reference for this path: data/test_result/40_DQN_GCN.csv
    (test.py, seed 0, DQN+GCN with the shipped weights, prob=0).  Fixtures under tests/golden were
    extracted by tests/golden/make_golden.py.  Landmark error and max localisation uncertainty are
    tracked to 1e-6 over the first 50 steps; the map entropy may differ by whole occupancy
    quanta on the knife-edge cells described in DESIGN.md, so it is checked to 0.25 nat."""
    gold = np.load(os.path.join(GOLD, "ref_40_DQN_GCN_seed0.npz"))
    sd = {k[3:]: torch.tensor(gold[k]) for k in gold.files if k.startswith("sd_")}
    model = gnn_ref.GCN()
    model.load_state_dict(sd)
    model.eval()
    cfg = EnvConfig(map_size=40)
    e = OracleEnv(cfg, 0)
    for _ in range(4):
        e.step(RESET_ODOM)
    step = 0
    with torch.no_grad():
        while step < 50:
            g = e.graph()
            data = gnn_ref.Graph(torch.tensor(g["features"], dtype=torch.float32), torch.tensor(g["edge_index"]),
                                 torch.tensor(g["edge_attr"], dtype=torch.float32))
            q = model(data, 0.0).view(-1).numpy()
            a = int(np.argmax(q[-g["fro_size"]:]))
            for act in e.line_plan(*g["frontier_xy"][a]):
                e.step(act)
                m = e.metrics()
                p = e.vmap()["prob"]
                ent = -(p * np.log(p)).sum() + 0.5 * np.log(0.5) * 1200   # test.py:61-76 map_entropy, S = 40
                assert abs(m["landmark_error"] - gold["landmark_error"][step]) < 1e-6 * gold["landmark_error"][step] + 1e-7, step
                assert abs(m["max_traj_uncertainty"] - gold["max_unc"][step]) < 1e-6 * gold["max_unc"][step], step
                assert abs(ent - gold["entropy"][step]) < 0.25, step
                step += 1
                if step >= 50:
                    break


# (map size, seed, rows that must be followed): a spread over the four result files of the episodes the oracle follows for
# >= 20 rows (tests/golden/oracle_golden_scan.json holds the full scan of all 200 episodes: 6218 rows followed, 144 episodes
# for >= 18 rows; regenerate with tests/golden/scan_golden.py).  Several go through the 'regenerate a environment' rule.
TRACKED = [(40, 1, 40), (40, 3, 30), (40, 8, 25), (40, 10, 35), (40, 12, 45), (40, 14, 40), (40, 24, 45), (40, 42, 45), (40, 46, 55),
           (60, 2, 30), (60, 4, 55), (60, 7, 55), (60, 17, 55), (60, 21, 45), (60, 38, 55), (60, 40, 50),
           (80, 1, 45), (80, 2, 45), (80, 6, 55), (80, 16, 55), (80, 24, 50), (80, 39, 45), (80, 49, 40),
           (100, 0, 45), (100, 6, 55), (100, 16, 55), (100, 26, 50), (100, 28, 50), (100, 37, 55), (100, 48, 55),
           # followed only because the nearest-frontier ties of the first decision are enumerated (scan_golden._policy_goals): without
           # that these stop after 1 / 1 / 5 / 15 rows
           (40, 5, 55), (60, 11, 30), (80, 45, 40), (100, 8, 55)]


@pytest.fixture(scope="module")
def golden_scan():
    sys.path.insert(0, GOLD)
    import scan_golden
    return scan_golden, scan_golden.load_model(), np.load(os.path.join(GOLD, "ref_DQN_GCN_multi.npz"))


@pytest.mark.parametrize("map_size,seed,n_rows", TRACKED)
def test_tracks_more_reference_episodes(golden_scan, map_size, seed, n_rows):
    """Further known-answer data of the reference: other seeds of 40_DQN_GCN.csv and the 60/80/100 m maps
    (tests/golden/ref_DQN_GCN_multi.npz, extracted by make_golden.py), followed end to end with the shipped weights."""
    scan, model, multi = golden_scan
    rows, worst, why = scan.follow(model, map_size, seed, multi[f"g_{map_size}_{seed}"], max_rows=n_rows)
    assert rows >= n_rows and worst < 1e-5, (rows, worst, why)


def test_reference_scan_order_is_the_old_libstdcxx_hashtable_order():
    """The order in which Simulator2D::measure visits the landmarks (= iteration order of the reference's unordered_map)
    decides which landmark gets which noise draw.  The oracle restates the GCC 5-7 hashtable; the host's GCC 13
    std::unordered_map gives 7,6,..,0 for 8 keys, which contradicts the golden files (oracle_golden_scan.json: with the
    restated order 5377 rows of the 200 episodes are followed, with the host's order 4007)."""
    e = OracleEnv(EnvConfig(map_size=40), 3)
    assert e.landmarks()["scan_id"].tolist() == [7, 6, 5, 4, 0, 1, 2, 3]
    e = OracleEnv(EnvConfig(map_size=60), 3)
    assert e.landmarks()["scan_id"].tolist() == [17, 16, 15, 14, 13, 12, 11, 10, 3, 2, 1, 0, 4, 5, 6, 7, 8, 9]
    scan = json.load(open(os.path.join(GOLD, "oracle_golden_scan.json")))
    assert sum(v["rows"] for v in scan["summary"].values()) >= 5300
    assert sum(v["episodes_ge_18_rows"] for v in scan["summary"].values()) >= 120


# ---- the reference's OTHER result files, followed without their policies (tests/golden/scan_guided.py) ----------------------
def _guided_cases():
    import json
    scan = json.load(open(os.path.join(GOLD, "oracle_guided_scan.json")))["episodes"]
    cases = []
    for cat in ("A2C_GG-NN", "Supervised_GCN", "Nearest_Frontier", "Random", "EM"):
        for ms in (40, 60, 80, 100):            # per category and map size: the episode the scan followed furthest
            key = max((k for k in scan if k.startswith(f"{cat}/{ms}_")), key=lambda k: (scan[k]["rows"], -int(k.split("_")[-1])))
            cases.append((cat, ms, int(key.split("_")[-1]), scan[key]["rows"], scan[key]["choices"]))
    return cases


@pytest.mark.parametrize("cat,map_size,seed,rows,choices", _guided_cases())
def test_follows_the_other_policies_episodes_without_their_policies(cat, map_size, seed, rows, choices):
    """Simulator + SLAM + virtual map + frontier detection + line planner against rows of the reference's A2C+GG-NN /
    Supervised+GCN / Nearest Frontier / Random / EM result files: at every decision the frontier whose rows reproduce the file
    is taken (the policies themselves cannot be recomputed).  20 of the 1000 scanned episodes (28 132 rows in all,
    tests/golden/oracle_guided_scan.json), each for >= 40 rows to <= 1e-5 relative."""
    sys.path.insert(0, GOLD)
    import scan_guided
    gold = np.load(os.path.join(GOLD, "ref_other_policies.npz"))[f"g_{cat}_{map_size}_{seed}"]
    got_rows, worst, got_choices, why = scan_guided.follow_guided(map_size, seed, gold)
    assert got_rows == rows >= 40 and worst <= 1e-5, (got_rows, rows, worst, why)
    assert " ".join(str(c) for c in got_choices) == choices
