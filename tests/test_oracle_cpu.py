"""CPU tests of the oracle itself (no GPU): structured vs dense SLAM solve, golden CSV tracking,
basic invariants.  These pin the checker before it is used to check the CUDA path."""
import csv
import math
import os

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


def _track_episode(model, map_size, seed, gold, n_steps):
    """test.py:78-150 on the oracle: reset (with the reference's 'regenerate a environment' rule,
    exploration_env.py:416-419: no landmark seen after the 4 forced steps -> env_index += 50), then DQN+GCN decisions."""
    cfg = EnvConfig(map_size=map_size)
    while True:
        e = OracleEnv(cfg, seed)
        for _ in range(4):
            e.step(RESET_ODOM)
        if int(np.sum(e.landmarks()["observed"])) >= 1:
            break
        seed += 50
    diff = {40: 1200, 60: 1600, 80: 2000, 100: 2400}[map_size]        # test.py:61-70
    step, worst = 0, 0.0
    with torch.no_grad():
        while step < n_steps:
            g = e.graph()
            data = gnn_ref.Graph(torch.tensor(g["features"], dtype=torch.float32), torch.tensor(g["edge_index"]),
                                 torch.tensor(g["edge_attr"], dtype=torch.float32))
            q = model(data, 0.0).view(-1).numpy()
            a = int(np.argmax(q[-g["fro_size"]:]))
            for act in e.line_plan(*g["frontier_xy"][a]):
                e.step(act)
                m = e.metrics()
                p = e.vmap()["prob"]
                ent = -(p * np.log(p)).sum() + 0.5 * np.log(0.5) * diff
                gl, ge, gm = gold[step]
                dl, dm = abs(m["landmark_error"] - gl) / gl, abs(m["max_traj_uncertainty"] - gm) / gm
                assert dl < 1e-5 and dm < 1e-5 and abs(ent - ge) < 0.3, (map_size, seed, step, dl, dm, ent - ge)
                worst = max(worst, dl, dm)
                step += 1
                if step >= n_steps:
                    break
    return worst


# (map size, seed, steps tracked): every episode of the reference's DQN+GCN result files examined so far that the oracle
# follows for >= 20 steps (17 of 20; 40/5, 40/10 and 60/0 leave the reference's noise stream within 4 steps -- DESIGN.md section 5).
# Seeds 7, 8, 9, 11 (S = 40) and 0 (S = 80, 100) go through the 'regenerate' rule.
TRACKED = [(40, 1, 40), (40, 2, 22), (40, 3, 30), (40, 4, 30), (40, 6, 20), (40, 8, 25), (40, 9, 18), (40, 11, 20), (60, 1, 20), (60, 2, 30), (60, 3, 20),
           (80, 0, 18), (80, 1, 24), (80, 2, 40), (100, 0, 18), (100, 1, 20)]


@pytest.mark.parametrize("map_size,seed,n_steps", TRACKED)
def test_tracks_more_reference_episodes(map_size, seed, n_steps):
    """Further known-answer data of the reference: other seeds of 40_DQN_GCN.csv and the 60/80/100 m maps
    (tests/golden/ref_DQN_GCN_multi.npz, extracted by make_golden.py), followed end to end with the shipped weights."""
    gold = np.load(os.path.join(GOLD, "ref_40_DQN_GCN_seed0.npz"))
    sd = {k[3:]: torch.tensor(gold[k]) for k in gold.files if k.startswith("sd_")}
    model = gnn_ref.GCN()
    model.load_state_dict(sd)
    model.eval()
    multi = np.load(os.path.join(GOLD, "ref_DQN_GCN_multi.npz"))
    worst = _track_episode(model, map_size, seed, multi[f"g_{map_size}_{seed}"], n_steps)
    assert worst < 1e-5
