"""Edge cases of the CUDA path: wide landmark borders (the k_slam code paths for more than 32 / 44 landmarks), an env
that sees no landmark at all, pose-capacity overflow, empty decision rounds, degenerate graphs in the GNN kernels."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import RESET_ODOM, choose_actions, make_oracles, world_arrays
from drl_graph_exploration_b200.config import EnvConfig
from test_env_parity_gpu import compare_state

pytestmark = pytest.mark.gpu


def _engine_and_oracles(cfg, seeds, max_poses=64):
    from drl_graph_exploration_b200.engine import Engine
    oracles = make_oracles(cfg, seeds)
    start, lm, scan, noise0 = world_arrays(oracles)
    eng = Engine(cfg, len(oracles), max_poses=max_poses)
    t = lambda a: torch.as_tensor(a, device=eng.device)
    eng.reset(seeds=t(np.asarray(list(seeds), dtype=np.int64)), start=t(start), landmarks=t(lm), scan=t(scan), noise=t(noise0))
    return eng, oracles, t


@pytest.mark.parametrize("n_lm", [40, 48, 64])
def test_wide_landmark_border_matches_oracle(n_lm):
    """2*Lt = 80 / 96 / 128 border columns: the two-tiles-per-thread and the global-memory Schur complement and the
    shared-memory Gauss-Jordan of k_slam (the benchmark configs stay in the <= 60-column fast paths)."""
    cfg = EnvConfig(map_size=20, num_landmarks=n_lm)
    eng, oracles, t = _engine_and_oracles(cfg, range(3), max_poses=48)
    B = len(oracles)

    def do_step(odoms, tag):
        noise = np.stack([o.step(od) for o, od in zip(oracles, odoms)])
        eng.step(t(np.asarray(odoms, dtype=np.float64)), noise=t(noise))
        torch.cuda.synchronize()
        compare_state(cfg, eng, oracles, tag)

    for i in range(4):
        do_step([RESET_ODOM] * B, f"reset-step {i}")
    rng = np.random.default_rng(5)
    for d in range(5):
        if max(o.sizes()["T"] for o in oracles) > 36:
            break
        plans = [choose_actions(o, rng) for o in oracles]
        for i in range(max(len(p) for p in plans)):
            do_step([p[i] if i < len(p) else np.array([0.0, 0.0, 0.1]) for p in plans], f"decision {d} action {i}")
    n_obs = max(int(np.sum(o.landmarks()["observed"])) for o in oracles)
    print(f"Lt = {n_lm}: up to {n_obs} landmarks observed (border of {2 * n_obs} columns, stride {2 * n_lm})")
    if n_lm == 64:
        assert n_obs > 32     # the > 64-column Gauss-Jordan was really taken
    eng.close()


def test_env_without_any_landmark_in_sight():
    """S = 60, seed 0: the nearest landmark is 8.6 m from the start (the reference regenerates such a world,
    exploration_env.py:416-419, AFTER stepping it): SLAM with an empty landmark border, graph with poses and frontiers only."""
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
    cfg = EnvConfig(map_size=60)
    eng, oracles, t = _engine_and_oracles(cfg, [0, 1], max_poses=32)
    for i in range(4):
        noise = np.stack([o.step(RESET_ODOM) for o in oracles])
        eng.step(t(np.tile(np.array(RESET_ODOM), (2, 1))), noise=t(noise))
        torch.cuda.synchronize()
        compare_state(cfg, eng, oracles, f"reset-step {i}")
    assert int(eng.state["observed"][0].sum()) == 0 and int(np.sum(oracles[0].landmarks()["observed"])) == 0
    assert int(eng.state["status"][0]) == 0 and abs(float(eng.state["metrics"][0, 4]) - oracles[0].metrics()["landmark_error"]) < 1e-12
    eng.close()
    # the same world through the graph kernels: K = T nodes + frontier nodes, edges = the odometry chain + the robot's frontier link
    env = VecExplorationEnv(1, cfg=cfg, max_poses=32)
    o = make_oracles(cfg, [0])[0]
    lm = o.landmarks()
    te = lambda a: torch.as_tensor(a, device=env.device)
    env.eng.reset(seeds=te(np.array([0], dtype=np.int64)), start=te(np.array([o.start])), landmarks=te(lm["true"][None]),
                  scan=te(lm["scan_id"][None].astype(np.int32)), noise=te(o.init_noise[None]))
    for _ in range(4):
        env.eng.step(te(np.array([RESET_ODOM])), noise=te(o.step(RESET_ODOM)[None]))
    g = env.build_graph(); ng, n, e = g.sync_sizes()
    r = o.graph()
    assert ng == 1 and int(g.key_size[0]) == 5 and int(g.fro_size[0]) == r["fro_size"] and n == r["features"].shape[0]
    assert np.array_equal(g.edge_index[:, :e].cpu().numpy(), r["edge_index"])
    env.close()


def test_pose_capacity_overflow_ends_the_episode_without_touching_other_envs():
    from drl_graph_exploration_b200.engine import Engine
    cfg = EnvConfig(map_size=20, num_landmarks=30)
    small, big = Engine(cfg, 2, max_poses=8), Engine(cfg, 2, max_poses=32)
    dev = small.device
    seeds = torch.arange(2, dtype=torch.int64, device=dev)
    odom = torch.tensor([[0.5, 0.0, 0.2]] * 2, dtype=torch.float64, device=dev)
    for eng in (small, big):
        eng.reset(seeds)
    mask = torch.tensor([1, 0], dtype=torch.uint8, device=dev)            # only env 0 moves
    for i in range(10):
        small.step(odom, mask=mask); big.step(odom, mask=mask)
    torch.cuda.synchronize()
    assert int(small.state["n_poses"][0]) == 8 and int(small.state["status"][0]) == -4 and int(small.state["done"][0]) == 1   # DGE_ECAP
    assert int(big.state["n_poses"][0]) == 11 and int(big.state["status"][0]) == 0
    # env 1 never stepped: identical in both engines, and env 0 agrees up to the capacity
    assert int(small.state["n_poses"][1]) == int(big.state["n_poses"][1]) == 1 and int(small.state["status"][1]) == 0
    assert torch.equal(small.state["est_pose"][1, :1], big.state["est_pose"][1, :1])
    small.close(); big.close()


def test_empty_decision_round():
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
    from drl_graph_exploration_b200.runner import HostPolicyLoop
    env = VecExplorationEnv(4, cfg=EnvConfig(map_size=20, num_landmarks=30), max_poses=32)
    env.reset()
    none = torch.zeros(4, dtype=torch.uint8, device=env.device)
    g = env.build_graph(none)
    assert g.sync_sizes() == (0, 0, 0)
    loop = HostPolicyLoop(env, Networks.GCN().to(env.device).eval(), overlap=False)
    loop.need[:] = 0
    mp = ctypes.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
    assert loop._L.dge_graph_host_packed_begin(env.eng._h, loop.t_need.data_ptr(), ctypes.byref(g.c), loop.a_pack.data_ptr(), loop.arena_cap, mp) == 0
    assert loop._L.dge_graph_host_packed_end(env.eng._h, loop.a_pack.data_ptr(), loop.a_host.data_ptr(), loop.arena_cap, ctypes.byref(loop._pk), mp) == 0
    assert (loop._pk.n_graphs, loop._pk.n_nodes, loop._pk.n_edges) == (0, 0, 0)
    env.close()


def test_gnn_kernels_on_degenerate_graphs():
    """A batch with a single-node graph, an edgeless graph and a normal one; and a batch without any edge."""
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.data import Data
    from oracle import gnn_ref
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = Networks.GCN().to(dev).eval()
    ref = gnn_ref.GCN().double().to(dev); ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()}); ref.eval()
    x = torch.randn(1 + 3 + 4, 5, device=dev)
    ei = torch.tensor([[4, 5, 5, 6, 6, 7], [5, 4, 6, 5, 7, 6]], device=dev)         # only the third graph (nodes 4..7) has edges
    w = torch.tensor([1.0, 1.0, 2.5, 2.5, 0.3, 0.3], device=dev)
    for e_index, e_w in ((ei, w), (ei[:, :0], w[:0])):
        with torch.no_grad():
            q = model(Data(x, e_index, e_w), 0.0).view(-1)
            r = ref(gnn_ref.Graph(x.double(), e_index, e_w.double()), 0.0).view(-1).float()
        assert torch.isfinite(q).all() and (q - r).abs().max() <= 1e-4 * r.abs().max()
        xg = x.clone().requires_grad_(True)                                           # training path (autograd kernels)
        out = model(Data(xg, e_index, e_w), 0.0)
        out.sum().backward()
        assert torch.isfinite(xg.grad).all()


def test_world_without_an_observed_landmark_is_regenerated():
    """exploration_env.py:416-419 for B envs: with the 8 landmarks of a 40x40 map a good part of the worlds show no landmark during the four forced steps.
    The eager reset replaces them (next seed of the env's sequence); the in-pipeline reset ends such an episode at once, uncounted."""
    import torch
    from drl_graph_exploration_b200.config import EnvConfig
    from drl_graph_exploration_b200.envs.exploration_env import RESET_ODOM, VecExplorationEnv
    cfg = EnvConfig(map_size=40)
    a = VecExplorationEnv(64, cfg=cfg, max_poses=32, seed0=0)
    a.reset(regenerate=False)
    blind0 = a.eng.state["observed"].sum(dim=1) == 0
    assert int(blind0.sum()) >= 2, int(blind0.sum())               # the situation exists
    seeds0 = a._seeds.clone()
    a.reset()                                                      # regenerate=True: every env ends up with a landmark in view
    assert int((a.eng.state["observed"].sum(dim=1) == 0).sum()) == 0
    assert bool((a._seeds[blind0] != seeds0[blind0]).all()) and torch.equal(a._seeds[~blind0], seeds0[~blind0])
    b = VecExplorationEnv(64, cfg=cfg, max_poses=32, seed0=0)
    b.eng.reset_queued(b._seeds, None, RESET_ODOM, 4)
    c0 = b.eng.state["counters"].clone()
    for _ in range(5):
        b.step_queued()
    torch.cuda.synchronize()
    assert torch.equal(b.eng.state["done"].bool(), blind0) and torch.equal(b.eng.state["counters"][:3], c0[:3])
    a.close(); b.close()
