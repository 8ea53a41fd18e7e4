"""Whole episodes, CUDA engine against the CPU oracle (same worlds, actions and noise), state compared after EVERY step:
* config C1 (20x20 map, 30 landmarks) from reset to `done` -- the episode's end (explored > 0.85, exploration_env.py:167-168) is
  reached on both sides at the same step;
* a 100x100 map (50 landmarks: the SLAM kernel's wide-border paths, 100 border columns) followed past 500 poses -- the chunked
  staging of k_slam / k_vmap_env and the workspace indexing at T >= 500 (max_steps of the reference is 5000).
fp64 state is held to 1e-5 here (1e-6 in the short-trajectory tests; contract 1e-4), covariance blocks relative to their largest
entry: over hundreds of poses the small off-diagonal entries of the marginals collect rounding differences between the two elimination orders.
Sorted last (zz): the long run takes the CPU oracle about a minute."""
import numpy as np
import pytest
import torch

from helpers import RESET_ODOM, choose_actions, make_oracles, world_arrays
from test_env_parity_gpu import compare_state
from drl_graph_exploration_b200.config import EnvConfig

pytestmark = pytest.mark.gpu


def _run(cfg, seeds, max_poses, stop, compare_every=1, map_every=1):
    from drl_graph_exploration_b200.engine import Engine
    B = len(seeds)
    oracles = make_oracles(cfg, seeds)
    start, lm, scan, noise0 = world_arrays(oracles)
    eng = Engine(cfg, B, max_poses=max_poses)
    t = lambda a: torch.as_tensor(a, device=eng.device)
    eng.reset(seeds=t(np.asarray(seeds, dtype=np.int64)), start=t(start), landmarks=t(lm), scan=t(scan), noise=t(noise0))
    rng = np.random.default_rng(11)
    n = 0

    def step(odoms):
        nonlocal n
        noise = np.stack([o.step(od) for o, od in zip(oracles, odoms)])
        eng.step(t(np.asarray(odoms, dtype=np.float64)), noise=t(noise))
        n += 1
        if n % compare_every == 0:
            torch.cuda.synchronize()
            compare_state(cfg, eng, oracles, f"step {n}", check_map=(n % map_every == 0), RTOL=1e-5, cov_matrix_rel=True)

    for _ in range(4):
        step([RESET_ODOM] * B)
    while not stop(oracles, eng):
        plans = [choose_actions(o, rng) for o in oracles]
        for i in range(max(len(p) for p in plans)):
            step([p[i] if i < len(p) else np.array([0.0, 0.0, 0.1]) for p in plans])
            if stop(oracles, eng):
                break
    torch.cuda.synchronize()
    compare_state(cfg, eng, oracles, "final", RTOL=1e-5, cov_matrix_rel=True)
    out = (n, [o.sizes()["T"] for o in oracles], eng.state["done"].cpu().numpy().copy(), [o.metrics()["done"] for o in oracles])
    eng.close()
    return out


def test_c1_episode_to_done():
    cfg = EnvConfig(map_size=20, num_landmarks=30)
    for seed in (0, 3):
        n, T, done_gpu, done_ref = _run(cfg, [seed], 400, lambda os, e: os[0].metrics()["done"] or os[0].sizes()["T"] >= 380)
        assert done_ref[0] and bool(done_gpu[0]), (seed, n, T)         # both sides end the episode, at the same step (compare_state holds `done` every step)
        assert T[0] >= 20


def test_long_trajectory_past_500_poses_on_the_100_map():
    cfg = EnvConfig(map_size=100)
    n, T, _, _ = _run(cfg, [1], 640, lambda os, e: os[0].sizes()["T"] >= 520, compare_every=5, map_every=25)
    assert T[0] >= 520
