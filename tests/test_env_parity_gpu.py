"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle on the same worlds,
actions and noise.  Integer outputs (trajectory length, measurement ids, visibility counts)
bit-exact; fp64 state within 1e-6 relative (the contract is 1e-4)."""
import math

import numpy as np
import pytest
import torch

from helpers import RESET_ODOM, borderline_cells, choose_actions, make_oracles, sym3_to_full, sym6_to_full, world_arrays
from drl_graph_exploration_b200.config import EnvConfig

pytestmark = pytest.mark.gpu

RTOL = 1e-6


def _close(a, b, rtol=RTOL, atol=1e-9, what=""):
    a, b = np.asarray(a), np.asarray(b)
    err = np.abs(a - b) / (atol / rtol + np.maximum(np.abs(a), np.abs(b)))
    assert np.all(err <= rtol), f"{what}: max rel err {err.max():.3e}"


def compare_state(cfg, eng, oracles, step_tag, check_map=True, RTOL=RTOL, cov_matrix_rel=False):
    st = {k: v.cpu().numpy() for k, v in eng.state.items()}
    n_border = 0
    for b, o in enumerate(oracles):
        s = o.sizes()
        T = s["T"]
        assert st["n_poses"][b] == T, f"{step_tag} env {b}: T"
        assert st["status"][b] == 0
        assert st["update_count"][b] == s["update_count"]
        P, Lm, F = o.poses(), o.landmarks(), o.factors()
        # measurement lists: ids exact, values tight
        assert np.array_equal(st["meas_ptr"][b, :T + 1], F["meas_ptr"]), f"{step_tag} env {b}: meas_ptr"
        M = F["meas_ptr"][-1]
        assert np.array_equal(st["meas_id"][b, :M], F["meas_id"]), f"{step_tag} env {b}: meas ids"
        _close(st["meas_bearing"][b, :M], F["meas_bearing"], 1e-11, 1e-13, "meas bearing")
        _close(st["meas_range"][b, :M], F["meas_range"], 1e-11, 1e-13, "meas range")
        _close(st["true_pose"][b], P["true"], 1e-11, 1e-13, "true pose")
        _close(st["est_pose"][b, :T], P["est"], RTOL, 1e-9, f"{step_tag} env {b} est pose")
        _close(st["lin_pose"][b, :T], P["lin"], RTOL, 1e-9, f"{step_tag} env {b} lin pose")
        if cov_matrix_rel:   # long trajectories: error relative to the 3x3 block's largest entry (an off-diagonal entry may cross zero)
            ga, ra = sym6_to_full(st["pose_cov"][b, :T]), P["cov"]
            merr = np.abs(ga - ra).max(axis=(-1, -2)) / np.abs(ra).max(axis=(-1, -2))
            assert merr.max() <= RTOL, f"{step_tag} env {b} pose cov: max matrix-relative err {merr.max():.3e}"
        else:
            _close(sym6_to_full(st["pose_cov"][b, :T]), P["cov"], RTOL, 1e-12, f"{step_tag} env {b} pose cov")
        assert np.array_equal(st["observed"][b], Lm["observed"])
        ob = Lm["observed"].astype(bool)
        _close(st["est_l"][b][ob], Lm["est"][ob], RTOL, 1e-9, "landmark est")
        if cov_matrix_rel:
            ga, ra = sym3_to_full(st["land_cov"][b])[ob], Lm["cov"][ob]
            if ob.any():
                merr = np.abs(ga - ra).max(axis=(-1, -2)) / np.abs(ra).max(axis=(-1, -2))
                assert merr.max() <= RTOL, f"{step_tag} env {b} landmark cov: max matrix-relative err {merr.max():.3e}"
        else:
            _close(sym3_to_full(st["land_cov"][b])[ob], Lm["cov"][ob], RTOL, 1e-12, "landmark cov")
        m = o.metrics()
        _close(st["metrics"][b, 4], m["landmark_error"], RTOL, 1e-9, "landmark error")
        _close(st["metrics"][b, 5], m["max_traj_uncertainty"], RTOL, 1e-12, "max traj uncertainty")
        if check_map:
            vm = o.vmap()
            border = borderline_cells(cfg, P["est"])
            n_border += int(border.sum())
            ok = ~border
            if not np.array_equal(st["seen"][b][ok], vm["seen"][ok]):
                bad = np.argwhere((st["seen"][b] != vm["seen"]) & ok)
                half = cfg.map_size / 2 + cfg.ext
                info = []
                for (r_, c_) in bad[:4]:
                    cx, cy = (c_ + 0.5) * cfg.resolution - half, (r_ + 0.5) * cfg.resolution - half
                    d = np.hypot(P["est"][:, 0] - cx, P["est"][:, 1] - cy)
                    k = int(np.argmin(np.abs(d - cfg.max_range)))
                    dg = np.hypot(st["est_pose"][b, :T, 0] - cx, st["est_pose"][b, :T, 1] - cy)
                    ep = st["est_pose"][b, :T]
                    qx = np.cos(ep[:, 2]) * (cx - ep[:, 0]) + np.sin(ep[:, 2]) * (cy - ep[:, 1])
                    qy = -np.sin(ep[:, 2]) * (cx - ep[:, 0]) + np.cos(ep[:, 2]) * (cy - ep[:, 1])
                    bear = np.degrees(np.arctan2(qy, qx))
                    vis = (dg < cfg.max_range) & (np.abs(bear) < cfg.max_bearing_deg)
                    edge = [(int(i), round(float(dg[i]), 6), round(float(bear[i]), 4)) for i in range(T) if abs(dg[i] - cfg.max_range) < 0.05 or abs(abs(bear[i]) - 179.9) < 0.3]
                    info.append((int(r_), int(c_), int(st["seen"][b][r_, c_]), int(vm["seen"][r_, c_]), k, float(d[k] - cfg.max_range), float(dg[k] - cfg.max_range),
                                 "numpy count on gpu poses", int(vis.sum()), "T", T, "edge poses (idx, range, bearing deg)", edge))
                raise AssertionError(f"{step_tag} env {b}: visibility counts differ at (row, col, gpu, oracle, nearest-pose, oracle margin, gpu margin): {info}")
            assert np.array_equal(st["prob"][b][ok], vm["prob"][ok]), f"{step_tag} env {b}: occupancy (bit-exact)"
            # 2x2 information matrices: error relative to the matrix scale (off-diagonals are ~1e-4 of the diagonal)
            ga, ra = sym3_to_full(st["vinfo"][b])[ok], vm["info"][ok]
            merr = np.abs(ga - ra).max(axis=(-1, -2)) / np.abs(ra).max(axis=(-1, -2))
            assert merr.max() <= 1e-5, f"{step_tag} env {b} cell information: max matrix-relative err {merr.max():.3e}"
            if not border.any():
                _close(st["metrics"][b, 0], m["explored"], 1e-12, 0, "explored")
                _close(st["metrics"][b, 1], m["utility0"], 1e-5, 1e-9, "utility")
                assert bool(st["done"][b]) == m["done"]
    return n_border


@pytest.mark.parametrize("map_size,n_lm,B,n_dec", [(20, 30, 6, 8), (40, None, 6, 10)])
def test_engine_tracks_oracle(map_size, n_lm, B, n_dec):
    from drl_graph_exploration_b200.engine import Engine

    cfg = EnvConfig(map_size=map_size, num_landmarks=n_lm)
    oracles = make_oracles(cfg, range(B))
    start, lm, scan, noise0 = world_arrays(oracles)
    eng = Engine(cfg, B, max_poses=128)
    dev = eng.device
    t = lambda a, dt=None: torch.as_tensor(a, device=dev) if dt is None else torch.as_tensor(a, device=dev, dtype=dt)
    eng.reset(seeds=t(np.arange(B, dtype=np.int64)), start=t(start), landmarks=t(lm), scan=t(scan), noise=t(noise0))
    torch.cuda.synchronize()
    compare_state(cfg, eng, oracles, "reset", check_map=False)

    def do_step(odoms, tag):
        noise = np.stack([o.step(od) for o, od in zip(oracles, odoms)])
        eng.step(t(np.asarray(odoms, dtype=np.float64)), noise=t(noise))
        torch.cuda.synchronize()
        return compare_state(cfg, eng, oracles, tag)

    nb = 0
    for i in range(4):
        nb += do_step([RESET_ODOM] * B, f"reset-step {i}")
    rng = np.random.default_rng(0)
    nsteps = 0
    for d in range(n_dec):
        plans = [choose_actions(o, rng) for o in oracles]
        for i in range(max(len(p) for p in plans)):
            # envs whose plan is exhausted keep rotating in place (still a full SLAM step)
            odoms = [p[i] if i < len(p) else np.array([0.0, 0.0, 0.1]) for p in plans]
            nb += do_step(odoms, f"decision {d} action {i}")
            nsteps += 1
            if any(o.sizes()["T"] >= 120 for o in oracles):
                break
        if any(o.sizes()["T"] >= 120 for o in oracles):
            break
    assert nsteps >= 10
    assert max(o.sizes()["update_count"] for o in oracles) >= 20   # past two relinearisation rounds
    print(f"steps={nsteps} borderline-cell exclusions={nb}")
    eng.close()


def test_step_host_roundtrip():
    """The host-buffer entry point (what a ctypes caller uses) gives the same state as the device one."""
    from drl_graph_exploration_b200.engine import Engine

    cfg = EnvConfig(map_size=20, num_landmarks=30)
    B = 4
    e1, e2 = Engine(cfg, B, max_poses=64), Engine(cfg, B, max_poses=64)
    seeds = torch.arange(B, dtype=torch.int64, device=e1.device)
    e1.reset(seeds); e2.reset(seeds)
    odom = np.tile(np.array(RESET_ODOM), (B, 1))
    done = np.zeros(B, dtype=np.uint8)
    obs = np.zeros((B, cfg.rows, cfg.cols))
    for _ in range(3):
        e1.step(torch.as_tensor(odom, device=e1.device))
        e2.step_host(odom, done, obs)
    torch.cuda.synchronize()
    for k in ("est_pose", "pose_cov", "prob", "vinfo", "seen", "n_poses"):
        assert torch.equal(e1.state[k], e2.state[k]), k
    assert np.array_equal(obs, e2.state["prob"].cpu().numpy())
    e1.close(); e2.close()
