"""ctypes wrapper over ``liboracle_dge.so`` -- the CPU restatement of the reference's
simulator / SLAM / virtual-map / graph path.  TEST INFRASTRUCTURE ONLY (see __init__)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from drl_graph_exploration_b200.config import DgeConfigStruct, EnvConfig, start_pose_for_seed

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_dge.so")
_lib = None


def build(force: bool = False) -> str:
    # always through make: it rebuilds only when a source is newer than the library (a stale checker is worse than a slow one)
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    except (OSError, subprocess.CalledProcessError):
        if not os.path.exists(_LIB_PATH):
            raise
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.POINTER(DgeConfigStruct)]
        L.orc_clone.restype = ctypes.c_void_p
        L.orc_clone.argtypes = [ctypes.c_void_p]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_utility.restype = ctypes.c_double
        L.orc_utility.argtypes = [ctypes.c_void_p, ctypes.c_double]
        L.orc_sim_reward.restype = ctypes.c_double
        assert L.orc_sizeof_config() == ctypes.sizeof(DgeConfigStruct)
        _lib = L
    return _lib


def _p(a, t=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


class OracleEnv:
    """One reference-faithful environment (Simulator2D + SLAM2D + VirtualMap + the
    ExplorationEnv orchestration), fp64, single-threaded like the reference."""

    def __init__(self, cfg: EnvConfig, seed: int, start=None, landmarks=None, dense: bool = False, record_noise: bool = True):
        self.cfg = cfg
        self._L = lib()
        self._cs = cfg.to_struct()
        self._h = ctypes.c_void_p(self._L.orc_create(ctypes.byref(self._cs)))
        self._L.orc_set_dense(self._h, int(dense))
        if start is None:
            start = start_pose_for_seed(seed, cfg.map_size, cfg.ext)
        self.start = tuple(float(v) for v in start)
        lt = cfg.n_landmarks if landmarks is None else len(landmarks)
        self.noise_len = 3 + 4 * lt
        self.init_noise = np.zeros(self.noise_len) if record_noise else None
        if landmarks is None:
            rc = self._L.orc_init(self._h, ctypes.c_uint32(seed), ctypes.c_double(self.start[0]), ctypes.c_double(self.start[1]),
                                  ctypes.c_double(self.start[2]), _p(self.init_noise))
        else:
            xy = np.ascontiguousarray(landmarks, dtype=np.float64)
            rc = self._L.orc_init_landmarks(self._h, ctypes.c_uint32(seed), ctypes.c_double(self.start[0]), ctypes.c_double(self.start[1]),
                                            ctypes.c_double(self.start[2]), _p(xy), ctypes.c_int(len(xy)), _p(self.init_noise))
        if rc:
            raise RuntimeError("oracle init failed")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_destroy(self._h)
            self._h = None

    def clone(self) -> "OracleEnv":
        o = object.__new__(OracleEnv)
        o.cfg, o._L, o._cs, o.start, o.noise_len, o.init_noise = self.cfg, self._L, self._cs, self.start, self.noise_len, None
        o._h = ctypes.c_void_p(self._L.orc_clone(self._h))
        return o

    # ---- stepping ----
    def step(self, odom, record_noise: bool = True):
        o = np.ascontiguousarray(odom, dtype=np.float64)
        rec = np.zeros(self.noise_len) if record_noise else None
        rc = self._L.orc_step(self._h, _p(o), _p(rec))
        if rc == 1:
            raise RuntimeError("oracle SLAM solve failed")
        return rec

    # ---- state views ----
    def sizes(self):
        out = np.zeros(8, dtype=np.int32)
        self._L.orc_sizes(self._h, _p(out, ctypes.c_int32))
        return dict(zip(["T", "Lt", "n_obs", "rows", "cols", "M", "sim_step", "update_count"], out.tolist()))

    def poses(self):
        T = self.sizes()["T"]
        est, lin, delta = np.zeros((T, 3)), np.zeros((T, 3)), np.zeros((T, 3))
        cov, info, true = np.zeros((T, 3, 3)), np.zeros((T, 3, 3)), np.zeros(3)
        self._L.orc_get_poses(self._h, _p(est), _p(lin), _p(delta), _p(cov), _p(info), _p(true))
        return dict(est=est, lin=lin, delta=delta, cov=cov, info=info, true=true)

    def landmarks(self):
        lt = self.sizes()["Lt"]
        obs = np.zeros(lt, dtype=np.uint8)
        est, lin, cov, info, true = np.zeros((lt, 2)), np.zeros((lt, 2)), np.zeros((lt, 2, 2)), np.zeros((lt, 2, 2)), np.zeros((lt, 2))
        scan = np.zeros(lt, dtype=np.uint32)
        self._L.orc_get_landmarks(self._h, _p(obs, ctypes.c_uint8), _p(est), _p(lin), _p(cov), _p(info), _p(true), _p(scan, ctypes.c_uint32))
        return dict(observed=obs, est=est, lin=lin, cov=cov, info=info, true=true, scan_id=scan)

    def factors(self):
        s = self.sizes()
        odom = np.zeros((max(s["T"] - 1, 0), 3))
        ptr = np.zeros(s["T"] + 1, dtype=np.int32)
        mid, mb, mr = np.zeros(s["M"], dtype=np.int32), np.zeros(s["M"]), np.zeros(s["M"])
        self._L.orc_get_factors(self._h, _p(odom), _p(ptr, ctypes.c_int32), _p(mid, ctypes.c_int32), _p(mb), _p(mr))
        return dict(odom=odom, meas_ptr=ptr, meas_id=mid, meas_bearing=mb, meas_range=mr)

    def vmap(self):
        s = self.sizes()
        V = s["rows"] * s["cols"]
        prob, vinfo, seen, trace = np.zeros(V), np.zeros((V, 4)), np.zeros(V, dtype=np.int32), np.zeros(V)
        self._L.orc_get_vmap(self._h, _p(prob), _p(vinfo), _p(seen, ctypes.c_int32), _p(trace))
        r, c = s["rows"], s["cols"]
        return dict(prob=prob.reshape(r, c), info=vinfo.reshape(r, c, 2, 2), seen=seen.reshape(r, c), trace=trace.reshape(r, c))

    def metrics(self):
        out = np.zeros(6)
        self._L.orc_metrics(self._h, _p(out))
        return dict(explored=out[0], utility0=out[1], landmark_error=out[2], max_traj_uncertainty=out[3], done=bool(out[4]), dist=out[5])

    def utility(self, distance: float) -> float:
        return self._L.orc_utility(self._h, ctypes.c_double(distance))

    # ---- graph / planning ----
    def graph(self):
        sz = np.zeros(6, dtype=np.int32)
        self._L.orc_graph_build(self._h, _p(sz, ctypes.c_int32))
        N, K, L, F, E, nc = sz.tolist()
        feat, ei, ew = np.zeros((N, 5)), np.zeros((2, E), dtype=np.int64), np.zeros(E)
        fxy, cells = np.zeros((F, 2)), np.zeros(nc, dtype=np.int32)
        self._L.orc_graph_fetch(_p(feat), _p(ei, ctypes.c_int64), _p(ew), _p(fxy), _p(cells, ctypes.c_int32))
        return dict(n_nodes=N, key_size=K, land_size=L, fro_size=F, features=feat, edge_index=ei, edge_attr=ew,
                    frontier_xy=fxy, all_frontier_cells=cells)

    def line_plan(self, gx: float, gy: float):
        out = np.zeros((256, 3))
        n = self._L.orc_line_plan(self._h, ctypes.c_double(gx), ctypes.c_double(gy), _p(out), ctypes.c_int(256))
        return out[:n].copy()

    def sim_reward(self, actions, noise=None) -> float:
        """EMPlanner2D::simulations_reward; `noise` [n_actions, 3+4*Lt] replaces the RNG streams (parity tests)."""
        a = np.ascontiguousarray(actions, dtype=np.float64)
        if noise is None:
            return self._L.orc_sim_reward(self._h, _p(a), ctypes.c_int(len(a)))
        nz = np.ascontiguousarray(noise, dtype=np.float64)
        assert nz.shape == (len(a), self.noise_len)
        self._L.orc_sim_reward_noise.restype = ctypes.c_double
        return self._L.orc_sim_reward_noise(self._h, _p(a), ctypes.c_int(len(a)), _p(nz))

    def rewards_all_goals(self, g=None):
        """exploration_env.py:145-162 (+ actions_all_goals :134-143); returns (rewards[N], loop_clo, actions)."""
        g = g or self.graph()
        K, F, N = g["key_size"], g["fro_size"], g["n_nodes"]
        actions = [self.line_plan(*g["frontier_xy"][f]) for f in range(F)]
        rewards = np.full(N, np.nan)
        for f in range(F):
            rewards[K + f] = self.sim_reward(actions[f])
        act_max = int(np.nanargmax(rewards))
        lo, hi = np.nanmin(rewards), np.nanmax(rewards)
        if act_max == K:   # is_nf: frontier 0 is the robot's nearest frontier
            loop_clo, out = False, np.interp(rewards, (lo, hi), (-1.0, 0.0))
        else:
            loop_clo, out = True, np.interp(rewards, (lo, hi), (-1.0, 1.0))
        out[np.isnan(out)] = 0
        return out, loop_clo, actions


def virtual_map_rebuild(cfg: EnvConfig, pose, info, lm):
    """Stand-alone a6+a7 rebuild (VirtualMap.cpp:61-84,256-316) on explicit inputs."""
    L = lib()
    cs = cfg.to_struct()
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    info = np.ascontiguousarray(info, dtype=np.float64)
    lm = np.ascontiguousarray(lm, dtype=np.float64).reshape(-1, 2)
    r, c = cfg.rows, cfg.cols
    prob, vinfo, seen = np.zeros(r * c), np.zeros((r * c, 4)), np.zeros(r * c, dtype=np.int32)
    L.orc_virtual_map_rebuild(ctypes.byref(cs), ctypes.c_int(len(pose)), _p(pose), _p(info), ctypes.c_int(len(lm)), _p(lm),
                              ctypes.c_int(r), ctypes.c_int(c), _p(prob), _p(vinfo), _p(seen, ctypes.c_int32))
    return prob.reshape(r, c), vinfo.reshape(r, c, 2, 2), seen.reshape(r, c)


def virtual_map_rebuild_batch(cfg: EnvConfig, pose, info, lm, threads: int):
    L = lib()
    cs = cfg.to_struct()
    n, T = pose.shape[0], pose.shape[1]
    nl = lm.shape[1]
    r, c = cfg.rows, cfg.cols
    prob, vinfo = np.zeros((n, r * c)), np.zeros((n, r * c, 4))
    L.orc_virtual_map_rebuild_batch(ctypes.byref(cs), ctypes.c_int(n), ctypes.c_int(threads), ctypes.c_int(T),
                                    _p(np.ascontiguousarray(pose)), _p(np.ascontiguousarray(info)), ctypes.c_int(nl),
                                    _p(np.ascontiguousarray(lm)), ctypes.c_int(r), ctypes.c_int(c), _p(prob), _p(vinfo))
    return prob, vinfo
