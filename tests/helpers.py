"""Shared helpers of the parity tests: drive CPU-oracle envs and the CUDA engine with the
same worlds, actions and noise."""
import math

import numpy as np

from drl_graph_exploration_b200.config import EnvConfig
from oracle.oracle import OracleEnv

RESET_ODOM = (1.0, 1.0, math.pi / 2.0)   # exploration_env.py:411-414


def make_oracles(cfg: EnvConfig, seeds):
    return [OracleEnv(cfg, int(s)) for s in seeds]


def world_arrays(oracles):
    """start poses, true landmarks (by id), scan order and init noise of a list of oracle envs."""
    start = np.array([o.start for o in oracles], dtype=np.float64)
    lms = [o.landmarks() for o in oracles]
    lm = np.stack([l["true"] for l in lms]).astype(np.float64)
    scan = np.stack([l["scan_id"] for l in lms]).astype(np.int32)
    noise = np.stack([o.init_noise for o in oracles]).astype(np.float64)
    return start, lm, scan, noise


def sym6_to_full(c6):
    """[...,6] (xx,xy,xt,yy,yt,tt) -> [...,3,3]"""
    c6 = np.asarray(c6)
    out = np.empty(c6.shape[:-1] + (3, 3))
    idx = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    for q, (i, j) in enumerate(idx):
        out[..., i, j] = c6[..., q]
        out[..., j, i] = c6[..., q]
    return out


def sym3_to_full(c3):
    c3 = np.asarray(c3)
    out = np.empty(c3.shape[:-1] + (2, 2))
    out[..., 0, 0] = c3[..., 0]; out[..., 0, 1] = c3[..., 1]; out[..., 1, 0] = c3[..., 1]; out[..., 1, 1] = c3[..., 2]
    return out


def borderline_cells(cfg: EnvConfig, poses, tol=1e-6):
    """Cells whose visibility predicate (range < max_range, range > min_range, rear blind wedge) is
    decided within `tol` by at least one pose -- there the integer outcome legitimately depends on
    the last bits of the pose estimate (the integer start positions of pyss2d.py:88-95 put four
    cell centres at exactly max_range; see DESIGN.md 'knife-edge predicates')."""
    r, c = cfg.rows, cfg.cols
    half = cfg.map_size / 2 + cfg.ext
    cx = (np.arange(c) + 0.5) * cfg.resolution - half
    cy = (np.arange(r) + 0.5) * cfg.resolution - half
    flag = np.zeros((r, c), dtype=bool)
    for x, y, th in poses:
        dx, dy = cx[None, :] - x, cy[:, None] - y
        d = np.sqrt(dx * dx + dy * dy)
        flag |= np.abs(d - cfg.max_range) < tol
        flag |= np.abs(d - cfg.min_range) < tol
        flag |= d < tol     # an odd-integer start pose IS a cell centre: the bearing of a ~1e-12 vector is noise
        qx = math.cos(th) * dx + math.sin(th) * dy
        qy = -math.sin(th) * dx + math.cos(th) * dy
        b = np.arctan2(qy, qx)
        lim = math.radians(cfg.max_bearing_deg)
        flag |= (np.abs(np.abs(b) - lim) < tol) & (d < cfg.max_range + tol)
    return flag


def choose_actions(oracle, rng):
    """A plausible decision: pick one of the oracle's frontiers at random, line-plan to it."""
    g = oracle.graph()
    if g["fro_size"] == 0:
        return [np.array([0.0, 0.0, 0.3])]
    f = int(rng.integers(g["fro_size"]))
    return list(oracle.line_plan(*g["frontier_xy"][f]))
