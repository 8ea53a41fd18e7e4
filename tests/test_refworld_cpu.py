"""The product's restatement of the reference's random streams (csrc/dge_refworld.cu, host code of libdge.so) against the oracle's
own (oracle/dge_oracle.cpp): same landmarks, same visiting order, bit-identical noise rows over whole trajectories -- two independent
implementations of pyss2d.py:89-119 / Simulator2D.cpp:161-173,436-463,505-527 / the GCC 5-7 hashtable order."""
import ctypes
import math

import numpy as np
import pytest

from drl_graph_exploration_b200.config import EnvConfig, start_pose_for_seed
from drl_graph_exploration_b200.engine import load_library
from oracle.oracle import OracleEnv


def _lib():
    L = load_library()
    L.dge_refworld_create.restype = ctypes.c_void_p
    L.dge_refworld_create.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
    L.dge_refworld_destroy.argtypes = [ctypes.c_void_p]
    L.dge_refworld_destroy.restype = None
    L.dge_refworld_world.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.dge_refworld_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return L


@pytest.mark.parametrize("map_size,n_lm,seed", [(20, 30, 0), (40, None, 0), (40, None, 7), (100, None, 3), (20, 64, 5)])
def test_streams_equal_the_oracles(map_size, n_lm, seed):
    L = _lib()
    cfg = EnvConfig(map_size=map_size, num_landmarks=n_lm)
    o = OracleEnv(cfg, seed)
    start = np.array(start_pose_for_seed(seed, map_size, cfg.ext), dtype=np.float64)
    assert np.array_equal(start, np.asarray(o.start, dtype=np.float64))
    cs = cfg.to_struct()
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    w = ctypes.c_void_p(L.dge_refworld_create(ctypes.byref(cs), ctypes.c_uint32(seed), vp(start)))
    assert w
    Lt = cfg.n_landmarks
    lm, scan, n0 = np.zeros((Lt, 2)), np.zeros(Lt, dtype=np.int32), np.zeros(3 + 4 * Lt)
    assert L.dge_refworld_world(w, vp(lm), vp(scan), vp(n0)) == 0
    ol = o.landmarks()
    assert np.array_equal(lm, ol["true"]) and np.array_equal(scan, ol["scan_id"].astype(np.int32))
    assert np.array_equal(n0, o.init_noise)
    rng = np.random.default_rng(seed)
    row = np.zeros(3 + 4 * Lt)
    for k in range(60):
        od = np.array([1.0, 1.0, math.pi / 2]) if k < 4 else np.array([rng.uniform(0, 2), 0.0, rng.uniform(-1, 1)])
        ref = o.step(od)
        assert L.dge_refworld_step(w, vp(od), vp(row)) == 0
        assert np.array_equal(row, ref), k
    L.dge_refworld_destroy(w)
