"""ctypes binding of ``libdge.so`` (include/dge.h) -- the batched B200 exploration engine.

PyTorch is used only for device memory, streams and (elsewhere) ``torch.distributed``; all
simulation / SLAM / virtual-map / graph arithmetic runs in the hand-written CUDA kernels of
``csrc/``.  There is **no CPU fallback**: importing this module on a machine without the built
library, or creating an engine without a GPU, raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

from .config import DgeConfigStruct, EnvConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdge.so")
_lib = None


class DgeError(RuntimeError):
    pass


class _StateView(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "n_poses", "sim_step", "update_count", "true_pose", "est_pose", "lin_pose", "delta_pose", "pose_cov", "pose_info",
        "odom", "meas_ptr", "meas_id", "meas_bearing", "meas_range", "lm_true", "scan_id", "observed", "est_l", "lin_l",
        "land_cov", "prob", "vinfo", "seen", "metrics", "done", "active", "status", "plan", "plan_cursor", "slam_clocks", "counters", "forced", "seed", "pending")]


class GraphOut(ctypes.Structure):
    """``struct dge_graph_out`` (include/dge.h)."""
    _fields_ = [("x", ctypes.c_void_p), ("edge_index", ctypes.c_void_p), ("edge_attr", ctypes.c_void_p), ("batch", ctypes.c_void_p),
                ("node_ptr", ctypes.c_void_p), ("edge_ptr", ctypes.c_void_p), ("key_size", ctypes.c_void_p), ("fro_size", ctypes.c_void_p),
                ("frontier_xy", ctypes.c_void_p), ("totals", ctypes.c_void_p), ("node_cap", ctypes.c_int64), ("edge_cap", ctypes.c_int64),
                ("csr_rowptr", ctypes.c_void_p), ("csr_perm", ctypes.c_void_p), ("gcn_norm", ctypes.c_void_p), ("gcn_selfnorm", ctypes.c_void_p)]


def load_library():
    """Loads libdge.so; raises if it has not been built (``python -m drl_graph_exploration_b200.build_ext``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DgeError(f"{LIB_PATH} is missing: build the CUDA extension first (python -m drl_graph_exploration_b200.build_ext). "
                           "There is no CPU fallback for the product path.")
        L = ctypes.CDLL(LIB_PATH)
        L.dge_last_error.restype = ctypes.c_char_p
        L.dge_create.argtypes = [ctypes.POINTER(DgeConfigStruct), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.dge_destroy.argtypes = [ctypes.c_void_p]
        L.dge_dims.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)]
        vp = ctypes.c_void_p
        L.dge_reset.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        L.dge_step.argtypes = [vp, vp, vp, vp, vp]
        L.dge_reset_queued.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        L.dge_reset_done_queued.argtypes = [vp, ctypes.c_uint64, vp, ctypes.c_int, vp]
        L.dge_mark_pending.argtypes = [vp, vp]
        L.dge_step_queued.argtypes = [vp, vp]
        L.dge_move_measure_queued.argtypes = [vp, vp]
        L.dge_set_counting.argtypes = [vp, ctypes.c_int]
        L.dge_move_measure.argtypes = [vp, vp, vp, vp, vp]
        L.dge_slam_optimize.argtypes = [vp, vp, vp]
        L.dge_virtual_map.argtypes = [vp, vp, vp]
        L.dge_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
        L.dge_get_state.argtypes = [vp, ctypes.POINTER(_StateView)]
        L.dge_virtual_map_rebuild.argtypes = [ctypes.POINTER(DgeConfigStruct), ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, vp, vp, vp, vp, vp, vp]
        L.dge_virtual_map_rebuild_ws_doubles.restype = ctypes.c_int64
        L.dge_virtual_map_rebuild_ws_doubles.argtypes = [ctypes.c_int, ctypes.c_int]
        L.dge_graph.argtypes = [vp, vp, vp, vp]
        L.dge_line_plan.argtypes = [vp, vp, vp, vp, vp]
        L.dge_select_and_plan.argtypes = [vp, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().dge_last_error()
        raise DgeError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


class _DevArray:
    """Zero-copy torch view of engine-owned device memory (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner


_TYPESTR = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    """B independent exploration environments resident in HBM (one CTA per env per kernel)."""

    def __init__(self, cfg: EnvConfig, n_envs: int, max_poses: int = 512, device: int | torch.device = 0):
        if not torch.cuda.is_available():
            raise DgeError("drl_graph_exploration_b200.Engine needs a CUDA device (no CPU fallback)")
        self.cfg = cfg
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        self._L = load_library()
        self._cs = cfg.to_struct()
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _check(self._L.dge_create(ctypes.byref(self._cs), n_envs, max_poses, self.device.index, ctypes.byref(h)), "dge_create")
        self._h = h
        dims = (ctypes.c_int32 * 8)()
        _check(self._L.dge_dims(self._h, dims), "dge_dims")
        (self.B, self.Tmax, self.Lt, self.rows, self.cols, self.Mmax, self.node_cap_env, self.edge_cap_env) = list(dims)
        self.V = self.rows * self.cols
        self.noise_len = 3 + 4 * self.Lt
        sv = _StateView()
        _check(self._L.dge_get_state(self._h, ctypes.byref(sv)), "dge_get_state")
        B, T, L, V, M = self.B, self.Tmax, self.Lt, self.V, self.Mmax
        spec = {
            "n_poses": ((B,), torch.int32), "sim_step": ((B,), torch.int32), "update_count": ((B,), torch.int32),
            "true_pose": ((B, 3), torch.float64), "est_pose": ((B, T, 3), torch.float64), "lin_pose": ((B, T, 3), torch.float64),
            "delta_pose": ((B, T, 3), torch.float64), "pose_cov": ((B, T, 6), torch.float64), "pose_info": ((B, T, 6), torch.float64),
            "odom": ((B, T, 3), torch.float64), "meas_ptr": ((B, T + 1), torch.int32), "meas_id": ((B, M), torch.int32),
            "meas_bearing": ((B, M), torch.float64), "meas_range": ((B, M), torch.float64), "lm_true": ((B, L, 2), torch.float64),
            "scan_id": ((B, L), torch.int32), "observed": ((B, L), torch.uint8), "est_l": ((B, L, 2), torch.float64),
            "lin_l": ((B, L, 2), torch.float64), "land_cov": ((B, L, 3), torch.float64), "prob": ((B, self.rows, self.cols), torch.float64),
            "vinfo": ((B, self.rows, self.cols, 3), torch.float64), "seen": ((B, self.rows, self.cols), torch.int32),
            "metrics": ((B, 8), torch.float64), "done": ((B,), torch.uint8), "active": ((B,), torch.uint8), "status": ((B,), torch.int32),
            "plan": ((B, 6), torch.float64), "plan_cursor": ((B,), torch.int32), "counters": ((8,), torch.int64), "slam_clocks": ((B, 12), torch.int64), "forced": ((B,), torch.int32), "seed": ((B,), torch.int64), "pending": ((B,), torch.uint8),
        }
        self.state = {}
        for name, (shape, dt) in spec.items():
            self.state[name] = torch.as_tensor(_DevArray(getattr(sv, name), shape, _TYPESTR[dt], self), device=self.device)

    def close(self):
        if getattr(self, "_h", None):
            self.state = {}
            self._L.dge_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ calls ---
    def reset(self, seeds: torch.Tensor, mask: Optional[torch.Tensor] = None, start: Optional[torch.Tensor] = None,
              landmarks: Optional[torch.Tensor] = None, scan: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None):
        """dge_reset: SS2D.__init__ for the masked envs (device-generated worlds unless given explicitly)."""
        self._chk(seeds, (self.B,), torch.int64); self._chk(mask, (self.B,), torch.uint8)
        self._chk(start, (self.B, 3), torch.float64); self._chk(landmarks, (self.B, self.Lt, 2), torch.float64)
        self._chk(scan, (self.B, self.Lt), torch.int32); self._chk(noise, (self.B, self.noise_len), torch.float64)
        _check(self._L.dge_reset(self._h, _ptr(mask), _ptr(seeds), _ptr(start), _ptr(landmarks), _ptr(scan), _ptr(noise),
                                 _stream_ptr(self.device)), "dge_reset")

    def reset_queued(self, seeds: torch.Tensor, mask: Optional[torch.Tensor], forced_odom, n_forced: int, start: Optional[torch.Tensor] = None):
        """dge_reset_queued: in-pipeline reset (the initial optimize and the forced steps ride along the next queued ticks)."""
        self._chk(seeds, (self.B,), torch.int64); self._chk(mask, (self.B,), torch.uint8); self._chk(start, (self.B, 3), torch.float64)
        fo = (ctypes.c_double * 3)(*[float(v) for v in forced_odom])
        _check(self._L.dge_reset_queued(self._h, _ptr(mask), _ptr(seeds), _ptr(start), None, None, fo, int(n_forced),
                                        _stream_ptr(self.device)), "dge_reset_queued")

    def step(self, odom: torch.Tensor, mask: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None):
        self._chk(odom, (self.B, 3), torch.float64); self._chk(mask, (self.B,), torch.uint8)
        self._chk(noise, (self.B, self.noise_len), torch.float64)
        _check(self._L.dge_step(self._h, _ptr(odom), _ptr(mask), _ptr(noise), _stream_ptr(self.device)), "dge_step")

    def step_queued(self):
        _check(self._L.dge_step_queued(self._h, _stream_ptr(self.device)), "dge_step_queued")

    def move_measure(self, odom, mask=None, noise=None):
        _check(self._L.dge_move_measure(self._h, _ptr(odom), _ptr(mask), _ptr(noise), _stream_ptr(self.device)), "dge_move_measure")

    def slam_optimize(self, mask=None):
        _check(self._L.dge_slam_optimize(self._h, _ptr(mask), _stream_ptr(self.device)), "dge_slam_optimize")

    def virtual_map(self, mask=None):
        _check(self._L.dge_virtual_map(self._h, _ptr(mask), _stream_ptr(self.device)), "dge_virtual_map")

    def step_host(self, odom_host: np.ndarray, done_host: np.ndarray, obs_host: Optional[np.ndarray] = None, mask_host: Optional[np.ndarray] = None):
        """dge_step_host: host buffers in, host buffers out (H2D/D2H inside the call)."""
        assert odom_host.dtype == np.float64 and odom_host.shape == (self.B, 3) and odom_host.flags.c_contiguous
        p = lambda a: None if a is None else ctypes.c_void_p(a.ctypes.data)
        _check(self._L.dge_step_host(self._h, p(odom_host), p(mask_host), p(done_host), p(obs_host), _stream_ptr(self.device)), "dge_step_host")

    def _chk(self, t, shape, dtype):
        if t is None:
            return
        if not (t.is_cuda and t.dtype == dtype and tuple(t.shape) == tuple(shape) and t.is_contiguous()):
            raise DgeError(f"expected contiguous cuda tensor {shape} {dtype}, got {tuple(t.shape)} {t.dtype} {t.device}")


def virtual_map_rebuild(cfg: EnvConfig, pose: torch.Tensor, cov6: torch.Tensor, landmarks: torch.Tensor, want_seen: bool = False):
    """Stand-alone a6+a7 rebuild (dge_virtual_map_rebuild) for n problems: pose [n,T,3], cov6 [n,T,6], landmarks [n,L,2]."""
    L = load_library()
    cs = cfg.to_struct()
    n, T = pose.shape[0], pose.shape[1]
    nl = landmarks.shape[1]
    V = cfg.rows * cfg.cols
    dev = pose.device
    prob = torch.empty(n, V, dtype=torch.float64, device=dev)
    vinfo = torch.empty(n, V, 3, dtype=torch.float64, device=dev)
    seen = torch.empty(n, V, dtype=torch.int32, device=dev) if want_seen else None
    ws = torch.empty(L.dge_virtual_map_rebuild_ws_doubles(n, T), dtype=torch.float64, device=dev)
    lm = landmarks.contiguous() if nl > 0 else torch.zeros(n, 1, 2, dtype=torch.float64, device=dev)
    _check(L.dge_virtual_map_rebuild(ctypes.byref(cs), n, T, _ptr(pose.contiguous()), _ptr(cov6.contiguous()), nl, _ptr(lm), _ptr(prob), _ptr(vinfo),
                                     _ptr(seen), _ptr(ws), _stream_ptr(dev)), "dge_virtual_map_rebuild")
    return prob.view(n, cfg.rows, cfg.cols), vinfo.view(n, cfg.rows, cfg.cols, 3), (None if seen is None else seen.view(n, cfg.rows, cfg.cols))
