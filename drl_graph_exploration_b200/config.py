"""Environment parameters of the exploration hot path.

Mirrors the values the reference reads from ``scripts/envs/exploration_env.ini`` through
``pyss2d.py:10-55`` (``read_*_params``) and the overrides ``ExplorationEnv.reset`` applies
(``exploration_env.py:399-407``).  ``EnvConfig.to_struct()`` yields the ``dge_config`` C
struct of ``include/dge.h`` (the CPU oracle under ``oracle/`` uses the same layout).
"""
from __future__ import annotations

import configparser
import ctypes
import math
from dataclasses import dataclass


class DgeConfigStruct(ctypes.Structure):
    """``struct dge_config`` (include/dge.h) -- field order is ABI."""

    _fields_ = [
        ("env_min_x", ctypes.c_double), ("env_max_x", ctypes.c_double),
        ("env_min_y", ctypes.c_double), ("env_max_y", ctypes.c_double),
        ("map_min_x", ctypes.c_double), ("map_max_x", ctypes.c_double),
        ("map_min_y", ctypes.c_double), ("map_max_y", ctypes.c_double),
        ("resolution", ctypes.c_double), ("sigma0", ctypes.c_double),
        ("bearing_noise", ctypes.c_double), ("range_noise", ctypes.c_double),
        ("min_bearing", ctypes.c_double), ("max_bearing", ctypes.c_double),
        ("min_range", ctypes.c_double), ("max_range", ctypes.c_double),
        ("trans_noise", ctypes.c_double), ("rot_noise", ctypes.c_double),
        ("sigma_x0", ctypes.c_double), ("sigma_y0", ctypes.c_double), ("sigma_theta0", ctypes.c_double),
        ("angle_weight", ctypes.c_double), ("dist_w0", ctypes.c_double), ("dist_w1", ctypes.c_double),
        ("max_edge_length", ctypes.c_double), ("occupancy_threshold", ctypes.c_double),
        ("max_steps", ctypes.c_double),
        ("relin_thresh", ctypes.c_double),
        ("relin_skip", ctypes.c_int32),
        ("num_landmarks", ctypes.c_int32),
    ]


def _rot2_theta(x: float) -> float:
    """``Rot2(x).theta()`` -- the reference setters normalise angles this way
    (include/em_exploration/Simulation2D.h:52-57,152)."""
    return math.atan2(math.sin(x), math.cos(x))


@dataclass
class EnvConfig:
    """Defaults = scripts/envs/exploration_env.ini (values, not the file)."""

    map_size: int = 40
    num_landmarks: int | None = None          # None -> int(map_size**2 * 0.005)  exploration_env.py:399
    # [Sensor Model]
    bearing_noise_deg: float = 0.5
    range_noise: float = 0.02
    min_bearing_deg: float = -179.9
    max_bearing_deg: float = 179.9
    min_range: float = 0.1
    max_range: float = 6.0
    # [Control Model]
    translation_noise: float = 0.1
    rotation_noise_deg: float = 0.2
    # [Environment]
    max_steps: float = 5000.0
    # [Virtual Map]
    resolution: float = 2.0
    sigma0: float = 1.0
    # [Simulator]
    sigma_x0: float = 0.05
    sigma_y0: float = 0.05
    sigma_theta0_deg: float = 0.01
    # [Planner]
    angle_weight: float = 0.4
    distance_weight0: float = 5.0
    distance_weight1: float = 2.0
    max_edge_length: float = 2.0
    occupancy_threshold: float = 0.4
    # gtsam::ISAM2Params defaults (SLAM2D.cpp:10-12)
    relinearize_threshold: float = 0.1
    relinearize_skip: int = 10
    ext: float = 20.0                          # pyss2d.py:48 read_map_params(ext=20.0)

    @classmethod
    def from_ini(cls, path: str, map_size: int, num_landmarks: int | None = None) -> "EnvConfig":
        """Accepts a file in the reference's ini format (utils.py:42-45 load_config)."""
        cp = configparser.ConfigParser(inline_comment_prefixes=";")
        cp.read(path)
        g = cp.getfloat
        return cls(
            map_size=map_size, num_landmarks=num_landmarks,
            bearing_noise_deg=g("Sensor Model", "bearing_noise"), range_noise=g("Sensor Model", "range_noise"),
            min_bearing_deg=g("Sensor Model", "min_bearing"), max_bearing_deg=g("Sensor Model", "max_bearing"),
            min_range=g("Sensor Model", "min_range"), max_range=g("Sensor Model", "max_range"),
            translation_noise=g("Control Model", "translation_noise"), rotation_noise_deg=g("Control Model", "rotation_noise"),
            max_steps=g("Environment", "max_steps"),
            resolution=g("Virtual Map", "resolution"), sigma0=g("Virtual Map", "sigma0"),
            sigma_x0=g("Simulator", "sigma_x0"), sigma_y0=g("Simulator", "sigma_y0"),
            sigma_theta0_deg=g("Simulator", "sigma_theta0"),
            angle_weight=g("Planner", "angle_weight"), distance_weight0=g("Planner", "distance_weight0"),
            distance_weight1=g("Planner", "distance_weight1"), max_edge_length=g("Planner", "max_edge_length"),
            occupancy_threshold=g("Planner", "occupancy_threshold"),
        )

    @property
    def n_landmarks(self) -> int:
        return int(self.map_size ** 2 * 0.005) if self.num_landmarks is None else int(self.num_landmarks)

    @property
    def max_plan_actions(self) -> int:
        """Upper bound of a line plan (Planner2D.cpp:982-1038): <= 2 rotation actions + floor(d / max_edge_length) forward steps + the
        remainder step, with d <= the diagonal of the map."""
        side = self.map_size + 2 * self.ext
        return 3 + int(math.hypot(side, side) / self.max_edge_length)

    @property
    def rows(self) -> int:   # VirtualMap.cpp:319-322
        return int(math.floor((self.map_size + 2 * self.ext) / self.resolution))

    @property
    def cols(self) -> int:
        return self.rows

    def to_struct(self) -> DgeConfigStruct:
        s = self.map_size
        c = DgeConfigStruct()
        c.env_min_x, c.env_max_x, c.env_min_y, c.env_max_y = -s / 2, s / 2, -s / 2, s / 2
        c.map_min_x, c.map_max_x = -s / 2 - self.ext, s / 2 + self.ext
        c.map_min_y, c.map_max_y = -s / 2 - self.ext, s / 2 + self.ext
        c.resolution, c.sigma0 = self.resolution, self.sigma0
        c.bearing_noise = _rot2_theta(math.radians(self.bearing_noise_deg))
        c.range_noise = self.range_noise
        c.min_bearing = _rot2_theta(math.radians(self.min_bearing_deg))
        c.max_bearing = _rot2_theta(math.radians(self.max_bearing_deg))
        c.min_range, c.max_range = self.min_range, self.max_range
        c.trans_noise = self.translation_noise
        c.rot_noise = _rot2_theta(math.radians(self.rotation_noise_deg))
        c.sigma_x0, c.sigma_y0 = self.sigma_x0, self.sigma_y0
        c.sigma_theta0 = math.radians(self.sigma_theta0_deg)
        c.angle_weight, c.dist_w0, c.dist_w1 = self.angle_weight, self.distance_weight0, self.distance_weight1
        c.max_edge_length, c.occupancy_threshold = self.max_edge_length, self.occupancy_threshold
        c.max_steps = self.max_steps
        c.relin_thresh, c.relin_skip = self.relinearize_threshold, self.relinearize_skip
        c.num_landmarks = self.n_landmarks
        return c


def start_pose_for_seed(lo: int, map_size: int, ext: float = 20.0):
    """Start pose exactly as pyss2d.py:88-95 draws it (legacy NumPy MT19937, q2):
    three separate ``np.random.seed(lo+k)`` calls and ``randint`` on the *map* bound."""
    import numpy as np

    max_x = map_size / 2 + ext
    st = np.random.get_state()
    try:
        np.random.seed(lo + 1)
        x0 = float(np.random.randint(int(max_x)) - max_x / 2)
        np.random.seed(lo + 2)
        y0 = float(np.random.randint(int(max_x)) - max_x / 2)
        np.random.seed(lo + 3)
        th0 = math.radians(float(np.random.randint(360)))
    finally:
        np.random.set_state(st)
    return x0, y0, th0
