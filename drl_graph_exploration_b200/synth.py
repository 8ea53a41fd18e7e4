"""Synthetic belief states for the covariance-propagation kernel (BASELINE config C4 / SURVEY 8(d)):
random-walk trajectories of 2 m steps inside the env bounds, random SPD pose covariances with
sigma_xy in [0.05, 0.5] m and sigma_theta in [0.002, 0.05] rad, uniform landmarks."""
import numpy as np


def synth_states(cfg, n, T, L, seed=0):
    rng = np.random.default_rng(seed)
    half = cfg.map_size / 2
    pose = np.zeros((n, T, 3))
    p = np.stack([rng.uniform(-half + 1, half - 1, n), rng.uniform(-half + 1, half - 1, n), rng.uniform(-np.pi, np.pi, n)], axis=1)
    for k in range(T):
        pose[:, k] = p
        th = p[:, 2] + rng.normal(0, 0.6, n)
        nx, ny = p[:, 0] + 2.0 * np.cos(th), p[:, 1] + 2.0 * np.sin(th)
        out = (np.abs(nx) > half) | (np.abs(ny) > half)
        th = np.where(out, th + np.pi, th)
        p = np.stack([np.clip(p[:, 0] + 2.0 * np.cos(th), -half, half), np.clip(p[:, 1] + 2.0 * np.sin(th), -half, half),
                      np.arctan2(np.sin(th), np.cos(th))], axis=1)
    A = rng.normal(size=(n, T, 3, 3))
    sig = np.stack([rng.uniform(0.05, 0.5, (n, T)), rng.uniform(0.05, 0.5, (n, T)), rng.uniform(0.002, 0.05, (n, T))], axis=-1)
    C = A @ A.transpose(0, 1, 3, 2) + 0.5 * np.eye(3)
    d = np.sqrt(np.diagonal(C, axis1=-2, axis2=-1))
    cov = C / (d[..., :, None] * d[..., None, :]) * (sig[..., :, None] * sig[..., None, :])
    info = np.linalg.inv(cov)
    lm = rng.uniform(-half, half, (n, L, 2))
    cov6 = np.stack([cov[..., 0, 0], cov[..., 0, 1], cov[..., 0, 2], cov[..., 1, 1], cov[..., 1, 2], cov[..., 2, 2]], axis=-1)
    return pose, cov, cov6, info, lm
