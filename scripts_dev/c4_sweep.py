"""BASELINE config C4: covariance-propagation kernel (k_vmap_prep + k_vmap_cells) roofline sweep.
1024 envs, 60x60 map (V = 2500 cells), 200 landmarks, T in {32..1024} synthetic belief states (tests/synth.py).
Algorithmic bytes per env-rebuild (SURVEY 8(d), fp64 state): 2 * (48 T + 20 V + 8 L).  L2 flushed between launches."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
from synth import synth_states
from drl_graph_exploration_b200.config import EnvConfig
from drl_graph_exploration_b200.engine import virtual_map_rebuild

peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
cfg = EnvConfig(map_size=60, num_landmarks=200)
n, L, V = 1024, 200, cfg.rows * cfg.cols
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(f"| T | ms/launch | pair visits/launch | algorithmic MB | GB/s | frac of measured HBM peak ({peak:.0f} GB/s) | Gvisit/s |")
print("|---|---|---|---|---|---|---|")
for T in (32, 64, 128, 256, 512, 1024):
    pose, cov, cov6, info, lm = synth_states(cfg, n, T, L, seed=T)
    tp, tc, tl = (torch.as_tensor(a, device=dev) for a in (pose, cov6, lm))
    for _ in range(3):
        prob, vinfo, seen = virtual_map_rebuild(cfg, tp, tc, tl, want_seen=True)
    visits = float(seen.clamp(min=0).sum())
    ts = []
    for _ in range(10):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); virtual_map_rebuild(cfg, tp, tc, tl); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    by = 2.0 * (48 * T + 20 * V + 8 * L) * n
    gbs = by / (ms * 1e-3) / 1e9
    print(f"| {T} | {ms:.3f} | {visits:.3g} | {by / 1e6:.1f} | {gbs:.1f} | {gbs / peak:.4f} | {visits / ms / 1e6:.2f} |")
