"""BASELINE config C4: covariance-propagation kernel (fused k_vmap_env) roofline sweep.
1024 envs, 60x60 map (V = 2500 cells), 200 landmarks, T in {32..1024} synthetic belief states (tests/synth.py).
Algorithmic bytes per env-rebuild (SURVEY 8(d), fp64 state): 2 * (48 T + 20 V + 8 L).  L2 flushed (write + read, bench.L2Flush) between launches;
the kernel is launched through the C ABI on pre-allocated buffers (no host work between the timing events)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
from synth import synth_states
from drl_graph_exploration_b200.config import EnvConfig
from drl_graph_exploration_b200.engine import load_library, _ptr, _stream_ptr

pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = EnvConfig(map_size=60, num_landmarks=200)
cs = cfg.to_struct()
L, V = 200, cfg.rows * cfg.cols
dev = torch.device("cuda")
lib = load_library()
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from bench import L2Flush
flush = L2Flush(dev)
print(f"| T | us/launch | pair visits/launch | algorithmic MB | GB/s | frac of measured HBM peak ({peak:.0f} GB/s) | Gvisit/s | per-CTA cycles: digest / fold / write-out |")
print("|---|---|---|---|---|---|---|---|")
for T in (32, 64, 128, 256, 512, 1024):
    pose, cov, cov6, info, lm = synth_states(cfg, n, T, L, seed=T)
    tp, tc, tl = (torch.as_tensor(a, device=dev).contiguous() for a in (pose, cov6, lm))
    prob = torch.empty(n, V, dtype=torch.float64, device=dev); vinfo = torch.empty(n, V, 3, dtype=torch.float64, device=dev)
    seen = torch.empty(n, V, dtype=torch.int32, device=dev)
    nws = lib.dge_virtual_map_rebuild_ws_doubles(n, T)
    ws = torch.zeros(nws, dtype=torch.float64, device=dev)
    call = lambda s: lib.dge_virtual_map_rebuild(ctypes.byref(cs), n, T, _ptr(tp), _ptr(tc), L, _ptr(tl), _ptr(prob), _ptr(vinfo), _ptr(s), _ptr(ws), _stream_ptr(dev))
    for _ in range(3):
        assert call(seen) == 0
    torch.cuda.synchronize()
    visits = float(seen.clamp(min=0).sum())
    ts = []
    for _ in range(10):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(None); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    nch = (T + 31) // 32
    nch16 = (T + 15) // 16
    clk = ws[n * T * 12:].view(torch.int64)[:4 * n].view(n, 4).cpu().numpy()
    d = np.diff(clk, axis=1).mean(axis=0)
    by = 2.0 * (48 * T + 20 * V + 8 * L) * n
    gbs = by / (ms * 1e-3) / 1e9
    ph = ws[n * T * 12:].view(torch.int64)[4 * n:12 * n].view(n, 8)[:, :5].double().mean(dim=0).cpu().numpy() / nch16
    print(f"    per chunk cycles (thread 0): gate {ph[0]:.0f} spd {ph[1]:.0f} wait {ph[2]:.0f} fold {ph[3]:.0f} rows+wait {ph[4]:.0f}", file=sys.stderr)
    print(f"| {T} | {ms * 1e3:.1f} | {visits:.3g} | {by / 1e6:.1f} | {gbs:.1f} | {gbs / peak:.4f} | {visits / ms / 1e6:.2f} | {d[0]:.0f} / {d[1]:.0f} / {d[2]:.0f} |")
