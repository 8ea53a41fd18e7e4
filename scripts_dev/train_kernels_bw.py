"""dev: the HBM-bound helper kernels of the families' training path, each timed alone (CUDA events, L2 flushed): achieved GB/s of their
algorithmic bytes against the measured HBM peak (MEASURED_PEAKS.json)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from drl_graph_exploration_b200 import gnn
dev = torch.device("cuda", 0)
pk, kind = bench.peaks()
peak = pk.get("hbm_gbs", 6544.0)
flush = bench.L2Flush(dev)
def timed(fn, reps=12):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e-3
N = 16896
rows = []
for C in (1000, 3000):
    x = torch.randn(N, C, device=dev)
    t = timed(lambda: gnn.split_transpose(x))
    rows.append((f"k_split_transpose [{N},{C}] -> hi, lo, hi^T, lo^T", 5 * N * C * 4, t))
    t = timed(lambda: gnn.split_transpose(x, plain=False))
    rows.append((f"k_split_transpose [{N},{C}] -> hi^T, lo^T", 3 * N * C * 4, t))
    t = timed(lambda: gnn.split_tf32(x))
    rows.append((f"k_split_tf32 [{N},{C}]", 3 * N * C * 4, t))
    t = timed(lambda: gnn.colsum(x))
    rows.append((f"k_colsum_part + final [{N},{C}]", N * C * 4, t))
C = 1000
rnn = torch.nn.GRUCell(C, C).to(dev)
gi, gh, h, go = torch.randn(N, 3 * C, device=dev), torch.randn(N, 3 * C, device=dev), torch.randn(N, C, device=dev), torch.randn(N, C, device=dev)
t = timed(lambda: gnn._GruGatesFn.forward(type("c", (), {"save_for_backward": lambda *a: None})(), gi, gh, rnn.bias_ih, rnn.bias_hh, h))
rows.append((f"k_gru_gates [{N},{C}]", (7 + 1) * N * C * 4, t))
L = gnn._gemm_lib()
dgi, dgh, dh = torch.empty_like(gi), torch.empty_like(gh), torch.empty_like(h)
bi, bh = rnn.bias_ih.detach().contiguous(), rnn.bias_hh.detach().contiguous()
t = timed(lambda: L.dge_gru_gates_bwd(N, C, gnn._p(gi), gnn._p(gh), gnn._p(bi), gnn._p(bh), gnn._p(h), gnn._p(go), gnn._p(dgi), gnn._p(dgh), gnn._p(dh), gnn._st(dev)))
rows.append((f"k_gru_gates_bwd [{N},{C}]", (8 + 7) * N * C * 4, t))
print(f"measured HBM peak {peak:.0f} GB/s ({kind})")
print("| kernel | algorithmic MB | us | GB/s | frac of peak |\n|---|---|---|---|---|")
for name, by, t in rows:
    print(f"| `{name}` | {by / 1e6:.1f} | {t * 1e6:.1f} | {by / t / 1e9:.0f} | {by / t / 1e9 / peak:.2f} |")
