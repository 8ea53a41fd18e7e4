"""dev: GG-NN forward at inference, native GRU path vs torch GRUCell (library) on a C5-shaped batch."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import bench
from drl_graph_exploration_b200 import Networks
from drl_graph_exploration_b200.data import Data
dev = torch.device("cuda"); rng = np.random.default_rng(0)
x, ei, w, bt = bench.synth_graph_batch(64, rng.choice(np.arange(8, 513, 8), size=64), rng, dev)
torch.manual_seed(0)
model = Networks.GGNN().to(dev).eval()
def run(mode, n=20):
    Networks.set_matmul_precision(mode)
    with torch.no_grad():
        for _ in range(3): model(Data(x, ei, w, bt), 0.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): q = model(Data(x, ei, w, bt), 0.0)
        b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, q
t_tc, q_tc = run("tc3"); t_lib, q_lib = run("fp32")
N = x.size(0)
print(f"GG-NN forward, {N} nodes: native (tcgen05 3xTF32 + dge_gru_gates) {t_tc:.3f} ms = {64 / t_tc * 1e3:.0f} graphs/s, {42e6 * N / t_tc / 1e9:.1f} TFLOP/s fp32-equiv | library fp32 GEMM + torch GRUCell {t_lib:.3f} ms = {64 / t_lib * 1e3:.0f} graphs/s | max |dq| {float((q_tc - q_lib).abs().max()):.2e} of {float(q_lib.abs().max()):.2e}")
