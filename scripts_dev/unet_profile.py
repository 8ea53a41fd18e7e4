"""dev: where a g-U-Net forward spends its time on a C5-shaped batch (torch profiler, CUDA time by kernel)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import bench
from drl_graph_exploration_b200 import Networks
from drl_graph_exploration_b200.data import Data
dev = torch.device("cuda"); rng = np.random.default_rng(0)
x, ei, w, bt = bench.synth_graph_batch(64, rng.choice(np.arange(8, 513, 8), size=64), rng, dev)
torch.manual_seed(0)
model = Networks.GraphUNet(5, 1000, 1000, depth=3).to(dev).eval()
with torch.no_grad():
    for _ in range(3): model(Data(x, ei, w, bt), 0.0, batch=bt)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): model(Data(x, ei, w, bt), 0.0, batch=bt)
    b.record(); torch.cuda.synchronize()
    print(f"g-U-Net forward, {x.size(0)} nodes, {ei.size(1)} edges: {a.elapsed_time(b) / 10:.3f} ms / batch = {64 / (a.elapsed_time(b) / 10) * 1e3:.0f} graphs/s")
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3): model(Data(x, ei, w, bt), 0.0, batch=bt)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
