"""dev: ONE forward + backward + clamp + Adam step of Networks.GGNN on a C5 batch (for an ncu launch list: which kernels a GG-NN training step runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from drl_graph_exploration_b200 import Networks
from drl_graph_exploration_b200.data import Data
from drl_graph_exploration_b200.dist import FlatGradBucket
dev = torch.device("cuda", 0)
rng = np.random.default_rng(1234)
x, ei, w, bt = bench.synth_graph_batch(64, rng.choice(np.arange(8, 513, 8), size=64), rng, dev)
torch.manual_seed(0)
net = Networks.GGNN().to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-5)
bucket = FlatGradBucket(net.parameters())
def step():
    bucket.zero_()
    q = net(Data(x, ei, w, bt), 0.5, batch=bt)
    ((q.view(-1) ** 2).sum() / 64).backward()
    bucket.clamp_(0.5); opt.step()
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("nodes", x.size(0))
