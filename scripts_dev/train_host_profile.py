"""dev: host-side profile of VecDQNTrainer.tick (C3 shapes)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from drl_graph_exploration_b200 import Networks
from drl_graph_exploration_b200.config import EnvConfig
from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
from drl_graph_exploration_b200.trainer import VecDQNTrainer
env = VecExplorationEnv(256, cfg=EnvConfig(map_size=40), max_poses=384, device=0)
env.reset()
torch.manual_seed(0)
pol, tgt = Networks.GCN().to(env.device), Networks.GCN().to(env.device)
tr = VecDQNTrainer(env, pol, tgt, observe=0, overlap=True)
for _ in range(40):
    tr.tick(learn=False)
for _ in range(10):
    tr.tick(learn=True)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(50):
    tr.tick(learn=True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue {1e3 * (t1 - t0) / 50:.2f} ms/tick, with drain {1e3 * (t2 - t0) / 50:.2f} ms/tick")
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    tr.tick(learn=True)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])
