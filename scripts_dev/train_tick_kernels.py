"""Device time per kernel over N training ticks (torch.profiler / CUPTI): what the GPU spends a C3 tick on (dev probe)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from drl_graph_exploration_b200 import Networks
from drl_graph_exploration_b200.config import EnvConfig
from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
from drl_graph_exploration_b200.trainer import VecDQNTrainer

B = 256
env = VecExplorationEnv(B, cfg=EnvConfig(map_size=40), max_poses=384, device=0, seed0=0, seed_stride=B)
env.reset()
torch.manual_seed(0)
pol, tgt = Networks.GCN().to(env.device), Networks.GCN().to(env.device)
tr = VecDQNTrainer(env, pol, tgt, observe=0, train_steps_per_tick=1, seed=0, overlap=("--seq" not in sys.argv))
for _ in range(40):
    tr.tick(learn=False)
for _ in range(60):
    tr.tick(learn=True)
torch.cuda.synchronize()
N = 30
d0, r0, k0 = tr.decisions, tr.rollout_steps, tr.rollout_clones
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        tr.tick(learn=True)
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f"{N} ticks, {1e3 * wall / N:.2f} ms / tick under the profiler; decisions/tick {(tr.decisions - d0) / N:.1f}, clone-engine steps/tick {(tr.rollout_steps - r0) / N:.1f}, clones/tick {(tr.rollout_clones - k0) / N:.1f}")
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:28]
tot = sum(e.device_time_total for e in prof.key_averages())
print(f"device time total {tot / N / 1e3:.2f} ms / tick")
for e in rows:
    print(f"{e.key[:90]:90s} n/tick {e.count / N:7.1f}  us/tick {e.device_time_total / N:9.1f}  avg us {e.device_time_total / max(e.count, 1):8.1f}  {100 * e.device_time_total / tot:5.1f}%")
