"""cProfile of the C3 training tick (trainer.VecDQNTrainer.tick): where the host time of a tick goes (dev probe)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
for _ in range(20):
    tr.tick(learn=True)
torch.cuda.synchronize()
N = 100
t0 = time.perf_counter()
for _ in range(N):
    tr.tick(learn=True)
torch.cuda.synchronize()
print(f"plain: {1e3 * (time.perf_counter() - t0) / N:.3f} ms / tick")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    tr.tick(learn=True)
pr.disable()
torch.cuda.synchronize()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
