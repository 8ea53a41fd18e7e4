"""dev (torchrun, 2 ranks): VecA2CTrainer with the per-tick segment-weighted all-reduce -- replicas must stay identical."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch, torch.distributed as dist
from drl_graph_exploration_b200 import Networks
from drl_graph_exploration_b200.config import EnvConfig
from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
from drl_graph_exploration_b200.policy import A2C
from drl_graph_exploration_b200.trainer import VecA2CTrainer
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
env = VecExplorationEnv(32, cfg=EnvConfig(map_size=20, num_landmarks=12), max_poses=128, device=local, seed0=1000 * rank)
env.reset()
torch.manual_seed(0)
actor, critic = Networks.PolicyGCN().to(env.device), Networks.ValueGCN().to(env.device)
a2c = A2C(); a2c.nstep = 4
tr = VecA2CTrainer(env, actor, critic, a2c=a2c, seed=rank)
for _ in range(60):
    tr.tick()
w = torch.cat([p.detach().flatten() for p in list(actor.parameters()) + list(critic.parameters())])
ws = [torch.empty_like(w) for _ in range(dist.get_world_size())]
dist.all_gather(ws, w)
same = all(torch.equal(ws[0], x) for x in ws)
steps = torch.tensor([float(tr.train_steps), float(tr.segments)], device=env.device)
all_steps = [torch.empty_like(steps) for _ in range(dist.get_world_size())]
dist.all_gather(all_steps, steps)
if rank == 0:
    print("A2C N=2: replicas identical:", same, "| (train steps, segments) per rank:", [t.tolist() for t in all_steps], "| loss", tr.last_loss)
    assert same and all_steps[0][0] == all_steps[1][0] and all_steps[0][0] > 0
dist.destroy_process_group()
