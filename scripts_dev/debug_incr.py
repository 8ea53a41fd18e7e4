"""dev: dump roll-out rewards + clone-engine state after a roll-out (run with DGE_SLAM_INCREMENTAL=0/1 and diff)."""
import sys, os, math
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from drl_graph_exploration_b200.config import EnvConfig
from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
out = sys.argv[1]
cfg = EnvConfig(map_size=20, num_landmarks=30)
B = 4
env = VecExplorationEnv(B, cfg=cfg, max_poses=96, device=0, seed0=0)
env.reset()
odom = torch.tensor([[1.0, 0.0, 0.3]] * B, dtype=torch.float64, device=env.device)
for _ in range(6):
    env.eng.step(odom)
g = env.build_graph(); g.sync_sizes()
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 17
slots = B * (env.eng.Lt + 1)
noise = torch.zeros(n_steps, slots, env.eng.noise_len, dtype=torch.float64, device=env.device)
raw, norm, clo = env.rollout_rewards(clone_slots=slots, noise=noise)
torch.cuda.synchronize()
st = env._roll.state
np.savez(out, raw=raw.cpu().numpy(), n_poses=st["n_poses"].cpu().numpy(), est=st["est_pose"].cpu().numpy(), cov=st["pose_cov"].cpu().numpy(),
         metrics=st["metrics"].cpu().numpy(), est_l=st["est_l"].cpu().numpy(), status=st["status"].cpu().numpy(), uc=st["update_count"].cpu().numpy(),
         fro=g.fro_size.cpu().numpy(), obs=st["observed"].cpu().numpy())
