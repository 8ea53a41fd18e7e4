"""debug: per-step comparison of GPU clone roll-outs vs oracle prefixes (run on the GPU box)."""
import sys, os, math
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
from drl_graph_exploration_b200.config import EnvConfig
from test_graph_gnn_parity_gpu import _drive
from helpers import borderline_cells

cfg = EnvConfig(map_size=20, num_landmarks=30)
B = 2
gen = _drive(cfg, B, 2)
env, oracles = next(gen); env, oracles = next(gen)
torch.cuda.synchronize()
g = env.build_graph(); g.sync_sizes()
slots = B * (env.eng.Lt + 1)
for nsteps in (0, 1, 2, 3, 4, 6, 9):
    noise = torch.zeros(max(nsteps, 1), slots, env.eng.noise_len, dtype=torch.float64, device=env.device)
    if nsteps == 0:
        noise = noise[:0]
    raw, norm, clo = env.rollout_rewards(clone_slots=slots, noise=noise)
    torch.cuda.synchronize()
    roll = env._roll
    cm = roll.state["metrics"].cpu().numpy(); rd = None
    o = oracles[0]; r = o.graph()
    print("nsteps", nsteps, "U0 gpu", env.eng.state["metrics"][0, 1].item(), "oracle", o.metrics()["utility0"])
    for f in range(r["fro_size"]):
        acts = o.line_plan(*r["frontier_xy"][f])
        k = min(nsteps, len(acts))
        ref = o.sim_reward(acts[:k], noise=np.zeros((k, o.noise_len))) if k > 0 else 0.0
        print("   f", f, "len", len(acts), "gpu raw", raw[0, f].item(), "ref(prefix)", ref, "clone T", int(roll.state["n_poses"][f]), "sumtrace", cm[f, 1], "w", cm[f, 2],
              "cursor", int(roll.state["plan_cursor"][f]), "plan", roll.state["plan"][f].cpu().numpy())
