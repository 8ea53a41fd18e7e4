import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
from drl_graph_exploration_b200.config import EnvConfig
from test_graph_gnn_parity_gpu import _drive
cfg = EnvConfig(map_size=20, num_landmarks=30)
B = 4
for it, (env, oracles) in enumerate(_drive(cfg, B, 3)):
    torch.cuda.synchronize()
    g = env.build_graph(); g.sync_sizes()
    n_steps = 3 + int(np.hypot(20, 20) / cfg.max_edge_length)
    slots = B * (env.eng.Lt + 1)
    noise = torch.zeros(n_steps, slots, env.eng.noise_len, dtype=torch.float64, device=env.device)
    raw, norm, clo = env.rollout_rewards(clone_slots=slots, noise=noise)
    torch.cuda.synchronize()
    raw = raw.cpu().numpy()
    st = env._roll.state
    print("it", it, "T src", env.eng.state["n_poses"].tolist(), "fro", g.fro_size.tolist()[:B])
    print("  clone n_poses", st["n_poses"].tolist()[:24], "status", st["status"].tolist()[:24], "uc", st["update_count"].tolist()[:24])
    for b, o in enumerate(oracles):
        r = o.graph()
        ref = np.array([o.sim_reward(o.line_plan(*r["frontier_xy"][f]), noise=np.zeros((len(o.line_plan(*r["frontier_xy"][f])), o.noise_len))) for f in range(r["fro_size"])])
        print("  env", b, "raw", np.round(raw[b, :len(ref)], 4), "ref", np.round(ref, 4))
    if it == int(os.environ.get("DBG_IT", "0")):
        np.savez(sys.argv[1], n_poses=st["n_poses"].cpu().numpy(), est=st["est_pose"].cpu().numpy(), cov=st["pose_cov"].cpu().numpy(), obs=st["observed"].cpu().numpy(),
                 src_obs=env.eng.state["observed"].cpu().numpy(), est_l=st["est_l"].cpu().numpy(), land_cov=st["land_cov"].cpu().numpy(), meas_ptr=st["meas_ptr"].cpu().numpy())
        break
