"""`ExplorationEnv(40, 0, True)` -- the product drop-in alone, no oracle, no test harness -- IS the world of the reference's
`test.py` seed-0 run: landmarks, visiting order and every noise draw come from the product's restatement of the reference's
libstdc++ streams (csrc/dge_refworld.cu), so driving it with test.py's loop (scripts/test.py:100-143: graph -> DQN+GCN with the
shipped weights -> arg-max frontier -> line plan -> steps) reproduces the rows of data/test_result/40_DQN_GCN.csv."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("map_size,seed,n_min", [(40, 0, 25), (40, 1, 25), (60, 2, 20)])
def test_drop_in_test_world_follows_the_reference_result_file(map_size, seed, n_min):
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.envs.exploration_env import ExplorationEnv
    from drl_graph_exploration_b200.policy import DeepQ

    g0 = np.load(os.path.join(GOLD, "ref_40_DQN_GCN_seed0.npz"))
    sd = {k[3:]: torch.tensor(g0[k]) for k in g0.files if k.startswith("sd_")}
    if (map_size, seed) == (40, 0):
        gold = np.stack([g0["landmark_error"], g0["entropy"], g0["max_unc"]], axis=1)
    else:
        gold = np.load(os.path.join(GOLD, "ref_DQN_GCN_multi.npz"))[f"g_{map_size}_{seed}"]
    env = ExplorationEnv(map_size, seed, True, max_poses=200)
    dev = env._vec.device
    model = Networks.GCN().to(dev)
    model.load_state_dict(sd)
    dq = DeepQ()
    step, worst, left = 0, 0.0, None
    while step < n_min + 10 and left is None:
        adj, feats, _, fro = env.graph_matrix()
        data = dq.data_process((adj, feats))
        q = dq.test(data, 0.0, dev, model).view(-1).cpu().numpy()
        key = env.get_key_size()
        a = int(np.argmax(q[-fro:])) if fro > 0 else 0                      # policy.py:109: arg-max over the frontier nodes
        for act in env.line_plan(key, env._frontier[a]):
            env.step(act)
            gl, _, gm = gold[step]
            dl, dm = abs(env.get_landmark_error() - gl) / gl, abs(env.max_uncertainty_of_trajectory() - gm) / gm
            if not (dl < 1e-5 and dm < 1e-5):
                left = (step, dl, dm)
                break
            worst = max(worst, dl, dm)
            step += 1
    print(f"ExplorationEnv({map_size}, {seed}, True): {step} rows of the reference's result file reproduced, worst rel err {worst:.2e}, left at {left}")
    env.close()
    assert step >= n_min and worst < 1e-5, (step, worst, left)
