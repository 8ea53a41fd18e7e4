"""The drop-in boundary itself: the reference's single-env class surface on the CUDA engine, driven
by the reference's own loops -- test.py:100-143 (DQN+GCN acting with the shipped weights) and policy.py:297-426
(A2C.running) -- with the reference's call sequence and return types."""
import math
import os

import numpy as np
import pytest
import torch

from drl_graph_exploration_b200.config import start_pose_for_seed

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _shipped_gcn(dev):
    from drl_graph_exploration_b200 import Networks
    g0 = np.load(os.path.join(GOLD, "ref_40_DQN_GCN_seed0.npz"))
    model = Networks.GCN().to(dev)
    model.load_state_dict({k[3:]: torch.tensor(g0[k]) for k in g0.files if k.startswith("sd_")})   # data/torch_weights/DQN_GCN/MyModel.pt
    return model


def _run_test_py_loop(n_decisions):
    """test.py:100-143 with `robot.ExplorationEnv`, `DeepQ.data_process` and `test_q` swapped for this package's."""
    import drl_graph_exploration_b200.envs.exploration_env as robot
    from drl_graph_exploration_b200.policy import DeepQ
    env = robot.ExplorationEnv(40, 0, True)
    device = env._vec.device
    model = _shipped_gcn(device)
    dq = DeepQ()
    log = []
    for _ in range(n_decisions):
        adjacency, features, globals_features, fro_size = env.graph_matrix()
        node_size = adjacency.shape[0]
        key_size = node_size - fro_size
        s_t = dq.data_process([adjacency, features])
        assert key_size == env.get_key_size()
        all_actions = env.actions_all_goals()
        readout_t = dq.test(s_t, 0, device, model).cpu().detach().numpy()
        action_index = int(np.argmax(readout_t[-fro_size:]))
        actions = all_actions[key_size + action_index]
        for act in actions:
            obs, done, info = env.step(act)
            log.append((env.get_landmark_error(), env.max_uncertainty_of_trajectory(), env.status(), float(obs.sum())))
            if done:
                break
        # ---- the shapes / types the callers rely on
        assert adjacency.shape == (node_size, node_size) and adjacency.dtype == np.float64 and np.array_equal(adjacency, adjacency.T)
        assert features.shape == (node_size, 5) and features.dtype == np.float64 and globals_features.shape == (1,)
        assert isinstance(fro_size, int) and 1 <= fro_size <= env.get_landmark_size() + 1
        assert len(all_actions) == node_size and all(a == [] for a in all_actions[:key_size]) and all(len(a) >= 2 for a in all_actions[key_size:])
        assert set(np.unique(features[:, 4])) <= {-1.0, 0.0, 1.0} and (features[key_size:, 4] == 1).all()   # node types: frontier = +1
        assert isinstance(done, bool) and info == {} and obs.shape == (env.leng_i_map, env.leng_j_map)
        assert readout_t.shape == (node_size, 1)
        if done:
            break
    res = dict(log=np.array(log), dist=env.get_dist(), key=env.get_key_size(), lm=env.get_landmark_size())
    # start pose of seed 0 is the reference's (pyss2d.py:88-95); the robot then made the 4 forced (1, 1, pi/2) moves
    assert start_pose_for_seed(0, 40) == (17.0, -5.0, math.radians(152.0))
    env.close()
    return res


def test_reference_test_py_loop_runs_on_the_drop_in_classes_and_is_reproducible():
    a = _run_test_py_loop(6)
    b = _run_test_py_loop(6)
    assert a["log"].shape[0] >= 12 and np.isfinite(a["log"]).all()
    assert np.array_equal(a["log"], b["log"]) and a["dist"] == b["dist"] and a["key"] == b["key"]     # test=True: seeded, deterministic
    assert a["log"][-1, 2] > a["log"][0, 2]          # the policy explores: ExplorationEnv.status() grows
    assert a["lm"] >= 1


def test_rewards_all_goals_and_a2c_running():
    import drl_graph_exploration_b200.envs.exploration_env as robot
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.policy import A2C
    env = robot.ExplorationEnv(40, 3, True)
    adjacency, features, _, fro_size = env.graph_matrix()
    n = adjacency.shape[0]
    all_actions = env.actions_all_goals()
    rewards = env.rewards_all_goals(all_actions)
    assert rewards.shape == (n,) and (rewards[:n - fro_size] == 0).all()
    fr = rewards[n - fro_size:]
    assert fr.min() >= -1.0 - 1e-12 and fr.max() <= 1.0 + 1e-12 and isinstance(env.loop_clo, bool)
    if fro_size > 1 and fr.max() > fr.min():
        assert abs(fr.min() + 1.0) < 1e-9 and abs(fr.max() - (1.0 if env.loop_clo else 0.0)) < 1e-9     # exploration_env.py:154-161
        assert env.loop_clo == (int(np.argmax(fr)) != 0)
    # the roll-outs ran on clones: the env itself did not move
    adjacency2, features2, _, fro2 = env.graph_matrix()
    assert np.array_equal(adjacency, adjacency2) and np.array_equal(features, features2) and fro2 == fro_size
    # A2C.running (policy.py:297-426) for a few decisions, one gradient step every nstep = 3
    dev = env._vec.device
    torch.manual_seed(0)
    actor, critic = Networks.PolicyGCN().to(dev), Networks.ValueGCN().to(dev)
    before = [p.detach().clone() for p in list(actor.parameters()) + list(critic.parameters())]
    a2c = A2C()
    a2c.nstep = 3
    logs = []
    a2c.running(actor, critic, env=env, epochs=7, device=dev, log=lambda *row: logs.append(row))
    assert a2c.step_t == 7 and len(logs) == 7 and len(a2c.buffer) == 1          # two gradient steps, one transition waiting
    assert np.isfinite(a2c.temp_loss) and np.isfinite(a2c.entro) and a2c.total_reward.shape == (7,)
    after = list(actor.parameters()) + list(critic.parameters())
    assert any(not torch.equal(b, a.detach()) for b, a in zip(before, after))
    env.close()
