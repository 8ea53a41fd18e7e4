"""Host logic of the DQN mirror (no GPU): data_process order, cost, target construction."""
import numpy as np
import torch

from drl_graph_exploration_b200.data import Data
from drl_graph_exploration_b200.policy import DeepQ


def _reference_scan(s_a):
    """policy.py:216-227 described literally: walk the dense matrix row-major, emit (i,j) then (j,i) at the
    first encounter of an undirected pair with non-zero weight."""
    ei, ea, seen = [], [], set()
    n = s_a.shape[0]
    for i in range(n):
        for j in range(n):
            if (i, j) in seen or (j, i) in seen or s_a[i][j] == 0:
                continue
            ei.append([i, j]); ea.append(s_a[i][j])
            if i != j:
                ei.append([j, i]); ea.append(s_a[j][i])
            seen.add((i, j)); seen.add((j, i))
    return np.array(ei).T, np.array(ea)


def test_data_process_order():
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 40):
        a = rng.uniform(0.1, 6, (n, n)) * (rng.uniform(size=(n, n)) < 0.3)
        a = np.triu(a, 1); a = a + a.T
        if n > 2:
            a[2, 2] = 1.5   # a self loop
        x = rng.normal(size=(n, 5))
        d = DeepQ().data_process([a, x])
        ei, ea = _reference_scan(a)
        if ei.size == 0:
            assert d.edge_index.numel() == 0
            continue
        assert np.array_equal(d.edge_index.numpy(), ei)
        assert np.allclose(d.edge_attr.numpy(), ea.astype(np.float32))
        assert d.x.dtype == torch.float32 and d.edge_index.dtype == torch.long


def test_cost_and_targets():
    dq = DeepQ()
    pred = torch.tensor([[1.0], [2.0], [3.0]]); a = torch.tensor([0.0, 1.0, 0.0]); y = torch.tensor([0.0, 0.5, 0.0])
    assert torch.isclose(dq.cost(pred, y, a), torch.tensor((2.0 - 0.5) ** 2 / 64))

    class Net(torch.nn.Module):            # target net stand-in: Q = first feature
        def forward(self, data, prob, batch=None):
            return data.x[:, :1]

    def graph(vals):
        n = len(vals)
        return Data(torch.tensor([[v, 0, 0, 0, 0] for v in vals], dtype=torch.float), torch.zeros(2, 0, dtype=torch.long), torch.zeros(0))

    mb = [(graph([0, 0, 0]), np.array([0.0, 0, 1]), 0.25, graph([9, 1, 4, 2]), False, 2),     # max over last 2 nodes = 4
          (graph([0, 0]), np.array([0.0, 1]), -1.0, graph([5, 7, 8]), True, 1)]                # terminal: y = r
    _, a, y = dq.build_targets(mb, torch.device("cpu"), Net())
    assert a.tolist() == [0, 0, 1, 0, 1]
    assert np.allclose(y.numpy(), [0, 0, 0.25 + 0.99 * 4, 0, -1.0])


def test_a2c_objectives_and_nstep_batch():
    """A2C host logic against the formulas of policy.py:361-393,452-472 written out literally."""
    from drl_graph_exploration_b200.policy import A2C
    ac = A2C()
    rng = np.random.default_rng(3)

    def graph(n):
        return Data(torch.tensor(rng.normal(size=(n, 5)), dtype=torch.float), torch.zeros(2, 0, dtype=torch.long), torch.zeros(0))

    sizes, fros = [4, 3, 5], [2, 1, 3]
    acts, rs, terms, vals = [], [0.5, -1.0, 0.25], [False, True, False], [0.1, -0.2, 0.3]
    for n, f in zip(sizes, fros):
        a = np.zeros(n); a[n - f + int(rng.integers(0, f))] = 1
        acts.append(a)
    for i in range(3):
        ac.buffer.append((graph(sizes[i]), acts[i], rs[i], graph(sizes[i]), terms[i], fros[i], vals[i]))
    batch, a, mask, returns, adv = ac.nstep_batch(last_value=2.0)
    g2 = 0.25 + 0.99 * 2.0; g1 = -1.0; g0 = 0.5 + 0.99 * g1           # terminal at i = 1 cuts the bootstrap
    assert np.allclose(returns, [g0, g1, g2])
    assert np.array_equal(a, np.concatenate(acts))
    assert mask.tolist() == [0, 0, 1, 1, 0, 0, 1, 0, 0, 1, 1, 1]
    exp_adv = np.concatenate([acts[i] * ([g0, g1, g2][i] - vals[i]) for i in range(3)])
    assert np.allclose(adv, exp_adv)
    assert batch.x.shape == (12, 5) and batch.batch.tolist() == [0] * 4 + [1] * 3 + [2] * 5
    # costs
    mk = torch.tensor(mask, dtype=torch.bool)
    prob = torch.tensor([0.3, 0.7, 1.0, 0.2, 0.5, 0.3])               # masked softmax output: frontier nodes only
    pc = ac.policy_cost(prob, torch.tensor(adv), torch.tensor(a), mk)
    lit = 0.0
    pa, pv = np.asarray(a)[mask.astype(bool)], np.asarray(adv)[mask.astype(bool)]
    for j in range(6):
        lit += -np.log(float(prob[j])) * pv[j] * pa[j]
    assert np.isclose(float(pc), lit / 40)
    assert np.isclose(float(ac.entropy_loss(prob)), -float((prob.log() * prob).sum()) / 40)
    assert np.isclose(float(ac.value_cost(torch.tensor([1.0, 2.0]), torch.tensor([0.0, 4.0]))), (1 + 4) / 2)


def test_state_dict_layouts_match_the_shipped_checkpoints():
    """Networks.py drop-in contract: the six shipped data/torch_weights/*/MyModel.pt load unchanged, i.e. parameter names and
    shapes are the reference's (fixture tests/golden/ref_state_dict_layouts.json, written by make_golden.py)."""
    import json
    import os
    from drl_graph_exploration_b200 import Networks
    lay = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_state_dict_layouts.json")))
    unet = dict(in_channels=5, hidden_channels=1000, out_channels=1000, depth=3)          # test.py:40,54
    cases = {"DQN_GCN": Networks.GCN(), "DQN_GG-NN": Networks.GGNN(), "DQN_g-U-Net": Networks.GraphUNet(**unet),
             "A2C_GCN": Networks.PolicyGCN(), "A2C_GG-NN": Networks.PolicyGGNN(), "A2C_g-U-Net": Networks.PolicyGraphUNet(**unet)}
    for name, model in cases.items():
        ours = {k: list(v.shape) for k, v in model.state_dict().items()}
        assert ours == lay[name], (name, sorted(set(ours) ^ set(lay[name])))
        # and what torch.save(policy_net.state_dict(), ...) writes (policy.py:192,205) has the same layout
        sd = {k: torch.zeros(s) for k, s in lay[name].items()}
        model.load_state_dict(sd)


def test_env_config_defaults_are_the_reference_ini_values(tmp_path):
    """EnvConfig() carries the values of scripts/envs/exploration_env.ini; from_ini reads a file in that format
    (utils.py:42-45 load_config) and gives the same configuration; ExplorationEnv.reset's overrides (exploration_env.py:399-407)
    are the map-size dependent bounds and the landmark count int(map_size^2 * 0.005)."""
    from drl_graph_exploration_b200.config import EnvConfig
    ini = tmp_path / "exploration_env.ini"
    ini.write_text("""
[Sensor Model]
bearing_noise = 0.5 ; (degree)
range_noise   = 0.02
min_bearing   = -179.9
max_bearing   =  179.9
min_range     = 0.1
max_range     = 6.0
[Control Model]
translation_noise = 0.1
rotation_noise    = 0.2
[Environment]
min_x = -20
max_x =  20
min_y = -20
max_y =  20
max_steps = 5000
safe_distance = 0.0
[Virtual Map]
resolution  = 2.0
sigma0      = 1.0
num_samples = 1
[Simulator]
seed   = 5
lo = 0
num    = 8
sigma_x0     = 0.05
sigma_y0     = 0.05
sigma_theta0 = 0.01
[Planner]
angle_weight = 0.4
distance_weight0 = 5.0
distance_weight1 = 2.0
max_edge_length = 2.0
occupancy_threshold = 0.4
""")
    for ms, n_lm, rows in ((40, 8, 40), (60, 18, 50), (80, 32, 60), (100, 50, 70), (20, 2, 30)):
        a, b = EnvConfig(map_size=ms), EnvConfig.from_ini(str(ini), ms)
        assert a == b
        assert a.n_landmarks == n_lm and a.rows == rows == a.cols
        c = a.to_struct()
        assert (c.env_min_x, c.env_max_x, c.map_min_x, c.map_max_x) == (-ms / 2, ms / 2, -ms / 2 - 20.0, ms / 2 + 20.0)
        assert abs(c.max_bearing - np.radians(179.9)) < 1e-15 and abs(c.bearing_noise - np.radians(0.5)) < 1e-15
        assert (c.relin_thresh, c.relin_skip) == (0.1, 10)                       # gtsam::ISAM2Params defaults (SLAM2D.cpp:10-12)


class _OneEnv:
    """Duck-typed stand-in for ExplorationEnv (graph_matrix / actions_all_goals / rewards_all_goals / step / status / reset)."""

    def __init__(self, seed=0):
        self.rng = np.random.default_rng(seed)
        self.loop_clo, self.steps, self.resets = False, 0, 0

    def graph_matrix(self):
        k, f = int(self.rng.integers(3, 6)), int(self.rng.integers(1, 4))
        n = k + f
        adj = np.zeros((n, n))
        for i in range(n - 1):
            adj[i, i + 1] = adj[i + 1, i] = self.rng.uniform(0.5, 2.0)
        self._n, self._k, self._f = n, k, f
        return adj, self.rng.normal(size=(n, 5)), np.array([0.0]), f

    def actions_all_goals(self):
        return [[]] * self._k + [["rot", "fwd", "fwd"][: int(self.rng.integers(2, 4))] for _ in range(self._f)]

    def rewards_all_goals(self, all_actions):
        r = np.zeros(self._n)
        r[self._k:] = self.rng.uniform(-1, 1, self._f)
        self.loop_clo = bool(self.rng.integers(0, 2))
        return r

    def step(self, act):
        self.steps += 1
        return None, self.steps % 17 == 0, {}

    def status(self):
        return 0.1

    def reset(self):
        self.resets += 1


class _NodeQ(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.l = torch.nn.Linear(5, 1)

    def forward(self, data, prob, batch=None):
        return self.l(data.x)


def test_deepq_running_follows_the_reference_loop():
    """DeepQ.running on a stand-in env: one transition per decision with a one-hot action on a frontier node, learning starts
    after OBSERVE decisions, epsilon decays per decision, the target net is refreshed every TARGET_UPDATE decisions, the env
    is reset when an episode ends."""
    import random
    random.seed(0); torch.manual_seed(0)
    dq = DeepQ()
    dq.OBSERVE, dq.BATCH, dq.TARGET_UPDATE, dq.EXPLORE, dq.REPLAY_MEMORY = 10, 4, 7, 100, 25
    pol, tgt = _NodeQ(), _NodeQ()
    rows = []
    env, losses = dq.running(pol, tgt, env=_OneEnv(), epochs=40, device=torch.device("cpu"), log=lambda *r: rows.append(r))
    assert dq.step_t == 40 and len(rows) == 40 and dq.total_reward.shape == (40,)
    assert len(dq.buffer) == 25                                        # popleft beyond REPLAY_MEMORY (policy.py:132-133)
    assert len(losses) == 30 and losses[0][0] == 11 and all(np.isfinite(l) for _, l in losses)     # learning from decision OBSERVE + 1 on
    assert abs(dq.epsilon - (0.9 - 30 * 0.9 / 100)) < 1e-12            # one decrement per decision after OBSERVE
    for s_t, a_t, r_t, s_t1, term, fro1 in dq.buffer:
        assert a_t.sum() == 1 and a_t.shape[0] == s_t.x.shape[0] and -1 <= r_t <= 1 and 1 <= fro1 <= 3
        assert s_t.edge_index.shape[1] == 2 * (s_t.x.shape[0] - 1)
    assert env.resets >= 1 and all(r[1] == "bayesian" for r in rows)
    # e-greedy branch
    dq2 = DeepQ(); dq2.OBSERVE, dq2.BATCH = 5, 4
    _, l2 = dq2.running(_NodeQ(), _NodeQ(), env=_OneEnv(1), epochs=12, device=torch.device("cpu"), method="e-greedy")
    assert len(l2) == 7


def test_a2c_running_follows_the_reference_loop():
    """A2C.running on the stand-in env: the actor's masked softmax is sampled, one gradient step every nstep decisions, the
    buffer is cleared after it (policy.py:358-395)."""
    from drl_graph_exploration_b200.policy import A2C
    from test_trainer_logic_cpu import _Actor, _Critic
    torch.manual_seed(0)
    ac = A2C(); ac.nstep = 4
    actor, critic = _Actor(), _Critic()
    before = torch.cat([p.detach().flatten().clone() for p in list(actor.parameters()) + list(critic.parameters())])
    rows = []
    env = ac.running(actor, critic, env=_OneEnv(2), epochs=14, device=torch.device("cpu"), log=lambda *r: rows.append(r))
    assert ac.step_t == 14 and len(rows) == 14 and len(ac.buffer) == 2            # three gradient steps, two transitions waiting
    assert np.isfinite(ac.temp_loss) and ac.entro > 0 and ac.total_reward.shape == (14,)
    after = torch.cat([p.detach().flatten() for p in list(actor.parameters()) + list(critic.parameters())])
    assert not torch.equal(before, after)
    for s_t, a_t, r_t, s_t1, term, fro, val in ac.buffer:
        assert a_t.sum() == 1 and int(np.argmax(a_t)) >= a_t.shape[0] - fro and np.isfinite(val)


def test_compact_line_plans_expand_the_same_way_on_both_host_routes():
    """The compact line plan (n_rot_pi, sign, rot_rem, n_fwd, fwd_rem, n_actions) of dge_line_plan is expanded in three places:
    envs.exploration_env.expand_plan (per-env API), HostPolicyLoop._next_actions (vectorised host loop) and the kernels.  The
    two host routes must produce the same action sequence (Planner2D.cpp:982-1038: pi-rotations, the remainder rotation, full
    forward edges, the remainder edge)."""
    import types
    from drl_graph_exploration_b200.envs.exploration_env import expand_plan
    from drl_graph_exploration_b200.runner import HostPolicyLoop
    rng = np.random.default_rng(5)
    B, edge = 64, 2.0
    nrot = rng.integers(0, 2, B).astype(float); sign = rng.choice([-1.0, 1.0], B); rrem = rng.uniform(0, np.pi, B)
    nfwd = rng.integers(0, 9, B).astype(float); frem = rng.uniform(0, edge, B)
    plans = np.stack([nrot, sign, rrem, nfwd, frem, nrot + 1 + nfwd + 1], axis=1)
    loop = types.SimpleNamespace(np=np, plans=plans, cursor=np.zeros(B, dtype=np.int64), odom=np.zeros((B, 3)),
                                 env=types.SimpleNamespace(cfg=types.SimpleNamespace(max_edge_length=edge)))
    want = [expand_plan(plans[b], edge) for b in range(B)]
    for step in range(int(plans[:, 5].max())):
        loop.cursor[:] = step
        HostPolicyLoop._next_actions(loop)
        for b in range(B):
            if step < int(plans[b, 5]):
                a = want[b][step]
                assert np.allclose(loop.odom[b], [a.x, a.y, a.theta], rtol=0, atol=0), (b, step)
    assert all(len(w) == int(plans[b, 5]) for b, w in enumerate(want))
