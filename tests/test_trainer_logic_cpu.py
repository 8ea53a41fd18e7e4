"""Host logic of trainer.VecA2CTrainer on CPU tensors: the CUDA engine is replaced by a small stand-in that has the same
surface (mark_pending / step_queued / build_graph / rollout_rewards / select_and_plan and the `done` state), so the
bookkeeping -- in-flight transitions, per-env n-step segments, graph-store serials, the training batch -- is exercised
without a GPU.  (The engine itself and the real kernels are covered by the -m gpu tests.)"""
import types

import numpy as np
import pytest
import torch

from drl_graph_exploration_b200 import trainer as trainer_mod
from drl_graph_exploration_b200.data import Data
from drl_graph_exploration_b200.policy import A2C


class _Graph:
    def __init__(self, B, fmax):
        self.B, self.fmax = B, fmax

    def fill(self, mask, rng):
        xs, eis, eas, bat, nptr, eptr, ks, fs = [], [], [], [], [0], [0], [], []
        for g, b in enumerate(np.nonzero(mask)[0]):
            k, f = int(rng.integers(3, 7)), int(rng.integers(1, self.fmax + 1))
            n = k + f
            src = np.arange(n - 1); dst = src + 1
            ei = np.stack([np.concatenate([src, dst]), np.concatenate([dst, src])]) + nptr[-1]
            xs.append(rng.normal(size=(n, 5))); eis.append(ei); eas.append(rng.uniform(0.1, 2.0, ei.shape[1])); bat.append(np.full(n, g))
            nptr.append(nptr[-1] + n); eptr.append(eptr[-1] + ei.shape[1]); ks.append(k); fs.append(f)
        self.n_graphs = len(ks)
        if self.n_graphs:
            self.x = torch.tensor(np.concatenate(xs), dtype=torch.float32)
            self.edge_index = torch.tensor(np.concatenate(eis, axis=1), dtype=torch.long)
            self.edge_attr = torch.tensor(np.concatenate(eas), dtype=torch.float32)
            self.batch = torch.tensor(np.concatenate(bat), dtype=torch.long)
        pad = lambda v: torch.tensor(v + [0] * (self.B + 1 - len(v)), dtype=torch.int32)
        self.node_ptr, self.edge_ptr = pad(nptr), pad(eptr)
        self.key_size, self.fro_size = pad(ks)[:self.B], pad(fs)[:self.B]

    def sync_sizes(self):
        return self.n_graphs, int(self.node_ptr[self.n_graphs]), int(self.edge_ptr[self.n_graphs])

    def data(self):
        d = Data(self.x, self.edge_index, self.edge_attr, self.batch)
        d.num_graphs = self.n_graphs
        return d


class _Env:
    """B envs whose 'plans' are just counters: a decision queues 2..4 steps, an episode ends with probability 0.04 per step."""

    def __init__(self, B, seed=0):
        self.B, self.device = B, torch.device("cpu")
        self.seed_stride = B                  # (VecExplorationEnv: an env's next world = its seed + seed_stride)
        self.rng = np.random.default_rng(seed)
        self.queue = np.zeros(B, dtype=np.int64)
        self.done = torch.zeros(B, dtype=torch.uint8)
        self.forced = np.zeros(B, dtype=np.int64)
        lib = types.SimpleNamespace(dge_reset_done_queued=self._reset_done)
        self.status = torch.zeros(B, dtype=torch.int32)
        self.eng = types.SimpleNamespace(Lt=3, node_cap_env=12, edge_cap_env=24, state={"done": self.done, "status": self.status}, _L=lib, _h=None, step_queued=self._step)
        from drl_graph_exploration_b200.config import EnvConfig
        self.cfg = EnvConfig(map_size=20)
        self.graph = _Graph(B, 3)
        self.rollout_steps = 0

    def _reset_done(self, *a):
        d = self.done.numpy().astype(bool)
        self.forced[d] = 2; self.queue[d] = 0; self.done[:] = 0
        return 0

    def _step(self):
        moving = (self.queue > 0) & (self.forced == 0)
        self.queue[moving] -= 1
        ends = moving & (self.rng.uniform(size=self.B) < 0.04)
        self.done[torch.tensor(ends)] = 1
        self.forced[self.forced > 0] -= 1

    def mark_pending(self):
        return torch.tensor((self.queue == 0) & (self.forced == 0) & ~self.done.numpy().astype(bool), dtype=torch.uint8)

    def build_graph(self, mask):
        self.graph.fill(mask.numpy().astype(bool), self.rng)
        self._mask = mask.numpy().astype(bool)
        return self.graph

    def rollout_rewards(self, mask, clone_slots=None, auto_steps=False):
        self.rollout_steps = 3
        norm = torch.tensor(self.rng.uniform(-1, 1, (self.B, 4)))
        return norm, norm, torch.tensor(self.rng.integers(0, 2, self.B), dtype=torch.uint8)

    def select_and_plan(self, q, mask):
        g, choice = self.graph, torch.zeros(self.B, dtype=torch.int32)
        for i, b in enumerate(np.nonzero(mask.numpy())[0]):
            lo = int(g.node_ptr[i]) + int(g.key_size[i])
            choice[b] = int(torch.argmax(q.view(-1)[lo:lo + int(g.fro_size[i])]))
            self.queue[b] = int(self.rng.integers(2, 5))
        return choice


class _Actor(torch.nn.Module):      # PolicyGCN surface: masked softmax over the frontier nodes of every graph
    def __init__(self):
        super().__init__()
        self.l = torch.nn.Linear(5, 1)

    def forward(self, data, mask, batch=None):
        s = torch.masked_select(self.l(data.x).view(-1), mask)
        idx = torch.masked_select(batch, mask)
        e = (s - s.max()).exp()
        den = torch.zeros(int(batch.max()) + 1).index_add_(0, idx, e)
        return e / den[idx]


class _Critic(torch.nn.Module):     # ValueGCN surface: one value per graph
    def __init__(self):
        super().__init__()
        self.l = torch.nn.Linear(5, 1)

    def forward(self, data, mask, batch=None):
        v = self.l(data.x).view(-1)
        n = int(batch.max()) + 1
        return torch.zeros(n).index_add_(0, batch, v) / torch.bincount(batch, minlength=n).clamp(min=1)


@pytest.fixture
def cpu_trainer(monkeypatch):
    monkeypatch.setattr(trainer_mod, "_stream_ptr", lambda dev: None)

    def make(B=12, nstep=3, seed=0):
        torch.manual_seed(seed)
        env = _Env(B, seed)
        a2c = A2C(); a2c.nstep = nstep
        return env, trainer_mod.VecA2CTrainer(env, _Actor(), _Critic(), a2c=a2c, lr=1e-2, seed=seed)
    return make


def test_a2c_bookkeeping_segments_serials_and_training(cpu_trainer):
    env, tr = cpu_trainer()
    before = torch.cat([p.detach().flatten().clone() for p in tr.params])
    for _ in range(120):
        tr.tick()
        assert int(tr.seg_len.max()) < 3 and int(tr.seg_len.min()) >= 0
        filled = torch.arange(3)[None, :] < tr.seg_len[:, None]
        s, a = tr.seg_slot[filled], tr.seg_a[filled]
        if s.numel():   # the sampled node is a frontier node of the stored graph, and that graph is still the one stored
            assert bool(((a >= tr.store.gk[s]) & (a < tr.store.gk[s] + tr.store.gf[s])).all())
            assert bool((tr.store.gserial[s] == tr.seg_serial[filled]).all())
    assert tr.segments >= 10 and tr.train_steps >= 5 and tr.decisions > 100
    assert np.isfinite(tr.last_loss) and tr.last_entropy > 0
    after = torch.cat([p.detach().flatten() for p in tr.params])
    assert not torch.equal(before, after) and torch.isfinite(after).all()
    assert bool(tr.seg_term.any())            # episode ends / loop closures were recorded as terminals


def test_a2c_overwritten_graph_is_an_error_not_a_silent_corruption(cpu_trainer):
    from drl_graph_exploration_b200.replay import GraphReplay
    env, tr = cpu_trainer(B=12, nstep=3, seed=1)
    tr.store = GraphReplay(6, env.eng.node_cap_env, env.eng.edge_cap_env, "cpu", slack=2)     # 8 slots for 12 envs x 3 steps: must wrap
    with pytest.raises(RuntimeError, match="overwritten"):
        for _ in range(200):
            tr.tick()


class _QNet(torch.nn.Module):       # Networks.GCN surface: Q per node
    def __init__(self):
        super().__init__()
        self.l = torch.nn.Linear(5, 1)

    def forward(self, data, prob, batch=None):
        return self.l(torch.nn.functional.dropout(data.x, p=float(prob) * 0.1))


def test_dqn_bookkeeping_transitions_targets_and_learning(monkeypatch):
    """VecDQNTrainer (sequential schedule) on the stand-in engine: every closed transition points at a frontier node of its
    stored s_t, s_t1 is the graph of the env's next decision (or s_t itself after a terminal), no stored graph is overwritten
    under a live transition, and gradient steps move the policy while the target net waits for its update period."""
    monkeypatch.setattr(trainer_mod, "_stream_ptr", lambda dev: None)
    torch.manual_seed(0)
    env = _Env(16, seed=3)
    pol, tgt = _QNet(), _QNet()
    tr = trainer_mod.VecDQNTrainer(env, pol, tgt, replay_capacity=64, observe=0, lr=1e-2, seed=0)
    tr.dqn.BATCH = 8
    tr.dqn.TARGET_UPDATE = 1000
    before = [p.detach().clone() for p in pol.parameters()]
    for _ in range(100):
        tr.tick()
    rp = tr.replay
    assert rp.size == 64 and tr.transitions > 64 and tr.train_steps > 50 and np.isfinite(tr.last_loss)
    C = rp.capacity                                                  # (position C of the transition arrays is append_masked's trash slot)
    s, a, r, s1, term = rp.t_s[:C], rp.t_a[:C], rp.t_r[:C], rp.t_s1[:C], rp.t_term[:C]
    assert bool(((a >= rp.gk[s]) & (a < rp.gk[s] + rp.gf[s])).all()) and bool((r.abs() <= 1).all())
    assert bool((rp.gserial[s] == rp.t_serial[:C, 0]).all()) and bool((rp.gserial[s1] == rp.t_serial[:C, 1]).all())
    assert bool(term.any()) and bool((~term).any())
    assert bool((s1[~term] != s[~term]).all())                       # a non-terminal transition ends in a NEW graph
    rp.sample(8, generator=tr.gen, check=True)
    assert any(not torch.equal(b, p.detach()) for b, p in zip(before, pol.parameters()))
    assert all(torch.equal(b, p.detach()) for b, p in zip(before, tgt.parameters()))      # target = the policy as of its last sync (step 0)
    # the pending table only holds envs that are between two decisions
    assert bool(((tr.pend_slot >= 0) | (tr.pend_slot == -1)).all())


def test_dqn_checkpoint_and_resume(monkeypatch, tmp_path):
    """checkpoint() / restore(): nets, optimizer, epsilon / step counters, sampling generator and the device replay survive a
    process boundary (the reference pickles its DeepQ object between subprocess chunks); environments start fresh episodes."""
    monkeypatch.setattr(trainer_mod, "_stream_ptr", lambda dev: None)

    def make(seed):
        torch.manual_seed(seed)
        env = _Env(16, seed=seed)
        tr = trainer_mod.VecDQNTrainer(env, _QNet(), _QNet(), replay_capacity=48, observe=0, lr=1e-2, seed=0)
        tr.dqn.BATCH = 8
        return env, tr
    _, a = make(5)
    for _ in range(60):
        a.tick()
    a.checkpoint(str(tmp_path / "ck.pt"))
    _, b = make(99)                                   # different weights, empty replay, different env
    b.restore(str(tmp_path / "ck.pt"))
    for (k, v), (_, w) in zip(a.policy_net.state_dict().items(), b.policy_net.state_dict().items()):
        assert torch.equal(v, w), k
    for k in ("decisions", "train_steps", "transitions"):
        assert getattr(a, k) == getattr(b, k)
    assert (a.dqn.step_t, a.dqn.epsilon) == (b.dqn.step_t, b.dqn.epsilon)
    ra, rb = a.replay, b.replay
    assert (ra.size, ra.head, ra.allocated) == (rb.size, rb.head, rb.allocated)
    for name in ra._TENSORS:
        assert torch.equal(getattr(ra, name), getattr(rb, name)), name
    assert bool((b.pend_slot == -1).all())
    # the same minibatch is drawn and the same gradient step is taken after the restore (generator and Adam moments restored)
    torch.manual_seed(123); la = a.learn()
    torch.manual_seed(123); lb = b.learn()
    assert la == lb
    for v, w in zip(a.policy_net.parameters(), b.policy_net.parameters()):
        assert torch.equal(v, w)
    for _ in range(10):
        b.tick()                                      # and it keeps running on its own envs
    assert b.train_steps > a.train_steps - 1


def test_dqn_run_leaves_the_reference_artefacts(monkeypatch, tmp_path):
    """VecDQNTrainer.run on the stand-in engine: the reference's training artefacts (policy.py:192-209) appear in a directory
    that does not exist yet -- MyModel.pt is written from inside the loop --, the CSV rows carry the decision count, and the
    saved state dicts load back into a fresh network."""
    monkeypatch.setattr(trainer_mod, "_stream_ptr", lambda dev: None)
    torch.manual_seed(0)
    tr = trainer_mod.VecDQNTrainer(_Env(16, seed=4), _QNet(), _QNet(), replay_capacity=64, observe=0, lr=1e-2, seed=0)
    tr.dqn.BATCH = 8
    out = tmp_path / "run" / "weights"
    rewards, losses = tr.run(60, out_dir=str(out), log_every=25, save_every=40)
    for name in ("temp_reward.csv", "temp_loss.csv", "reward_data.csv", "MyModel.pt", "Model_Policy.pt", "Model_Target.pt"):
        assert (out / name).exists(), name
    assert len(rewards) >= 2 and len(losses) == tr.train_steps > 10
    rw = np.loadtxt(out / "temp_reward.csv", delimiter=",").reshape(-1, 2)
    assert np.all(np.diff(rw[:, 0]) > 0) and rw[-1, 0] <= tr.dqn.step_t and np.all(np.abs(rw[:, 1]) <= 1)
    assert (out / "reward_data.csv").read_text().splitlines()[0] == "Step,Reward"
    net = _QNet()
    net.load_state_dict(torch.load(out / "Model_Policy.pt"))
    for a, b in zip(net.parameters(), tr.policy_net.parameters()):
        assert torch.equal(a, b)
