"""Vectorised DQN trainer on the GPU engine (BASELINE config C3 at test size): the acting loop fills the device replay
with well-formed transitions, the minibatch targets equal the literal restatement of policy.py:153-178 on the same
transitions, and gradient steps change the policy while the target net stays put."""
import numpy as np
import pytest
import torch

from drl_graph_exploration_b200.config import EnvConfig

pytestmark = pytest.mark.gpu


def _make(B=24, map_size=20, n_lm=12, seed=0, **kw):
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
    from drl_graph_exploration_b200.trainer import VecDQNTrainer

    env = VecExplorationEnv(B, cfg=EnvConfig(map_size=map_size, num_landmarks=n_lm), max_poses=128, seed0=seed)
    env.reset()
    torch.manual_seed(0)
    pol, tgt = Networks.GCN().to(env.device), Networks.GCN().to(env.device)
    return env, VecDQNTrainer(env, pol, tgt, replay_capacity=512, observe=0, **kw)


def test_acting_loop_fills_replay_with_well_formed_transitions():
    env, tr = _make()
    rp = tr.replay
    for _ in range(40):
        tr.tick(learn=False)
    torch.cuda.synchronize()
    assert tr.decisions > 24 and tr.transitions > 24 and rp.size == min(tr.transitions, rp.capacity)
    n = rp.size
    s, a, r, s1, term = rp.t_s[:n], rp.t_a[:n], rp.t_r[:n], rp.t_s1[:n], rp.t_term[:n]
    # the chosen node is one of the frontier nodes of s_t (policy.py:109-110: a_t[key_size + action_index] = 1)
    assert bool(((a >= rp.gk[s]) & (a < rp.gk[s] + rp.gf[s])).all())
    assert bool((rp.gn[s] == rp.gk[s] + rp.gf[s]).all()) and bool((rp.gn[s1] > 0).all())
    # rewards are the min-max normalised look-ahead rewards of exploration_env.py:154-161
    assert bool(((r >= -1.0) & (r <= 1.0)).all()) and float(r.abs().max()) > 0
    # no stored graph was overwritten under a live transition
    assert bool((rp.gserial[s] == rp.t_serial[:n, 0]).all()) and bool((rp.gserial[s1] == rp.t_serial[:n, 1]).all())
    assert int(term.sum()) > 0 and int((~term).sum()) > 0
    # stored graphs are the engine's graphs: symmetric COO with local node ids inside the graph
    b, nn, off = rp.gather(s[:16])
    ei = b.edge_index
    assert int(ei.min()) >= 0 and int(ei.max()) < b.x.size(0)
    assert bool((b.batch[ei[0]] == b.batch[ei[1]]).all())
    N = b.x.size(0)                                                # data_process emits both directions (policy.py:222-226)
    assert torch.equal(torch.sort(ei[0] * N + ei[1]).values, torch.sort(ei[1] * N + ei[0]).values)
    assert bool((b.x[:, 4].abs() <= 1).all())                     # node type column: -1 / 0 / +1
    env.close()


def test_minibatch_targets_equal_literal_restatement_and_learning_moves_the_policy():
    from drl_graph_exploration_b200.data import Data
    from drl_graph_exploration_b200.trainer import dqn_targets

    env, tr = _make(seed=100)
    for _ in range(30):
        tr.tick(learn=False)
    dq, rp = tr.dqn, tr.replay
    k = min(dq.BATCH, rp.size)
    assert k >= 16
    s, a, r, s1, term = rp.sample(k, generator=tr.gen, check=True)
    b_s, n_s, off_s = rp.gather(s)
    b_s1, n_s1, off_s1 = rp.gather(s1)
    with torch.no_grad():
        q1 = dq.test(b_s1, 0.0, tr.dev, tr.target_net).view(-1)
    act, y = dqn_targets(q1, b_s1.batch, a, r, term, off_s, n_s1, off_s1, rp.gf[s1], b_s.x.size(0), dq.GAMMA)
    # policy.py:153-178 literally, on the host
    q1h, a_ref, y_ref, start = q1.cpu().numpy(), [], [], 0
    for i in range(k):
        n0, n1 = int(n_s[i]), int(n_s1[i])
        onehot = np.zeros(n0); onehot[int(a[i])] = 1
        ty = np.zeros(n0)
        if bool(term[i]):
            ty[int(a[i])] = float(r[i])
        else:
            ty[int(a[i])] = float(r[i]) + dq.GAMMA * np.max(q1h[start:start + n1][-int(rp.gf[s1[i]]):])
        start += n1
        a_ref.append(onehot); y_ref.append(ty)
    assert np.array_equal(act.cpu().numpy(), np.concatenate(a_ref))
    assert np.allclose(y.cpu().numpy(), np.concatenate(y_ref), rtol=1e-5, atol=1e-6)
    # gradient steps: loss finite, policy moves, target stays until TARGET_UPDATE
    before = {k_: v.clone() for k_, v in tr.policy_net.state_dict().items()}
    tgt_before = {k_: v.clone() for k_, v in tr.target_net.state_dict().items()}
    dq.BATCH = k
    losses = [tr.learn(check=True) for _ in range(3)]
    assert all(np.isfinite(l) for l in losses)
    assert any(not torch.equal(before[k_], v) for k_, v in tr.policy_net.state_dict().items())
    assert all(torch.equal(tgt_before[k_], v) for k_, v in tr.target_net.state_dict().items())
    # the clamp of policy.py:251-252 was applied to what Adam saw
    assert float(tr.dqn._bucket.flat.abs().max()) <= dq.max_grad_norm + 1e-7
    # keeps running with learning inside the tick
    for _ in range(5):
        tr.tick()
    assert tr.train_steps >= 8
    env.close()


def test_overlapped_learning_equals_sequential_learning():
    """overlap=True runs the gradient step of a tick on a second stream beside the roll-out kernels.  It is the sequential
    schedule shifted by one tick, so k sequential learning ticks and (one acting tick + k overlapped ticks) must leave
    bit-identical weights -- a missing event between the streams (weights, replay) would break the equality."""
    def run(overlap):
        torch.manual_seed(7); torch.cuda.manual_seed(7)
        env, tr = _make(B=16, seed=500, overlap=overlap)
        tr.dqn.BATCH = 16
        for _ in range(25):
            tr.tick(learn=False)
        assert tr.replay.size >= 16
        if overlap:
            tr.tick(learn=False)
        for _ in range(6):
            tr.tick(learn=True)
        torch.cuda.synchronize()
        w = torch.cat([p.detach().flatten().clone() for p in tr.policy_net.parameters()])
        out = (w, tr.train_steps, tr.last_loss)
        env.close()
        return out
    w_seq, n_seq, l_seq = run(False)
    w_ovl, n_ovl, l_ovl = run(True)
    assert n_seq == n_ovl == 6
    assert torch.equal(w_seq, w_ovl) and l_seq == l_ovl


def test_tail_schedule_equals_sequential_learning():
    """overlap="tail": the gradient step is issued at the end of its tick on the second stream and runs beside the next tick's step /
    roll-out kernels; the order of operations is the sequential one, so the weights after k ticks are bit-identical -- and the loss read
    right after a tick is that tick's (the read joins the step in flight)."""
    def run(overlap):
        torch.manual_seed(7); torch.cuda.manual_seed(7)
        env, tr = _make(B=16, seed=500, overlap=overlap)
        tr.dqn.BATCH = 16
        for _ in range(25):
            tr.tick(learn=False)
        losses = []
        for _ in range(6):
            tr.tick(learn=True)
            losses.append(tr.last_loss)
        torch.cuda.synchronize()
        w = torch.cat([p.detach().flatten().clone() for p in tr.policy_net.parameters()])
        env.close()
        return w, tr.train_steps, losses
    w_seq, n_seq, l_seq = run(False)
    w_tail, n_tail, l_tail = run("tail")
    assert n_seq == n_tail == 6 and l_seq == l_tail
    assert torch.equal(w_seq, w_tail)


def test_vectorised_a2c_trainer_runs_segments_and_learns():
    """VecA2CTrainer (A2C.running for B envs): segments of nstep decisions close, one gradient step per tick with closed
    segments, actor and critic move, sampled actions are frontier nodes, values / rewards stored per transition are finite."""
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
    from drl_graph_exploration_b200.policy import A2C
    from drl_graph_exploration_b200.trainer import VecA2CTrainer

    env = VecExplorationEnv(16, cfg=EnvConfig(map_size=20, num_landmarks=12), max_poses=128, seed0=40)
    env.reset()
    torch.manual_seed(0)
    actor, critic = Networks.PolicyGCN().to(env.device), Networks.ValueGCN().to(env.device)
    before = torch.cat([p.detach().flatten().clone() for p in list(actor.parameters()) + list(critic.parameters())])
    a2c = A2C(); a2c.nstep = 4
    tr = VecA2CTrainer(env, actor, critic, a2c=a2c)
    seen_partial = False
    for _ in range(70):
        tr.tick()
        ln = tr.seg_len
        assert int(ln.max()) < 4 and int(ln.min()) >= 0          # closed segments are consumed in the tick they close
        seen_partial |= bool((ln > 0).any())
        filled = torch.arange(4, device=env.device)[None, :] < ln[:, None]
        s, a = tr.seg_slot[filled], tr.seg_a[filled]
        if s.numel():
            assert bool(((a >= tr.store.gk[s]) & (a < tr.store.gk[s] + tr.store.gf[s])).all())      # the sampled node is a frontier node
            assert torch.isfinite(tr.seg_val[filled]).all() and bool((tr.seg_r[filled].abs() <= 1.0 + 1e-6).all())
    torch.cuda.synchronize()
    assert seen_partial and tr.segments >= 4 and tr.train_steps >= 2 and tr.decisions > 40
    assert np.isfinite(tr.last_loss) and np.isfinite(tr.last_entropy) and tr.last_entropy > 0
    after = torch.cat([p.detach().flatten() for p in list(actor.parameters()) + list(critic.parameters())])
    assert not torch.equal(before, after) and torch.isfinite(after).all()
    env.close()


def test_training_run_leaves_the_reference_artefacts_and_resumes(tmp_path):
    """VecDQNTrainer.run writes what DeepQ.running writes (policy.py:192-209): reward / loss CSVs and state dicts that load back
    into the drop-in modules (and, by the layout test, into the reference's)."""
    from drl_graph_exploration_b200 import Networks
    env, tr = _make(B=16, seed=900)
    tr.dqn.BATCH = 8
    rewards, losses = tr.run(45, out_dir=str(tmp_path), log_every=20)
    assert len(losses) > 0 and len(rewards) > 0 and all(-1.0 <= r <= 1.0 for _, r in rewards)
    for name in ("temp_reward.csv", "temp_loss.csv", "reward_data.csv", "Model_Policy.pt", "Model_Target.pt"):
        assert (tmp_path / name).exists(), name
    assert (tmp_path / "reward_data.csv").read_text().splitlines()[0] == "Step,Reward"
    assert np.loadtxt(tmp_path / "temp_loss.csv", delimiter=",").reshape(-1, 2).shape[0] == len(losses)
    sd = torch.load(tmp_path / "Model_Policy.pt")
    fresh = Networks.GCN()
    fresh.load_state_dict(sd)
    assert all(torch.equal(sd[k].cpu(), v.detach().cpu()) for k, v in tr.policy_net.state_dict().items())
    tr.load(str(tmp_path / "Model_Policy.pt"))
    assert all(torch.equal(a, b) for a, b in zip(tr.policy_net.state_dict().values(), tr.target_net.state_dict().values()))
    env.close()


def test_lag_schedule_takes_the_same_number_of_steps_and_learns():
    """overlap="lag": the acting forward reads the weights as of the previous update, the tick's update is issued behind it (a whole
    tick of slack for the gradient all-reduce).  The acting policy is one update staler, so trajectories differ from the other
    schedules; what must hold: one gradient step per learning tick, finite losses, weights that moved, a replay that stays intact,
    and an update that is never skipped (the step counter of the optimizer equals the number of learning ticks)."""
    torch.manual_seed(7); torch.cuda.manual_seed(7)
    env, tr = _make(B=16, seed=500, overlap="lag")
    tr.dqn.BATCH = 16
    for _ in range(25):
        tr.tick(learn=False)
    w0 = torch.cat([p.detach().flatten().clone() for p in tr.policy_net.parameters()])
    losses = []
    for _ in range(8):
        tr.tick(learn=True)
        losses.append(tr.last_loss)
    tr.tick(learn=False)                       # a tick without a step still orders its acting forward behind the last update
    torch.cuda.synchronize()
    w1 = torch.cat([p.detach().flatten() for p in tr.policy_net.parameters()])
    assert tr.train_steps == 8 and int(tr.optimizer.step_t) == 8 and tr.dqn._pending_apply is None
    assert np.isfinite(losses).all() and not torch.equal(w0, w1) and torch.isfinite(w1).all()
    tr.replay.assert_intact()
    env.close()
