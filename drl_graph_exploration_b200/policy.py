"""Drop-in for the DQN half of the reference's ``scripts/policy.py`` (class ``DeepQ``,
policy.py:16-259): same hyper-parameters, same ``data_process / cost / train / test`` methods and the
same target computation (policy.py:141-181), on the CUDA GNN kernels.

Differences that are deliberate (B200-first), none of which change a result:
* ``data_process`` is vectorised (the reference scans the dense adjacency with a Python double loop);
  it emits exactly the reference's COO order.  Graphs that come from ``VecExplorationEnv`` are already in
  that layout on the device and skip it.
* targets are built with a segmented max on the device instead of a NumPy loop.
* multi-GPU: one flat-bucket all-reduce of the gradients before the clamp (``dist.FlatGradBucket``).
The environment-facing DQN loop for B envs per GPU (roll-out rewards, device replay) is ``trainer.VecDQNTrainer``;
``DeepQ.running`` / ``A2C.running`` below are the reference's single-env loops on ``ExplorationEnv`` (B = 1 view).

``A2C`` mirrors policy.py:262-497 (n-step advantage actor-critic on the Policy*/Value* heads of ``Networks``).
"""
from __future__ import annotations

import random
from collections import deque

import numpy as np
import torch
import torch.nn.functional as F

from .data import Batch, Data
from . import gnn
from .dist import FlatGradBucket


class DeepQ(object):
    def __init__(self, case_path="DQN_GCN/", model_name="GCN"):
        self.case_path = case_path
        # RL parameters (policy.py:33-58)
        self.BATCH = 64
        self.REPLAY_MEMORY = 1e4
        self.GAMMA = 0.99
        self.OBSERVE = 5e3
        self.EXPLORE = 1e6
        self.epoch = 1e4
        self.TARGET_UPDATE = 15000 if model_name == "GCN" else 9000
        self.FINAL_EPSILON = 0
        self.INITIAL_EPSILON = 0.9
        self.max_grad_norm = 0.5
        self.map_size = 40
        self.buffer = deque()
        self.step_t = 0
        self.epsilon = self.INITIAL_EPSILON
        self.temp_loss = 0
        self.total_reward = np.empty([0, 0])
        self._bucket = None
        self.native_steps = 0      # gradient steps taken on the native path (keys the dropout stream together with torch.initial_seed())
        self._pending_apply = None  # (optimizer, clamp, gscale) of a train(..., apply=False) step whose update has not been issued yet

    # ------------------------------------------------------------------ data ---
    def data_process(self, data):
        """policy.py:211-232: dense adjacency [N,N] + features [N,5] -> Data(x, edge_index, edge_attr);
        edges in row-major first-encounter order, both directions, zero weight = no edge."""
        s_a, s_x = data
        s_a = np.asarray(s_a)
        # the reference meets every undirected pair first at (min, max) when walking the matrix row-major
        # (exploration graphs are symmetric: exploration_env.py:217-224, SLAM2D.cpp:246-259)
        iu, ju = np.nonzero(np.triu(s_a != 0))          # row-major order of the upper triangle (incl. diagonal)
        loops = iu == ju
        n_dir = 2 * len(iu) - int(loops.sum())
        rows, cols, w = np.empty(n_dir, np.int64), np.empty(n_dir, np.int64), np.empty(n_dir, np.float64)
        pos = np.cumsum(np.where(loops, 1, 2)) - np.where(loops, 1, 2)
        rows[pos], cols[pos], w[pos] = iu, ju, s_a[iu, ju]
        nl = ~loops
        rows[pos[nl] + 1], cols[pos[nl] + 1], w[pos[nl] + 1] = ju[nl], iu[nl], s_a[ju[nl], iu[nl]]
        edge_index = torch.tensor(np.stack([rows, cols]), dtype=torch.long).view(2, -1)
        return Data(x=torch.tensor(np.asarray(s_x), dtype=torch.float), edge_index=edge_index,
                    edge_attr=torch.tensor(w, dtype=torch.float))

    # ------------------------------------------------------------- objective ---
    def cost(self, pred, target, action):
        """policy.py:234-239: sum((Q*a - y)^2) / BATCH."""
        readout_action = torch.mul(pred.view(-1), action)
        return torch.pow(readout_action - target.view(-1), 2).sum() / self.BATCH

    def apply_pending(self):
        """The clamp + Adam update of a ``train(..., apply=False)`` step (the trainer issues it behind the acting forward of the tick)."""
        if self._pending_apply is not None:
            optimizer, clamp, gscale = self._pending_apply
            self._pending_apply = None
            optimizer.step(clamp=clamp, gscale=gscale)

    def train(self, data, action, y, device, model, optimizer, apply: bool = True):
        """policy.py:241-253 (model.train(), dropout p=0.5, clamp +-0.5, Adam step).  With ``dist.NativeAdam`` as the optimizer and
        the DQN Q-network itself (``Networks.GCN`` on CUDA) the whole step runs on the hand-written kernels: forward + cost +
        backward in one native call (``gnn.gcn_train_step``), the gradient all-reduce, then clamp + Adam in one kernel; the loss
        stays on the device (``temp_loss`` is a 0-d tensor then: ``float()`` it to read it)."""
        model.train()
        data = data.to(device)
        from .dist import NativeAdam
        if isinstance(optimizer, NativeAdam) and data.x.is_cuda and gnn.QForwardPlan.eligible(model):
            from . import Networks
            import torch.distributed as tdist
            gs = Networks._structure(data, data.x.size(0))
            y = torch.as_tensor(y, device=device, dtype=torch.float32)
            action = torch.as_tensor(action, device=device, dtype=torch.float32)
            seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self.native_steps) & 0xFFFFFFFFFFFFFFFF
            self.native_steps += 1
            self._bucket = optimizer.bucket
            loss, _ = gnn.gcn_train_step(model, data.x, gs, action, y, 1.0 / self.BATCH, 0.5, seed=seed)
            world = tdist.get_world_size() if tdist.is_available() and tdist.is_initialized() else 1
            if world > 1:
                tdist.all_reduce(optimizer.bucket.flat, op=tdist.ReduceOp.SUM)     # the one collective of the path
            if apply:
                optimizer.step(clamp=self.max_grad_norm, gscale=1.0 / world)      # mean over ranks, clamp, Adam: one kernel
            else:                                                                 # (native path only: the caller applies the update later)
                self._pending_apply = (optimizer, self.max_grad_norm, 1.0 / world)
            self.temp_loss = loss
            return loss
        if self._bucket is None or self._bucket.params[0] is not next(p for p in model.parameters() if p.requires_grad):
            self._bucket = FlatGradBucket(model.parameters())
        self._bucket.zero_()
        out = model(data, 0.5, batch=data.batch)
        y = torch.as_tensor(y, device=device, dtype=out.dtype)
        action = torch.as_tensor(action, device=device, dtype=out.dtype)
        loss = self.cost(out, y, action)
        self.temp_loss = loss.item()
        loss.backward()
        self._bucket.all_reduce_mean()            # the one collective of the path (no-op at world size 1)
        self._bucket.clamp_(self.max_grad_norm)   # policy.py:251-252, after averaging
        optimizer.step()
        return self.temp_loss

    @torch.no_grad()
    def test(self, data, prob, device, model):
        """policy.py:255-259"""
        model.eval()
        return model(data.to(device), prob)

    # --------------------------------------------------------------- targets ---
    @torch.no_grad()
    def build_targets(self, minibatch, device, target_net):
        """policy.py:141-181: y = r (+ gamma * max over the next state's last `fro_size1` nodes if not
        terminal) at the chosen action's node, 0 elsewhere.  minibatch items are
        (s_t, a_t, r_t, s_t1, terminal, fro_size1) like the reference's replay tuples."""
        s_j = Batch.from_data_list([d[0] for d in minibatch]).to(device)
        s_j1 = Batch.from_data_list([d[3] for d in minibatch]).to(device)
        q1 = self.test(s_j1, 0.0, device, target_net).view(-1)
        sizes1 = torch.tensor([d[3].x.size(0) for d in minibatch], device=device)
        fro1 = torch.tensor([d[5] for d in minibatch], device=device)
        ends = torch.cumsum(sizes1, 0)
        idx = torch.arange(q1.numel(), device=device)
        seg = s_j1.batch
        in_tail = idx >= (ends - fro1)[seg]
        neg = torch.full_like(q1, -float("inf"))
        max_q = torch.full((len(minibatch),), -float("inf"), device=device).scatter_reduce(0, seg, torch.where(in_tail, q1, neg), "amax")
        r = torch.tensor([d[2] for d in minibatch], device=device, dtype=q1.dtype)
        term = torch.tensor([bool(d[4]) for d in minibatch], device=device)
        yv = torch.where(term, r, r + self.GAMMA * max_q)
        a = torch.cat([torch.as_tensor(d[1], dtype=q1.dtype) for d in minibatch]).to(device)
        y = torch.zeros_like(a)
        starts = torch.cumsum(torch.tensor([0] + [len(d[1]) for d in minibatch[:-1]], device=device), 0)
        act_idx = torch.stack([torch.as_tensor(np.argmax(d[1])) for d in minibatch]).to(device) + starts
        y[act_idx] = yv
        return s_j, a, y

    def learn_step(self, device, policy_net, target_net, optimizer):
        """One DQN gradient step on a sampled minibatch (policy.py:136-182)."""
        if self.step_t % self.TARGET_UPDATE == 0:
            target_net.load_state_dict(policy_net.state_dict())
        minibatch = random.sample(self.buffer, self.BATCH)
        s_j, a, y = self.build_targets(minibatch, device, target_net)
        return self.train(s_j, a, y, device, policy_net, optimizer)

    def remember(self, transition):
        self.buffer.append(transition)
        if len(self.buffer) > self.REPLAY_MEMORY:
            self.buffer.popleft()

    def running(self, model, modelTarget, test=False, env=None, epochs=None, device=None, method="bayesian", log=None):
        """policy.py:60-209 on ``ExplorationEnv`` (the B = 1 view of the CUDA engine; ``trainer.VecDQNTrainer`` is the B-env form):
        per decision the exploration graph, the look-ahead reward of every frontier, a frontier chosen from the dropout-perturbed
        Q-values ("bayesian", dropout p = epsilon) or epsilon-greedily, the line plan executed step by step, the transition
        ``(s_t, a_t, r_t, s_t1, done or loop_clo, fro_size1)`` stored, and -- after OBSERVE decisions -- one gradient step on BATCH
        sampled transitions with the target network refreshed every TARGET_UPDATE decisions.  File outputs (reward CSVs,
        ``MyModel.pt``) are the caller's business here: ``log(step_t, state, epsilon, q_max, explored, reward, terminal)`` is called
        once per decision and the per-decision rewards are kept in ``self.total_reward``."""
        from .envs.exploration_env import ExplorationEnv
        env = env or ExplorationEnv(self.map_size, 0, test)
        device = device or env._vec.device
        policy_net, target_net = model, modelTarget
        target_net.eval()
        optimizer = torch.optim.Adam(policy_net.parameters(), lr=1e-5)
        losses = []
        for _ in range(int(self.epoch if epochs is None else epochs)):
            self.step_t += 1
            if self.epsilon > self.FINAL_EPSILON and self.step_t > self.OBSERVE:
                self.epsilon -= (self.INITIAL_EPSILON - self.FINAL_EPSILON) / self.EXPLORE
            adjacency, features, _, fro_size = env.graph_matrix()
            node_size = adjacency.shape[0]
            key_size = node_size - fro_size
            s_t = self.data_process([adjacency, features])
            all_actions = env.actions_all_goals()
            rewards = env.rewards_all_goals(all_actions)
            if method == "e-greedy":
                q = self.test(s_t, 0.0, device, policy_net).view(-1).cpu().numpy()
                explore = random.random() <= self.epsilon
                action_index = random.randrange(fro_size) if explore else int(np.argmax(q[node_size - fro_size:]))
                state = "explore" if explore else "exploit"
            else:
                q = self.test(s_t, self.epsilon, device, policy_net).view(-1).cpu().numpy()
                action_index = int(np.argmax(q[node_size - fro_size:]))
                state = "bayesian"
            a_t = np.zeros(node_size)
            a_t[key_size + action_index] = 1
            r_t = float(rewards[key_size + action_index])
            done = False
            for act in all_actions[key_size + action_index]:
                _, done, _ = env.step(act)
            current_done = done or env.loop_clo
            adjacency, features, _, fro_size1 = env.graph_matrix()
            s_t1 = self.data_process([adjacency, features])
            self.remember((s_t, a_t, r_t, s_t1, current_done, fro_size1))
            if self.step_t > self.OBSERVE and len(self.buffer) >= self.BATCH:
                if self.step_t % self.TARGET_UPDATE == 0:
                    target_net.load_state_dict(policy_net.state_dict())
                minibatch = random.sample(self.buffer, self.BATCH)
                s_j, a, y = self.build_targets(minibatch, device, target_net)
                losses.append([self.step_t, self.train(s_j, a, y, device, policy_net, optimizer)])
            if log is not None:
                log(self.step_t, state, self.epsilon, float(np.max(q)), env.status(), r_t, current_done)
            if done:
                env.reset()
            self.total_reward = np.append(self.total_reward, r_t)
        return env, losses


class A2C(object):
    """policy.py:262-497: n-step advantage actor-critic.  Same hyper-parameters, ``data_process / policy_cost / value_cost /
    entropy_loss / train / test`` and the same n-step batch construction (policy.py:361-393, ``nstep_batch``)."""

    def __init__(self, case_path="A2C_GCN/"):
        self.case_path = case_path
        self.GAMMA = 0.99          # policy.py:279-285
        self.EXPLORE = 1e6
        self.epoch = 1e4
        self.nstep = 40
        self.ent_coef = 0.01
        self.vf_coef = 0.25
        self.max_grad_norm = 0.5
        self.buffer = deque()
        self.map_size = 40
        self.step_t = 0
        self.temp_loss = 0
        self.entro = 0
        self.total_reward = np.empty([0, 0])
        self._bucket = None

    # ------------------------------------------------------------------ data ---
    def data_process(self, data, device):
        """policy.py:428-450: the DQN conversion plus an all-zero ``batch`` vector on the device."""
        state = DeepQ.data_process(self, data)
        return state, torch.zeros(np.shape(data[0])[0], dtype=torch.long, device=device)

    # ------------------------------------------------------------ objectives ---
    def policy_cost(self, prob, advantages, action, mask):
        """policy.py:452-460: -sum(log pi(a) * A) / nstep; ``prob`` holds the frontier nodes only (masked softmax)."""
        adv = torch.masked_select(advantages.view(-1), mask)
        act = torch.masked_select(action, mask)
        return (-(prob.view(-1).log() * adv) * act).sum() / self.nstep

    def value_cost(self, pred, target):
        """policy.py:462-466"""
        return F.mse_loss(pred.view(-1), target.view(-1))

    def entropy_loss(self, prob):
        """policy.py:468-472 (detached: it is reported and scaled into the loss value, it carries no gradient)."""
        p = prob.view(-1).detach()
        entro = -(p.log() * p).sum() / self.nstep
        self.entro = entro.item()
        return entro

    def nstep_batch(self, last_value):
        """policy.py:361-393 on ``self.buffer`` (items ``(s_t, a_t, r_t, s_t1, terminal, fro_size, value)``):
        discounted returns bootstrapped from ``last_value``, one-hot actions, frontier masks and the advantage vector
        (return - value at the chosen node, 0 elsewhere).  Returns (Batch, a, mask, returns, y_adv)."""
        buf = list(self.buffer)
        ret, returns = float(last_value), []
        for d in reversed(buf):                                     # ret = r + gamma * ret * (1 - terminal)
            ret = d[2] + self.GAMMA * ret * (1.0 - float(bool(d[4])))
            returns.append(ret)
        returns = returns[::-1]
        a, mask, adv = [], [], []
        for d, g in zip(buf, returns):
            act = np.asarray(d[1], dtype=np.float64)
            m = np.zeros(act.size); m[act.size - int(d[5]):] = 1
            y = np.zeros(act.size); y[int(np.argmax(act))] = g - d[6]
            a.append(act); mask.append(m); adv.append(y)
        return Batch.from_data_list([d[0] for d in buf]), np.concatenate(a), np.concatenate(mask), np.asarray(returns), np.concatenate(adv)

    def train(self, data, action, mask, dis_reward, y_adv, device, modelA, modelC, optimizer):
        """policy.py:474-497; multi-GPU: one all-reduce of the joint actor+critic gradient bucket before the clamp."""
        modelA.train(); modelC.train()
        data = data.to(device)
        mask = torch.as_tensor(np.asarray(mask), dtype=torch.bool, device=device) if not torch.is_tensor(mask) else mask.to(device).bool()
        params = [p for p in list(modelA.parameters()) + list(modelC.parameters()) if p.requires_grad]
        if self._bucket is None or self._bucket.params[0] is not params[0]:
            self._bucket = FlatGradBucket(params)
        self._bucket.zero_()
        actor_out = modelA(data, mask, batch=data.batch) + 1e-35
        critic_out = modelC(data, mask, batch=data.batch)
        dt = actor_out.dtype
        y_adv = torch.as_tensor(y_adv, device=device).to(dt)
        dis_reward = torch.as_tensor(dis_reward, device=device).to(critic_out.dtype)
        action = torch.as_tensor(action, device=device).to(dt)
        loss = (self.policy_cost(actor_out, y_adv, action, mask) - self.entropy_loss(actor_out) * self.ent_coef
                + self.value_cost(critic_out, dis_reward) * self.vf_coef)
        self.temp_loss = loss.item()
        loss.backward()
        self._bucket.all_reduce_mean()
        self._bucket.clamp_(self.max_grad_norm)
        optimizer.step()
        return self.temp_loss

    @torch.no_grad()
    def test(self, data, batch, mask, device, model):
        """policy.py:499-504"""
        model.eval()
        mask = torch.as_tensor(np.asarray(mask), dtype=torch.bool, device=device) if not torch.is_tensor(mask) else mask.to(device).bool()
        return model(data.to(device), mask, batch)

    def running(self, actor, critic, env=None, test=False, epochs=None, device=None, log=None):
        """policy.py:297-426 on ``ExplorationEnv`` (the B = 1 view of the CUDA engine): sample a frontier from the actor's
        masked softmax, follow its line plan, and every ``nstep`` decisions take one actor-critic gradient step."""
        from .envs.exploration_env import ExplorationEnv
        env = env or ExplorationEnv(self.map_size, 0, test)
        device = device or env._vec.device
        optimizer = torch.optim.Adam(list(actor.parameters()) + list(critic.parameters()), lr=1e-5)
        rng = np.random.default_rng()
        for _ in range(int(self.epoch if epochs is None else epochs)):
            self.step_t += 1
            adjacency, features, _, fro_size = env.graph_matrix()
            node_size = adjacency.shape[0]
            key_size = node_size - fro_size
            s_t, b_t = self.data_process([adjacency, features], device)
            mask = np.zeros(node_size); mask[node_size - fro_size:] = 1
            all_actions = env.actions_all_goals()
            rewards = env.rewards_all_goals(all_actions)
            pi = self.test(s_t, b_t, mask, device, actor).view(-1).double().cpu().numpy()
            val = float(self.test(s_t, b_t, mask, device, critic).item())
            action_index = key_size + int(rng.choice(fro_size, p=pi / pi.sum()))
            a_t = np.zeros(node_size); a_t[action_index] = 1
            r_t = float(rewards[action_index])
            done = False
            for act in all_actions[action_index]:
                _, done, _ = env.step(act)
            current_done = done or env.loop_clo
            adjacency, features, _, fro_size1 = env.graph_matrix()
            s_t1, b_t1 = self.data_process([adjacency, features], device)
            m1 = np.zeros(adjacency.shape[0]); m1[adjacency.shape[0] - fro_size1:] = 1
            last_value = float(self.test(s_t1, b_t1, m1, device, critic).item())
            self.buffer.append((s_t, a_t, r_t, s_t1, current_done, fro_size, val))
            if len(self.buffer) == self.nstep:
                batch, a, mk, returns, adv = self.nstep_batch(last_value)
                self.train(batch, a, mk, returns, adv, device, actor, critic, optimizer)
                self.buffer.clear()
            if log is not None:
                log(self.step_t, self.temp_loss, self.entro, env.status(), r_t, current_done)
            if done:
                env.reset()
            self.total_reward = np.append(self.total_reward, r_t)
        return env
