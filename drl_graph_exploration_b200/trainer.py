"""Vectorised DQN training loop: ``DeepQ.running`` (policy.py:60-209) for B environments per GPU.

The reference trains on ONE env: per decision it builds the exploration graph, evaluates the look-ahead reward of
every frontier on cloned simulators (``rewards_all_goals``), picks the arg-max of the dropout-perturbed Q-values
("bayesian" exploration, policy.py:104-110), drives the env along the line plan, stores
``(s_t, a_t, r_t, s_t1, done or loop_clo, fro_size1)`` and does one gradient step on 64 sampled transitions.

Here the same per-env sequence runs for B envs on the batched engine (BASELINE config C3):

* a **tick** = every env with a queued action executes one simulator step; every env whose queue ran empty gets a
  decision: graph batch (``dge_graph``) -> roll-out rewards of all its frontiers on the clone engine
  (``dge_rollout_*``) -> Q with functional dropout p = epsilon -> arg-max frontier + line plan on the device
  (``dge_select_and_plan``);
* the graph batch of a decision round is stored once in the device replay (``replay.GraphReplay``): it is ``s_t`` of
  the transitions that start now and ``s_t1`` of the ones that end now;
* ``train_steps_per_tick`` gradient steps per tick (``DeepQ.train``: loss sum((Q a - y)^2)/64, clamp +-0.5, Adam 1e-5)
  with ONE all-reduce of the flat gradient bucket per step when ``torch.distributed`` is initialised -- every rank
  takes the same number of steps per tick, so the collective never waits on data-dependent control flow.

Deliberate differences, none of which change a per-env result: an episode that reaches ``done`` in the middle of a
line plan restarts at once (the reference finishes the plan on the finished env, policy.py:117-118); an env left
without frontier ends its episode (the reference would raise, quirk q15); epsilon decays per decision like the
reference, counted over all envs.
"""
from __future__ import annotations

import ctypes

import torch

from .engine import _check, _stream_ptr
from .envs.exploration_env import RESET_ODOM, VecExplorationEnv
from .policy import DeepQ
from .replay import GraphReplay


class VecDQNTrainer:
    def __init__(self, env: VecExplorationEnv, policy_net: torch.nn.Module, target_net: torch.nn.Module, dqn: DeepQ | None = None,
                 replay_capacity: int | None = None, train_steps_per_tick: int = 1, observe: int | None = None, lr: float = 1e-5,
                 clone_slots: int | None = None, seed: int = 0, overlap: bool | str = False):
        self.env, self.policy_net, self.target_net = env, policy_net, target_net
        self.dqn = dqn or DeepQ()
        self.dev = env.device
        B, eng = env.B, env.eng
        cap = int(replay_capacity if replay_capacity is not None else self.dqn.REPLAY_MEMORY)
        # graphs allocated between a transition's s_t and its completion: <= B per tick x (longest line plan + reset phase) ticks
        # (decisions without a frontier and transitions dropped by restore() allocate graphs too -- hence per tick, not per transition)
        self.replay = GraphReplay(cap, eng.node_cap_env, eng.edge_cap_env, self.dev, slack=B * (env.cfg.max_plan_actions + 8) + 64)
        # the DQN Q-network on CUDA trains on the native step (gnn.gcn_train_step + dist.NativeAdam); any other net on autograd + torch Adam
        from . import gnn
        from .dist import NativeAdam
        native = gnn.QForwardPlan.eligible(policy_net)
        self.optimizer = NativeAdam(policy_net.parameters(), lr=lr) if native else torch.optim.Adam(policy_net.parameters(), lr=lr)
        self._sync_replicas()                # rank 0's weights everywhere: the replicas must not depend on every caller seeding alike
        self.target_net.load_state_dict(policy_net.state_dict())
        self.target_net.eval()
        self.train_steps_per_tick = int(train_steps_per_tick)
        self.observe = int(self.dqn.OBSERVE if observe is None else observe)   # decisions before learning starts
        self.clone_slots = clone_slots
        # overlap: the gradient step of a tick runs on a second stream beside the roll-out kernels of the same tick (it trains
        # on the replay as of the previous tick and finishes before this tick's Q forward reads the weights)
        # overlap="tail": the gradient step is issued at the END of its tick (exactly the sequential order of operations: it trains on
        # the replay including this tick's transitions) but on the second stream -- it runs beside the NEXT tick's step pipeline and
        # roll-out kernels, and the next tick's Q forward waits for it.  The host issues the step -- and, with several ranks, its
        # gradient all-reduce -- half a tick earlier than in the "beside" schedule, which is slack for ranks whose ticks differ in length.
        # Measured (bench.py --workload train --train-schedule tail, N = 1): 8.8-10.3 ms per tick against 7.4 ms for "beside" -- the minibatch
        # gather still reads two totals back (replay.gather), and at the end of a tick that read waits for the tick's whole backlog on the
        # device.  Bit-identical to the sequential schedule (tests/test_trainer_gpu.py); not the default until the gather is sync-free.
        # overlap="lag": like True, but the acting forward of a tick does not wait for the tick's gradient step: it reads the weights as
        # of the PREVIOUS update, and the update of this tick (clamp + Adam) is issued behind it.  The gradient steps themselves are the
        # same sequence (each is taken at the weights the previous one left); the acting policy is one update staler, which gives the
        # gradient all-reduce a whole tick before anything waits for it -- ranks whose ticks differ in length stop waiting for each other.
        self.lag = overlap == "lag"
        self.tail = overlap == "tail"
        self.overlap = bool(overlap)
        self.s_learn = torch.cuda.Stream(self.dev) if self.overlap else None
        self.ev_tick, self.ev_learn = (torch.cuda.Event(), torch.cuda.Event()) if self.overlap else (None, None)
        self.ev_gathered, self.ev_q = (torch.cuda.Event(), torch.cuda.Event()) if self.overlap else (None, None)
        if self.lag and (int(train_steps_per_tick) != 1 or not native):
            raise ValueError("VecDQNTrainer(overlap='lag') takes one gradient step per tick on the native training path (Networks.GCN + dist.NativeAdam)")
        i64 = lambda v: torch.full((B,), v, dtype=torch.int64, device=self.dev)
        self.pend_slot, self.pend_a = i64(-1), i64(0)                  # in-flight transition of every env
        self.pend_serial = i64(-1)                                     # allocation serial of its s_t, read when the graph was stored
        self.ecap_episodes = torch.zeros((), dtype=torch.int64, device=self.dev)   # episodes cut short by the pose capacity (excluded from the replay)
        self.pend_r = torch.zeros(B, dtype=torch.float32, device=self.dev)
        self.pend_clo = torch.zeros(B, dtype=torch.bool, device=self.dev)
        self._fo = (ctypes.c_double * 3)(*RESET_ODOM)
        self.gen = torch.Generator(device=self.dev); self.gen.manual_seed(seed)
        self.decisions = self.train_steps = self.ticks = 0
        self.rollout_steps = self.rollout_clones = 0    # clone-engine ticks launched / clones evaluated
        self._loss = float("nan")
        # transitions closed / their reward sum: accumulated on the device (a tick closes transitions without the host counting them)
        self._transitions = torch.zeros((), dtype=torch.int64, device=self.dev)
        self._reward_sum = torch.zeros((), dtype=torch.float64, device=self.dev)
        self._learning = False                          # see _learning_started
        self._learn_pending = False                     # overlap="tail": a gradient step is in flight on the second stream

    @property
    def transitions(self) -> int:
        """transitions stored so far (reads the device counter: a host sync -- logging, tests, checkpoints)"""
        return int(self._transitions)

    @transitions.setter
    def transitions(self, v):
        self._transitions.fill_(int(v))

    @property
    def reward_sum(self) -> float:
        return float(self._reward_sum)

    @reward_sum.setter
    def reward_sum(self, v):
        self._reward_sum.fill_(float(v))

    # ---------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _act(self, side_work=None):
        """Step pipeline + decision round of one tick.  Returns the number of decisions taken.  ``side_work``: callable run
        (on its own stream) once the step and roll-out kernels of this tick are queued, before the Q forward."""
        env, eng, rp, dqn = self.env, self.env.eng, self.replay, self.dqn
        st, dev = eng.state, self.dev
        main = torch.cuda.current_stream(dev) if side_work is not None else None
        if side_work is not None:
            self.ev_tick.record(main)                  # everything of the previous tick (replay writes, weight reads) is before this
        done_prev = st["done"].bool().clone()          # episodes that ended in the previous tick's step
        ecap_prev = done_prev & (st["status"] == -4)   # ... because the trajectory buffer was full (DGE_ECAP): not a terminal state of the MDP
        self.ecap_episodes = self.ecap_episodes + ecap_prev.sum()
        need = env.mark_pending().bool().clone()       # empty queue, not done, not in the reset phase
        # ---- step pipeline: restart finished episodes, one simulator step for every env with a queued action ----
        _check(eng._L.dge_reset_done_queued(eng._h, env.seed_stride, self._fo, 4, _stream_ptr(dev)), "dge_reset_done_queued")
        eng.step_queued()
        # ---- transitions that ended with the episode: terminal, s_t1 is not used by the target (policy.py:166-167) ----
        ended = done_prev & (self.pend_slot >= 0) & ~ecap_prev      # a capacity stop drops its in-flight transition instead of storing y = r
        self.pend_slot = torch.where(ecap_prev, torch.full_like(self.pend_slot, -1), self.pend_slot)
        # ---- decision round ----
        g = env.build_graph(need.to(torch.uint8))
        ng, n, e = g.sync_sizes()                      # host sync (sizes the GNN's GEMMs)
        slot_new = torch.full_like(self.pend_slot, -1)
        if ng > 0:
            ordinal = torch.cumsum(need.long(), 0) - 1                      # graph ordinal of env b (graphs are in env order)
            # roll-outs: one clone per (env, frontier); the decision round is cut into chunks that fit the clone engine
            slots_c = self.clone_slots or min(env.B * (eng.Lt + 1), max(4 * env.B, 512))
            fro_host = g.fro_size[:ng].tolist()                             # host sync (tiny)
            lo = acc = 0
            for i, f in enumerate(fro_host + [slots_c + 1]):
                if acc + f > slots_c:
                    m = (need & (ordinal >= lo) & (ordinal < i)).to(torch.uint8)
                    _, norm, clo = env.rollout_rewards(m, clone_slots=slots_c, auto_steps=True)
                    self.rollout_steps += env.rollout_steps; self.rollout_clones += acc
                    lo, acc = i, 0
                acc += f
        if self.lag:
            cur = torch.cuda.current_stream(dev)
            if side_work is not None:
                side_work()                            # forward / backward / all-reduce of this tick's gradient step; its update waits (tick())
                cur.wait_event(self.ev_gathered)       # the replay is the step's until its minibatch is gathered
            if self._learn_pending:
                cur.wait_event(self.ev_learn)          # the weights are the PREVIOUS update's until it is done
        elif side_work is not None:
            side_work()
            main.wait_event(self.ev_learn)             # the weights and the replay are the gradient step's until it is done
        elif self.tail and self._learn_pending:
            torch.cuda.current_stream(dev).wait_event(self.ev_learn)   # the previous tick's gradient step (issued at its end) owns weights and replay until done
            self._learn_pending = False
        if ng > 0:
            d = g.data()
            q = self.policy_net(d, float(self.epsilon))                     # functional dropout: "bayesian" exploration
            choice = env.select_and_plan(q, need.to(torch.uint8)).long()    # [B], valid where need
            slots = rp.store_graphs(d.x, d.edge_index, d.edge_attr, d.batch, g.node_ptr, g.edge_ptr, g.key_size, g.fro_size, ng)
            o = ordinal.clamp(0, ng - 1)
            slot_new = torch.where(need, slots[o], slot_new)
            fro = torch.where(need, g.fro_size[:ng].long()[o], torch.zeros_like(o))
            key = g.key_size[:ng].long()[o]
        else:
            fro = torch.zeros_like(self.pend_slot); key = fro; choice = fro
            norm = clo = None
        # ---- close the transitions of the envs that decided now (s_t1 = the graph just stored) and of the ended episodes ----
        closing = need & (self.pend_slot >= 0)
        fin = closing | ended
        # one row per env + the mask: the ring places the closing rows behind its head on the device (no host count, no sync)
        s1 = torch.where(closing, slot_new, self.pend_slot)
        term = ended | self.pend_clo | (closing & (fro <= 0))
        ser1 = torch.where(closing, rp.gserial[slot_new.clamp(min=0)], self.pend_serial)     # s_t1 was stored in this tick (or is s_t itself)
        cnt = rp.append_masked(fin, self.pend_slot, self.pend_a, self.pend_r, s1, term, serial_s=self.pend_serial, serial_s1=ser1)
        self._transitions += cnt
        self._reward_sum += torch.where(fin, self.pend_r, torch.zeros_like(self.pend_r)).double().sum()
        self.pend_slot = torch.where(ended, torch.full_like(self.pend_slot, -1), self.pend_slot)
        # ---- open the transitions of the envs that decided now ----
        if ng > 0:
            start = need & (fro > 0)
            r = norm.gather(1, choice.clamp(0, norm.size(1) - 1).view(-1, 1)).view(-1).float()
            self.pend_slot = torch.where(start, slot_new, torch.where(need, torch.full_like(slot_new, -1), self.pend_slot))
            self.pend_serial = torch.where(start, rp.gserial[slot_new.clamp(min=0)], self.pend_serial)
            self.pend_a = torch.where(start, key + choice, self.pend_a)
            self.pend_r = torch.where(start, r, self.pend_r)
            self.pend_clo = torch.where(start, clo.bool(), self.pend_clo)
        self.decisions += ng
        dqn.step_t += ng
        if dqn.epsilon > dqn.FINAL_EPSILON and dqn.step_t > self.observe:   # policy.py:78-79, per decision
            dqn.epsilon = max(dqn.FINAL_EPSILON, dqn.epsilon - ng * (dqn.INITIAL_EPSILON - dqn.FINAL_EPSILON) / dqn.EXPLORE)
        return ng

    @property
    def epsilon(self):
        return self.dqn.epsilon

    # ---------------------------------------------------------------------------------------------------------------
    def minibatch(self, k: int | None = None, check: bool = False):
        """(s_j batch, action one-hot [N], y [N]) of policy.py:141-178, built on the device."""
        dqn, rp = self.dqn, self.replay
        k = k or dqn.BATCH
        s, a, r, s1, term = rp.sample(k, generator=self.gen, check=check)
        b_s, n_s, off_s = rp.gather(s)
        b_s1, n_s1, off_s1 = rp.gather(s1)
        return b_s, b_s1, (a, r, term, off_s, n_s1, off_s1, rp.gf[s1])

    def learn(self, check: bool = False, apply: bool = True):
        """One gradient step (policy.py:136-182); collective inside ``DeepQ.train``.  ``apply=False`` (the "lag" schedule): forward,
        backward and all-reduce only -- ``DeepQ.apply_pending`` issues the update."""
        dqn = self.dqn
        if dqn.TARGET_UPDATE and self.train_steps % int(dqn.TARGET_UPDATE) == 0:
            self.target_net.load_state_dict(self.policy_net.state_dict())
        b_s, b_s1, (a, r, term, off_s, n_s1, off_s1, fro1) = self.minibatch(check=check)
        if not apply:
            self.ev_gathered.record(torch.cuda.current_stream(self.dev))
        with torch.no_grad():
            q1 = dqn.test(b_s1, 0.0, self.dev, self.target_net).view(-1)
        act, y = dqn_targets(q1, b_s1.batch, a, r, term, off_s, n_s1, off_s1, fro1, b_s.x.size(0), dqn.GAMMA)
        self._loss = dqn.train(b_s, act, y, self.dev, self.policy_net, self.optimizer, apply=apply)    # (a device scalar on the native path: no sync here)
        self.train_steps += 1
        return self.last_loss if check else self._loss

    def _join_learn(self):
        """overlap="tail": orders the current stream behind the gradient step in flight (weights, loss, optimizer state)."""
        if self._learn_pending:
            torch.cuda.current_stream(self.dev).wait_event(self.ev_learn)

    @property
    def last_loss(self) -> float:
        """Loss of the last gradient step (reading it synchronises with the step's stream on the native path)."""
        self._join_learn()
        return float(self._loss)

    def _learn_beside(self):
        with torch.enable_grad(), torch.cuda.stream(self.s_learn):
            self.s_learn.wait_event(self.ev_tick)
            for _ in range(self.train_steps_per_tick):
                self.learn(apply=not self.lag)
            if not self.lag:
                self.ev_learn.record(self.s_learn)

    def tick(self, learn: bool | None = None):
        """One tick.  Sequential mode: act, then ``train_steps_per_tick`` gradient steps.  Overlap mode: the gradient steps run
        beside this tick's roll-outs on the replay as of the previous tick -- the same sequence of operations as the
        sequential mode shifted by one tick (a sequential run of k learning ticks == one acting tick + k overlapped ticks)."""
        if learn is None:
            learn = self._learning_started()
        if self.tail:
            ng = self._act()
            self.ticks += 1
            if learn:
                self.ev_tick.record(torch.cuda.current_stream(self.dev))     # this tick's replay writes and weight reads are before this
                self._learn_beside()
                self._learn_pending = True
            return ng
        if self.lag:
            ng = self._act(self._learn_beside if learn else None)
            self.ticks += 1
            if learn:     # the update of this tick's step, behind the acting forward that read the old weights
                self.ev_q.record(torch.cuda.current_stream(self.dev))
                with torch.cuda.stream(self.s_learn):
                    self.s_learn.wait_event(self.ev_q)
                    self.dqn.apply_pending()
                    self.ev_learn.record(self.s_learn)
                self._learn_pending = True
            return ng
        if self.overlap:
            ng = self._act(self._learn_beside if learn else None)
            self.ticks += 1
            return ng
        ng = self._act()
        self.ticks += 1
        if learn:
            for _ in range(self.train_steps_per_tick):
                self.learn()
        return ng

    def _sync_replicas(self):
        """Broadcast the policy net's parameters and buffers from rank 0 (no-op on one rank).  Gradients are all-reduced every step, so
        replicas that start equal stay equal; this makes them start equal whatever the ranks' seeds or checkpoint files are."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            with torch.no_grad():
                for t in list(self.policy_net.parameters()) + list(self.policy_net.buffers()):
                    dist.broadcast(t.data, src=0)

    def _learning_started(self) -> bool:
        """policy.py:136: learning starts once OBSERVE decisions have passed (and, here, one minibatch of transitions exists).
        Both counts are rank-local, and a rank that started alone would sit in the gradient all-reduce of ``DeepQ.train``
        without a partner: with several ranks the start is agreed on (one scalar MIN all-reduce per tick until every rank is
        ready -- both conditions are monotone, so nothing is exchanged afterwards)."""
        if self._learning:
            return True
        import torch.distributed as dist
        ready = self.dqn.step_t > self.observe and self.replay.size >= self.dqn.BATCH
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            t = torch.tensor([1.0 if ready else 0.0], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ready = bool(t.item() >= 1.0)
        self._learning = ready
        return ready

    def save(self, path: str):
        """``torch.save(policy_net.state_dict(), .../MyModel.pt)`` like policy.py:192 -- loadable by the reference."""
        self._join_learn()
        torch.save({k: v.detach().cpu() for k, v in self.policy_net.state_dict().items()}, path)

    def run(self, n_ticks: int, out_dir: str | None = None, log_every: int = 100, save_every: int = 50000):
        """The outer loop of ``DeepQ.running`` (policy.py:72-209) in ticks: trains for ``n_ticks`` and, if ``out_dir`` is given,
        leaves the reference's artefacts there -- ``temp_reward.csv`` (decision count, mean reward of the closed transitions
        since the last row), ``temp_loss.csv`` (decision count, loss), ``reward_data.csv`` (Step, Reward), ``MyModel.pt`` every
        ``save_every`` decisions, ``Model_Policy.pt`` / ``Model_Target.pt`` at the end (policy.py:192-209).  With several ranks the
        replicas are identical: pass ``out_dir`` on rank 0 only (the reward rows are those of the rank's own envs)."""
        import os
        if out_dir is not None:
            os.makedirs(out_dir, exist_ok=True)     # before the loop: MyModel.pt is written from inside it
        rewards, losses, rows = [], [], []
        r0, n0, next_log, next_save = self.reward_sum, self.transitions, log_every, save_every
        for _ in range(int(n_ticks)):
            steps_before = self.train_steps
            self.tick()
            if self.train_steps > steps_before:
                losses.append([self.dqn.step_t, self.last_loss])
            if self.dqn.step_t >= next_log:
                self.replay.assert_intact()         # the per-minibatch serial test, read back once per log row
                m = self.transitions - n0
                if m > 0:
                    rewards.append([self.dqn.step_t, (self.reward_sum - r0) / m])
                    rows.append([self.dqn.step_t, (self.reward_sum - r0) / m])
                r0, n0, next_log = self.reward_sum, self.transitions, next_log + log_every
            if out_dir is not None and self.dqn.step_t >= next_save:
                self.save(os.path.join(out_dir, "MyModel.pt"))
                next_save += save_every
        if out_dir is not None:
            import numpy as np
            np.savetxt(os.path.join(out_dir, "temp_reward.csv"), np.asarray(rewards).reshape(-1, 2), delimiter=",")
            np.savetxt(os.path.join(out_dir, "temp_loss.csv"), np.asarray(losses).reshape(-1, 2), delimiter=",")
            with open(os.path.join(out_dir, "reward_data.csv"), "w") as f:
                f.write("Step,Reward\n")
                for step, rew in rows:
                    f.write(f"{step},{rew}\n")
            self.save(os.path.join(out_dir, "Model_Policy.pt"))
            torch.save({k: v.detach().cpu() for k, v in self.target_net.state_dict().items()}, os.path.join(out_dir, "Model_Target.pt"))
        return rewards, losses

    def checkpoint(self, path: str, include_replay: bool = True):
        """The reference pickles the whole ``DeepQ`` object between its 10 000-step subprocess chunks (hyper-parameters, step_t,
        epsilon, the replay deque: train.py:33-35, run_training.py:13-16,63-64) next to the two state dicts.  Same content
        here, as one torch file: nets, optimizer, counters, epsilon, sampling generator and (optionally) the device replay.
        Like in the reference the environments are NOT part of it: a resumed run starts fresh episodes."""
        self._join_learn()
        ck = {"policy": self.policy_net.state_dict(), "target": self.target_net.state_dict(), "optimizer": self.optimizer.state_dict(),
              "dqn": {"step_t": self.dqn.step_t, "epsilon": self.dqn.epsilon, "native_steps": self.dqn.native_steps},
              "counters": {k: getattr(self, k) for k in ("decisions", "train_steps", "ticks", "transitions", "rollout_steps", "rollout_clones", "reward_sum")},
              "gen": self.gen.get_state(), "replay": self.replay.state_dict() if include_replay else None}
        torch.save(ck, path)

    def restore(self, path: str):
        ck = torch.load(path, map_location=self.dev, weights_only=False)
        self.policy_net.load_state_dict(ck["policy"]); self.target_net.load_state_dict(ck["target"])
        self.optimizer.load_state_dict(ck["optimizer"])
        self.dqn.step_t, self.dqn.epsilon = ck["dqn"]["step_t"], ck["dqn"]["epsilon"]
        self.dqn.native_steps = ck["dqn"].get("native_steps", 0)
        for k, v in ck["counters"].items():
            setattr(self, k, v)
        self.gen.set_state(ck["gen"].cpu())
        if ck["replay"] is not None:
            self.replay.load_state_dict(ck["replay"])
        self.pend_slot.fill_(-1)              # in-flight transitions belonged to the episodes of the old process
        self._learning = False                # re-agreed on by the ranks of the new process group
        self._sync_replicas()

    def load(self, path: str):
        """Resume from a ``MyModel.pt`` / ``Model_Policy.pt`` state dict (the reference's or ours)."""
        sd = torch.load(path, map_location=self.dev)
        self.policy_net.load_state_dict(sd)
        self._sync_replicas()
        self.target_net.load_state_dict(self.policy_net.state_dict())


def dqn_targets(q1, batch1, a, r, term, off_s, n_s1, off_s1, fro1, n_nodes_s: int, gamma: float):
    """policy.py:153-178 without the Python loop: y = r (+ gamma * max of the next state's last ``fro1`` Q-values unless
    terminal) at the chosen node, 0 elsewhere; the action vector is one-hot at the same node."""
    k = a.numel()
    node = torch.arange(q1.numel(), device=q1.device)
    in_tail = node >= (off_s1 + n_s1 - fro1)[batch1]
    neg = torch.full_like(q1, -float("inf"))
    max_q = torch.full((k,), -float("inf"), device=q1.device, dtype=q1.dtype).scatter_reduce(0, batch1, torch.where(in_tail, q1, neg), "amax")
    yv = torch.where(term, r.to(q1.dtype), r.to(q1.dtype) + gamma * max_q)
    at = off_s + a
    act = torch.zeros(n_nodes_s, dtype=q1.dtype, device=q1.device); act[at] = 1.0
    y = torch.zeros_like(act); y[at] = yv
    return act, y


class VecA2CTrainer:
    """``A2C.running`` (policy.py:297-426) for B environments per GPU.

    Per env the sequence is the reference's: at a decision the actor's masked softmax over the frontier nodes is sampled
    (``np.random.choice(fro_size, p=readout_t)``, policy.py:326 -- here a Gumbel-max draw fed to the device-side read-out), the
    critic's value of the state is recorded, the look-ahead reward of the chosen frontier comes from the batched roll-outs, the
    env follows the line plan, and ``(s_t, a_t, r_t, done or loop_clo, fro_size, V(s_t))`` joins the env's n-step segment.  When an
    env has ``nstep`` transitions its segment is closed with ``last_value = V(s_t+1)`` (the value at the env's next decision, 0
    after a terminal), the discounted returns and advantages are formed like policy.py:361-393 and the segment is trained on.
    All segments that close in the same tick form ONE batch (losses averaged over segments), so a tick takes at most one
    gradient step; with ``torch.distributed`` every rank joins the all-reduce of that tick (zero gradient if it has no segment),
    weighted by the number of segments, so ranks never wait on each other's episodes.
    """

    def __init__(self, env: VecExplorationEnv, actor: torch.nn.Module, critic: torch.nn.Module, a2c=None, lr: float = 1e-5,
                 clone_slots: int | None = None, seed: int = 0):
        from .dist import FlatGradBucket
        from .policy import A2C
        self.env, self.actor, self.critic = env, actor, critic
        self.a2c = a2c or A2C()
        self.dev = env.device
        B, eng, n = env.B, env.eng, int(self.a2c.nstep)
        # graph store only.  Slots are handed out round-robin over ALL envs, and a graph must survive until its env's segment
        # trains (<= n more decisions of that env): 3 n slots per env leave room for an env that decides at half the average
        # rate; the allocation serials kept with every transition turn an overwritten graph into an error (see _train)
        self.store = GraphReplay(B * 3 * n, eng.node_cap_env, eng.edge_cap_env, self.dev, slack=B * 8)
        self.params = [p for p in list(actor.parameters()) + list(critic.parameters()) if p.requires_grad]
        self.optimizer = torch.optim.Adam(self.params, lr=lr)
        self.bucket = FlatGradBucket(self.params)
        self.clone_slots = clone_slots
        z = lambda dt, *shape: torch.zeros(*shape, dtype=dt, device=self.dev)
        self.seg_slot, self.seg_a = z(torch.int64, B, n), z(torch.int64, B, n)
        self.seg_r, self.seg_val, self.seg_term = z(torch.float32, B, n), z(torch.float32, B, n), z(torch.bool, B, n)
        self.seg_len = z(torch.int64, B)
        self.seg_serial, self.pend_serial = z(torch.int64, B, n), z(torch.int64, B)
        self.pend_slot, self.pend_a = torch.full((B,), -1, dtype=torch.int64, device=self.dev), z(torch.int64, B)
        self.pend_r, self.pend_val, self.pend_clo = z(torch.float32, B), z(torch.float32, B), z(torch.bool, B)
        self._ar = torch.arange(B, device=self.dev)
        self._fo = (ctypes.c_double * 3)(*RESET_ODOM)
        self.gen = torch.Generator(device=self.dev); self.gen.manual_seed(seed)
        self.decisions = self.train_steps = self.segments = self.ticks = 0
        self.last_loss, self.last_entropy, self.reward_sum = float("nan"), float("nan"), 0.0

    @torch.no_grad()
    def tick(self):
        env, eng, a2c, st, dev = self.env, self.env.eng, self.a2c, self.env.eng.state, self.dev
        B, n = env.B, int(a2c.nstep)
        done_prev = st["done"].bool().clone()
        need = env.mark_pending().bool().clone()
        need_u8 = need.to(torch.uint8)
        _check(eng._L.dge_reset_done_queued(eng._h, env.seed_stride, self._fo, 4, _stream_ptr(dev)), "dge_reset_done_queued")
        eng.step_queued()
        ended = done_prev & (self.pend_slot >= 0)
        g = env.build_graph(need_u8)
        ng, nn, ne = g.sync_sizes()
        v_now = torch.zeros(B, dtype=torch.float32, device=dev)
        if ng > 0:
            ordinal = (torch.cumsum(need.long(), 0) - 1).clamp(0, ng - 1)
            slots_c = self.clone_slots or min(B * (eng.Lt + 1), max(4 * B, 512))
            fro_host = g.fro_size[:ng].tolist()
            lo = acc = 0
            for i, f in enumerate(fro_host + [slots_c + 1]):      # roll-outs in chunks that fit the clone engine
                if acc + f > slots_c:
                    m = (need & (ordinal >= lo) & (ordinal < i)).to(torch.uint8)
                    _, norm, clo = env.rollout_rewards(m, clone_slots=slots_c, auto_steps=True)
                    lo, acc = i, 0
                acc += f
            d = g.data()
            key_g, fro_g, nptr = g.key_size[:ng].long(), g.fro_size[:ng].long(), g.node_ptr[:ng + 1].long()
            local = torch.arange(nn, device=dev) - nptr[d.batch]
            mask = local >= key_g[d.batch]                                    # frontier nodes (policy.py:316-317)
            pi = a2c.test(d, d.batch, mask, dev, self.actor).view(-1)         # masked softmax, functional dropout 0.5 (q19)
            val = a2c.test(d, d.batch, mask, dev, self.critic).view(-1)       # V(s) per graph
            u = torch.rand(pi.numel(), device=dev, generator=self.gen).clamp_(1e-12, 1 - 1e-12)
            score = torch.zeros(nn, dtype=torch.float32, device=dev)
            score[mask] = (pi + 1e-35).log() - (-u.log()).log()               # Gumbel-max == np.random.choice(fro_size, p=pi)
            choice = env.select_and_plan(score, need_u8).long()
            slots = self.store.store_graphs(d.x, d.edge_index, d.edge_attr, d.batch, g.node_ptr, g.edge_ptr, g.key_size, g.fro_size, ng)
            slot_new = torch.where(need, slots[ordinal], torch.full_like(self.pend_slot, -1))
            fro = torch.where(need, fro_g[ordinal], torch.zeros_like(ordinal))
            key = key_g[ordinal]
            v_now = torch.where(need, val[ordinal].float(), v_now)
        else:
            fro = torch.zeros_like(self.pend_slot); key = fro; choice = fro; slot_new = torch.full_like(self.pend_slot, -1)
        # ---- the in-flight transition of every env that decides now, or whose episode ended, joins its segment ----
        closing = need & (self.pend_slot >= 0)
        fin = closing | ended
        pos = self.seg_len.clamp(max=n - 1)
        put = lambda buf, v: buf.__setitem__((self._ar, pos), torch.where(fin, v, buf[self._ar, pos]))
        put(self.seg_slot, self.pend_slot); put(self.seg_a, self.pend_a); put(self.seg_r, self.pend_r); put(self.seg_val, self.pend_val)
        put(self.seg_serial, self.pend_serial)
        put(self.seg_term, ended | self.pend_clo | (closing & (fro <= 0)))
        self.seg_len = self.seg_len + fin.long()
        self.reward_sum += float(torch.where(fin, self.pend_r, torch.zeros_like(self.pend_r)).sum())
        self.pend_slot = torch.where(fin, torch.full_like(self.pend_slot, -1), self.pend_slot)
        # ---- segments that reached nstep: last_value = V(s_t+1) at this decision (0 after the episode's end) ----
        complete = self.seg_len >= n
        n_seg = int(complete.sum())                                           # host sync
        self._train(complete, n_seg, torch.where(ended, torch.zeros_like(v_now), v_now))
        # ---- open the transitions of the envs that decided now ----
        if ng > 0:
            start = need & (fro > 0)
            r = norm.gather(1, choice.clamp(0, norm.size(1) - 1).view(-1, 1)).view(-1).float()
            self.pend_slot = torch.where(start, slot_new, self.pend_slot)
            self.pend_serial = torch.where(start, self.store.gserial[slot_new.clamp(min=0)], self.pend_serial)
            self.pend_a = torch.where(start, key + choice, self.pend_a)
            self.pend_r = torch.where(start, r, self.pend_r)
            self.pend_val = torch.where(start, v_now, self.pend_val)
            self.pend_clo = torch.where(start, clo.bool(), self.pend_clo)
        self.decisions += ng
        self.a2c.step_t += ng
        self.ticks += 1
        return ng

    def _train(self, complete, n_seg, last_value):
        """policy.py:361-393 + 474-497 for the segments that closed in this tick (one batch, one gradient step)."""
        import torch.distributed as dist
        a2c, dev, n = self.a2c, self.dev, int(self.a2c.nstep)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        total = n_seg
        if world > 1:      # every rank joins the collective of a tick in which ANY rank has a segment
            t = torch.tensor([float(n_seg)], device=dev)
            dist.all_reduce(t)
            total = int(t.item())
        if total == 0:
            return
        self.bucket.zero_()
        if n_seg > 0:
            rows = complete.nonzero().view(-1)
            r, term, vals = self.seg_r[rows], self.seg_term[rows].float(), self.seg_val[rows]
            returns = nstep_returns(r, term, last_value[rows], a2c.GAMMA)
            slots = self.seg_slot[rows].view(-1)                              # segment-major, time order inside a segment
            if not bool((self.store.gserial[slots] == self.seg_serial[rows].view(-1)).all()):
                raise RuntimeError("VecA2CTrainer: a stored graph was overwritten before its segment trained -- enlarge the graph store")
            batch, n_nodes, off = self.store.gather(slots)
            N = batch.x.size(0)
            node = torch.arange(N, device=dev)
            mask = node - off[batch.batch] >= (n_nodes - self.store.gf[slots])[batch.batch]
            at = off + self.seg_a[rows].view(-1)
            action = torch.zeros(N, device=dev); action[at] = 1.0
            y_adv = torch.zeros(N, device=dev); y_adv[at] = (returns - vals).view(-1)
            with torch.enable_grad():
                self.actor.train(); self.critic.train()
                actor_out = self.actor(batch, mask, batch=batch.batch) + 1e-35
                critic_out = self.critic(batch, mask, batch=batch.batch)
                ent = a2c.entropy_loss(actor_out)
                # per segment the reference's loss; averaged over the segments of all ranks
                loss = ((a2c.policy_cost(actor_out, y_adv, action, mask) - ent * a2c.ent_coef) / n_seg
                        + a2c.value_cost(critic_out, returns.view(-1)) * a2c.vf_coef) * (n_seg / total)
                loss.backward()
            self.last_loss, self.last_entropy = float(loss) * total / n_seg, float(ent) / n_seg
            self.seg_len = torch.where(complete, torch.zeros_like(self.seg_len), self.seg_len)
            self.segments += n_seg
        if world > 1:
            dist.all_reduce(self.bucket.flat)
        self.bucket.clamp_(a2c.max_grad_norm)                                 # policy.py:493-495
        self.optimizer.step()
        self.train_steps += 1


def nstep_returns(r: torch.Tensor, term: torch.Tensor, last_value: torch.Tensor, gamma: float) -> torch.Tensor:
    """policy.py:366-372 for S segments at once: r, term [S, nstep], last_value [S] -> discounted returns [S, nstep]
    (``ret = r_i + gamma * ret * (1 - terminal_i)`` walking the segment backwards from ``last_value``)."""
    ret, out = last_value, torch.zeros_like(r)
    for i in range(r.size(1) - 1, -1, -1):
        ret = r[:, i] + gamma * ret * (1.0 - term[:, i])
        out[:, i] = ret
    return out
