"""Host-side mirror of the reference's ``scripts/envs/exploration_env.py`` on the CUDA engine.

* ``VecExplorationEnv`` -- B environments stepped together (the B200-native entry point):
  device-resident graph batches for the GNN, device-side policy read-out / line planning, and a
  host-buffer API (``step_host`` / ``graph_host``) that mirrors what ``policy.py`` / ``test.py`` do
  per env (NumPy in, NumPy out).
* ``ExplorationEnv(map_size, env_index, test)`` -- the reference's single-env class
  (exploration_env.py:22-422) as a B = 1 view: ``reset / step / graph_matrix / frontier /
  actions_all_goals / status / done / get_landmark_error / max_uncertainty_of_trajectory ...``
  with the reference's return shapes (``step`` returns the 3-tuple ``(obs, done, {})``, q17).

All simulation, SLAM, virtual-map and graph arithmetic happens in ``libdge.so``; this file only
moves buffers and reshapes results.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import numpy as np
import torch

from ..config import EnvConfig, start_pose_for_seed
from ..data import Data
from ..engine import DgeError, Engine, GraphOut, _check, _ptr, _stream_ptr

RESET_ODOM = (1.0, 1.0, math.pi / 2.0)   # exploration_env.py:411-414: 4 x simulate((1, 1, pi/2))


class Pose2:
    """Duck-type of ss2d.Pose2 for the actions handed to ``step`` (x, y, theta)."""
    __slots__ = ("x", "y", "theta")

    def __init__(self, x=0.0, y=0.0, theta=0.0):
        self.x, self.y, self.theta = float(x), float(y), float(theta)

    def __repr__(self):
        return f"Pose2({self.x:.6g}, {self.y:.6g}, {self.theta:.6g})"


def expand_plan(plan_row, max_edge_length: float):
    """(n_rot_pi, sign, rot_rem, n_fwd, fwd_rem, n_actions) -> list[Pose2]  (Planner2D.cpp:982-1038)"""
    nrot, sign, rrem, nfwd, frem, _ = [float(v) for v in plan_row]
    acts = [Pose2(0, 0, sign * math.pi) for _ in range(int(nrot))]
    acts.append(Pose2(0, 0, sign * rrem))
    acts += [Pose2(max_edge_length, 0, 0) for _ in range(int(nfwd))]
    acts.append(Pose2(frem, 0, 0))
    return acts


class GraphBatch:
    """Device-resident batched exploration graph (PyG DataLoader layout) + per-graph bookkeeping."""

    def __init__(self, eng: Engine, n_graph_cap: int):
        dev = eng.device
        self.node_cap = n_graph_cap * eng.node_cap_env
        self.edge_cap = n_graph_cap * eng.edge_cap_env
        self.x = torch.empty(self.node_cap, 5, dtype=torch.float32, device=dev)
        self.edge_index = torch.empty(2, self.edge_cap, dtype=torch.int64, device=dev)
        self.edge_attr = torch.empty(self.edge_cap, dtype=torch.float32, device=dev)
        self.batch = torch.empty(self.node_cap, dtype=torch.int64, device=dev)
        self.node_ptr = torch.zeros(eng.B + 1, dtype=torch.int32, device=dev)
        self.edge_ptr = torch.zeros(eng.B + 1, dtype=torch.int32, device=dev)
        self.key_size = torch.zeros(eng.B, dtype=torch.int32, device=dev)
        self.fro_size = torch.zeros(eng.B, dtype=torch.int32, device=dev)
        self.frontier_xy = torch.zeros(eng.B, eng.Lt + 1, 2, dtype=torch.float64, device=dev)
        self.totals = torch.zeros(8, dtype=torch.int32, device=dev)
        self.totals_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.csr_rowptr = torch.zeros(self.node_cap + 1, dtype=torch.int32, device=dev)
        self.csr_perm = torch.zeros(self.edge_cap, dtype=torch.int32, device=dev)
        self.gcn_norm = torch.zeros(self.edge_cap, dtype=torch.float32, device=dev)
        self.gcn_selfnorm = torch.zeros(self.node_cap, dtype=torch.float32, device=dev)
        self.c = GraphOut(self.x.data_ptr(), self.edge_index.data_ptr(), self.edge_attr.data_ptr(), self.batch.data_ptr(),
                          self.node_ptr.data_ptr(), self.edge_ptr.data_ptr(), self.key_size.data_ptr(), self.fro_size.data_ptr(),
                          self.frontier_xy.data_ptr(), self.totals.data_ptr(), self.node_cap, self.edge_cap,
                          self.csr_rowptr.data_ptr(), self.csr_perm.data_ptr(), self.gcn_norm.data_ptr(), self.gcn_selfnorm.data_ptr())
        self.n_graphs = self.n_nodes = self.n_edges = self.n_done = 0

    def sync_sizes(self):
        """One small D2H (pinned) so the host can size the dense GEMMs of the GNN."""
        self.totals_host.copy_(self.totals, non_blocking=True)
        torch.cuda.current_stream(self.x.device).synchronize()
        g, n, e, ovf, self.n_done = self.totals_host.tolist()[:5]
        if ovf:
            raise DgeError("graph batch capacity exceeded")
        self.n_graphs, self.n_nodes, self.n_edges = g, n, e
        return g, n, e

    def data(self) -> Data:
        """``Data(x, edge_index, edge_attr)`` views of the valid prefix (no copies)."""
        d = Data(self.x[:self.n_nodes], self.edge_index[:, :self.n_edges], self.edge_attr[:self.n_edges], self.batch[:self.n_nodes])
        d.num_graphs = self.n_graphs
        # the graph kernel already produced the destination-sorted CSR and the GCN normalisation:
        # hand them to the GNN so that it launches no preprocessing kernels
        from ..gnn import GraphStructure
        d._dge_structure = GraphStructure.from_csr(d.edge_index, d.edge_attr, self.n_nodes, self.csr_rowptr[:self.n_nodes + 1],
                                                   self.csr_perm[:max(self.n_edges, 1)], self.gcn_norm[:max(self.n_edges, 1)],
                                                   self.gcn_selfnorm[:self.n_nodes])
        return d


class VecExplorationEnv:
    """B exploration environments resident on one B200."""

    def __init__(self, n_envs: int, map_size: int = 40, cfg: Optional[EnvConfig] = None, max_poses: int = 512, device=0,
                 seed0: int = 0, test: bool = True, seed_stride: Optional[int] = None):
        self.cfg = cfg or EnvConfig(map_size=map_size)
        self.map_size = self.cfg.map_size
        self.eng = Engine(self.cfg, n_envs, max_poses=max_poses, device=device)
        self.B, self.device = n_envs, self.eng.device
        self.test = test
        self.graph = GraphBatch(self.eng, n_envs)
        self._seeds = torch.arange(seed0, seed0 + n_envs, dtype=torch.int64, device=self.device)
        # an env's next world = its seed + seed_stride.  One process: B.  Env-sharded over W ranks: seed0 = rank * B and seed_stride = W * B keep the
        # ranks' world sequences disjoint for ever
        self.seed_stride = int(seed_stride or n_envs)
        self._next_seed = seed0 + n_envs
        self._reset_odom = torch.tensor([RESET_ODOM] * n_envs, dtype=torch.float64, device=self.device)
        self._choice = torch.zeros(n_envs, dtype=torch.int32, device=self.device)
        self._ones = torch.ones(n_envs, dtype=torch.uint8, device=self.device)
        self.episodes_done = 0

    # ---------------------------------------------------------------- reset ---
    def reset(self, mask: Optional[torch.Tensor] = None, seeds: Optional[torch.Tensor] = None, reference_worlds: bool = False, regenerate: bool = True):
        """exploration_env.py:389-422.  ``reference_worlds`` reproduces the reference's start poses
        (legacy NumPy RNG, pyss2d.py:88-95) on the host; landmarks/noise stay Philox (device)."""
        eng = self.eng
        if seeds is not None:
            self._seeds = seeds.to(self.device, torch.int64).contiguous()
        start = None
        if reference_worlds:
            sp = np.array([start_pose_for_seed(int(s), self.map_size, self.cfg.ext) for s in self._seeds.tolist()], dtype=np.float64)
            start = torch.as_tensor(sp, device=self.device)
        eng._L.dge_set_counting(eng._h, 0)          # the 4 forced steps are not policy steps
        eng.reset(self._seeds, mask=mask, start=start)
        for _ in range(4):
            eng.step(self._reset_odom, mask=mask)
        eng._L.dge_set_counting(eng._h, 1)
        if regenerate and not reference_worlds:
            # exploration_env.py:416-419: a world whose forced steps saw no landmark is regenerated (the in-pipeline resets apply the same
            # rule on the device: csrc/dge_vmap.cu ends such an episode at once and the next reset draws the env's next world)
            for _ in range(64):
                blind = eng.state["observed"].sum(dim=1) == 0
                if mask is not None:
                    blind &= mask.bool()
                if not bool(blind.any()):            # (one host sync per eager reset)
                    break
                self._seeds = torch.where(blind, self._seeds + self.seed_stride, self._seeds)
                bm = blind.to(torch.uint8)
                eng._L.dge_set_counting(eng._h, 0)
                eng.reset(self._seeds, mask=bm)
                for _ in range(4):
                    eng.step(self._reset_odom, mask=bm)
                eng._L.dge_set_counting(eng._h, 1)
        return eng.state["prob"]

    def reset_done(self, in_pipeline: bool = False):
        """Re-seed and reset the envs whose episode ended (device-side mask, no host sync).
        ``in_pipeline``: only the world generation is launched now; the initial optimize() and the 4 forced
        steps are executed by the next 5 ``step_queued`` ticks together with the other envs' steps (same
        per-env operation sequence and Philox draws as the eager reset, one launch instead of 21)."""
        done = self.eng.state["done"].clone()
        self._seeds = torch.where(done.bool(), self._seeds + self.seed_stride, self._seeds)
        if in_pipeline:
            self.eng.reset_queued(self._seeds, done, RESET_ODOM, 4)
        else:
            self.reset(mask=done)
        return done

    # ----------------------------------------------------------------- step ---
    def step(self, odom: torch.Tensor, mask: Optional[torch.Tensor] = None):
        self.eng.step(odom, mask=mask)
        return self.eng.state["prob"], self.eng.state["done"]

    def step_queued(self):
        self.eng.step_queued()
        return self.eng.state["done"]

    def needs_decision(self) -> torch.Tensor:
        """[B] u8: envs whose action queue is empty (and are not done)."""
        st = self.eng.state
        return ((st["plan_cursor"] >= st["plan"][:, 5].to(torch.int32)) & (st["done"] == 0) & (st["forced"] == 0)).to(torch.uint8)

    # ---------------------------------------------------------------- graph ---
    def build_graph(self, mask: Optional[torch.Tensor] = None) -> GraphBatch:
        eng = self.eng
        _check(eng._L.dge_graph(eng._h, _ptr(mask), ctypes.byref(self.graph.c), _stream_ptr(self.device)), "dge_graph")
        return self.graph

    def mark_pending(self) -> torch.Tensor:
        """``needs_decision`` evaluated by one small kernel into the engine-owned ``pending`` mask."""
        eng = self.eng
        _check(eng._L.dge_mark_pending(eng._h, _stream_ptr(self.device)), "dge_mark_pending")
        return eng.state["pending"]

    def select_and_plan(self, q: torch.Tensor, mask: Optional[torch.Tensor] = None):
        """arg-max over each graph's frontier nodes + line plan into the env queues (device-side)."""
        eng = self.eng
        q = q.contiguous().view(-1).float()
        _check(eng._L.dge_select_and_plan(eng._h, ctypes.byref(self.graph.c), _ptr(q), _ptr(mask), _ptr(self._choice),
                                          _stream_ptr(self.device)), "dge_select_and_plan")
        return self._choice

    def line_plan(self, goals: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        eng = self.eng
        out = torch.zeros(self.B, 6, dtype=torch.float64, device=self.device)
        _check(eng._L.dge_line_plan(eng._h, _ptr(goals.contiguous()), _ptr(mask), _ptr(out), _stream_ptr(self.device)), "dge_line_plan")
        return out

    # ------------------------------------------------------------- roll-outs ---
    def rollout_rewards(self, mask: Optional[torch.Tensor] = None, clone_slots: Optional[int] = None, noise: Optional[torch.Tensor] = None,
                        auto_steps: bool = False):
        """rewards_all_goals (exploration_env.py:145-162) for every env selected in the last
        ``build_graph(mask)``: one clone per (env, frontier) in a second engine, all clones stepped
        together through their line plans.  Returns (raw [B,Fmax], normalised [B,Fmax], loop_clo [B]).
        ``noise`` [n_steps, clone_slots, 3+4*Lt] makes the roll-out noise explicit (parity tests).
        ``auto_steps``: read the longest clone plan back (one host sync) and launch exactly that many clone steps instead of
        the upper bound of a line plan (typically 4-10 instead of 3 + diagonal / max_edge_length).
        The caller keeps sum(fro_size) of the selected envs <= clone_slots (``VecDQNTrainer`` chunks the decision round)."""
        eng = self.eng
        max_actions = self.cfg.max_plan_actions                    # <= 2 rotations + floor(d / max_edge) forward steps + remainder
        if getattr(self, "_roll", None) is not None and clone_slots is not None and clone_slots != self._roll.B:
            raise DgeError(f"rollout_rewards: the clone engine was created with {self._roll.B} slots, not {clone_slots}")
        if getattr(self, "_roll", None) is None:
            slots = clone_slots or min(self.B * (eng.Lt + 1), max(4 * self.B, 512))
            if slots < eng.Lt + 1:
                raise DgeError(f"rollout_rewards: clone_slots = {slots} cannot hold the frontiers of one env (up to {eng.Lt + 1})")
            # a clone of an env near its pose capacity must still be able to run its whole line plan
            self._roll = Engine(self.cfg, slots, max_poses=eng.Tmax + max_actions, device=self.device)
            self._roll_totals = torch.zeros(2, dtype=torch.int32, device=self.device)
            self._roll_totals_host = torch.zeros(2, dtype=torch.int32).pin_memory()
            fm = eng.Lt + 1
            self._roll_raw = torch.zeros(self.B, fm, dtype=torch.float64, device=self.device)
            self._roll_norm = torch.zeros(self.B, fm, dtype=torch.float64, device=self.device)
            self._roll_clo = torch.zeros(self.B, dtype=torch.uint8, device=self.device)
            L = eng._L
            vp = ctypes.c_void_p
            L.dge_rollout_prepare.argtypes = [vp, vp, vp, vp, vp, vp]
            L.dge_rollout_rewards.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
            L.dge_step_queued_noise.argtypes = [vp, vp, vp]
        roll, sp = self._roll, _stream_ptr(self.device)
        roll._L.dge_set_counting(roll._h, 0)
        _check(eng._L.dge_rollout_prepare(roll._h, eng._h, ctypes.byref(self.graph.c), _ptr(mask), _ptr(self._roll_totals), sp), "dge_rollout_prepare")
        n_steps = max_actions
        if noise is not None:
            n_steps = noise.shape[0]
        elif auto_steps:     # one host sync: the longest clone plan and the overflow flag of the clone map
            self._roll_totals_host.copy_(self._roll_totals, non_blocking=True)
            n_steps = min(n_steps, int(roll.state["plan"][:, 5].max().item()))
            if int(self._roll_totals_host[1]):
                raise DgeError("rollout_rewards: more (env, frontier) clones than the clone engine has slots; chunk the decision round or raise clone_slots")
        self.rollout_steps = n_steps
        for i in range(n_steps):
            _check(roll._L.dge_step_queued_noise(roll._h, _ptr(None if noise is None else noise[i].contiguous()), sp), "dge_step_queued_noise")
        _check(eng._L.dge_rollout_rewards(roll._h, eng._h, ctypes.byref(self.graph.c), _ptr(mask), _ptr(self._roll_raw), _ptr(self._roll_norm),
                                          _ptr(self._roll_clo), sp), "dge_rollout_rewards")
        return self._roll_raw, self._roll_norm, self._roll_clo

    def rollout_overflowed(self) -> bool:
        """True if the last ``rollout_rewards`` asked for more clones than the clone engine holds (host sync).  ``auto_steps``
        checks this by itself; the explicit-noise / fixed-step modes leave the check to the caller."""
        return bool(int(self._roll_totals[1].item()))

    def capacity_overflows(self) -> torch.Tensor:
        """[B] bool: envs whose episode was ended by a full trajectory buffer (status DGE_ECAP) rather than by the reference's
        own rules (explored > 0.85 or step > max_steps, exploration_env.py:166-168)."""
        return self.eng.state["status"] == -4

    # ------------------------------------------------------------- host API ---
    def step_host(self, odom_host: np.ndarray, done_host: np.ndarray, obs_host: Optional[np.ndarray] = None):
        self.eng.step_host(odom_host, done_host, obs_host)

    def graph_host(self, mask: Optional[torch.Tensor] = None):
        """Per-env NumPy graphs like ``graph_matrix`` + ``data_process`` produce them (D2H copies)."""
        g = self.build_graph(mask)
        ng, n, e = g.sync_sizes()
        x = g.x[:n].cpu().numpy(); ei = g.edge_index[:, :e].cpu().numpy(); ea = g.edge_attr[:e].cpu().numpy()
        nptr = g.node_ptr[:ng + 1].cpu().numpy(); eptr = g.edge_ptr[:ng + 1].cpu().numpy()
        ks = g.key_size[:ng].cpu().numpy(); fs = g.fro_size[:ng].cpu().numpy()
        out = []
        for i in range(ng):
            sl, se = slice(nptr[i], nptr[i + 1]), slice(eptr[i], eptr[i + 1])
            out.append(dict(x=x[sl], edge_index=ei[:, se] - nptr[i], edge_attr=ea[se], key_size=int(ks[i]), fro_size=int(fs[i])))
        return out

    def metrics(self):
        m = self.eng.state["metrics"]
        return dict(explored=m[:, 0], utility0=m[:, 1], landmark_error=m[:, 4], max_traj_uncertainty=m[:, 5], dist=m[:, 6])

    def close(self):
        self.eng.close()


class ExplorationEnv:
    """The reference's gym-style single env (exploration_env.py:22) on the CUDA engine (B = 1)."""

    def __init__(self, map_size, env_index, test, num_landmarks: Optional[int] = None, max_poses: Optional[int] = None, device=0):
        self.map_size, self.env_index, self.test = map_size, env_index, test
        self.dist = 0.0
        self._cfg = EnvConfig(map_size=map_size, num_landmarks=num_landmarks)
        if max_poses is None:     # the reference runs an episode until step > max_steps (exploration_env.py:168): no earlier cap here
            max_poses = int(self._cfg.max_steps) + 8
        self._vec = VecExplorationEnv(1, cfg=self._cfg, max_poses=max_poses, device=device, test=test)
        self._np_random = np.random.default_rng()   # q20: training envs are seeded from an unseeded RNG
        self._max_steps = self._cfg.max_steps
        self.max_step = self._max_steps
        self.map_resolution = self._cfg.resolution
        self.leng_i_map, self.leng_j_map = self._cfg.rows, self._cfg.cols
        self.ext = self._cfg.ext
        self.loop_clo = False
        self.nearest_frontier_point = 0
        self._frontier, self._frontier_index = [], []
        self._done = False
        self._obs = self.reset()

    # -- helpers --
    def _st(self, name):
        return self._vec.eng.state[name][0]

    def _get_obs(self):
        self._obs = self._st("prob").cpu().numpy()
        return self._obs

    def reset(self):
        self._done = False
        while True:
            if not self.test:
                seed = int(self._np_random.integers(0, np.iinfo(np.int32).max))
            else:
                seed = int(self.env_index)
            seeds = torch.tensor([seed], dtype=torch.int64, device=self._vec.device)
            if self.test:
                self._reset_reference_world(seed, seeds)
            else:
                self._vec.reset(seeds=seeds, reference_worlds=True)
            if int(self._st("observed").sum()) < 1:    # exploration_env.py:416-419
                print("regenerate a environment")
                self.env_index = self.env_index + 50
                continue
            self.dist = 0.0
            return self._get_obs()

    # -- test=True: the reference's own world and noise streams (exploration_env.py:389-407 seeds everything with env_index) --
    def _ref_noise(self, odom3):
        row = np.zeros((1, self._vec.eng.noise_len))
        od = (ctypes.c_double * 3)(*[float(v) for v in odom3])
        _check(self._vec.eng._L.dge_refworld_step(self._ref, od, row.ctypes.data_as(ctypes.c_void_p)), "dge_refworld_step")
        return torch.as_tensor(row, device=self._vec.device)

    def _reset_reference_world(self, seed, seeds):
        """SS2D.__init__ + the four forced steps with the landmarks, visiting order and noise of the reference's libstdc++
        streams (``dge_refworld_*``, csrc/dge_refworld.cu): ``ExplorationEnv(40, 0, True)`` IS the world of test.py's seed-0 run."""
        eng, L = self._vec.eng, self._vec.eng._L
        if not hasattr(L, "_refworld_ready"):
            L.dge_refworld_create.restype = ctypes.c_void_p
            L.dge_refworld_create.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
            L.dge_refworld_destroy.argtypes = [ctypes.c_void_p]
            L.dge_refworld_destroy.restype = None
            L.dge_refworld_world.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
            L.dge_refworld_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
            L._refworld_ready = True
        if getattr(self, "_ref", None):
            L.dge_refworld_destroy(self._ref)
        start = np.array(start_pose_for_seed(seed, self.map_size, self._cfg.ext), dtype=np.float64)
        cs = self._cfg.to_struct()
        self._ref = ctypes.c_void_p(L.dge_refworld_create(ctypes.byref(cs), ctypes.c_uint32(seed), start.ctypes.data_as(ctypes.c_void_p)))
        if not self._ref:
            raise DgeError("dge_refworld_create failed")
        lm, scan, n0 = np.zeros((1, eng.Lt, 2)), np.zeros((1, eng.Lt), dtype=np.int32), np.zeros((1, eng.noise_len))
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _check(L.dge_refworld_world(self._ref, vp(lm), vp(scan), vp(n0)), "dge_refworld_world")
        t = lambda a: torch.as_tensor(a, device=self._vec.device)
        L.dge_set_counting(eng._h, 0)
        eng.reset(seeds, start=t(start[None]), landmarks=t(lm), scan=t(scan), noise=t(n0))
        for _ in range(4):
            eng.step(self._vec._reset_odom, noise=self._ref_noise(RESET_ODOM))
        L.dge_set_counting(eng._h, 1)

    def step(self, action):
        odom = torch.tensor([[action.x, action.y, action.theta]], dtype=torch.float64, device=self._vec.device)
        if self.test:
            self._vec.eng.step(odom, noise=self._ref_noise((action.x, action.y, action.theta)))
        else:
            self._vec.step(odom)
        if int(self._st("status")) == -4:      # DGE_ECAP: the trajectory buffer is full -- not an episode end the reference knows
            raise DgeError(f"ExplorationEnv: pose capacity ({self._vec.eng.Tmax}) exhausted; create the env with a larger max_poses")
        self.dist = self.dist + math.sqrt(action.x ** 2 + action.y ** 2)
        return self._get_obs(), self.done(), {}

    def status(self):
        return float(self._st("metrics")[0])

    def done(self):
        return self._done or bool(self._st("done"))

    def get_landmark_error(self, sigma0=1.0):
        return float(self._st("metrics")[4])

    def max_uncertainty_of_trajectory(self):
        return float(self._st("metrics")[5])

    def get_dist(self):
        return self.dist

    def get_landmark_size(self):
        return int(self._st("observed").sum())

    def get_key_size(self):
        return self.get_landmark_size() + int(self._st("n_poses"))

    def graph_matrix(self):
        """(adjacency [N,N] f64, features [N,5] f64, global_features [1], fro_size) like exploration_env.py:196-281.
        The dense adjacency is assembled on the host from the device COO only for API compatibility."""
        g = self._vec.graph_host()[0]
        n, k, f = g["x"].shape[0], g["key_size"], g["fro_size"]
        adj = np.zeros((n, n))
        adj[g["edge_index"][0], g["edge_index"][1]] = g["edge_attr"]
        feats = g["x"].astype(np.float64)
        fxy = self._vec.graph.frontier_xy[0, :f].cpu().numpy()
        self._frontier = [list(p) for p in fxy]
        self.nearest_frontier_point = k
        self._last_graph = g
        land = self.get_landmark_size()
        glob = np.array([np.mean(feats[1:land + 1][:, 0])]) if land > 0 else np.array([np.nan])
        return adj, feats, glob, f

    get_graph = graph_matrix   # BASELINE.json's name for it

    def frontier(self):
        self.graph_matrix()
        return self._frontier

    def line_plan(self, goal_key, fro=(0, 0)):
        goal = torch.tensor([[fro[0], fro[1]]], dtype=torch.float64, device=self._vec.device)
        plan = self._vec.line_plan(goal)[0].cpu().numpy()
        return expand_plan(plan, self._cfg.max_edge_length)

    def actions_all_goals(self):
        key_size, fro_size = self.get_key_size(), len(self._frontier)
        all_actions = [[]] * (key_size + fro_size)
        for i, vi in enumerate(self._frontier):
            all_actions[i + key_size] = self.line_plan(key_size, vi)
        return all_actions

    def rewards_all_goals(self, all_actions=None):
        """exploration_env.py:145-162: look-ahead reward of every frontier (roll-outs on cloned simulator / SLAM / virtual
        map, batched on the clone engine), min-max normalised; 0 for the SLAM nodes; sets ``loop_clo``."""
        key_size, fro_size = self.get_key_size(), len(self._frontier)
        rewards = np.zeros(key_size + fro_size)
        if fro_size > 0:
            _, norm, clo = self._vec.rollout_rewards(clone_slots=self._vec.eng.Lt + 1, auto_steps=True)
            rewards[key_size:] = norm[0, :fro_size].cpu().numpy()
            self.loop_clo = bool(clo[0])
        return rewards

    def is_nf(self, id):
        return self.nearest_frontier_point == id

    def index2coor(self, matrix_i, matrix_j):
        half = self.map_size / 2 + self.ext
        return [(matrix_j + 0.5) * self.map_resolution - half, (matrix_i + 0.5) * self.map_resolution - half]

    def coor2index(self, x, y):
        half = self.map_size / 2 + self.ext
        return [int(round((y + half) / self.map_resolution - 0.5)), int(round((x + half) / self.map_resolution - 0.5))]

    def close(self):
        if getattr(self, "_ref", None):
            self._vec.eng._L.dge_refworld_destroy(self._ref)
            self._ref = None
        self._vec.close()
