"""Device-resident acting loop: B environments follow a GNN policy with NO host synchronisation per tick.

This is the batched counterpart of the reference's ``test.py`` / ``DeepQ.test`` inner loop
(test.py:100-143, policy.py:236-306: get graph -> Q-values -> arg-max frontier -> ``env.step`` along the planned line),
scheduled for a GPU instead of one env at a time:

* every tick, each env with a queued action executes exactly one simulator step
  (move + association scan, SLAM update, virtual-map rebuild -- the *step pipeline*), and each env whose
  queue ran empty gets its exploration graph built, scored by the policy and a new line plan queued
  (the *policy pipeline*).  A freshly planned env starts moving at the next tick;
* the two pipelines touch disjoint env sets, so they run on two CUDA streams;
* finished episodes restart in-pipeline (``dge_reset_done_queued``): no extra launches, no host branch;
* for the DQN Q-network (``Networks.GCN``) the whole tick is ONE native call, ``dge_policy_tick`` (include/dge.h): the size of
  the decision batch never leaves the device (launches are sized by capacity, kernels read the live counts), so the launch
  sequence is fixed and is replayed as a CUDA graph -- one ``cudaGraphLaunch`` per tick.  Other policies (GG-NN, g-U-Net) go
  through ``model(data, 0.0)`` and pay one size sync per tick.

Per env the sequence of operations (and of Philox draws) is the same as in a sequential run of the
reference loop; only the interleaving between envs differs.
"""
from __future__ import annotations

import ctypes

import torch

from .engine import _check, _ptr
from .envs.exploration_env import RESET_ODOM, VecExplorationEnv


class GcnPolicy(ctypes.Structure):
    """``struct dge_gcn_policy`` (include/dge.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("W1", "b1", "W2t_hi", "W2t_lo", "b2", "head_w", "head_b_dev", "ws", "q", "choice")] + \
               [("node_cap", ctypes.c_int64), ("Cin", ctypes.c_int32), ("C", ctypes.c_int32)]


class DeviceTick:
    """Arguments of ``dge_policy_tick`` for one (env, Networks.GCN) pair, prepared once; re-derived when a parameter changes."""

    def __init__(self, env: VecExplorationEnv, model, seed_stride: int, graph: bool = True, overlap: bool = True, node_cap: int | None = None):
        from . import gnn
        self.env, self.model = env, model
        self.dev = env.device
        L = env.eng._L
        L.dge_policy_tick.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        self._L = L
        self.cin, self.C = (int(v) for v in model.conv1.weight.shape)
        g = env.graph
        self.node_cap = int(min(g.node_cap, node_cap or g.node_cap))
        self.ws = torch.empty(3 * self.node_cap * self.C, dtype=torch.float32, device=self.dev)
        self.q = torch.zeros(self.node_cap, dtype=torch.float32, device=self.dev)
        self.choice = torch.zeros(env.B, dtype=torch.int32, device=self.dev)
        self.flags = (1 if graph else 0) | (0 if overlap else 2)
        self.seed_stride = int(seed_stride)
        self._fo = (ctypes.c_double * 3)(*RESET_ODOM)
        self._sig = self._pol = self._keep = None
        self._gnn = gnn

    def _policy(self):
        m = self.model
        params = (m.conv1.weight, m.conv1.bias, m.conv2.weight, m.conv2.bias, m.fully_con1.weight, m.fully_con1.bias)
        sig = tuple((t.data_ptr(), t._version) if t is not None else None for t in params)
        if sig != self._sig:
            w1, b1, w2, b2, hw, hb = params
            margs, self._keep = self._gnn._q_forward_model_args(w1, b1, w2, b2, hw[0], hb)
            ptr = lambda a: None if a is None else a.value
            self._pol = GcnPolicy(*(ptr(a) for a in margs), self.ws.data_ptr(), self.q.data_ptr(), self.choice.data_ptr(), self.node_cap, self.cin, self.C)
            self._sig = sig
        return self._pol

    def __call__(self):
        env = self.env
        pol = self._policy()
        st = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        _check(self._L.dge_policy_tick(env.eng._h, ctypes.byref(env.graph.c), ctypes.byref(pol), self.seed_stride, self._fo, 4, self.flags, st), "dge_policy_tick")


class PolicyLoop:
    def __init__(self, env: VecExplorationEnv, model: torch.nn.Module, overlap: bool = True, seed_stride: int | None = None,
                 device_tick: bool | None = None, cuda_graph: bool = True):
        """``device_tick``: None = use ``dge_policy_tick`` whenever the model is the DQN Q-network (Networks.GCN on the tcgen05
        GEMM), False = the per-launch route with one size sync per tick (A/B, and the only route for other model families)."""
        from . import Networks, gnn
        self.env, self.model, self.overlap = env, model, overlap
        self.dev = env.device
        self.seed_stride = int(seed_stride or getattr(env, "seed_stride", env.B))
        self._fo = (ctypes.c_double * 3)(*RESET_ODOM)
        self.s_step = torch.cuda.Stream(self.dev) if overlap else None
        self.ev_need = torch.cuda.Event()
        self.ev_move = torch.cuda.Event()
        self.ev_step = torch.cuda.Event()
        self._launches = 0         # kernels of this package launched by the per-launch route (libdge.so + gnn kernels)
        self._graphs = 0
        self.ticks = 0
        self.stage_events = None   # optional {"slam": [], "vmap": []} of (start, end) CUDA events on the step stream (per-launch route only)
        eligible = gnn.QForwardPlan.eligible(model) and Networks._PRECISION == "tc3"
        if device_tick and not eligible:
            raise ValueError("PolicyLoop(device_tick=True) needs a Networks.GCN Q-network on the GPU with matmul precision 'tc3'")
        self.device = DeviceTick(env, model, self.seed_stride, graph=cuda_graph, overlap=overlap) if (eligible and device_tick is not False) else None
        self._c0 = None
        # Networks.GCN at inference: the Q forward's argument list is prepared once (gnn.QForwardPlan) -- every pointer of the call
        # is constant between ticks; other models (and other precision modes) go through model(g.data(), 0.0)
        self._plan = _make_plan(model, env.graph)

    # kernels per device tick: mark_pending, reset, move_measure, slam, vmap, graph count/scan/fill, conv_small, gemm, aggregate, select_plan
    # (+ split_groups and a second slam / vmap pair when the tick runs its heavy and light envs as two chains: csrc/dge_tick.cu, DGE_TICK_HEAVY)
    import os as _os
    DEVICE_TICK_LAUNCHES = 12 if _os.environ.get("DGE_TICK_HEAVY", "0") == "0" else 15

    @property
    def launches(self):
        return self._launches + (self.DEVICE_TICK_LAUNCHES * self.ticks if self.device is not None else 0)

    @launches.setter
    def launches(self, v):
        self._launches = int(v)
        if self.device is not None:
            self.ticks = 0

    @property
    def graphs(self):
        """graphs scored so far (device route: the engine's work counter, read back on demand -- a host sync)"""
        if self.device is None:
            return self._graphs
        c = int(self.env.eng.state["counters"][4].item())
        return c - (self._c0 or 0)

    @graphs.setter
    def graphs(self, v):
        self._graphs = int(v)
        if self.device is not None:
            self._c0 = int(self.env.eng.state["counters"][4].item()) - int(v)

    @torch.no_grad()
    def tick(self):
        if self.device is not None and self.stage_events is None:
            self.device()
            self.ticks += 1
            return None
        return self._tick_launches()

    @torch.no_grad()
    def _tick_launches(self):
        from . import Networks, gnn
        env, eng = self.env, self.env.eng
        L, h, st = eng._L, eng._h, eng.state
        main = torch.cuda.current_stream(self.dev)
        need = env.mark_pending()                    # before the step pipeline starts moving the queues
        if self.overlap:
            self.ev_need.record(main)
            self.s_step.wait_event(self.ev_need)     # orders the step pipeline after the previous tick's select_and_plan too
            s1 = self.s_step
        else:
            s1 = main
        sp = ctypes.c_void_p(s1.cuda_stream)
        # ---- step pipeline -------------------------------------------------------------------------------
        _check(L.dge_reset_done_queued(h, self.seed_stride, self._fo, 4, sp), "dge_reset_done_queued")
        _check(L.dge_move_measure_queued(h, sp), "dge_move_measure_queued")
        if self.overlap:
            self.ev_move.record(s1)
        timed = self.stage_events is not None
        if timed:
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(s1)
        _check(L.dge_slam_optimize(h, _ptr(st["active"]), sp), "dge_slam_optimize")
        if timed:
            e1.record(s1)
        _check(L.dge_virtual_map(h, _ptr(st["active"]), sp), "dge_virtual_map")
        if timed:
            e2.record(s1)
            self.stage_events["slam"].append((e0, e1)); self.stage_events["vmap"].append((e1, e2))
        if self.overlap:
            self.ev_step.record(s1)
        self._launches += 6
        # ---- policy pipeline -----------------------------------------------------------------------------
        g = env.build_graph(need); self._launches += 4
        ng, nn, _ = g.sync_sizes()                   # the tick's only host sync (main stream only)
        if ng > 0:
            l0 = gnn.launch_count
            q = None
            if self._plan is not None and Networks._PRECISION == "tc3":
                q = _planned_q(self, nn, None)
            if q is None:
                q = self.model(g.data(), 0.0)
            if self.overlap:
                main.wait_event(self.ev_move)        # plans are rewritten only after this tick's move kernel has read them
            env.select_and_plan(q)                   # the envs of this graph batch
            self._launches += gnn.launch_count - l0 + 1
            if self.device is None:
                self._graphs += ng
        if self.overlap:
            main.wait_event(self.ev_step)            # join: the tick ends when both pipelines are done
        return ng


class GraphPacked(ctypes.Structure):
    """``struct dge_graph_packed`` (include/dge.h): byte offsets of the sections of a packed graph batch + its totals."""
    _fields_ = [(n, ctypes.c_int64) for n in ("total_bytes", "x", "edge_index", "edge_attr", "node_ptr", "edge_ptr", "key_size", "fro_size",
                                              "frontier_xy", "csr_rowptr", "csr_perm", "gcn_norm", "gcn_selfnorm", "frontier_plan")] + \
               [(n, ctypes.c_int32) for n in ("n_graphs", "n_nodes", "n_edges", "n_done")]


class GraphHostOut(ctypes.Structure):
    """``struct dge_graph_host_out`` (include/dge.h): the caller's host buffers of ``dge_graph_host``."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("x", "edge_index", "edge_attr", "node_ptr", "edge_ptr", "key_size", "fro_size", "frontier_xy", "totals",
                                               "csr_rowptr", "csr_perm", "gcn_norm", "gcn_selfnorm")]


class HostLoop(ctypes.Structure):
    """``struct dge_host_loop`` (include/dge.h): host state, pinned buffers and device arenas of ``dge_host_policy_tick``."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("plans", "cursor", "phase", "mask", "done", "need", "obs")] + [("obs_bytes", ctypes.c_int64)] + \
               [(n, ctypes.c_void_p) for n in ("metrics", "arena_host", "q_host", "plan_host", "choice_host", "arena_pack", "arena_dev")] + \
               [(n, ctypes.c_int64) for n in ("arena_cap", "prefetch_guess", "n_stepped", "n_graphs", "n_nodes", "h2d_bytes", "d2h_bytes", "launches")]


def _make_plan(model, batch):
    """``gnn.QForwardPlan`` for a ``Networks.GCN`` Q-network on the GPU, ``None`` for every other model."""
    from . import gnn
    try:
        return gnn.QForwardPlan(model, batch) if gnn.QForwardPlan.eligible(model) else None
    except Exception as exc:      # noqa: BLE001
        import warnings
        warnings.warn(f"gnn.QForwardPlan not available ({type(exc).__name__}: {exc}); using model(data, 0.0)")
        return None


def _planned_q(loop, n_nodes: int, graph_args):
    """Q-values through the loop's ``gnn.QForwardPlan``.  The plan is glue around the same native call the module route makes; if
    the glue itself fails (anything but the native call's own error), the loop says so once and continues on the module route --
    both routes run the same kernels with the same arguments."""
    from .engine import DgeError
    try:
        return loop._plan(n_nodes, graph_args)
    except DgeError:
        raise
    except Exception as exc:      # noqa: BLE001
        import warnings
        warnings.warn(f"gnn.QForwardPlan disabled for this loop ({type(exc).__name__}: {exc}); using model(data, 0.0)")
        loop._plan = None
        return None


def packed_graph_data(arena: torch.Tensor, pk, n: int, e: int):
    """``Data`` views of a packed graph batch (``dge_graph_packed`` layout, include/dge.h) in a device arena, with the batch's
    destination CSR and GCN normalisation adopted (no preprocessing launches)."""
    from . import gnn
    from .data import Data
    f32, i32, i64 = torch.float32, torch.int32, torch.int64
    dv = lambda off, cnt, dt, sz: arena[off:off + cnt * sz].view(dt)
    x, ei, ea = dv(pk.x, n * 5, f32, 4).view(n, 5), dv(pk.edge_index, 2 * e, i64, 8).view(2, e), dv(pk.edge_attr, e, f32, 4)
    data = Data(x, ei, ea)
    data._dge_structure = gnn.GraphStructure.from_csr(ei, ea, n, dv(pk.csr_rowptr, n + 1, i32, 4), dv(pk.csr_perm, max(e, 1), i32, 4),
                                                      dv(pk.gcn_norm, max(e, 1), f32, 4), dv(pk.gcn_selfnorm, n, f32, 4))
    return data


def packed_graph_args(arena: torch.Tensor, pk):
    """The graph pointers of ``dge_gcn_q_forward`` (x, destination CSR row pointers and permutation, edge sources, GCN norm and
    self-norm) for the same arena: what ``packed_graph_data`` + ``Networks.GCN.forward`` would pass, as plain addresses."""
    base = arena.data_ptr()
    return [ctypes.c_void_p(base + int(off)) for off in (pk.x, pk.csr_rowptr, pk.csr_perm, pk.edge_index, pk.gcn_norm, pk.gcn_selfnorm)]


class HostPolicyLoop:
    """The same acting loop driven from the HOST through the host-buffer C ABI -- the batched form of what
    the reference's ``test.py:100-143`` does per env: every observation crosses to the host and every
    action comes from it.

    Per tick (B envs):
      * stepping envs: ``dge_step_host_async`` -- actions H2D from pinned memory, one simulator step, done flags +
        ``ExplorationEnv.status`` metrics + occupancy maps D2H (exploration_env.py:98-105);
      * envs whose action list ran empty: ``dge_graph_host`` (graph_matrix + data_process, host arrays out) ->
        ``Data(...).to(device)`` -> ``model`` -> Q-values D2H -> arg-max over each graph's frontier nodes on the host
        (test.py:112) -> ``dge_line_plan_host`` (actions_all_goals for the chosen frontier) -> host action lists;
      * finished episodes restart in-pipeline (``dge_reset_done_queued``); the host tracks the reset phase.
    The two pipelines work on disjoint env sets, so the policy side runs on a second stream while the step is in
    flight; NumPy work is vectorised over envs (no per-env Python).
    """

    def __init__(self, env: VecExplorationEnv, model: torch.nn.Module, overlap: bool = True, read_obs: bool = True, native: bool | None = None):
        """``native``: None = the whole tick in one native call (``dge_host_policy_tick``) whenever the model is the DQN Q-network on
        the tcgen05 GEMM and the transfer is packed; False = the same sequence of C-ABI calls issued from Python (A/B, per-section
        host timing, and the only route for other model families)."""
        import numpy as np
        from .engine import load_library
        self.np = np
        self.env, self.model, self.overlap = env, model, overlap
        eng, g, B = env.eng, env.graph, env.B
        self.dev = env.device
        L = self._L = load_library()
        vp = ctypes.c_void_p
        L.dge_step_host_async.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        L.dge_graph_host.argtypes = [vp, vp, vp, vp, vp]
        L.dge_line_plan_host.argtypes = [vp, vp, vp, vp, vp]
        L.dge_graph_packed_capacity.restype = ctypes.c_int64
        L.dge_graph_packed_capacity.argtypes = [vp, vp]
        L.dge_graph_host_packed_begin.argtypes = [vp, vp, vp, vp, ctypes.c_int64, vp]
        L.dge_graph_host_packed_end.argtypes = [vp, vp, vp, ctypes.c_int64, vp, vp]
        L.dge_graph_host_packed_prefetch.argtypes = [vp, vp, vp, ctypes.c_int64, vp]
        L.dge_graph_host_packed_end_prefetched.argtypes = [vp, vp, vp, ctypes.c_int64, ctypes.c_int64, vp, vp]
        L.dge_step_host_plans_async.argtypes = [vp, vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
        L.dge_select_plan_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()
        f64, f32, i32, i64, u8 = torch.float64, torch.float32, torch.int32, torch.int64, torch.uint8
        # step-side host buffers
        self.t_odom, self.t_mask, self.t_done = pin((B, 3), f64), pin((B,), u8), pin((B,), u8)
        self.t_metrics = pin((B, 8), f64)
        self.t_obs = pin((B, eng.rows, eng.cols), f64) if read_obs else None
        # policy-side host buffers (capacity of the device staging batch)
        self.t_need = pin((B,), u8)
        self.t_x, self.t_ei, self.t_ea = pin((g.node_cap, 5), f32), pin((2 * g.edge_cap,), i64), pin((g.edge_cap,), f32)
        self.t_nptr, self.t_eptr, self.t_ks, self.t_fs = pin((B + 1,), i32), pin((B + 1,), i32), pin((B,), i32), pin((B,), i32)
        self.t_fxy, self.t_tot = pin((B, eng.Lt + 1, 2), f64), pin((8,), i32)
        self.t_goal, self.t_plan = pin((B, 2), f64), pin((B, 6), f64)
        # packed graph transfer: one pinned arena, one device arena the pack kernel fills, one device arena the policy reads
        self.arena_cap = int(L.dge_graph_packed_capacity(eng._h, ctypes.byref(g.c)))
        self.a_host = torch.zeros(self.arena_cap, dtype=u8).pin_memory()
        self.a_pack = torch.zeros(self.arena_cap, dtype=u8, device=self.dev)
        self.a_dev = torch.zeros(self.arena_cap, dtype=u8, device=self.dev)

        self._pk = GraphPacked()
        self.packed = True
        self.t_q = pin((g.node_cap,), f32)
        self.t_choice = pin((B,), i32)
        self.t_rowptr, self.t_perm = pin((g.node_cap + 1,), i32), pin((g.edge_cap,), i32)
        self.t_norm, self.t_selfnorm = pin((g.edge_cap,), f32), pin((g.node_cap,), f32)

        self._ho = GraphHostOut(*(t.data_ptr() for t in (self.t_x, self.t_ei, self.t_ea, self.t_nptr, self.t_eptr, self.t_ks, self.t_fs, self.t_fxy, self.t_tot,
                                                     self.t_rowptr, self.t_perm, self.t_norm, self.t_selfnorm)))
        for name in ("odom", "mask", "done", "metrics", "need", "nptr", "ks", "fs", "fxy", "goal", "plan", "q", "choice"):
            setattr(self, name, getattr(self, "t_" + name).numpy())
        # host-side action lists in the compact form of dge_line_plan: (n_rot_pi, sign, rot_rem, n_fwd, fwd_rem, n_actions)
        self.plans = np.zeros((B, 6)); self.cursor = np.zeros(B, dtype=np.int64)
        self.phase = np.zeros(B, dtype=np.int64)       # ticks of reset work (initial optimize + forced steps) still to run
        self.s_step = torch.cuda.Stream(self.dev) if overlap else None
        self._fo = (ctypes.c_double * 3)(*RESET_ODOM)
        self.seed_stride = int(getattr(env, "seed_stride", B))
        self.steps = 0                                  # policy env-steps executed
        self.h2d = self.d2h = 0                         # bytes moved
        self.launches = 0
        self.graphs = 0
        self._frange = np.arange(eng.Lt + 1)
        self.timing = None                              # optional dict: host seconds per section (dev profiling)
        self._plan = _make_plan(model, g)
        self.native = native
        self.prefetch = True                            # native tick: the batch follows its header to the host at once (sized by the previous batch)
        self._native = None                             # (DeviceTick for the policy struct, HostLoop struct), built on first use
        L.dge_host_policy_tick.argtypes = [vp, vp, vp, vp, ctypes.c_uint64, vp, ctypes.c_int, vp, vp]

    def _native_tick(self):
        """One tick through ``dge_host_policy_tick``: the sequence of calls below, issued natively (include/dge.h)."""
        env, L = self.env, self._L
        if self._native is None:
            dt = DeviceTick(env, self.model, self.seed_stride, graph=False, overlap=self.overlap)
            self.plans = self.np.ascontiguousarray(self.plans, dtype=self.np.float64)
            hl = HostLoop(plans=self.plans.ctypes.data, cursor=self.cursor.ctypes.data, phase=self.phase.ctypes.data, mask=self.t_mask.data_ptr(),
                          done=self.t_done.data_ptr(), need=self.t_need.data_ptr(), obs=None if self.t_obs is None else self.t_obs.data_ptr(),
                          obs_bytes=0 if self.t_obs is None else self.t_obs.nbytes, metrics=self.t_metrics.data_ptr(), arena_host=self.a_host.data_ptr(),
                          q_host=self.t_q.data_ptr(), plan_host=self.t_plan.data_ptr(), choice_host=self.t_choice.data_ptr(),
                          arena_pack=self.a_pack.data_ptr(), arena_dev=self.a_dev.data_ptr(), arena_cap=self.arena_cap,
                          prefetch_guess=65536 if self.prefetch else 0)
            self._native = (dt, hl, (self.plans, self.cursor, self.phase))      # (the arrays whose addresses the struct holds)
        dt, hl, held = self._native
        if held[0] is not self.plans or held[1] is not self.cursor or held[2] is not self.phase:
            hl.plans, hl.cursor, hl.phase = self.plans.ctypes.data, self.cursor.ctypes.data, self.phase.ctypes.data
            self._native = (dt, hl, (self.plans, self.cursor, self.phase))
        pol = dt._policy()
        main = torch.cuda.current_stream(self.dev)
        sp = ctypes.c_void_p(self.s_step.cuda_stream) if self.overlap else None
        _check(L.dge_host_policy_tick(env.eng._h, ctypes.byref(env.graph.c), ctypes.byref(pol), ctypes.byref(hl), self.seed_stride, self._fo, 4,
                                      ctypes.c_void_p(main.cuda_stream), sp), "dge_host_policy_tick")
        self.steps += hl.n_stepped; self.graphs += hl.n_graphs
        self.h2d += hl.h2d_bytes; self.d2h += hl.d2h_bytes; self.launches += hl.launches
        return int(hl.n_stepped)

    def _next_actions(self):
        """Vectorised expansion of action `cursor` of every env's line plan (Planner2D.cpp:982-1038) into odom[B,3]."""
        np = self.np
        pl, cur = self.plans, self.cursor
        nrot, nfwd = pl[:, 0].astype(np.int64), pl[:, 3].astype(np.int64)
        od = self.odom
        od[:] = 0.0
        od[:, 2] = np.where(cur < nrot, pl[:, 1] * np.pi, np.where(cur == nrot, pl[:, 1] * pl[:, 2], 0.0))
        fwd = cur > nrot
        od[:, 0] = np.where(fwd, np.where(cur < nrot + 1 + nfwd, self.env.cfg.max_edge_length, pl[:, 4]), 0.0)

    @torch.no_grad()
    def tick(self):
        np = self.np
        if self.native is not False and self.packed and self.timing is None and self._plan is not None:
            from . import Networks
            if Networks._PRECISION == "tc3":
                return self._native_tick()
            if self.native:
                raise ValueError("HostPolicyLoop(native=True) needs matmul precision 'tc3'")
        elif self.native:
            raise ValueError("HostPolicyLoop(native=True) needs a Networks.GCN Q-network on the GPU and the packed transfer")
        env, eng, L = self.env, self.env.eng, self._L
        B = env.B
        tm = self.timing
        if tm is not None:
            import time
            t_prev = [time.perf_counter()]

            def lap(name):
                t = time.perf_counter(); tm[name] = tm.get(name, 0.0) + t - t_prev[0]; t_prev[0] = t
        else:
            def lap(name):
                pass
        main = torch.cuda.current_stream(self.dev)
        s1 = self.s_step if self.overlap else main
        nact = self.plans[:, 5].astype(np.int64)
        in_reset = self.phase > 0
        has_act = (self.cursor < nact) & ~in_reset
        need = ~has_act & ~in_reset
        if self.overlap:
            s1.wait_stream(main)
        mp = ctypes.c_void_p(main.cuda_stream)
        do_policy = bool(need.any())
        if do_policy and self.packed:
            # ---- policy pipeline, part 1 (main stream, async): graph kernels + pack; they run while the host launches the step ----
            self.need[:] = need
            _check(L.dge_graph_host_packed_begin(eng._h, self.t_need.data_ptr(), ctypes.byref(env.graph.c), self.a_pack.data_ptr(), self.arena_cap, mp),
                   "dge_graph_host_packed_begin")
            self.launches += 5
            self.h2d += B
            lap("policy: graph launch")
        # ---- step pipeline (async on s1): reset finished episodes, then one simulator step from host actions ------
        sp = ctypes.c_void_p(s1.cuda_stream)
        _check(L.dge_reset_done_queued(eng._h, self.seed_stride, self._fo, 4, sp), "dge_reset_done_queued")
        self.mask[:] = has_act | in_reset
        if self.packed:   # the action lists stay in their compact form; action `cursor` is expanded inside the call
            _check(L.dge_step_host_plans_async(eng._h, self.plans.ctypes.data, self.cursor.ctypes.data, self.t_mask.data_ptr(), self.t_done.data_ptr(),
                                               None if self.t_obs is None else self.t_obs.data_ptr(), self.t_metrics.data_ptr(), 1 | 2, sp),
                   "dge_step_host_plans_async")
        else:
            self._next_actions()
            _check(L.dge_step_host_async(eng._h, self.t_odom.data_ptr(), self.t_mask.data_ptr(), self.t_done.data_ptr(),
                                         None if self.t_obs is None else self.t_obs.data_ptr(), self.t_metrics.data_ptr(), 1 | 2, sp), "dge_step_host_async")
        self.launches += 6
        self.h2d += self.t_odom.nbytes + B
        self.d2h += B + self.t_metrics.nbytes + (0 if self.t_obs is None else self.t_obs.nbytes)
        n_stepped = int(has_act.sum())
        self.cursor[has_act] += 1
        self.phase[in_reset] -= 1
        lap("step: host prep + async launches")
        # ---- policy pipeline (main stream, host in the loop) ----------------------------------------------------
        if do_policy:
            from . import Networks, gnn
            from .data import Data
            if self.packed:
                pk = self._pk
                _check(L.dge_graph_host_packed_end(eng._h, self.a_pack.data_ptr(), self.a_host.data_ptr(), self.arena_cap, ctypes.byref(pk), mp),
                       "dge_graph_host_packed_end")
                ng, n, e, tot = pk.n_graphs, pk.n_nodes, pk.n_edges, pk.total_bytes
                lap("policy: graph D2H (2 syncs)")
                self.d2h += tot
            else:
                self.need[:] = need
                _check(L.dge_graph_host(eng._h, self.t_need.data_ptr(), ctypes.byref(env.graph.c), ctypes.byref(self._ho), mp), "dge_graph_host")
                ng, n, e = (int(v) for v in self.t_tot[:3])
                lap("policy: dge_graph_host (2 syncs)")
                self.launches += 4
                self.h2d += B
                self.d2h += 32 + n * 20 + e * 20 + (2 * ng + 2) * 4 + 2 * ng * 4 + self.t_fxy.nbytes + (n + 1) * 4 + e * 8 + n * 4
            if ng > 0:
                # the policy gets the HOST graph batch, like DeepQ.test: data.to(device) -> model -> Q back on the host
                use_plan = self.packed and self._plan is not None and Networks._PRECISION == "tc3"
                if self.packed:
                    self.a_dev[:tot].copy_(self.a_host[:tot], non_blocking=True)          # ONE H2D for the whole batch
                    if not use_plan:
                        data = packed_graph_data(self.a_dev, pk, n, e)
                    self.h2d += tot
                else:
                    up = lambda t: t.to(self.dev, non_blocking=True)
                    x, ei, ea = up(self.t_x[:n]), up(self.t_ei[:2 * e].view(2, e)), up(self.t_ea[:e])
                    data = Data(x, ei, ea)
                    # the batch's CSR + GCN normalisation travelled with it (dge_graph_host_out): adopt instead of rebuilding
                    data._dge_structure = gnn.GraphStructure.from_csr(ei, ea, n, up(self.t_rowptr[:n + 1]), up(self.t_perm[:max(e, 1)]),
                                                                      up(self.t_norm[:max(e, 1)]), up(self.t_selfnorm[:n]))
                    self.h2d += n * 20 + e * 20 + (n + 1) * 4 + e * 8 + n * 4
                    ks, fs, nptr = self.ks[:ng].astype(np.int64), self.fs[:ng].astype(np.int64), self.nptr[:ng].astype(np.int64)
                l0 = gnn.launch_count
                lap("policy: H2D graph")
                # Networks.GCN: the prepared call (gnn.QForwardPlan) on the arena's sections; any other model: Data views + module
                q = _planned_q(self, n, packed_graph_args(self.a_dev, pk)) if use_plan else None
                if q is None:
                    q = self.model(data if not use_plan else packed_graph_data(self.a_dev, pk, n, e), 0.0).view(-1)
                lap("policy: model launches")
                self.t_q[:n].copy_(q, non_blocking=True)
                main.synchronize()
                lap("policy: Q D2H (sync)")
                self.launches += gnn.launch_count - l0
                self.d2h += n * 4
                envs = np.nonzero(need)[0]
                if self.packed:
                    # arg-max over the last fro_size nodes of every graph (test.py:112) + line plan of the chosen frontier, one native call
                    _check(L.dge_select_plan_host(eng._h, self.a_host.data_ptr(), ctypes.byref(pk), self.t_q.data_ptr(), self.t_need.data_ptr(),
                                                  self.t_plan.data_ptr(), self.t_choice.data_ptr(), mp), "dge_select_plan_host")
                    lap("policy: dge_select_plan_host (sync)")
                    choice = self.choice[envs].astype(np.int64)
                    self.phase[envs[choice < 0]] = 5          # no frontier left (q15): episode over
                    # (the plans came with the batch -- its frontier_plan section: no launch, no transfer here)
                    self.plans[envs] = self.plan[envs]
                    self.cursor[envs] = 0
                    self.last_choice = (envs, choice)
                    self.graphs += ng
                    ng = 0                                    # (skips the per-array path below)
            if ng > 0:
                # arg-max over the last fro_size nodes of every graph (test.py:112), vectorised with a padded gather
                idx = (nptr + ks)[:, None] + self._frange[None, :]
                valid = self._frange[None, :] < fs[:, None]
                vals = np.where(valid, self.q[np.minimum(idx, n - 1)], -np.inf)
                choice = vals.argmax(axis=1)
                choice = np.where(fs > 0, choice, 0)
                self.goal[envs] = self.fxy[envs, choice]
                nofro = envs[fs <= 0]
                if nofro.size:                    # no frontier left (q15): episode over -- mask value 2 sets the done flag
                    self.need[nofro] = 2
                    self.phase[nofro] = 5
                lap("policy: host arg-max")
                _check(L.dge_line_plan_host(eng._h, self.t_goal.data_ptr(), self.t_need.data_ptr(), self.t_plan.data_ptr(), mp), "dge_line_plan_host")
                lap("policy: dge_line_plan_host (sync)")
                self.launches += 1
                self.h2d += self.t_goal.nbytes + B
                self.d2h += self.t_plan.nbytes
                self.plans[envs] = self.plan[envs]
                self.cursor[envs] = 0
                self.last_choice = (envs, choice)
                self.graphs += ng
        # ---- join: the step's host buffers are valid after this ---------------------------------------------------
        s1.synchronize()
        lap("join: wait for the step stream")
        done = self.done.astype(bool)
        if done.any():
            self.phase[done] = 5          # initial optimize + 4 forced steps, executed by the next 5 ticks
            self.plans[done, 5] = 0; self.cursor[done] = 0
        self.steps += n_stepped
        return n_stepped
