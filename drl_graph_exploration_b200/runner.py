"""Device-resident acting loop: B environments follow a GNN policy with one host sync per tick.

This is the batched counterpart of the reference's ``test.py`` / ``DeepQ.test`` inner loop
(policy.py:236-306: get graph -> Q-values -> arg-max frontier -> ``env.step`` along the planned line),
scheduled for a GPU instead of one env at a time:

* every tick, each env with a queued action executes exactly one simulator step
  (move + association scan, SLAM update, virtual-map rebuild -- the *step pipeline*), and each env whose
  queue ran empty gets its exploration graph built, scored by the policy and a new line plan queued
  (the *policy pipeline*).  A freshly planned env starts moving at the next tick;
* the two pipelines touch disjoint env sets, so they run on two CUDA streams: the host-side size sync of
  the graph batch and the GNN's dense GEMMs hide under the latency-bound SLAM kernel;
* finished episodes restart in-pipeline (``dge_reset_done_queued``): no extra launches, no host branch.

Per env the sequence of operations (and of Philox draws) is the same as in a sequential run of the
reference loop; only the interleaving between envs differs.
"""
from __future__ import annotations

import ctypes

import torch

from .engine import _check, _ptr
from .envs.exploration_env import RESET_ODOM, VecExplorationEnv


class PolicyLoop:
    def __init__(self, env: VecExplorationEnv, model: torch.nn.Module, overlap: bool = True, seed_stride: int | None = None):
        self.env, self.model, self.overlap = env, model, overlap
        self.dev = env.device
        self.seed_stride = int(seed_stride or env.B)
        self._fo = (ctypes.c_double * 3)(*RESET_ODOM)
        self.s_step = torch.cuda.Stream(self.dev) if overlap else None
        self.ev_need = torch.cuda.Event()
        self.ev_move = torch.cuda.Event()
        self.ev_step = torch.cuda.Event()
        self.launches = 0          # kernels of this package launched so far (libdge.so + gnn kernels; cuBLAS GEMMs not counted)
        self.graphs = 0            # graphs scored so far
        self.stage_events = None   # optional {"slam": [], "vmap": []} of (start, end) CUDA events on the step stream

    @torch.no_grad()
    def tick(self):
        from . import gnn
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
        self.launches += 6
        # ---- policy pipeline -----------------------------------------------------------------------------
        g = env.build_graph(need); self.launches += 4
        ng, _, _ = g.sync_sizes()                    # the tick's only host sync (main stream only)
        if ng > 0:
            l0 = gnn.launch_count
            q = self.model(g.data(), 0.0)
            if self.overlap:
                main.wait_event(self.ev_move)        # plans are rewritten only after this tick's move kernel has read them
            env.select_and_plan(q)                   # the envs of this graph batch
            self.launches += gnn.launch_count - l0 + 1
            self.graphs += ng
        if self.overlap:
            main.wait_event(self.ev_step)            # join: the tick ends when both pipelines are done
        return ng
