"""The reference path on host cores, tick-structured like the GPU arm of bench.py:
CPU-oracle envs (each single-threaded, like the reference) spread over a persistent C++ worker
pool + the pure-PyTorch GCN restatement on CPU.  TEST INFRASTRUCTURE (bench.py cpu_baseline /
--impl reference only)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from drl_graph_exploration_b200.config import EnvConfig, start_pose_for_seed
from oracle import gnn_ref
from oracle.oracle import lib as _lib

MAX_ACTIONS = 64


def cpu_reference(map_size: int, n_landmarks: int, n_envs: int, threads: int, max_poses: int, seed0: int = 0, gnn_threads: int | None = None,
                  overlap: bool = True):
    """All ``n_envs`` envs live on the pool (n_envs / threads per worker).  One tick, like the GPU arm: the envs whose action list
    ran empty get a decision -- graphs built on the pool, ONE GCN forward over the whole decision batch on the torch threads -- while
    (``overlap``) the other envs are already stepping on the pool; the deciding envs step when their plans are there."""
    import threading
    cfg = EnvConfig(map_size=map_size, num_landmarks=n_landmarks)
    L = _lib()
    cs = cfg.to_struct()
    torch.manual_seed(0)
    model = gnn_ref.GCN().eval()
    torch.set_num_threads(gnn_threads or min(threads, 64))
    vp = ctypes.c_void_p
    handles = [L.orc_create(ctypes.byref(cs)) for _ in range(n_envs)]
    p = lambda a: a.ctypes.data_as(vp)
    state = {"next_seed": seed0, "steps": 0, "sumT": 0}
    harr = lambda idx: (vp * len(idx))(*[handles[i] for i in idx])

    def fresh(idx):
        idx = np.asarray(idx, dtype=np.int64)
        seeds = np.arange(state["next_seed"], state["next_seed"] + len(idx), dtype=np.uint32)
        state["next_seed"] += len(idx)
        starts = np.array([start_pose_for_seed(int(s), map_size, cfg.ext) for s in seeds], dtype=np.float64)
        L.orc_batch_fresh(harr(idx), p(seeds), p(starts), len(idx), threads)

    fresh(range(n_envs))
    queues = [[] for _ in range(n_envs)]
    Ts = np.zeros(n_envs, dtype=np.int32)

    def step(idx):
        """one simulator step of the envs ``idx`` on the pool; finished episodes restart"""
        if not idx:
            return
        odoms = np.stack([queues[i].pop(0) for i in idx]).astype(np.float64)
        done = np.zeros(len(idx), dtype=np.uint8); T = np.zeros(len(idx), dtype=np.int32)
        L.orc_batch_step(harr(idx), p(odoms), len(idx), threads, p(done), p(T))
        state["steps"] += len(idx)
        Ts[idx] = T
        redo = [i for k, i in enumerate(idx) if done[k] or T[k] >= max_poses - 1]
        if redo:
            fresh(redo)
            for i in redo:
                queues[i] = []
                Ts[i] = 5

    def run_tick():
        need = [i for i in range(n_envs) if not queues[i]]
        ready = [i for i in range(n_envs) if queues[i]]
        if need:
            hs = harr(need)
            sizes = np.zeros((len(need), 4), dtype=np.int32)
            L.orc_batch_graph_build(hs, len(need), threads, p(sizes))
            ntot, etot = int(sizes[:, 0].sum()), int(sizes[:, 3].sum())
            x = np.zeros((ntot, 5), dtype=np.float32); ei = np.zeros((2, etot), dtype=np.int64); ea = np.zeros(etot, dtype=np.float32)
            L.orc_batch_graph_fetch(len(need), p(x), p(ei), p(ea), ctypes.c_int64(etot))
            box = {}

            def forward():
                with torch.no_grad():
                    box["q"] = model(gnn_ref.Graph(torch.from_numpy(x), torch.from_numpy(ei), torch.from_numpy(ea)), 0.0).view(-1).numpy()

            if overlap:       # the GCN forward (torch threads, GIL released inside the kernels) beside the pool stepping the other envs
                th = threading.Thread(target=forward); th.start()
                step(ready); ready = []
                th.join()
            else:
                forward()
            q = box["q"]
            choice = np.full(len(need), -1, dtype=np.int32)
            off = 0
            for k in range(len(need)):
                n, K, F = sizes[k, 0], sizes[k, 1], sizes[k, 2]
                if F > 0:
                    choice[k] = int(np.argmax(q[off + K: off + n]))
                off += n
            plans = np.zeros((len(need), MAX_ACTIONS, 3)); counts = np.zeros(len(need), dtype=np.int32)
            L.orc_batch_line_plan(hs, len(need), p(choice), MAX_ACTIONS, p(plans), p(counts))
            for k, i in enumerate(need):
                queues[i] = [plans[k, a] for a in range(counts[k])] if counts[k] > 0 else [np.array([0.0, 0.0, 0.5])]
        step(ready + need)
        state["sumT"] = int(Ts.sum())

    run_tick.mean_poses = lambda: state["sumT"] / n_envs
    return run_tick, (lambda: state["steps"])


def timed_reference(map_size, n_landmarks, n_envs, threads, max_poses, preroll, warmup, min_ticks, min_seconds, max_seconds, seed0=0):
    """Pre-roll (episodes de-synchronise: the trajectory-length distribution must be the steady state the GPU arm is timed in), warm up,
    then time ticks until BOTH ``min_ticks`` and ``min_seconds`` are reached (or ``max_seconds``).  -> dict"""
    import time
    run_tick, count = cpu_reference(map_size, n_landmarks, n_envs, threads, max_poses, seed0=seed0)
    hist = []
    for i in range(preroll):
        run_tick()
        if (i + 1) % 50 == 0:
            hist.append(round(run_tick.mean_poses(), 1))
    for _ in range(warmup):
        run_tick()
    c0, t0, nt, sT = count(), time.perf_counter(), 0, 0.0
    while True:
        run_tick(); nt += 1; sT += run_tick.mean_poses()
        dt = time.perf_counter() - t0
        if (nt >= min_ticks and dt >= min_seconds) or dt >= max_seconds:
            break
    return {"value": (count() - c0) / dt, "ticks": nt, "seconds": dt, "mean_poses": sT / nt, "preroll_mean_poses": hist, "n_envs": n_envs, "threads": threads}


def cpu_gnn_baseline(batches, threads: int, budget_s: float = 8.0):
    """GNN samples/sec of the reference's GCN on host cores: the pure-PyTorch restatement of the PyG layers (gnn_ref.GCN, fp32, MKL
    on ``threads`` threads) on the bench's own C5 batches -- forward, and forward + backward + Adam like DeepQ.train (policy.py:241-253).
    ``batches`` = [(x, edge_index, edge_attr, batch)] CPU tensors; each leg runs for about ``budget_s`` seconds (>= 1 pass)."""
    import time
    torch.set_num_threads(max(1, min(threads, 32)))
    torch.manual_seed(0)
    model = gnn_ref.GCN()
    opt = torch.optim.Adam(model.parameters(), lr=1e-5)
    graphs = [gnn_ref.Graph(x, ei, w) for x, ei, w, _ in batches]
    n_graphs = [int(bt.max()) + 1 for _, _, _, bt in batches]

    def forward(i):
        with torch.no_grad():
            model(graphs[i % len(graphs)], 0.0)

    def forward_backward(i):
        opt.zero_grad()
        q = model(graphs[i % len(graphs)], 0.5)
        ((q.view(-1) ** 2).sum() / n_graphs[i % len(graphs)]).backward()
        for p in model.parameters():
            p.grad.clamp_(-0.5, 0.5)
        opt.step()

    out = {"cores": int(torch.get_num_threads()), "kind": "port", "sample": f"{len(batches)} of the C5 batches, ~{budget_s:.0f} s per leg, gnn_ref.GCN (PyG restatement) fp32"}
    for name, fn in (("forward", forward), ("forward_backward", forward_backward)):
        model.eval() if name == "forward" else model.train()
        fn(0)
        t0, n, done = time.perf_counter(), 0, 0
        while True:
            fn(n); done += n_graphs[n % len(graphs)]; n += 1
            if time.perf_counter() - t0 >= budget_s:
                break
        out[name + "_graphs_per_s"] = done / (time.perf_counter() - t0)
    return out
