#!/usr/bin/env python
"""bench.py -- headline benchmark of the exploration hot path on B200.

Workload (BASELINE.json configs[1]): 256 parallel envs per GPU, 20x20 map, 30 landmarks,
GNN (GCN) policy inference only.  One "step" = one tick of the batched engine: envs whose
action queue is empty get a decision (exploration graph -> GCN forward -> arg-max over frontier
nodes -> line plan), then every env executes one simulator step (motion + association scan +
SLAM solve + marginals + virtual-map rebuild); finished episodes are reset.
metric = policy-driven env-steps/sec (reset steps are not counted).

  python bench.py --gpus N --steps K --warmup W        # N > 1: launched under torchrun, one rank per GPU
  python bench.py --impl reference ...                 # the reference path on host cores (CPU oracle + torch-CPU GCN)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv and int(os.environ.get("RANK", "0")) == 0:
    # the reference arm runs on rank 0 alone with all the host threads it can use: torchrun's default OMP_NUM_THREADS=1
    # (set for every worker when the variable is unset) must not throttle its torch-CPU GCN -- before torch / MKL load
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ONE JSON line on stdout: native libraries may print to fd 1 (NCCL's version banner does, under torchrun), so fd 1 is pointed at
# stderr for the life of the process and the line is written to a duplicate of the original stdout
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _JSON_OUT.write(json.dumps(obj) + "\n")
    _JSON_OUT.flush()


MAP_SIZE, N_LANDMARKS, ENVS_PER_GPU, MAX_POSES = 20, 30, 256, 192
WORKLOAD = f"{ENVS_PER_GPU} envs/GPU, {MAP_SIZE}x{MAP_SIZE} map, {N_LANDMARKS} landmarks, GCN policy inference (BASELINE configs[1])"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the simulator kernels in this workload, from the committed
    `ncu --set full` capture of the CURRENT kernels (profiles/traffic.json, written next to the capture's summary; a number measured
    under the profiler is evidence for traffic, never a timing).  No file, or a file for another kernel version: null."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons DURING the timed region: NVML every 10 ms (nvidia-smi every 200 ms if NVML
    is unavailable).  Reported: median SM clock under load, max SM clock, active slowdown reasons."""

    def __init__(self, index=0, period=0.01):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.period = float(os.environ.get("DGE_BENCH_SAMPLER_PERIOD", period))   # seconds between NVML samples
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n, h = self.nvml, self.h
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        flags = []
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            flags.append("Active" if r & bit else "Not Active")
        return [str(sm), str(mx), "0", flags[0], flags[3], flags[2], flags[1]]

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(self.period if self.nvml is not None else max(self.period, 0.2))

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


class L2Flush:
    """Evicts the 126 MB L2 between timed iterations: a 256 MiB write (the recipe's flush) followed by a 256 MiB read of
    a second buffer.  The read matters: after the write alone L2 is full of DIRTY lines, and every store miss of the
    timed kernels would first have to wait for a write-back (measured: an SM then sustains ~5 GB/s of stores), which
    times the flush, not the kernels.  After the read pass L2 holds clean lines of a foreign buffer -- cold for us."""
    HOW = "flushed between timed ticks: 256 MiB write, then 256 MiB read (leaves L2 clean and cold)"

    def __init__(self, dev):
        self.w = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.r = torch.ones(32 << 20, dtype=torch.int64, device=dev)
        self.sink = None

    def __call__(self):
        self.w.fill_(1)
        self.sink = self.r.sum()


# ------------------------------------------------------------------ CPU reference arm ---
REF_PREROLL = 200      # ticks before either CPU leg is timed: the episodes de-synchronise and the mean trajectory length becomes stationary


def timed_reference(**k):
    """The reference path on host cores (oracle/cpu_loop.py): all 256 CPU-oracle envs on a C++ worker pool + ONE torch-CPU GCN forward
    per decision round, overlapped with the stepping of the other envs."""
    from oracle.cpu_loop import timed_reference as impl
    return impl(MAP_SIZE, N_LANDMARKS, ENVS_PER_GPU, os.cpu_count() or 1, MAX_POSES, **k)


def run_reference(args):
    """--impl reference: a step = `tps` ticks of all 256 envs, tps chosen from the warm-up so that the K timed steps last >= 10 s."""
    threads = os.cpu_count() or 1
    from oracle.cpu_loop import cpu_reference
    run_tick, count = cpu_reference(MAP_SIZE, N_LANDMARKS, ENVS_PER_GPU, threads, MAX_POSES)
    for _ in range(int(os.environ.get("DGE_BENCH_REF_PREROLL", REF_PREROLL))):
        run_tick()
    t0 = time.perf_counter()
    for _ in range(args.warmup):
        run_tick()
    tick_s = (time.perf_counter() - t0) / args.warmup
    min_s = float(os.environ.get("DGE_BENCH_REF_SECONDS", "10"))          # (tests shorten the run; the contract's run keeps the default)
    tps = int(max(1, min(50, math.ceil(min_s / (args.steps * tick_s)))))   # >= 10 s timed whatever --steps is (a decision round makes single ticks uneven)
    c0, t0, sT = count(), time.perf_counter(), 0.0
    for _ in range(args.steps * tps):
        run_tick(); sT += run_tick.mean_poses()
    dt = time.perf_counter() - t0
    val = (count() - c0) / dt
    sample = (f"all {ENVS_PER_GPU} envs x {args.steps * tps} ticks ({dt:.1f} s, {tps} ticks per step) after {REF_PREROLL} pre-roll ticks, mean trajectory length "
              f"{sT / (args.steps * tps):.1f}: CPU-oracle envs on a {threads}-thread C++ pool + torch-CPU GCN per decision round (overlapped with the stepping)")
    emit({"impl": "reference", "metric": "env-steps/sec", "value": val, "unit": "env-steps/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "sample": sample, "ticks_per_step": tps, "timed_seconds": dt},
                      "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def finish(world, dist):
    """End of a run.  One rank: normal interpreter exit.  Several ranks: everything is flushed, the ranks meet at a last barrier
    and leave with os._exit(0) -- the NCCL / CUDA teardown (destroy_process_group + context destruction at interpreter
    shutdown) is skipped on purpose: at N = 4 it did not return within 6 minutes AFTER the complete JSON line had been
    written (r01), and nothing is left to release that the process exit does not release."""
    sys.stdout.flush(); sys.stderr.flush()
    if world > 1:
        try:
            dist.barrier()
            torch.cuda.synchronize()
        except Exception:
            pass
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


# ---------------------------------------------------------------- extra workloads ---
def _dist_setup():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def run_train(args):
    rank, world, local, dist = _dist_setup()
    out = measure_train(args, rank, world, local, dist, args.steps, args.warmup)
    if rank == 0:
        emit(out)
    finish(world, dist)


TRAIN_PREROLL = 200


def measure_train(args, rank, world, local, dist, steps, warmup):
    """BASELINE configs[2] (C3): 256 envs per GPU, 40x40 map, DQN training with batched roll-out rewards, device replay and
    one NCCL all-reduce of the flat gradient bucket per gradient step.  A step = one tick of trainer.VecDQNTrainer."""
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.config import EnvConfig
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
    from drl_graph_exploration_b200.trainer import VecDQNTrainer
    ms, B = 40, ENVS_PER_GPU
    cfg = EnvConfig(map_size=ms)
    if args.train_gemm in ("fp32", "tc3"):
        Networks.set_matmul_precision("tc3", train=args.train_gemm)
    env = VecExplorationEnv(B, cfg=cfg, max_poses=384, device=local, seed0=rank * B, seed_stride=world * B)
    env.reset()
    torch.manual_seed(0)                      # identical replicas on every rank
    pol, tgt = Networks.GCN().to(env.device), Networks.GCN().to(env.device)
    tr = VecDQNTrainer(env, pol, tgt, observe=0, train_steps_per_tick=args.train_steps_per_tick, seed=rank,
                       overlap=False if args.no_overlap else (args.train_schedule if args.train_schedule in ("tail", "lag") else True))
    if args.train_gemm != "native":            # A/B: the autograd path with torch's Adam
        tr.optimizer = torch.optim.Adam(pol.parameters(), lr=1e-5)
    # untimed pre-roll: fills the replay (every rank needs one minibatch of transitions) and, like the policy loop's pre-roll,
    # de-synchronises the episodes -- right after the common reset all 256 envs decide in the same ticks (bursts of 256 graphs and
    # ~1000 roll-out clones, then ticks without a decision); ~200 ticks later a tick sees its stationary ~40 decisions
    for _ in range(TRAIN_PREROLL):
        tr.tick(learn=False)
    assert tr.replay.size >= tr.dqn.BATCH, "prefill too short"
    for _ in range(warmup):
        tr.tick(learn=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # (a training tick is ~150 launches issued from the host: NVML queries take driver locks the launches need, so this leg samples
    #  the clocks at 5 Hz instead of the policy loop's 100 Hz -- still several samples inside the ~1 s timed region)
    sampler = ClockSampler(local, period=0.2) if rank == 0 else None
    if sampler:
        sampler.start()
    c0 = [int(v) for v in env.eng.state["counters"].tolist()[:3]]
    d0, t0, r0, k0 = tr.decisions, tr.train_steps, tr.rollout_steps, tr.rollout_clones
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc
    gc.collect(); gc.freeze()          # the objects of the earlier legs of this process leave the collector's generations (the tick is host-bound)
    ms0 = torch.cuda.memory_stats(env.device)
    a.record()
    t_host = time.perf_counter()
    for _ in range(steps):
        tr.tick(learn=True)
    t_host = time.perf_counter() - t_host
    b.record()
    ms1 = torch.cuda.memory_stats(env.device)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = a.elapsed_time(b)
    c1 = [int(v) for v in env.eng.state["counters"].tolist()[:3]]
    t = torch.tensor([ms_total], dtype=torch.float64, device=env.device)
    v = torch.tensor([c1[0] - c0[0], tr.decisions - d0, tr.train_steps - t0, tr.rollout_steps - r0, tr.rollout_clones - k0], dtype=torch.float64, device=env.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(v, op=dist.ReduceOp.SUM)
    # the one collective of the path, timed alone: the 4.0 MB flat gradient bucket, CUDA events, max over ranks
    ar_us = None
    if world > 1:
        flat = (tr.dqn._bucket or tr.optimizer.bucket).flat
        for _ in range(5):
            dist.all_reduce(flat)
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dist.all_reduce(flat)
        b.record(); torch.cuda.synchronize()
        tt = torch.tensor([a.elapsed_time(b) / 20 * 1e3], dtype=torch.float64, device=env.device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ar_us = float(tt.item())
    out = None
    if rank == 0:
        sec = float(t.item()) / 1e3
        env_steps, dec, tsteps, rsteps, clones = (float(x) for x in v.tolist())
        bsz = tr.dqn.BATCH
        out = {"metric": "env-steps/sec (DQN training)", "value": env_steps / sec, "unit": "env-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 simulator / f32 GNN",
               "data": "synthetic",
               "config": {"workload": f"{B} envs/GPU, {ms}x{ms} map, {cfg.n_landmarks} landmarks, DQN+GCN training (BASELINE configs[2])", "batch_graphs_per_rank": bsz,
                          "train_steps_per_tick": args.train_steps_per_tick, "train_gemm": args.train_gemm, "preroll_ticks": TRAIN_PREROLL,
                          "schedule": ("sequential" if args.no_overlap else "gradient step issued at the end of its tick on a second stream: runs beside the next tick's step / roll-out kernels (sequential order of operations)"
                                       if args.train_schedule == "tail" else "gradient step on a second stream beside the roll-out kernels; its update is applied behind the tick's acting forward (the acting policy is one update staler; the all-reduce has a whole tick before anything waits for it)"
                                       if args.train_schedule == "lag" else "gradient step on a second stream beside the roll-out kernels of the same tick"), "replay": f"device ring, {tr.replay.capacity} transitions, {tr.replay.nbytes() / 2**30:.2f} GiB",
                          "collective": "one all-reduce of the 4.0 MB flat gradient bucket per gradient step" if world > 1 else "none (1 rank)"},
               "decisions_per_s": dec / sec, "train_steps_per_s": tsteps / sec / world, "rollout_clone_steps_per_s": None,
               "gnn_samples_per_s": {"forward_acting": dec / sec, "forward_target": tsteps * bsz / sec, "forward_backward": tsteps * bsz / sec},
               "rollout": {"clones_per_s": clones / sec, "clone_engine_ticks_per_s": rsteps / sec}, "loss": tr.last_loss, "epsilon": tr.epsilon, "clocks": clocks,
               "allreduce_us": ar_us, "host_issue_ms_per_tick": 1e3 * t_host / steps,
               "allocator": {k: ms1.get(k, 0) - ms0.get(k, 0) for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated")}}
    env.close()
    return out


def synth_graph_batch(n_graphs, sizes, rng, device, n_landmarks=8):
    """C5 topology: per graph a pose chain, every pose observing Poisson(1.5) of the L landmarks, F <= L+1
    frontier stubs; x ~ N(0,1), edge_attr ~ U(0.1, 6); both edge directions, PyG DataLoader layout."""
    src, dst, wts, bat, off = [], [], [], [], 0
    for g in range(n_graphs):
        n = int(sizes[g])
        L = min(n_landmarks, max(1, n // 4)); F = min(L + 1, max(1, n // 8)); T = max(1, n - L - F)
        L = n - T - F
        e = [(L + k, L + k + 1) for k in range(T - 1)]
        for k in range(T):
            if L > 0:
                for l in rng.choice(L, size=min(L, int(rng.poisson(1.5))), replace=False):
                    e.append((int(l), L + k))
        for f in range(F):
            e.append((int(rng.integers(0, L + T)), L + T + f))
        e = np.array(sorted(set(e)), dtype=np.int64).reshape(-1, 2)
        s_, d_ = np.concatenate([e[:, 0], e[:, 1]]), np.concatenate([e[:, 1], e[:, 0]])
        wh = rng.uniform(0.1, 6.0, e.shape[0])
        src.append(s_ + off); dst.append(d_ + off); wts.append(np.concatenate([wh, wh])); bat.append(np.full(n, g)); off += n
    ei = torch.tensor(np.stack([np.concatenate(src), np.concatenate(dst)]), dtype=torch.long, device=device)
    w = torch.tensor(np.concatenate(wts), dtype=torch.float32, device=device)
    x = torch.tensor(rng.normal(size=(off, 5)), dtype=torch.float32, device=device)
    return x, ei, w, torch.tensor(np.concatenate(bat), dtype=torch.long, device=device)


def run_gnn(args):
    rank, world, local, dist = _dist_setup()
    out = measure_gnn(args, rank, world, local, dist)
    if rank == 0:
        emit(out)
    finish(world, dist)


def measure_gnn(args, rank, world, local, dist, steps=None):
    """BASELINE configs[4] (C5): batches of 64 graphs with 8..512 nodes (mixed), GCN forward and forward+backward+Adam."""
    from drl_graph_exploration_b200 import Networks, gnn
    from drl_graph_exploration_b200.data import Data
    n_steps = steps or args.steps
    dev = torch.device("cuda", local)
    rng = np.random.default_rng(1234 + rank)
    nb, G = 8, 64
    batches = []
    for _ in range(nb):
        sizes = rng.choice(np.arange(8, 513, 8), size=G)
        batches.append(synth_graph_batch(G, sizes, rng, dev))
    torch.manual_seed(0)
    model = Networks.GCN().to(dev)
    from drl_graph_exploration_b200.dist import FlatGradBucket, NativeAdam
    native = args.train_gemm == "native"
    if native:
        opt = NativeAdam(model.parameters(), lr=1e-5)
        bucket = opt.bucket
    else:
        Networks.set_matmul_precision("tc3", train=args.train_gemm)
        opt = torch.optim.Adam(model.parameters(), lr=1e-5)
        bucket = FlatGradBucket(model.parameters())
    flush = None if args.no_flush_l2 else L2Flush(dev)
    nodes = sum(b[0].size(0) for b in batches) / nb
    edges = sum(b[1].size(1) for b in batches) / nb

    def fwd(i):
        x, ei, w, bt = batches[i % nb]
        with torch.no_grad():
            return model(Data(x, ei, w, bt), 0.0)

    def fwd_bwd(i):
        x, ei, w, bt = batches[i % nb]
        if native:   # DeepQ.train on the native path: forward (dropout 0.5) + cost + backward in one call, all-reduce, clamp + Adam in one kernel
            d = Data(x, ei, w, bt)
            gnn.gcn_train_step(model, x, Networks._structure(d, x.size(0)), None, None, 1.0 / G, 0.5)
            if world > 1:
                dist.all_reduce(bucket.flat)
            opt.step(clamp=0.5, gscale=1.0 / world)
            return
        bucket.zero_()
        q = model(Data(x, ei, w, bt), 0.5, batch=bt)
        loss = (q.view(-1) ** 2).sum() / G
        loss.backward()
        bucket.all_reduce_mean(); bucket.clamp_(0.5)
        opt.step()

    res = {}
    for name, fn in (("forward", fwd), ("forward_backward", fwd_bwd)):
        model.eval() if name == "forward" else model.train()
        for i in range(max(args.warmup, nb)):      # every batch at least once: the caching allocator has seen all sizes before the timed steps
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evs, l0 = [], gnn.launch_count
        for i in range(n_steps):
            if flush is not None:
                flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(i); b.record(); evs.append((a, b))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = (float(t.item()) / 1e3, gnn.launch_count - l0)
    # ---- the other two families of Networks.py on the same batches ("GCN (primary), GG-NN, g-U-Net; forward and
    # forward+backward"): module forward at inference, and forward (dropout 0.5) + backward + gradient all-reduce + clamp + Adam under
    # autograd with the dense products, the GRU cell and the aggregations on this package's kernels (Networks.set_matmul_precision 'tc3')
    fam = {}
    if not getattr(args, "no_families", False):
        from drl_graph_exploration_b200.dist import FlatGradBucket as _Bucket
        if args.train_gemm in ("fp32", "tc3"):
            Networks.set_matmul_precision("tc3", train=args.train_gemm)
        for fname, make in (("GGNN", lambda: Networks.GGNN()), ("GraphUNet", lambda: Networks.GraphUNet(5, 1000, 1000, depth=3))):
            torch.manual_seed(0)
            net = make().to(dev)
            f_opt = torch.optim.Adam(net.parameters(), lr=1e-5)
            f_bucket = _Bucket(net.parameters())

            def f_fwd(i, net=net):
                x, ei, w, bt = batches[i % nb]
                with torch.no_grad():
                    return net(Data(x, ei, w, bt), 0.0, batch=bt)

            def f_fwd_bwd(i, net=net, f_opt=f_opt, f_bucket=f_bucket):
                x, ei, w, bt = batches[i % nb]
                f_bucket.zero_()
                q = net(Data(x, ei, w, bt), 0.5, batch=bt)
                ((q.view(-1) ** 2).sum() / G).backward()
                f_bucket.all_reduce_mean(); f_bucket.clamp_(0.5)
                f_opt.step()

            k = max(3, min(n_steps, 10))
            row = {}
            for name, fn in (("forward", f_fwd), ("forward_backward", f_fwd_bwd)):
                net.eval() if name == "forward" else net.train()
                for i in range(nb):           # every batch once: the allocator has seen all sizes, the weight operands are split
                    fn(i)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                evs, l0 = [], gnn.launch_count
                for i in range(k):
                    if flush is not None:
                        flush()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(i); b.record(); evs.append((a, b))
                torch.cuda.synchronize()
                t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t.item()) / 1e3
                row[name + "_graphs_per_s"] = world * G * k / sec
                row[name + "_ms_per_batch"] = 1e3 * sec / k
                row[name + "_launches_per_batch"] = (gnn.launch_count - l0) / k
            row["steps"] = k
            fam[fname] = row
            del net, f_opt, f_bucket
        torch.cuda.empty_cache()
    if rank == 0:
        pk, pk_kind = peaks()
        sec_f, sec_b = res["forward"][0], res["forward_backward"][0]
        gemm_flops = 2.0 * nodes * 1000 * 1000                     # the one [N,1000]x[1000,1000] product of a forward pass
        out = {"metric": "GNN samples/sec", "value": world * G * n_steps / sec_f, "unit": "graphs/s", "n_gpus": world, "steps": n_steps, "warmup": args.warmup,
               "ms_per_step": 1e3 * sec_f / n_steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (3xTF32 tensor-core GEMM, fp32 accumulate)", "train_path": "native (gnn.gcn_train_step + dist.NativeAdam)" if native else "autograd",
               "data": "synthetic",
               "config": {"workload": "64 graphs/batch, 8..512 nodes mixed, GCN (BASELINE configs[4])", "mean_nodes_per_batch": nodes, "mean_edges_per_batch": edges,
                          "train_gemm": args.train_gemm, "l2": L2Flush.HOW if flush is not None else "not flushed"},
               "forward": {"graphs_per_s": world * G * n_steps / sec_f, "nodes_per_s": world * nodes * n_steps / sec_f, "ms_per_batch": 1e3 * sec_f / n_steps,
                           "fp32_equiv_tflops_whole_pass": (2.012e6 * nodes * n_steps / sec_f) / 1e12},
               "forward_backward": {"graphs_per_s": world * G * n_steps / sec_b, "nodes_per_s": world * nodes * n_steps / sec_b, "ms_per_batch": 1e3 * sec_b / n_steps,
                                    "fp32_equiv_tflops_whole_pass": (3 * 2.012e6 * nodes * n_steps / sec_b) / 1e12, "includes": "backward, gradient all-reduce, clamp, Adam"},
               "gpu_launches": res["forward"][1] + res["forward_backward"][1],
               "roofline": {"bound": "tensor", "kernel": "k_gemm_tf32x3 inside the forward pass", "achieved": gemm_flops * n_steps / sec_f / 1e12, "peak": pk.get("bf16_tflops_sustained", pk.get("bf16_tflops")),
                            "peak_kind": pk_kind + " (sustained dense bf16: the GEMM is timed inside the whole pass)", "unit": "TFLOP/s",
                            "frac": gemm_flops * n_steps / sec_f / 1e12 / pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1.0)),
                            "note": "whole forward pass time in the denominator (GEMM + 2 aggregations + head), fp32-equivalent flops; the kernel alone: profiles/r01_k_gemm_tf32x3_ncu.md",
                            "traffic": None}}
        out["families"] = fam
        # the same two legs on the host cores (the PyG restatement of the oracle, bounded sample): the "vs CPU ref" half of the metric
        out["cpu_baseline"] = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle.cpu_loop import cpu_gnn_baseline
                out["cpu_baseline"] = cpu_gnn_baseline([tuple(t.cpu() for t in b) for b in batches[:2]], os.cpu_count() or 1, budget_s=5.0)
            except Exception as exc:      # noqa: BLE001 -- a reported baseline must never take the measured line down
                out["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"}
        return out
    return None


def measure_c4_sweep(dev, n=1024, Ts=(32, 64, 128, 256, 512, 1024), reps=10):
    """BASELINE configs[3] (C4): the covariance-propagation kernel (k_vmap_env: occupancy + ordered covariance-intersection fold) alone,
    1024 envs, 60x60 map (V = 2500 cells), 200 landmarks, T synthetic poses per env (drl_graph_exploration_b200/synth.py), through
    the C ABI (dge_virtual_map_rebuild) on pre-allocated buffers.  Algorithmic bytes per env-rebuild (SURVEY 8(d), fp64 state):
    2 (48 T + 20 V + 8 L).  L2 flushed before every timed launch; median of `reps` launches, CUDA events on the launching stream."""
    import ctypes
    from drl_graph_exploration_b200.config import EnvConfig
    from drl_graph_exploration_b200.engine import _ptr, _stream_ptr, load_library
    from drl_graph_exploration_b200.synth import synth_states
    pk, pk_kind = peaks()
    cfg = EnvConfig(map_size=60, num_landmarks=200)
    cs = cfg.to_struct()
    L, V = 200, cfg.rows * cfg.cols
    lib = load_library()
    flush = L2Flush(dev)
    rows = []
    for T in Ts:
        pose, _, cov6, _, lm = synth_states(cfg, n, T, L, seed=T)
        tp, tc, tl = (torch.as_tensor(a, device=dev).contiguous() for a in (pose, cov6, lm))
        prob = torch.empty(n, V, dtype=torch.float64, device=dev); vinfo = torch.empty(n, V, 3, dtype=torch.float64, device=dev)
        seen = torch.empty(n, V, dtype=torch.int32, device=dev)
        ws = torch.zeros(lib.dge_virtual_map_rebuild_ws_doubles(n, T), dtype=torch.float64, device=dev)
        call = lambda s: lib.dge_virtual_map_rebuild(ctypes.byref(cs), n, T, _ptr(tp), _ptr(tc), L, _ptr(tl), _ptr(prob), _ptr(vinfo), _ptr(s), _ptr(ws), _stream_ptr(dev))
        for _ in range(3):
            assert call(seen) == 0
        torch.cuda.synchronize()
        visits = float(seen.clamp(min=0).sum())
        ts = []
        for _ in range(reps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); call(None); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        by = 2.0 * (48 * T + 20 * V + 8 * L) * n
        gbs = by / (ms * 1e-3) / 1e9
        rows.append({"T": T, "us": 1e3 * ms, "algorithmic_MB": by / 1e6, "GBps": gbs, "frac": gbs / pk["hbm_gbs"], "pair_visits": visits,
                     "Gvisits_per_s": visits / ms / 1e6})
    return {"workload": f"{n} envs, 60x60 map (V = {V}), {L} landmarks, k_vmap_env alone (BASELINE configs[3])", "peak_GBps": pk["hbm_gbs"], "peak_kind": pk_kind,
            "bytes_per_env": "2 (48 T + 20 V + 8 L)", "l2": L2Flush.HOW, "rows": rows}


# ------------------------------------------------------------------------- GPU arm ---
class GpuLoop:
    def __init__(self, device, seed0, overlap=True, device_tick=True, seed_stride=None):
        from drl_graph_exploration_b200 import Networks, gnn
        from drl_graph_exploration_b200.config import EnvConfig
        from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv

        self.gnn = gnn
        cfg = EnvConfig(map_size=MAP_SIZE, num_landmarks=N_LANDMARKS)
        self.env = VecExplorationEnv(ENVS_PER_GPU, cfg=cfg, max_poses=MAX_POSES, device=device, seed0=seed0, seed_stride=seed_stride)
        torch.manual_seed(0)
        self.model = Networks.GCN().to(self.env.device).eval()
        self.env.reset()
        self.dev = self.env.device
        from drl_graph_exploration_b200.runner import PolicyLoop
        self.runner = PolicyLoop(self.env, self.model, overlap=overlap, device_tick=None if device_tick else False)
        self.ev = {k: [] for k in ("slam", "vmap")}

    def counters(self):
        """(policy env-steps, sum of trajectory lengths, sum of measurement counts) accumulated by the engine."""
        return [int(v) for v in self.env.eng.state["counters"].tolist()[:3]]

    def tick(self, timed=False):
        """One tick of drl_graph_exploration_b200.runner.PolicyLoop (step pipeline || policy pipeline)."""
        self.runner.stage_events = self.ev if timed else None
        self.runner.tick()


def e2e_loop(loop: GpuLoop, overlap=True):
    """The same tick through the reference-facing host-buffer API (drl_graph_exploration_b200.runner.HostPolicyLoop):
    actions H2D from pinned memory, done flags + status metrics + occupancy maps D2H every step; graphs D2H -> host ->
    H2D for the GNN, Q D2H, arg-max on the host, goals H2D, line plans D2H."""
    from drl_graph_exploration_b200.runner import HostPolicyLoop
    return HostPolicyLoop(loop.env, loop.model, overlap=overlap)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-flush-l2", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-python", action="store_true", help="e2e leg: issue the host tick's C-ABI calls from Python instead of dge_host_policy_tick (A/B)")
    ap.add_argument("--no-gnn", action="store_true", help="skip the C5 GNN samples/sec measurement appended to the default line")
    ap.add_argument("--no-families", action="store_true", help="GNN leg: skip the GG-NN / g-U-Net rows (C5's other two families)")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 covariance-propagation roofline sweep appended to the default line (N = 1)")
    ap.add_argument("--no-train", action="store_true", help="skip the C3 DQN-training measurement appended to the default line")
    ap.add_argument("--no-overlap", action="store_true", help="run the step and policy pipelines on one stream (A/B of the overlap)")
    ap.add_argument("--per-launch", action="store_true", help="policy loop: the per-launch schedule with one size sync per tick instead of dge_policy_tick (A/B)")
    ap.add_argument("--ticks-per-step", type=int, default=50, help="policy loop: ticks per bench step (every tick is timed separately)")
    ap.add_argument("--preroll", type=int, default=600, help="untimed ticks before warm-up (de-synchronises the episodes; independent of --warmup)")
    ap.add_argument("--workload", default="policy", choices=["policy", "train", "gnn"],
                    help="policy = BASELINE configs[1] (the headline line); train = configs[2] DQN training; gnn = configs[4] GNN fwd / fwd+bwd")
    ap.add_argument("--train-steps-per-tick", type=int, default=1)
    ap.add_argument("--train-schedule", default="beside", choices=["tail", "beside", "lag"],
                    help="C3: where the gradient step of a tick is issued (trainer.VecDQNTrainer overlap=True / 'tail').  'tail' measured slower (8.8-10.3 vs 7.4 ms per tick at N = 1): "
                         "the two size syncs of the minibatch gather then wait for the whole tick's backlog on the device")
    ap.add_argument("--train-gemm", default="native", choices=["native", "fp32", "tc3"],
                    help="training step of the GCN: native = gnn.gcn_train_step + dist.NativeAdam (hand-written kernels, tcgen05 3xTF32 GEMMs; the product path); "
                         "fp32 / tc3 = autograd with the library fp32 GEMM / the tcgen05 GEMM for forward and grad-input (A/B)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return

    if args.workload == "train":
        return run_train(args)
    if args.workload == "gnn":
        return run_gnn(args)
    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = local
    torch.cuda.set_device(device)
    # env-sharded worlds: rank r owns seeds r B .. r B + B - 1 and walks them with stride W B (disjoint from every other rank's for ever)
    loop = GpuLoop(device, seed0=rank * ENVS_PER_GPU, overlap=not args.no_overlap, device_tick=not args.per_launch, seed_stride=world * ENVS_PER_GPU)
    flush = None if args.no_flush_l2 else L2Flush(loop.dev)

    # untimed pre-roll, independent of --warmup: all envs were reset together, so the first ticks see T = 5..30 only.  Run until
    # the episodes have de-synchronised (several episode lengths) and the mean trajectory length is stationary.
    n_poses = loop.env.eng.state["n_poses"]
    hist = []
    for i in range(args.preroll):
        loop.tick()
        if (i + 1) % 100 == 0:
            hist.append(float(n_poses.float().mean().item()))
    for _ in range(args.warmup):
        loop.tick()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(device) if rank == 0 else None
    if sampler:
        sampler.start()
    loop.runner.launches = 0; loop.runner.graphs = 0
    c_start = loop.counters()
    tick_events = []
    t_host = 0.0
    n_ticks = args.steps * args.ticks_per_step            # a step = ticks_per_step ticks, every tick timed on its own
    for _ in range(n_ticks):
        if flush is not None:
            flush()                  # L2 flush (untimed) before every timed tick
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); th = time.perf_counter(); loop.tick(); t_host += time.perf_counter() - th; b.record()
        tick_events.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    tick_ms = np.array([a.elapsed_time(b) for a, b in tick_events])
    total_ms = float(tick_ms.sum())
    c_end = loop.counters()
    steps_rank, sumT, sumM = (c_end[i] - c_start[i] for i in range(3))
    graphs_rank, launches_rank = loop.runner.graphs, loop.runner.launches
    # per-kernel times of the two simulator kernels: CUDA events on the step stream around each launch, on the ticks that follow the
    # timed region (same steady state; a replayed CUDA graph has no place for events, so these ticks use the per-launch schedule)
    cs0 = loop.counters()
    n_stage = min(max(n_ticks, 20), 100)
    for _ in range(n_stage):
        if flush is not None:
            flush()
        loop.tick(timed=True)
    torch.cuda.synchronize()
    cs1 = loop.counters()
    t = torch.tensor([total_ms], dtype=torch.float64, device=loop.dev)
    s = torch.tensor([steps_rank, graphs_rank], dtype=torch.float64, device=loop.dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(s, op=dist.ReduceOp.SUM)
    total_ms = float(t.item()); steps_all, graphs_all = float(s[0].item()), float(s[1].item())

    out = None
    if rank == 0:
        pk, pk_kind = peaks()
        value = steps_all / (total_ms / 1e3)
        # roofline of the dominant simulator kernels (algorithmic fp64 bytes per SURVEY 8(d) x 2 for fp64)
        T_mean = sumT / max(steps_rank, 1); M_mean = sumM / max(steps_rank, 1)
        st_steps = max(cs1[0] - cs0[0], 1)
        T_stage, M_stage = (cs1[1] - cs0[1]) / st_steps, (cs1[2] - cs0[2]) / st_steps
        eng = loop.env.eng
        ms_slam = sum(a.elapsed_time(b) for a, b in loop.ev["slam"]) / max(len(loop.ev["slam"]), 1)
        ms_vmap = sum(a.elapsed_time(b) for a, b in loop.ev["vmap"]) / max(len(loop.ev["vmap"]), 1)
        envs_per_launch = st_steps / n_stage
        bytes_vmap = 2 * (48 * T_stage + 20 * eng.V + 8 * eng.Lt) * envs_per_launch
        bytes_slam = 2 * (72 * T_stage + 16 * M_stage + 32 * eng.Lt) * envs_per_launch
        # the honest bound of these two kernels is the fp64 pipe, not HBM: algorithmic fp64 flops of the SLAM
        # solve per env-step = Schur complement 3 T n^2 + marginals 3 T n^2 + inverse n^3 FMAs (n = 2 x observed landmarks; the
        # O(T n) chain phases are left out), against the DFMA rate measured on this GPU (profiles/r01_fp64_latency_b200.txt)
        n2 = 2.0 * float(loop.env.eng.state["observed"].sum(dim=1).float().mean().item())
        flops_slam = 2.0 * (6.0 * T_stage * n2 * n2 + n2 ** 3) * envs_per_launch
        fp64_peak = 35.2e12
        fp64 = {"kernel": "k_slam", "achieved_tflops": flops_slam / (ms_slam * 1e-3) / 1e12, "peak_tflops": fp64_peak / 1e12,
                "frac": flops_slam / (ms_slam * 1e-3) / fp64_peak, "peak_source": "profiles/r01_fp64_latency_b200.txt (17.6 T DFMA/s measured)",
                "mean_border_columns": n2}
        dom = "slam" if ms_slam >= ms_vmap else "vmap"
        ach = (bytes_slam / (ms_slam * 1e-3) if dom == "slam" else bytes_vmap / (ms_vmap * 1e-3)) / 1e9
        traffic = ncu_traffic().get(dom, {})
        roof = {"bound": "hbm", "kernel": "k_slam" if dom == "slam" else "k_vmap_env", "achieved": ach,
                "peak": pk["hbm_gbs"], "peak_kind": pk_kind, "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": traffic.get("bytes"),
                "traffic_source": traffic.get("source"),
                "sample": f"{n_stage} ticks of the per-launch schedule right after the timed region, CUDA events on the step stream around each launch",
                "ms_per_launch": {"slam": ms_slam, "vmap": ms_vmap}, "mean_poses": T_stage, "mean_measurements": M_stage,
                "vmap": {"achieved": bytes_vmap / (ms_vmap * 1e-3) / 1e9, "frac": bytes_vmap / (ms_vmap * 1e-3) / 1e9 / pk["hbm_gbs"]},
                "slam": {"achieved": bytes_slam / (ms_slam * 1e-3) / 1e9, "frac": bytes_slam / (ms_slam * 1e-3) / 1e9 / pk["hbm_gbs"]},
                "fp64": fp64}
        out = {"metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": {"workload": WORKLOAD, "envs_per_gpu": ENVS_PER_GPU, "map_size": MAP_SIZE, "landmarks": N_LANDMARKS,
                                               "policy": "GCN fp32, random init", "parallelism": f"env-sharded x{world}, no data-path collective",
                                               "step": f"{args.ticks_per_step} ticks of the acting loop, each timed by its own pair of CUDA events (L2 flushed before every tick)",
                                               "tick": ("dge_policy_tick: one native call, replayed as a CUDA graph, no host sync" if not args.per_launch
                                                        else "per-launch schedule, one size sync per tick"),
                                               "l2": L2Flush.HOW if flush is not None else "not flushed"},
               "gnn_graphs_per_s": graphs_all / (total_ms / 1e3), "gpu_launches": launches_rank, "clocks": clocks, "roofline": roof,
               "tick_ms": {"median": float(np.median(tick_ms)), "p10": float(np.percentile(tick_ms, 10)), "p90": float(np.percentile(tick_ms, 90)),
                           "host_issue_ms": 1e3 * t_host / n_ticks, "ticks": n_ticks},
               "mean_poses": T_mean, "envs_stepping_per_tick": steps_rank / n_ticks,
               "preroll": {"ticks": args.preroll, "mean_poses_every_100_ticks": hist}}
    # e2e: every rank drives its own envs through the host-buffer API; whole-job steps / max wall time over ranks
    if not args.no_e2e:
        loop.env.reset()                      # fresh episodes: the host loop owns the action lists from here on
        hl = e2e_loop(loop, overlap=not args.no_overlap)
        if args.e2e_python:
            hl.native = False
        n_e2e = max(20, n_ticks)
        for _ in range(max(args.warmup, 12)):   # past the first decision and a few restarts
            hl.tick()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        s0, h0, d0, t0 = hl.steps, hl.h2d, hl.d2h, time.perf_counter()
        for _ in range(n_e2e):
            hl.tick()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        es = torch.tensor([hl.steps - s0, hl.h2d - h0, hl.d2h - d0], dtype=torch.float64, device=loop.dev)
        et = torch.tensor([dt], dtype=torch.float64, device=loop.dev)
        if world > 1:
            dist.all_reduce(es, op=dist.ReduceOp.SUM); dist.all_reduce(et, op=dist.ReduceOp.MAX)
        if rank == 0:
            dt = float(et.item())
            out["e2e"] = {"value": float(es[0].item()) / dt, "unit": "env-steps/s", "h2d_bytes_per_step": float(es[1].item()) / n_e2e,
                          "d2h_bytes_per_step": float(es[2].item()) / n_e2e, "ticks": n_e2e, "ms_per_tick": 1e3 * dt / n_e2e,
                          "api": ("HostPolicyLoop.tick = dge_host_policy_tick (one native call per tick: " if hl.native is not False else "HostPolicyLoop.tick issued from Python (A/B: ") +
                                 "dge_step_host_plans_async + dge_graph_host_packed_begin/end + H2D + dge_gcn_q_forward + Q D2H + dge_select_plan_host; pinned host buffers), wall clock, max over ranks"}
    elif rank == 0:
        out["e2e"] = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        r = timed_reference(preroll=REF_PREROLL, warmup=3, min_ticks=10, min_seconds=12.0, max_seconds=30.0, seed0=777)
        out["cpu_baseline"] = {"value": r["value"], "unit": "env-steps/s", "cores": threads, "kind": "port",
                               "sample": f"all {ENVS_PER_GPU} envs x {r['ticks']} ticks ({r['seconds']:.1f} s) after {REF_PREROLL} pre-roll ticks, mean trajectory length "
                                         f"{r['mean_poses']:.1f}: CPU-oracle envs on a {threads}-thread C++ pool + torch-CPU GCN per decision round (overlapped with the stepping)"}
    elif rank == 0:
        out["cpu_baseline"] = None
    if not args.no_gnn:
        # the second half of BASELINE's metric ("+ GNN samples/sec"): C5 batches through the GCN, forward and forward+backward
        loop.env.close()
        g = measure_gnn(args, rank, world, local, dist, steps=20)
        if rank == 0:
            out["gnn_c5"] = {"workload": g["config"]["workload"], "mean_nodes_per_batch": g["config"]["mean_nodes_per_batch"],
                             "forward_graphs_per_s": g["forward"]["graphs_per_s"], "forward_ms_per_batch": g["forward"]["ms_per_batch"],
                             "forward_backward_graphs_per_s": g["forward_backward"]["graphs_per_s"],
                             "forward_backward_ms_per_batch": g["forward_backward"]["ms_per_batch"], "train_gemm": args.train_gemm,
                             "tensor_roofline": g["roofline"], "cpu_baseline": g.get("cpu_baseline"), "families": g.get("families")}
    if not args.no_c4 and world == 1:
        try:
            loop.env.close()
        except Exception:
            pass
        out["c4_sweep"] = measure_c4_sweep(loop.dev)
    if not args.no_train:
        # BASELINE configs[2]: DQN training (the path's one collective, the gradient all-reduce, is in this leg when N > 1)
        try:
            loop.env.close()
        except Exception:
            pass
        tr = measure_train(args, rank, world, local, dist, steps=100, warmup=10)
        if rank == 0:
            out["train_c3"] = {k: tr[k] for k in ("value", "unit", "ms_per_step", "host_issue_ms_per_tick", "decisions_per_s", "train_steps_per_s", "allreduce_us", "loss", "allocator")}
            out["train_c3"].update(workload=tr["config"]["workload"], schedule=tr["config"]["schedule"], collective=tr["config"]["collective"], ticks=100)
    if rank == 0:
        emit(out)
    finish(world, dist)


if __name__ == "__main__":
    main()
