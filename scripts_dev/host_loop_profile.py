"""dev: host-side time of one HostPolicyLoop tick by section (Python-issued route), after a pre-roll that de-synchronises the episodes;
   `--no-overlap`: step and policy pipelines on one stream; `--policy-only`: the step pipeline's kernels are skipped for the timing of the chain."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
loop = bench.GpuLoop(0, 0)
hl = bench.e2e_loop(loop, overlap="--no-overlap" not in sys.argv)
for _ in range(600): hl.tick()
torch.cuda.synchronize()
N = 500
for mode in ("native", "python"):
    hl.native = None if mode == "native" else False
    hl.timing = None
    for _ in range(20): hl.tick()
    torch.cuda.synchronize()
    s0 = hl.steps; t0 = time.perf_counter()
    for _ in range(N): hl.tick()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{mode}: tick {1e6 * dt / N:.1f} us, {(hl.steps - s0) / dt:.0f} env-steps/s")
hl.native = False
hl.timing = {}
g0 = hl.graphs; s0 = hl.steps; t0 = time.perf_counter()
for _ in range(N): hl.tick()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"python + section timers: tick {1e6 * dt / N:.1f} us, {(hl.steps - s0) / dt:.0f} env-steps/s, graphs/tick {(hl.graphs - g0) / N:.1f}")
for k, v in hl.timing.items(): print(f"  {k:40s} {1e6 * v / N:8.1f} us/tick")
# device time of the policy-side kernels alone (no step pipeline running beside them): graph build + pack, GCN forward
env = loop.env
import ctypes
from drl_graph_exploration_b200.runner import packed_graph_args
L, pk = hl._L, hl._pk
mp = ctypes.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
hl.need[:] = 0; hl.need[:30] = 1
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tg = tq = 0.0
for it in range(50):
    ev[0].record()
    L.dge_graph_host_packed_begin(env.eng._h, hl.t_need.data_ptr(), ctypes.byref(env.graph.c), hl.a_pack.data_ptr(), hl.arena_cap, mp)
    ev[1].record()
    L.dge_graph_host_packed_end(env.eng._h, hl.a_pack.data_ptr(), hl.a_host.data_ptr(), hl.arena_cap, ctypes.byref(pk), mp)
    hl.a_dev[:pk.total_bytes].copy_(hl.a_host[:pk.total_bytes], non_blocking=True)
    ev[2].record()
    q = hl._plan(pk.n_nodes, packed_graph_args(hl.a_dev, pk))
    ev[3].record(); torch.cuda.synchronize()
    if it >= 10:
        tg += ev[0].elapsed_time(ev[1]); tq += ev[2].elapsed_time(ev[3])
print(f"alone on the GPU: graph build + pack of 30 envs {1e3 * tg / 40:.1f} us, GCN forward on {pk.n_nodes} nodes {1e3 * tq / 40:.1f} us")
