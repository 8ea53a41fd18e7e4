"""dev: host-side time of one PolicyLoop tick, split by section (perf_counter), next to the GPU tick time."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from drl_graph_exploration_b200 import runner

loop = bench.GpuLoop(0, 0)
for _ in range(60):
    loop.tick()
torch.cuda.synchronize()
# monkey-patch timing into the pieces
import ctypes
env = loop.env; r = loop.runner
acc = {}
def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter(); out = fn(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return out
    return w
env.mark_pending = timed("mark_pending", env.mark_pending)
env.build_graph = timed("build_graph", env.build_graph)
env.graph.sync_sizes = timed("sync_sizes(wait)", env.graph.sync_sizes)
env.select_and_plan = timed("select_and_plan", env.select_and_plan)
r.model = timed("model", r.model)
L = env.eng._L
class LW:
    def __init__(self, L): self._L = L
    def __getattr__(self, n):
        f = getattr(self._L, n)
        return timed("L." + n, f)
env.eng._L = LW(L)
N = 200
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(N):
    r.tick()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"wall per tick {1e6 * (t2 - t0) / N:.1f} us (host loop {1e6 * (t1 - t0) / N:.1f} us)")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"  {k:28s} {1e6 * v / N:8.1f} us/tick")
