"""dev: host-side time of one HostPolicyLoop tick by section."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
loop = bench.GpuLoop(0, 0)
hl = bench.e2e_loop(loop, overlap="--no-overlap" not in sys.argv)
for _ in range(40): hl.tick()
torch.cuda.synchronize()
hl.timing = {}
N = 300; s0 = hl.steps; t0 = time.perf_counter()
for _ in range(N): hl.tick()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"tick {1e6 * dt / N:.1f} us, {(hl.steps - s0) / dt:.0f} env-steps/s, graphs/tick {hl.graphs / (N + 40):.1f}")
for k, v in hl.timing.items(): print(f"  {k:40s} {1e6 * v / N:8.1f} us/tick")
