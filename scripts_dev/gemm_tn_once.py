"""dev: the weight-gradient product of a GG-NN GRU transform on a C5 batch -- dW [3000,1000] = dgi^T m over K = 16 896 nodes -- through the MN-major
GEMM (operands as stored) and through the route over transposed copies (two split-transposes + K-major GEMM): CUDA-event times (L2 flushed), and
one profiled launch of the MN-major kernel for ncu (`--profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from drl_graph_exploration_b200 import gnn
dev = torch.device("cuda", 0)
flush = bench.L2Flush(dev)
K, Ca, Cb = 16896, 3000, 1000
g = torch.Generator(device="cpu").manual_seed(0)
a = torch.randn(K, Ca, generator=g).to(dev); b = torch.randn(K, Cb, generator=g).to(dev)
def timed(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2] * 1e3
sa, sb = gnn.split_tf32(a), gnn.split_tf32(b)
t_tn = timed(lambda: gnn.tc_gemm_tn(sa, sb))
t_split = timed(lambda: (gnn.split_tf32(a), gnn.split_tf32(b)))
def old():
    (_, _), at = gnn.split_transpose(a, plain=False); (_, _), bt = gnn.split_transpose(b, plain=False)
    return gnn._tc_gemm_over_rows(at, bt, K)
(_, _), at = gnn.split_transpose(a, plain=False); (_, _), bt = gnn.split_transpose(b, plain=False)
t_kmajor = timed(lambda: gnn._tc_gemm_over_rows(at, bt, K))
t_old = timed(old)
fl = 2.0 * K * Ca * Cb
print(f"dW [{Ca},{Cb}] over K = {K}: MN-major GEMM {t_tn:.1f} us ({fl / t_tn / 1e6:.1f} TFLOP/s fp32-equivalent, {3 * fl / t_tn / 1e6:.1f} TF32-issued)"
      f" | K-major GEMM on transposed copies {t_kmajor:.1f} us | the two transposes + that GEMM {t_old:.1f} us | plain splits of both operands (shared with the other products) {t_split:.1f} us")
ref = a.double().t() @ b.double()
c = gnn.tc_gemm_tn(sa, sb)
print("max error / max entry:", float((c.double() - ref).abs().max() / ref.abs().max()))
torch.cuda.synchronize()
torch.cuda.profiler.start()
gnn.tc_gemm_tn(sa, sb)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
