"""dev: time the tcgen05 3xTF32 GEMM against the library fp32 / tf32 GEMM at the node-MLP shapes."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from drl_graph_exploration_b200 import gnn

def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

L = gnn._gemm_lib()
for M in (2048, 4096, 8192, 16384, 32768):
    K = N = 1000
    x = torch.randn(M, K, device="cuda"); w = torch.randn(K, N, device="cuda") / 30
    hi, lo = gnn._weight_operand(w, True)
    ah, al = gnn.split_tf32(x)
    c = torch.empty(M, N, device="cuda")
    t_core = timeit(lambda: L.dge_gemm_tf32x3(M, None, N, K, gnn._p(ah), gnn._p(al), gnn._p(hi), gnn._p(lo), gnn._p(c), N, gnn._st(x.device)))
    t_full = timeit(lambda: gnn.tc_matmul(x, w))
    torch.backends.cuda.matmul.allow_tf32 = False
    t_fp32 = timeit(lambda: x @ w)
    torch.backends.cuda.matmul.allow_tf32 = True
    t_tf32 = timeit(lambda: x @ w)
    torch.backends.cuda.matmul.allow_tf32 = False
    fl = 2.0 * M * N * K
    print(f"M={M:6d}: tcgen05 3xTF32 kernel {t_core:8.1f} us ({3 * fl / t_core / 1e6:7.1f} TF/s tf32-issued, {fl / t_core / 1e6:6.1f} TF/s fp32-equivalent) | "
          f"with split {t_full:8.1f} us | cuBLAS fp32 {t_fp32:8.1f} us ({fl / t_fp32 / 1e6:6.1f} TF/s) | cuBLAS tf32 {t_tf32:8.1f} us")
