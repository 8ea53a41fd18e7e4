import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from drl_graph_exploration_b200 import gnn
L = gnn._gemm_lib()
for M in (1000, 2048, 2432, 2560, 4097, 8192):
    K = N = 1000
    torch.manual_seed(M)
    a = torch.randn(M, K, device="cuda"); w = torch.randn(K, N, device="cuda") / 30
    hi, lo = gnn._weight_operand(w, True); ah, al = gnn.split_tf32(a)
    for rep in range(2):
        c = torch.full((M, N), float("nan"), device="cuda")
        L.dge_gemm_tf32x3(M, None, N, K, gnn._p(ah), gnn._p(al), gnn._p(hi), gnn._p(lo), gnn._p(c), N, gnn._st(a.device))
        torch.cuda.synchronize()
        ref = (a.double() @ w.double())
        bad = ~((c.double() - ref).abs() <= 1e-4 * (1 + ref.abs()))
        nb = int(bad.sum())
        tiles = sorted({(int(r) // 128, int(cc) // 128) for r, cc in bad.nonzero()[:200000:97].tolist()})
        print(f"M={M} rep={rep} ctas={8 * ((M + 127) // 128)} bad={nb} nan={int(torch.isnan(c).sum())} tiles(sample)={tiles[:24]}")
        if nb:
            r, cc = bad.nonzero()[0].tolist()
            print("   first bad", r, cc, float(c[r, cc]), float(ref[r, cc]), " row bad cols:", bad[r].nonzero().flatten()[:12].tolist(), " col bad rows:", bad[:, cc].nonzero().flatten()[:12].tolist())
