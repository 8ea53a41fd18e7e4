"""dev: gradient error of the GCN per parameter against the fp64 restatement, for the train-GEMM modes."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np, torch
from drl_graph_exploration_b200 import Networks, gnn
from drl_graph_exploration_b200.data import Data
from oracle import gnn_ref
from test_graph_gnn_parity_gpu import _random_graph_batch
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
for ng in (12, 64):
    batch = _random_graph_batch(np.random.default_rng(0), ng, dev)
    torch.manual_seed(1)
    model = Networks.GCN().to(dev)
    ref = gnn_ref.GCN().double().to(dev)
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    d64 = gnn_ref.Graph(batch.x.double(), batch.edge_index, batch.edge_attr.double())
    out_ref = ref(d64, 0.0, batch=batch.batch)
    a = torch.zeros(out_ref.numel(), device=dev); a[::7] = 1.0
    y = torch.randn_like(a)
    ((out_ref.view(-1) * a.double() - y.double()) ** 2).sum().div(64).backward()
    print("graphs", ng, "nodes", batch.x.size(0))
    for mode in ("fp32", "tc3"):
        Networks.set_matmul_precision("tc3", train=mode)
        model.zero_grad()
        out = model(Data(batch.x, batch.edge_index, batch.edge_attr), 0.0, batch=batch.batch)
        ((out.view(-1) * a - y) ** 2).sum().div(64).backward()
        errs = {n1: float((p1.grad.double() - p2.grad).abs().max() / p2.grad.abs().max()) for (n1, p1), (_, p2) in zip(model.named_parameters(), ref.named_parameters())}
        print(" ", mode, "out err", float((out.double() - out_ref).abs().max() / out_ref.abs().max()), {k: f"{v:.2e}" for k, v in errs.items()})
    # the three products alone, on this batch's own operands
    h = torch.randn(batch.x.size(0), 1000, device=dev).relu(); g = torch.randn(batch.x.size(0), 1000, device=dev) * (torch.rand(batch.x.size(0), 1, device=dev) < 0.15)
    W = model.conv2.weight.detach()
    for name, A, B in (("h @ W", h, W), ("g @ W^T", g, W.t().contiguous()), ("h^T @ g", h.t().contiguous(), g)):
        r64 = A.double() @ B.double()
        e32 = float(((A @ B).double() - r64).abs().max() / r64.abs().max())
        etc = float((gnn.tc_matmul(A.contiguous(), B.contiguous()).double() - r64).abs().max() / r64.abs().max())
        print(f"  product {name}: K = {A.size(1)}: fp32 {e32:.2e}  tc3 {etc:.2e}")
