"""Which piece of the GG-NN training path moves the gradients away from fp64?  (dev probe)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from drl_graph_exploration_b200 import Networks, gnn
from drl_graph_exploration_b200.data import Data
from oracle import gnn_ref
from test_graph_gnn_parity_gpu import _random_graph_batch

dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = False
batch = _random_graph_batch(np.random.default_rng(21), 10, dev)
torch.manual_seed(5)
model = Networks.GGNN().to(dev).train()
ref = gnn_ref.GGNN().double().to(dev)
ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
N = batch.x.size(0)
a = torch.zeros(N, device=dev); a[::5] = 1.0
y = torch.randn(N, device=dev)
out_ref = ref(gnn_ref.Graph(batch.x.double(), batch.edge_index, batch.edge_attr.double()), 0.0, batch=batch.batch)
(((out_ref.view(-1) * a.double() - y.double()) ** 2).sum() / 64).backward()
gref = {n: p.grad for n, p in ref.named_parameters()}

def run(tag):
    model.zero_grad(set_to_none=True)
    out = model(Data(batch.x, batch.edge_index, batch.edge_attr), 0.0, batch=batch.batch)
    (((out.view(-1) * a - y) ** 2).sum() / 64).backward()
    errs = {n: float((p.grad.double() - gref[n]).abs().max() / gref[n].abs().max()) for n, p in model.named_parameters()}
    print(tag, "out", float((out.detach().double() - out_ref.detach()).abs().max() / out_ref.abs().max()), {k: f"{v:.2e}" for k, v in errs.items()})

Networks.set_matmul_precision("tc3", train="fp32"); run("lib        ")
Networks.set_matmul_precision("tc3", train="tc3"); run("tc3 all    ")
orig_cell, orig_lin, orig_mm = gnn.gru_cell_train, gnn.tc_linear, gnn.tc_matmul
gnn.gru_cell_train = lambda m, h, rnn: rnn(m, h); run("tc3 mm only (torch GRUCell)")
gnn.gru_cell_train = orig_cell
gnn.tc_matmul = lambda x, w: x @ w; run("native GRU, lib h@W")
gnn.tc_linear = lambda x, w: torch.nn.functional.linear(x, w); run("gate kernels only (lib products)")
gnn.tc_linear = orig_lin
# the gate kernels on lib products but tc products for the weight gradient only
class _L(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w); return torch.nn.functional.linear(x, w)
    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        (_, _), gt = gnn.split_transpose(gy, plain=False); (_, _), xt = gnn.split_transpose(x, plain=False)
        gw = gnn._tc_gemm_over_rows(gt, xt, x.shape[0])
        print("   dW tc vs fp64 product of the same operands:", float((gw.double() - gy.double().t() @ x.double()).abs().max() / (gy.double().t() @ x.double()).abs().max()),
              " sum|a||b| / max|c|:", float((gy.double().abs().t() @ x.double().abs()).max() / (gy.double().t() @ x.double()).abs().max()))
        return gy @ w, gw
gnn.tc_linear = lambda x, w: _L.apply(x, w); run("lib fwd/gx, tc dW")
