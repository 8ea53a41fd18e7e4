"""tcgen05 3xTF32 GEMM (csrc/dge_gemm.cu) against an fp64 reference of the same product.

Tolerance: the contract on Q-values is 1e-4 relative in fp32 (BASELINE north_star); the kernel is required to be
fp32-GEMM quality: |C - C64| <= 4e-6 * (|A| |B|) element-wise (the dropped lo*lo term and the TF32 rounding of the lo
parts bound it near 3 * 2^-21 = 1.4e-6; a single-pass TF32 GEMM sits near 1e-3, the library fp32 GEMM near 1e-7)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(M, K, N, seed=0, m_dev=None):
    from drl_graph_exploration_b200 import gnn
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(M, K, device="cuda", generator=g) * torch.rand(M, 1, device="cuda", generator=g) * 3
    w = torch.randn(K, N, device="cuda", generator=g) / K ** 0.5
    hi, lo = gnn._weight_operand(w, True)
    md = None if m_dev is None else torch.tensor([m_dev], dtype=torch.int32, device="cuda")
    c = torch.full((M, N), float("nan"), device="cuda")
    ah, al = gnn.split_tf32(a)
    L = gnn._gemm_lib()
    rc = L.dge_gemm_tf32x3(M, gnn._p(md), N, K, gnn._p(ah), gnn._p(al), gnn._p(hi), gnn._p(lo), gnn._p(c), N, gnn._st(a.device))
    assert rc == 0
    torch.cuda.synchronize()
    live = M if m_dev is None else m_dev
    ref = a.double() @ w.double()
    scale = a.double().abs() @ w.double().abs()
    err = ((c[:live].double() - ref[:live]).abs() / scale[:live]).max().item()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = ((a @ w).double()[:live] - ref[:live]).abs().div(scale[:live]).max().item()
    torch.backends.cuda.matmul.allow_tf32 = prev
    assert err <= 4e-6, (M, K, N, err)
    assert lib <= 4e-6
    if live < M:
        assert torch.isnan(c[live:]).all()          # rows beyond the device-side count are left untouched
    return err, lib


@pytest.mark.parametrize("M", [1, 31, 128, 129, 1000, 4097])
def test_gemm_node_mlp_shape(M):
    _check(M, 1000, 1000, seed=M)


def test_gemm_other_shapes():
    _check(300, 1000, 3000, seed=1)      # GRUCell transform [N,1000] x [1000,3000]
    _check(257, 64, 100, seed=2)         # ragged N (value heads: Linear(1000,100)), short K
    _check(130, 36, 8, seed=3)           # K tail inside one 32-wide block, N < one tile


def test_gemm_device_side_row_count():
    _check(1024, 1000, 1000, seed=5, m_dev=700)
    _check(1024, 1000, 1000, seed=6, m_dev=1)


def test_matmul_autograd_and_gcn_forward_parity():
    """tc_matmul inside autograd (grad-input on the tcgen05 kernel) and a GCN forward in 'tc3' vs 'fp32' mode."""
    from drl_graph_exploration_b200 import Networks, gnn
    from drl_graph_exploration_b200.data import Data
    torch.manual_seed(0)
    x = torch.randn(200, 1000, device="cuda", requires_grad=True)
    w = (torch.randn(1000, 1000, device="cuda") / 30).requires_grad_()
    y = gnn.tc_matmul(x, w)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr, wr = x.detach().double().requires_grad_(), w.detach().double().requires_grad_()
    (xr @ wr).backward(gy.double())
    # element-wise error of the 3xTF32 product is ~1e-6 of sum|a||w| (~21 here): 1e-4 absolute on O(1) values
    assert torch.allclose(y.double(), (xr @ wr).detach(), rtol=1e-5, atol=1e-4)
    assert torch.allclose(x.grad.double(), xr.grad, rtol=1e-5, atol=1e-4)
    assert torch.allclose(w.grad.double(), wr.grad, rtol=1e-4, atol=1e-3)
    # GCN Q-values: tensor-core path vs library fp32 path, 1e-4 relative (the contract), observed ~1e-6
    n = 300
    ei = torch.randint(0, n, (2, 900), device="cuda")
    ei = torch.cat([ei, ei.flip(0)], dim=1)
    ea = torch.rand(900, device="cuda").repeat(2) * 5 + 0.1
    data = Data(torch.randn(n, 5, device="cuda"), ei, ea)
    model = Networks.GCN().cuda().eval()
    with torch.no_grad():
        Networks.set_matmul_precision("tc3"); q1 = model(data, 0.0)
        Networks.set_matmul_precision("fp32"); q0 = model(data, 0.0)
        Networks.set_matmul_precision("tc3")
    assert torch.allclose(q1, q0, rtol=1e-4, atol=1e-5), (q1 - q0).abs().max()
