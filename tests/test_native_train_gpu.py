"""The native DQN training step of the GCN (gnn.gcn_train_step + dist.NativeAdam: csrc/dge_train.cu, no autograd, the three dense
products on the tcgen05 3xTF32 GEMM) against the fp64 restatement of PyG's layers (oracle/gnn_ref.py) under autograd.

ReLU is not differentiable at 0: a pre-activation within the forward error of zero (3xTF32: ~1e-5 of the largest entry) may sit on
the other side of the kink in fp64, and ONE flipped unit moves a bias gradient by a whole summand (~1e-2 of its largest entry) --
for any fp32 implementation, with a probability proportional to its forward error.  So, like the knife-edge cells of the occupancy
grid, the comparison is made kink by kink: the fp64 reference takes the ReLU / dropout pattern of the CUDA path (the flips are
counted and each must be a genuine knife edge), and then every gradient has to agree to 2e-4 of its largest entry."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_with_pattern(ref, d64, m1, m2s):
    """fp64 forward of Networks.py:18-28 with the activation pattern given: h1 = z1 * m1, dropout(relu(z2)) = z2 * m2s."""
    z1 = ref.conv1(d64.x, d64.edge_index, d64.edge_attr)
    z2 = ref.conv2(z1 * m1, d64.edge_index, d64.edge_attr)
    return ref.fully_con1(z2 * m2s), z1, z2


@pytest.mark.parametrize("n_graphs,drop_p", [(12, 0.0), (12, 0.5), (40, 0.5)])
def test_native_step_gradients_match_fp64_reference(n_graphs, drop_p):
    from drl_graph_exploration_b200 import Networks, gnn
    from drl_graph_exploration_b200.data import Data
    from drl_graph_exploration_b200.dist import FlatGradBucket
    from oracle import gnn_ref
    from test_graph_gnn_parity_gpu import _random_graph_batch

    dev = torch.device("cuda")
    batch = _random_graph_batch(np.random.default_rng(n_graphs), n_graphs, dev)
    N = batch.x.size(0)
    torch.manual_seed(3)
    model = Networks.GCN().to(dev)
    with torch.no_grad():
        model.conv1.bias.normal_(0, 0.05); model.conv2.bias.normal_(0, 0.05)      # (zero-initialised by the constructor: make them count)
    ref = gnn_ref.GCN().double().to(dev)
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    bucket = FlatGradBucket(model.parameters())
    a = torch.zeros(N, device=dev); a[::7] = 1.0
    y = torch.randn(N, device=dev)
    d32 = Data(batch.x, batch.edge_index, batch.edge_attr)
    gs = Networks._structure(d32, N)
    loss, q, dbg = gnn.gcn_train_step(model, batch.x, gs, a, y, 1.0 / 64, drop_p, seed=1234, return_debug=True)
    torch.cuda.synchronize()
    h1, d2 = dbg["h1"].clone(), dbg["d2"].clone()
    scale = 1.0 / (1.0 - drop_p)
    m1, m2s = (h1 > 0).double(), (d2 != 0).double() * scale
    if drop_p > 0:   # the dropout keeps about half of the active units
        active = int((m2s > 0).sum())
        d64 = gnn_ref.Graph(batch.x.double(), batch.edge_index, batch.edge_attr.double())
        with torch.no_grad():
            _, _, z2 = _ref_with_pattern(ref, d64, m1, m2s)
        frac = active / max(int((z2 > 0).sum()), 1)
        assert abs(frac - (1.0 - drop_p)) < 0.02, frac
    d64 = gnn_ref.Graph(batch.x.double(), batch.edge_index, batch.edge_attr.double())
    out_ref, z1, z2 = _ref_with_pattern(ref, d64, m1, m2s)
    # the pattern the CUDA path took is the fp64 one except on knife edges
    flip1 = (z1 > 0).double() != m1
    assert flip1.float().mean() < 1e-4 and (not flip1.any() or float(z1[flip1].abs().max()) < 1e-4 * float(z1.abs().max()))
    kept = m2s > 0
    assert bool((z2[kept] > -1e-4 * float(z2.abs().max())).all()), "a kept unit must be active (up to a knife edge)"
    if drop_p == 0:
        flip2 = (z2 > 0) != kept
        assert flip2.float().mean() < 1e-4 and (not flip2.any() or float(z2[flip2].abs().max()) < 1e-4 * float(z2.abs().max()))
    loss_ref = ((out_ref.view(-1) * a.double() - y.double()) ** 2).sum() / 64
    loss_ref.backward()
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    assert float((q.double() - out_ref.view(-1)).abs().max()) <= 1e-4 * float(out_ref.abs().max())
    for (n1, p1), (n2, p2) in zip(model.named_parameters(), ref.named_parameters()):
        g = p2.grad.abs().max()
        err = float((p1.grad.double() - p2.grad).abs().max() / g)
        assert err <= 2e-4, (n1, err)


def test_native_adam_equals_torch_adam_with_the_clamp():
    from drl_graph_exploration_b200.dist import NativeAdam

    dev = torch.device("cuda")
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(1000, 37, device=dev)), torch.nn.Parameter(torch.randn(53, device=dev))]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    nat, ref = NativeAdam(ps, lr=1e-3), torch.optim.Adam(qs, lr=1e-3)
    for it in range(5):
        grads = [torch.randn_like(p) * (3.0 if it % 2 else 0.3) for p in ps]
        for p, q_, g in zip(ps, qs, grads):
            p.grad.copy_(g)
            q_.grad = (g * 0.5).clamp(-0.5, 0.5)
        v0 = ps[0]._version
        nat.step(clamp=0.5, gscale=0.5)
        ref.step()
        assert ps[0]._version > v0                                     # version-keyed caches see the raw-pointer update
        assert float(nat.bucket.flat.abs().max()) <= 0.5              # the gradient Adam saw
        for p, q_ in zip(ps, qs):
            assert float((p - q_).abs().max()) <= 2e-6 * float(q_.abs().max())
    sd = nat.state_dict()
    nat2 = NativeAdam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1e-3)
    nat2.load_state_dict(sd)
    assert int(nat2.step_t) == 5 and torch.equal(nat2.m, nat.m)


def test_deepq_train_takes_the_native_route_and_learns():
    from drl_graph_exploration_b200 import Networks, gnn
    from drl_graph_exploration_b200.dist import NativeAdam
    from drl_graph_exploration_b200.policy import DeepQ
    from test_graph_gnn_parity_gpu import _random_graph_batch

    dev = torch.device("cuda")
    batch = _random_graph_batch(np.random.default_rng(5), 16, dev)
    N = batch.x.size(0)
    torch.manual_seed(1)
    model = Networks.GCN().to(dev)
    opt = NativeAdam(model.parameters(), lr=1e-3)
    dq = DeepQ(); dq.BATCH = 16
    a = torch.zeros(N, device=dev); a[::11] = 1.0
    y = torch.ones(N, device=dev)
    l0 = gnn.launch_count
    losses = [float(dq.train(batch, a, y, dev, model, opt)) for _ in range(12)]
    assert gnn.launch_count > l0 and dq.native_steps == 12
    assert np.isfinite(losses).all() and np.mean(losses[-3:]) < np.mean(losses[:3])     # fits the constant target
