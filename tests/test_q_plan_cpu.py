"""gnn.QForwardPlan (the acting loop's prepared Q forward) against the generic route, without a GPU: the C entry point is replaced
by a recorder, and both routes must hand it the SAME argument list -- same pointers in the same order, same sizes -- for the same
graph batch.  (The values themselves are compared on the GPU in tests/test_zz_q_plan_gpu.py.)"""
import contextlib
import ctypes
import types

import pytest
import torch

from drl_graph_exploration_b200 import Networks, gnn
from drl_graph_exploration_b200.envs.exploration_env import GraphBatch


class _Recorder:
    def __init__(self):
        self.calls = []
        self._q_forward_ready = True

    def dge_gcn_q_forward(self, *args):
        self.calls.append([a.value if isinstance(a, ctypes.c_void_p) else a for a in args])
        return 0

    def dge_gemm_prep_weight(self, *args):
        return 0


@pytest.fixture
def fake_gpu(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(gnn, "_gemm_lib", lambda: rec)
    monkeypatch.setattr(gnn, "_need_cuda", lambda t, what: None)
    monkeypatch.setattr(gnn, "_st", lambda dev: ctypes.c_void_p(1234))
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: types.SimpleNamespace(cuda_stream=1234))
    monkeypatch.setattr(torch.cuda, "current_device", lambda: None)       # == torch.device("cpu").index: the plan's device is "current"
    return rec


def _batch(n_nodes, n_edges, node_cap=64, edge_cap=256):
    """The buffers of envs.exploration_env.GraphBatch (engine-owned in the product), filled by hand."""
    b = types.SimpleNamespace(node_cap=node_cap, edge_cap=edge_cap, n_nodes=n_nodes, n_edges=n_edges, n_graphs=2)
    b.x = torch.randn(node_cap, 5)
    b.edge_index = torch.randint(0, n_nodes, (2, edge_cap))
    b.edge_attr = torch.rand(edge_cap)
    b.batch = torch.zeros(node_cap, dtype=torch.int64)
    b.csr_rowptr = torch.zeros(node_cap + 1, dtype=torch.int32)
    b.csr_perm = torch.zeros(edge_cap, dtype=torch.int32)
    b.gcn_norm = torch.rand(edge_cap)
    b.gcn_selfnorm = torch.rand(node_cap)
    return b


def _generic(model, b):
    """What Networks.GCN._trunk does at inference (tc3 mode) with GraphBatch.data()."""
    d = GraphBatch.data(b)
    gs = Networks._structure(d, d.x.size(0))
    return gnn.gcn_q_forward(d.x, gs, model.conv1.weight, model.conv1.bias, model.conv2.weight, model.conv2.bias,
                             model.fully_con1.weight[0], model.fully_con1.bias)


def test_plan_passes_the_same_arguments_as_the_generic_route(fake_gpu):
    torch.manual_seed(0)
    model = Networks.GCN()
    b = _batch(23, 80)
    q_gen = _generic(model, b)
    plan = gnn.QForwardPlan(model, b)
    q_plan = plan(23)
    gen, pl = fake_gpu.calls
    assert len(gen) == len(pl) == 19
    ws_g, q_g, ws_p, q_p = gen[16], gen[17], pl[16], pl[17]
    assert gen[:16] == pl[:16] and gen[18] == pl[18]             # sizes, graph pointers, model pointers, stream
    assert q_g == q_gen.data_ptr() and q_p == plan.q.data_ptr() and q_plan.data_ptr() == plan.q.data_ptr() and q_plan.numel() == 23
    assert ws_g and ws_p                                         # each route owns a workspace of >= 3 N C floats
    assert plan._ws.numel() >= 3 * 23 * 1000
    # pointers are what they claim to be
    assert pl[3] == b.x.data_ptr() and pl[4] == b.csr_rowptr.data_ptr() and pl[5] == b.csr_perm.data_ptr() and pl[6] == b.edge_index.data_ptr()
    assert pl[7] == b.gcn_norm.data_ptr() and pl[8] == b.gcn_selfnorm.data_ptr()
    assert pl[9] == model.conv1.weight.data_ptr() and pl[10] == model.conv1.bias.data_ptr() and pl[13] == model.conv2.bias.data_ptr()
    assert pl[14] == model.fully_con1.weight.data_ptr() and pl[15] == model.fully_con1.bias.data_ptr()
    assert (pl[0], pl[1], pl[2]) == (23, 5, 1000)

    # the next tick: another node count, same pointers, no re-preparation
    keep = plan._keep
    b.n_nodes, b.n_edges = 40, 120
    _generic(model, b); plan(40)
    gen2, pl2 = fake_gpu.calls[2:]
    assert gen2[:16] == pl2[:16] and pl2[0] == 40 and plan._keep is keep


def test_plan_follows_weight_updates_and_refuses_other_models(fake_gpu):
    torch.manual_seed(0)
    model = Networks.GCN()
    b = _batch(10, 30)
    plan = gnn.QForwardPlan(model, b)
    plan(10)
    hi_before = fake_gpu.calls[-1][11]
    with torch.no_grad():
        model.conv2.weight.mul_(0.5)                              # an optimizer step / load_state_dict bumps the version counter
    _generic(model, b); plan(10)
    gen, pl = fake_gpu.calls[-2:]
    assert gen[:16] == pl[:16] and pl[11] != hi_before            # a fresh TF32 split of W2, the same one the generic route uses
    with pytest.raises(gnn.DgeError):
        plan(b.node_cap + 1)
    torch.cuda.current_device = lambda: 7                         # another device is current: the call is wrapped in a device guard
    plan(10)
    assert fake_gpu.calls[-1][:16] == pl[:16]
    # eligibility: only the DQN Q-network on a GPU (here: CPU parameters -> not eligible; other families never)
    assert not gnn.QForwardPlan.eligible(model)
    assert not gnn.QForwardPlan.eligible(Networks.PolicyGCN()) and not gnn.QForwardPlan.eligible(Networks.GGNN())
    assert not gnn.QForwardPlan.eligible(torch.nn.Linear(5, 1))


def test_packed_arena_pointers_equal_the_views_of_the_generic_route(fake_gpu):
    """The host loop's packed graph batch (dge_graph_packed layout): runner.packed_graph_args hands the plan the addresses that
    runner.packed_graph_data + the generic route pass for the same arena."""
    from drl_graph_exploration_b200 import runner
    torch.manual_seed(1)
    model = Networks.GCN()
    n, e = 17, 52
    al = lambda v: (v + 15) // 16 * 16
    off, pk = 128, types.SimpleNamespace()
    for name, nbytes in (("x", n * 20), ("edge_index", 2 * e * 8), ("edge_attr", e * 4), ("node_ptr", 12), ("edge_ptr", 12), ("key_size", 8), ("fro_size", 8),
                         ("frontier_xy", 64), ("csr_rowptr", (n + 1) * 4), ("csr_perm", e * 4), ("gcn_norm", e * 4), ("gcn_selfnorm", n * 4)):
        setattr(pk, name, off)
        off = al(off + nbytes)
    arena = torch.zeros(off + 64, dtype=torch.uint8)
    data = runner.packed_graph_data(arena, pk, n, e)
    gs = Networks._structure(data, n)
    gnn.gcn_q_forward(data.x, gs, model.conv1.weight, model.conv1.bias, model.conv2.weight, model.conv2.bias, model.fully_con1.weight[0], model.fully_con1.bias)
    plan = gnn.QForwardPlan(model, _batch(8, 8))
    plan(n, runner.packed_graph_args(arena, pk))
    gen, pl = fake_gpu.calls
    assert gen[:16] == pl[:16] and pl[3] == arena.data_ptr() + pk.x and pl[6] == arena.data_ptr() + pk.edge_index


def test_glue_errors_fall_back_to_the_module_route_loudly_native_errors_do_not():
    from drl_graph_exploration_b200 import runner

    def broken(n, gargs=None):
        raise ValueError("glue")

    loop = types.SimpleNamespace(_plan=broken)
    with pytest.warns(UserWarning, match="QForwardPlan disabled"):
        assert runner._planned_q(loop, 5, None) is None
    assert loop._plan is None

    def native_failure(n, gargs=None):
        raise gnn.DgeError("dge_gcn_q_forward failed (-2)")

    loop = types.SimpleNamespace(_plan=native_failure)
    with pytest.raises(gnn.DgeError):
        runner._planned_q(loop, 5, None)
    assert loop._plan is native_failure
    assert runner._make_plan(torch.nn.Linear(5, 1), None) is None        # not the DQN Q-network: no plan, no warning
