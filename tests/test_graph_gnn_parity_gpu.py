"""GPU parity of the exploration-graph builder, the line planner / policy read-out and the GNN
kernels (through the C ABI) against the CPU oracle / the pure-PyTorch GNN reference.

Integer topology (node counts, frontier cells, COO edge list) bit-exact; f32 node features and
edge weights within 1 f32 ulp of the oracle's f64 values; Q-values within 1e-4 relative."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import RESET_ODOM, choose_actions, make_oracles, world_arrays
from drl_graph_exploration_b200.config import EnvConfig

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _drive(cfg, B, n_dec, max_poses=160):
    """Engine + oracles advanced together (same worlds / noise / actions)."""
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv

    oracles = make_oracles(cfg, range(B))
    start, lm, scan, noise0 = world_arrays(oracles)
    env = VecExplorationEnv(B, cfg=cfg, max_poses=max_poses)
    dev = env.device
    t = lambda a: torch.as_tensor(a, device=dev)
    env.eng.reset(seeds=t(np.arange(B, dtype=np.int64)), start=t(start), landmarks=t(lm), scan=t(scan), noise=t(noise0))

    def step(odoms):
        noise = np.stack([o.step(od) for o, od in zip(oracles, odoms)])
        env.eng.step(t(np.asarray(odoms, dtype=np.float64)), noise=t(noise))

    for _ in range(4):
        step([RESET_ODOM] * B)
    rng = np.random.default_rng(5)
    yield env, oracles
    for d in range(n_dec):
        plans = [choose_actions(o, rng) for o in oracles]
        for i in range(max(len(p) for p in plans)):
            step([p[i] if i < len(p) else np.array([0.0, 0.0, 0.1]) for p in plans])
        yield env, oracles


@pytest.mark.parametrize("map_size,n_lm", [(20, 30), (40, None)])
def test_graph_matches_oracle(map_size, n_lm):
    cfg = EnvConfig(map_size=map_size, num_landmarks=n_lm)
    B = 6
    compared = skipped = 0
    for env, oracles in _drive(cfg, B, 7):
        torch.cuda.synchronize()
        graphs = env.graph_host()
        seen = env.eng.state["seen"].cpu().numpy()
        fxy_all = env.graph.frontier_xy.cpu().numpy()
        q = torch.randn(env.graph.n_nodes, device=env.device)
        choice = env.select_and_plan(q).cpu().numpy()
        plans = env.eng.state["plan"].cpu().numpy()
        qh = q.cpu().numpy()
        nptr = env.graph.node_ptr.cpu().numpy()
        for b, o in enumerate(oracles):
            if not np.array_equal(seen[b], o.vmap()["seen"]):   # knife-edge cell flipped: topology may differ
                skipped += 1
                continue
            g, r = graphs[b], o.graph()
            assert g["x"].shape[0] == r["n_nodes"] and g["key_size"] == r["key_size"] and g["fro_size"] == r["fro_size"]
            assert np.array_equal(g["edge_index"], r["edge_index"]), "COO edge list (bit-exact, data_process order)"
            assert np.allclose(g["edge_attr"], r["edge_attr"].astype(np.float32), rtol=2e-7, atol=0)
            assert np.array_equal(fxy_all[b, :r["fro_size"]], r["frontier_xy"]), "frontier cells"
            ref_x = r["features"].astype(np.float32)
            K, F = r["key_size"], r["fro_size"]
            assert np.array_equal(g["x"][:, 4], ref_x[:, 4]) and np.array_equal(g["x"][:, 3], ref_x[:, 3])
            assert np.allclose(g["x"][:, 1], ref_x[:, 1], rtol=3e-6, atol=1e-6)
            dth = np.abs(g["x"][:, 2].astype(np.float64) - ref_x[:, 2])      # bearing difference in [0, 2pi): a node dead ahead
            far = ref_x[:, 1] > 1e-6    # a node at the robot's own position (in-place rotation, dead reckoning) has no direction:
            assert np.all(np.minimum(dth, 2 * math.pi - dth)[far] < 5e-6)     # atan2 of a ~1e-17 vector is noise in the reference too
            assert np.allclose(g["x"][:K, 0], ref_x[:K, 0], rtol=3e-6, atol=1e-9) and np.allclose(g["x"][K:, 0], ref_x[K:, 0], rtol=1e-4)  # frontier trace comes from the CI fold
            # policy read-out + line plan
            K, F = r["key_size"], r["fro_size"]
            a = int(np.argmax(qh[nptr[b] + K: nptr[b] + K + F]))
            assert choice[b] == a
            acts = o.line_plan(*r["frontier_xy"][a])
            assert int(plans[b, 5]) == len(acts)
            nrot, nfwd = int(plans[b, 0]), int(plans[b, 3])
            assert nfwd == sum(1 for v in acts if v[0] == cfg.max_edge_length) and nrot + nfwd + 2 == len(acts)
            assert abs(plans[b, 1] * plans[b, 2] - acts[nrot][2]) < 1e-7
            assert abs(plans[b, 4] - acts[-1][0]) < 1e-7
            compared += 1
    # env-steps where a knife-edge cell (range exactly 6.0 from the integer start pose, or a pose sitting on a
    # cell centre) flipped are skipped; with 6 envs x 8 graphs at least a third must remain comparable
    print(f"graph parity: compared {compared}, skipped {skipped}")
    assert compared >= 16 and compared >= 0.33 * (compared + skipped)
    env.close()


def _ref_graph_batch(oracles, device, dtype=torch.float32):
    from oracle import gnn_ref
    xs, eis, eas, off, sizes = [], [], [], 0, []
    for o in oracles:
        g = o.graph()
        xs.append(torch.tensor(g["features"], dtype=dtype)); eis.append(torch.tensor(g["edge_index"]) + off)
        eas.append(torch.tensor(g["edge_attr"], dtype=dtype)); off += g["n_nodes"]; sizes.append((g["n_nodes"], g["key_size"], g["fro_size"]))
    return gnn_ref.Graph(torch.cat(xs).to(device), torch.cat(eis, dim=1).to(device), torch.cat(eas).to(device)), sizes


def _random_graph_batch(rng, n_graphs, device):
    """BASELINE config 5: mixed 8..512-node graphs: pose chain + landmark observations + frontier stubs."""
    from drl_graph_exploration_b200.data import Batch, Data
    items = []
    for _ in range(n_graphs):
        n = int(rng.choice(np.arange(8, 513, 8)))
        L = max(1, n // 8); F = max(1, min(L + 1, n // 10)); T = n - L - F
        edges = [(L + k, L + k + 1) for k in range(T - 1)]
        for k in range(T):
            for j in rng.choice(L, size=min(L, rng.poisson(1.5)), replace=False):
                edges.append((int(j), L + k))
        edges += [(int(rng.integers(L + T)), L + T + f) for f in range(F)]
        edges = sorted(set((min(a, b), max(a, b)) for a, b in edges))
        w = rng.uniform(0.1, 6.0, size=len(edges)).astype(np.float32)
        ei = np.array([[a, b] for (a, b) in edges for _ in (0, 1)]).T.copy()
        ei[:, 1::2] = ei[::-1, 1::2]
        items.append(Data(torch.tensor(rng.normal(size=(n, 5)).astype(np.float32)), torch.tensor(ei, dtype=torch.long), torch.tensor(np.repeat(w, 2))))
    return Batch.from_data_list(items).to(device)


def test_gcn_forward_matches_reference_with_shipped_weights():
    """DQN+GCN with the reference's trained weights on real exploration graphs."""
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.data import Data
    from oracle import gnn_ref

    gold = np.load(os.path.join(GOLD, "ref_40_DQN_GCN_seed0.npz"))
    sd = {k[3:]: torch.tensor(gold[k]) for k in gold.files if k.startswith("sd_")}
    dev = torch.device("cuda")
    model = Networks.GCN().to(dev); model.load_state_dict(sd); model.eval()
    ref32 = gnn_ref.GCN().to(dev); ref32.load_state_dict(sd); ref32.eval()
    ref64 = gnn_ref.GCN().double().to(dev); ref64.load_state_dict({k: v.double() for k, v in sd.items()}); ref64.eval()
    cfg = EnvConfig(map_size=40)
    oracles = make_oracles(cfg, range(8))
    rng = np.random.default_rng(2)
    for o in oracles:
        for _ in range(4):
            o.step(RESET_ODOM)
        for _ in range(3):
            for a in choose_actions(o, rng):
                o.step(a)
    g32, sizes = _ref_graph_batch(oracles, dev)
    g64, _ = _ref_graph_batch(oracles, dev, torch.float64)
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        q = model(Data(g32.x, g32.edge_index, g32.edge_attr), 0.0).view(-1)            # fused inference path
        with torch.enable_grad():
            q_train_path = model(Data(g32.x, g32.edge_index, g32.edge_attr), 0.0).view(-1)   # unfused path
        r32 = ref32(g32, 0.0).view(-1); r64 = ref64(g64, 0.0).view(-1).float()
    scale = r64.abs().max()
    assert (q - r64).abs().max() <= 1e-4 * scale, float((q - r64).abs().max() / scale)
    assert (q_train_path.detach() - r64).abs().max() <= 1e-4 * scale
    # regression guard, tighter than the contract: the 3xTF32 tensor-core GEMM adds ~1e-6 of sum|a||w| per product, which
    # the trained weights' cancellation turns into ~1.3e-5 of the Q scale (library fp32: ~3e-7)
    assert (q - r64).abs().max() <= 4e-5 * scale
    off = 0
    for n, k, f in sizes:   # the decision (arg-max over frontier nodes) agrees
        if f > 0:
            assert int(torch.argmax(q[off + k: off + n])) == int(torch.argmax(r64[off + k: off + n]))
        off += n


def test_gcn_backward_and_other_families_match_reference():
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.data import Data
    from oracle import gnn_ref

    dev = torch.device("cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    rng = np.random.default_rng(0)
    batch = _random_graph_batch(rng, 12, dev)
    torch.manual_seed(1)
    # The GCN's gradients are compared entry by entry with fp64 below, without matching ReLU patterns: that holds for the autograd glue +
    # aggregation kernels on the library fp32 GEMM (the A/B route, 'fp32').  The default training products (tcgen05 3xTF32) are held
    # kink by kink in test_native_train_gpu.py and op by op in test_family_train_gpu.py.
    Networks.set_matmul_precision("tc3", train="fp32")
    try:
        _families(Networks, Data, gnn_ref, batch, dev)
    finally:
        Networks.set_matmul_precision("tc3", train="tc3")


def _families(Networks, Data, gnn_ref, batch, dev):
    for name, kwargs in (("GCN", {}), ("GGNN", {}), ("GraphUNet", dict(in_channels=5, hidden_channels=1000, out_channels=1000, depth=3))):
        model = getattr(Networks, name)(**kwargs).to(dev)
        ref = getattr(gnn_ref, name)(**kwargs).double().to(dev)
        ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
        d32 = Data(batch.x, batch.edge_index, batch.edge_attr)
        d64 = gnn_ref.Graph(batch.x.double(), batch.edge_index, batch.edge_attr.double())
        out = model(d32, 0.0, batch=batch.batch)
        out_ref = ref(d64, 0.0, batch=batch.batch)
        scale = out_ref.abs().max()
        assert out.shape == out_ref.shape and torch.isfinite(out).all()
        if name == "GraphUNet":
            # top-k pooling selects nodes by score order: an fp32 / fp64 near-tie of two scores may legitimately be ordered the other
            # way, so the fp64 reference pools the nodes the CUDA path kept (checked to BE a top-k selection up to such ties), and then
            # every output has to agree
            rec = []
            out_ref = ref(d64, 0.0, batch=batch.batch, forced_perms=model.last_perms, record=rec)
            for (score, perm64, bt), perm32 in zip(rec, model.last_perms):
                assert perm32.numel() == perm64.numel()
                s32, s64 = set(perm32.tolist()), set(perm64.tolist())
                for node in s32 ^ s64:      # a node only one side kept: its score ties (to 1e-5) with the weakest score the fp64 side kept in its graph
                    gsel = perm64[bt[perm64] == bt[node]]
                    assert abs(float(score[node]) - float(score[gsel].min())) < 1e-5, "pooled set differs beyond a score near-tie"
                assert len(s32 ^ s64) <= max(2, perm32.numel() // 200)
            scale = out_ref.abs().max()
            assert (out.double() - out_ref).abs().max() <= 1e-4 * scale, (name, float((out.double() - out_ref).abs().max() / scale))
            continue
        tol = 1e-4 if name == "GCN" else 2e-3   # 3 stacked GRU layers: looser for the "next" family
        assert (out.double() - out_ref).abs().max() <= tol * scale, (name, float((out.double() - out_ref).abs().max() / scale))
        if name == "GCN":
            a = torch.zeros_like(out).view(-1); a[::7] = 1.0
            y = torch.randn_like(a)
            loss = ((out.view(-1) * a - y) ** 2).sum() / 64          # DeepQ.cost  policy.py:234-239
            loss.backward()
            loss_ref = ((out_ref.view(-1) * a.double() - y.double()) ** 2).sum() / 64
            loss_ref.backward()
            for (n1, p1), (n2, p2) in zip(model.named_parameters(), ref.named_parameters()):
                gs = p2.grad.abs().max()
                assert (p1.grad.double() - p2.grad).abs().max() <= 2e-4 * gs + 1e-9, (n1, float((p1.grad.double() - p2.grad).abs().max() / gs))


def test_policy_and_value_heads_run():
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.data import Data
    dev = torch.device("cuda")
    batch = _random_graph_batch(np.random.default_rng(3), 4, dev)
    mask = torch.zeros(batch.x.size(0), dtype=torch.bool, device=dev)
    mask[-5:] = True; mask[:3] = True
    for cls in (Networks.PolicyGCN, Networks.ValueGCN, Networks.PolicyGGNN, Networks.ValueGGNN):
        out = cls().to(dev)(Data(batch.x, batch.edge_index, batch.edge_attr), mask, batch=batch.batch)
        assert torch.isfinite(out).all()


def test_ggnn_inference_on_native_gru_path_matches_reference():
    """GG-NN at inference: h W_i, the two GRUCell transforms (tcgen05 3xTF32 GEMM) and dge_gru_gates against the fp64
    restatement of PyG's GatedGraphConv + torch GRUCell, with random and with the A2C-style dropout-free trunk."""
    from drl_graph_exploration_b200 import Networks, gnn
    from drl_graph_exploration_b200.data import Data
    from oracle import gnn_ref

    dev = torch.device("cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    batch = _random_graph_batch(np.random.default_rng(4), 10, dev)
    torch.manual_seed(2)
    model = Networks.GGNN().to(dev).eval()
    ref = gnn_ref.GGNN().double().to(dev)
    ref.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    d32 = Data(batch.x, batch.edge_index, batch.edge_attr)
    d64 = gnn_ref.Graph(batch.x.double(), batch.edge_index, batch.edge_attr.double())
    with torch.no_grad():
        l0 = gnn.launch_count
        q = model(d32, 0.0).view(-1)                           # native path
        launches = gnn.launch_count - l0
        r = ref(d64, 0.0).view(-1)
    q_autograd = model(Data(batch.x, batch.edge_index, batch.edge_attr), 0.0).view(-1).detach()   # torch GRUCell path
    scale = r.abs().max()
    assert (q.double() - r).abs().max() <= 1e-4 * scale, float((q.double() - r).abs().max() / scale)
    assert (q - q_autograd).abs().max() <= 1e-4 * scale
    assert launches >= 3 * (1 + 2 * 2 + 1 + 1)               # per layer: (split + GEMM) x 3 transforms, aggregate, gates


def test_csr_build_and_gcn_norm_on_a_large_batch():
    """gnn.GraphStructure on 5 000 nodes / 40 000 edges (several scan tiles, rows longer than a warp): CSR against a stable
    torch sort, improved-GCN normalisation against its definition (PyG 1.x GCNConv.norm with add_remaining_self_loops)."""
    from drl_graph_exploration_b200 import gnn
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(3)
    N, E = 5000, 40000
    src = torch.randint(0, N, (E,), generator=g); dst = torch.randint(0, N, (E,), generator=g)
    dst[:300] = 7                                                     # a hub: one row of 300+ incoming edges
    src[300:420] = 11                                                 # and one node with 120+ outgoing edges
    src[500:520] = dst[500:520]                                       # a few explicit self loops
    w = torch.rand(E, generator=g) * 5 + 0.1
    ei = torch.stack([src, dst]).to(dev); wd = w.to(dev)
    gs = gnn.GraphStructure(ei, wd, N)
    order = torch.sort(ei[1], stable=True)[1]
    rowptr = torch.zeros(N + 1, dtype=torch.int64, device=dev); rowptr[1:] = torch.cumsum(torch.bincount(ei[1], minlength=N), 0)
    assert torch.equal(gs.rowptr_dst.long(), rowptr) and torch.equal(gs.perm_dst.long(), order)
    norm, selfnorm = gs.gcn_norm(True)
    loop = ei[0] == ei[1]
    deg = torch.zeros(N, dtype=torch.float64, device=dev).index_add_(0, ei[0][~loop], wd[~loop].double())
    selfw = torch.full((N,), 2.0, dtype=torch.float64, device=dev)
    has = torch.zeros(N, dtype=torch.bool, device=dev); has[ei[0][loop]] = True
    # nodes with ONE explicit loop keep its weight (the random draw above may give a node two: skip those in the comparison)
    cnt = torch.bincount(ei[0][loop], minlength=N)
    lw = torch.zeros(N, dtype=torch.float64, device=dev).index_add_(0, ei[0][loop], wd[loop].double())
    selfw = torch.where(cnt == 1, lw, selfw)
    ok = cnt <= 1
    deg = deg + selfw
    dis = deg.pow(-0.5)
    ref_norm = torch.where(loop, torch.zeros_like(deg[ei[0]]), dis[ei[0]] * wd.double() * dis[ei[1]])
    eok = ok[ei[0]] & ok[ei[1]]
    assert torch.allclose(norm.double()[eok], ref_norm[eok], rtol=2e-6, atol=1e-9)
    assert torch.allclose(selfnorm.double()[ok], (dis * selfw * dis)[ok], rtol=2e-6, atol=1e-9)


def test_native_augment_adj_equals_sparse_matmul_path():
    """gnn.augment_adj (row-accumulator kernels) against the torch.sparse restatement of GraphUNet.augment_adj
    (add_self_loops -> spspmm(A, A) coalesced -> remove_self_loops): same (row, col) list in the same order, same values."""
    from drl_graph_exploration_b200 import Networks
    dev = torch.device("cuda")
    batch = _random_graph_batch(np.random.default_rng(8), 9, dev)
    ei, ew, bt, n = batch.edge_index, batch.edge_attr, batch.batch, batch.x.size(0)
    for level in range(2):                                       # the second round squares an already squared (denser) adjacency
        nat_i, nat_w = Networks._augment_adj(ei, ew, n, bt)
        loop = torch.arange(n, device=dev)
        A = torch.sparse_coo_tensor(torch.cat([ei, torch.stack([loop, loop])], dim=1), torch.cat([ew, ew.new_ones(n)]), (n, n)).coalesce()
        A2 = torch.sparse.mm(A, A).coalesce()
        keep = A2.indices()[0] != A2.indices()[1]
        ref_i, ref_w = A2.indices()[:, keep], A2.values()[keep]
        assert torch.equal(nat_i, ref_i), level
        assert torch.allclose(nat_w, ref_w, rtol=1e-5, atol=1e-6), level
        assert bool((bt[nat_i[0]] == bt[nat_i[1]]).all())        # stays block diagonal
        ei, ew = nat_i, nat_w
    # a single graph without a batch vector
    one = bt == 0
    n1 = int(one.sum())
    e1 = one[batch.edge_index[0]]
    i1, w1 = Networks._augment_adj(batch.edge_index[:, e1], batch.edge_attr[e1], n1, None)
    assert i1.size(1) > 0 and int(i1.max()) < n1 and bool((i1[0] != i1[1]).all())


def test_native_topk_pooling_equals_the_sort_based_path():
    """gnn.topk_pool / gnn.filter_adj against the two-key stable sort + boolean-mask restatement of PyG topk / filter_adj:
    same perm (order included, ties by lower index), same pooled edge list in the same order."""
    from drl_graph_exploration_b200 import gnn
    dev = torch.device("cuda")
    b = _random_graph_batch(np.random.default_rng(12), 11, dev)
    n = b.x.size(0)
    g = torch.Generator(device="cpu").manual_seed(0)
    score = torch.tanh(torch.randn(n, generator=g)).to(dev)
    score[5:9] = score[5]                                          # exact ties inside a graph
    score[-3:] = score[-3]
    counts = torch.bincount(b.batch)
    k = torch.ceil(0.5 * counts.float()).long()
    zero = counts.new_zeros(1)
    perm, newid = gnn.topk_pool(score, torch.cat([zero, counts.cumsum(0)]), torch.cat([zero, k.cumsum(0)]), int(counts.max()), int(k.sum()))
    order = torch.sort(score, descending=True, stable=True)[1]
    order = order[torch.sort(b.batch[order], stable=True)[1]]
    start = counts.cumsum(0) - counts
    rank = torch.arange(n, device=dev) - start[b.batch[order]]
    perm_ref = order[rank < k[b.batch[order]]]
    assert torch.equal(perm, perm_ref)
    mask = perm_ref.new_full((n,), -1); mask[perm_ref] = torch.arange(perm_ref.numel(), device=dev)
    assert torch.equal(newid, mask)
    ei2, ew2 = gnn.filter_adj(b.edge_index, b.edge_attr, newid)
    row, col = mask[b.edge_index[0]], mask[b.edge_index[1]]
    keep = (row >= 0) & (col >= 0)
    assert torch.equal(ei2, torch.stack([row[keep], col[keep]])) and torch.equal(ew2, b.edge_attr[keep])
