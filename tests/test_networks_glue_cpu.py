"""The PyTorch glue of Networks.py that needs no kernel (the torch formulations of top-k selection / augment_adj for shapes
the kernels do not take, segment softmax, mean pool) against the oracle's restatement of the PyG-1.x functions, on CPU tensors.
(The CUDA paths of the same functions are compared with these in the -m gpu tests; the layer entry points refuse CPU tensors.)"""
import numpy as np
import pytest
import torch

from drl_graph_exploration_b200 import Networks
from drl_graph_exploration_b200.engine import DgeError
from oracle import gnn_ref


def _batch(rng, sizes):
    xs, eis, ws, bt, off = [], [], [], [], 0
    for g, n in enumerate(sizes):
        e = [(i, i + 1) for i in range(n - 1)] + [(int(rng.integers(0, n)), int(rng.integers(0, n))) for _ in range(n)]
        e = sorted(set((min(a, b), max(a, b)) for a, b in e if a != b))
        w = rng.uniform(0.1, 5.0, len(e))
        src = [a for a, b in e] + [b for a, b in e]; dst = [b for a, b in e] + [a for a, b in e]
        eis.append(np.array([src, dst]) + off); ws.append(np.concatenate([w, w]))
        xs.append(rng.normal(size=(n, 6))); bt.append(np.full(n, g)); off += n
    return (torch.tensor(np.concatenate(xs), dtype=torch.float32), torch.tensor(np.concatenate(eis, axis=1), dtype=torch.long),
            torch.tensor(np.concatenate(ws), dtype=torch.float32), torch.tensor(np.concatenate(bt), dtype=torch.long))


def test_segment_softmax_and_mean_pool():
    rng = np.random.default_rng(0)
    src = torch.tensor(rng.normal(size=40), dtype=torch.float32)
    idx = torch.tensor(np.sort(rng.integers(0, 6, 40)), dtype=torch.long)
    idx[-1] = 5
    out = Networks._segment_softmax(src, idx)
    assert torch.allclose(out, gnn_ref.segment_softmax(src, idx), rtol=1e-6, atol=1e-8)
    sums = torch.zeros(6).index_add_(0, idx, out)
    present = torch.bincount(idx, minlength=6) > 0
    assert torch.allclose(sums[present], torch.ones(int(present.sum())), atol=1e-5)      # softmax(src, index): sums to 1 per segment
    x = torch.tensor(rng.normal(size=(40, 7)), dtype=torch.float32)
    assert torch.allclose(Networks._global_mean_pool(x, idx), gnn_ref.global_mean_pool(x, idx), rtol=1e-6, atol=1e-7)


def test_topk_pooling_and_augment_adj_generic_paths():
    """The torch formulations TopKPooling / _augment_adj use on the device for the shapes the kernels do not take (graphs above
    1024 nodes, edge weights with a gradient), against the oracle's PyG restatement; the entry points themselves refuse CPU
    tensors (no CPU fallback in the product path)."""
    rng = np.random.default_rng(1)
    x, ei, w, bt = _batch(rng, [9, 4, 16, 1, 7])
    pool = Networks.TopKPooling(6, ratio=0.5)
    with pytest.raises(DgeError):
        pool(x, ei, w, bt)
    with pytest.raises(DgeError):
        Networks._augment_adj(ei, w, x.size(0))
    with torch.no_grad():
        score = torch.tanh((x * pool.weight).sum(dim=-1) / pool.weight.norm(p=2, dim=-1))
    counts = torch.bincount(bt)
    k = torch.ceil(0.5 * counts.float()).long()
    perm, eio, wo = Networks._topk_select_sorted(score, counts, k, ei, w, bt)
    xo, bo = x[perm] * score[perm].view(-1, 1), bt[perm]
    xr, eir, wr, br, permr = gnn_ref.topk_pool(x, ei, w, bt, pool.weight.detach(), 0.5)
    assert torch.equal(perm, permr) and torch.equal(bo, br)
    assert torch.allclose(xo, xr, rtol=1e-6, atol=1e-7) and torch.equal(eio, eir) and torch.equal(wo, wr)
    assert torch.equal(torch.bincount(bo, minlength=5), k)                                               # k = ceil(ratio n) per graph
    a_i, a_w = Networks._augment_adj_sparse(ei, w, x.size(0))
    r_i, r_w = gnn_ref.augment_adj(ei, w, x.size(0))
    assert torch.equal(a_i, r_i) and torch.allclose(a_w, r_w, rtol=1e-6, atol=1e-7)
    assert bool((a_i[0] != a_i[1]).all()) and bool((bt[a_i[0]] == bt[a_i[1]]).all())
    # (A + I)^2 contains A: every original edge is still there
    key = lambda e: set((e[0] * 1000 + e[1]).tolist())
    assert key(ei) <= key(a_i)
