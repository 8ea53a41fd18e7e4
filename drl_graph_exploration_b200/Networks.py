"""Drop-in for the reference's ``scripts/Networks.py``: the same nine ``nn.Module`` classes, the
same constructor arguments, the same ``forward(data, prob_or_mask, batch=None)`` signatures and
the same parameter names / shapes as the shipped state dicts (``data/torch_weights/*/MyModel.pt``:
``conv1.weight (5,1000)``, ``gconv1.rnn.weight_ih``, ``down_convs.0.weight``, ``pools.0.weight`` ...),
so ``model.load_state_dict(torch.load('MyModel.pt'))`` and ``train.py`` / ``test.py`` work unchanged.

What differs is underneath: PyG / torch_scatter / torch_sparse are gone.  Neighbour aggregation
(GCNConv.propagate, GatedGraphConv.propagate) runs on the hand-written gather kernel of
``csrc/dge_gnn.cu`` (forward and backward); the edge structure is preprocessed once per batch
into deterministic CSR (``gnn.GraphStructure``) and cached on the ``data`` object.  Dense X@W
products go to the tensor-core GEMM selected by ``set_matmul_precision``.  ``F.dropout`` stays
functional (= always active, quirk q19 of SURVEY.md) exactly like the reference.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import gnn

_PRECISION = "tc3"        # inference (autograd off)
_PRECISION_TRAIN = "tc3"   # under autograd


def set_matmul_precision(mode: str, train: str | None = None):
    """Dense node-MLP GEMMs.  ``mode`` (inference, autograd off): 'tc3' (default) = hand-written tcgen05 kernel, 3xTF32
    split, fp32 accumulation in tensor memory -- ~1e-6 of sum|a||w| per product, Q-values within ~1e-5 of fp64 (contract
    1e-4); 'fp32' = library SGEMM (A/B reference); 'tf32' / 'bf16' = library single-pass tensor-core modes (do not meet
    the contract).  ``train`` (under autograd; default 'tc3' too): forward, grad-input and grad-weight (x^T dy, contracted
    over the nodes with split-K) of every family's dense layers run on the tcgen05 kernel, the GG-NN's GRU cell on
    ``gnn.gru_cell_train``; 'fp32' = library SGEMM + torch GRUCell (A/B reference).  The products agree with fp64 to ~1e-6
    of sum|a||b|; what differs between ANY two fp32-grade implementations is on which side of a ReLU kink a pre-activation
    within that error of zero falls (tests/test_native_train_gpu.py compares kink by kink)."""
    global _PRECISION, _PRECISION_TRAIN
    assert mode in ("tc3", "fp32", "tf32", "bf16") and train in (None, "tc3", "fp32", "tf32", "bf16")
    _PRECISION = mode
    if train is not None:
        _PRECISION_TRAIN = train


def _mm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    mode = _PRECISION_TRAIN if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad) else _PRECISION
    return _mm_mode(a, b, mode)


def _mm_mode(a: torch.Tensor, b: torch.Tensor, _PRECISION: str) -> torch.Tensor:
    if _PRECISION == "tc3" and a.is_cuda and a.dim() == 2 and b.dim() == 2 and gnn.tc_supported(a.shape[1], b.shape[1]):
        return gnn.tc_matmul(a, b)
    if _PRECISION == "bf16":
        return (a.bfloat16() @ b.bfloat16()).float()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = _PRECISION == "tf32"   # 'tc3' on an unsupported shape (K = 5 input layer): fp32
    try:
        return a @ b
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def _structure(data, num_nodes: int, edge_index=None, edge_weight=None) -> gnn.GraphStructure:
    """CSR of the batch, built once and cached on the data object (same batch is used by the
    target-net forward, the policy forward and its backward in DeepQ.train)."""
    if edge_index is None:
        gs = getattr(data, "_dge_structure", None)
        if gs is None or gs.N != num_nodes or gs.src.device != data.edge_index.device:
            gs = gnn.GraphStructure(data.edge_index, data.edge_attr, num_nodes)
            try:
                data._dge_structure = gs
            except Exception:
                pass
        return gs
    return gnn.GraphStructure(edge_index, edge_weight, num_nodes)


class GCNConv(torch.nn.Module):
    """PyG-1.x ``GCNConv(in, out, improved=True)`` parameter layout: ``weight (in,out)``, ``bias (out)``."""

    def __init__(self, in_channels, out_channels, improved=True):
        super().__init__()
        self.in_channels, self.out_channels, self.improved = in_channels, out_channels, improved
        self.weight = torch.nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = torch.nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        bound = math.sqrt(6.0 / (self.in_channels + self.out_channels))   # glorot
        torch.nn.init.uniform_(self.weight, -bound, bound)
        torch.nn.init.zeros_(self.bias)

    def forward(self, x, gs: gnn.GraphStructure, relu: bool = False):
        return gnn.gcn_aggregate(_mm(x, self.weight), self.bias, gs, self.improved, relu)


class GatedGraphConv(torch.nn.Module):
    """PyG-1.x ``GatedGraphConv(out_channels, num_layers)``: ``weight (L,C,C)`` + ``rnn`` = GRUCell(C,C)."""

    def __init__(self, out_channels, num_layers):
        super().__init__()
        self.out_channels, self.num_layers = out_channels, num_layers
        self.weight = torch.nn.Parameter(torch.empty(num_layers, out_channels, out_channels))
        self.rnn = torch.nn.GRUCell(out_channels, out_channels)
        bound = 1.0 / math.sqrt(out_channels)
        torch.nn.init.uniform_(self.weight, -bound, bound)

    def forward(self, x, gs: gnn.GraphStructure):
        h = x if x.size(1) == self.out_channels else torch.cat([x, x.new_zeros(x.size(0), self.out_channels - x.size(1))], dim=1)
        ok = h.is_cuda and self.out_channels % 4 == 0 and self.rnn.bias
        native = ok and not torch.is_grad_enabled() and _PRECISION == "tc3"
        native_train = ok and torch.is_grad_enabled() and _PRECISION_TRAIN == "tc3"
        for i in range(self.num_layers):
            m = gnn.weighted_aggregate(_mm(h, self.weight[i]), gs)
            # the GRU's two transforms on the tcgen05 GEMM + one gate kernel; under autograd the same with their backward
            # (gnn.gru_cell_train); torch's GRUCell only in the 'fp32' A/B mode
            h = gnn.gru_cell_inference(m, h, self.rnn) if native else (gnn.gru_cell_train(m, h, self.rnn) if native_train else self.rnn(m, h))
        return h


class TopKPooling(torch.nn.Module):
    """PyG-1.x ``TopKPooling(in_channels, ratio)``: ``weight (1,C)``; score = tanh(x.w/||w||)."""

    def __init__(self, in_channels, ratio=0.5):
        super().__init__()
        self.in_channels, self.ratio = in_channels, ratio
        self.weight = torch.nn.Parameter(torch.empty(1, in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.in_channels)
        torch.nn.init.uniform_(self.weight, -bound, bound)

    def forward(self, x, edge_index, edge_attr, batch):
        gnn.require_cuda(x, "TopKPooling")
        score = torch.tanh((x * self.weight).sum(dim=-1) / self.weight.norm(p=2, dim=-1))
        n_graphs = int(batch.max()) + 1 if batch.numel() else 0
        counts = torch.bincount(batch, minlength=n_graphs)
        k = torch.ceil(self.ratio * counts.to(score.dtype)).long()
        if n_graphs > 0 and not edge_attr.requires_grad:
            max_n, n_kept = (int(v) for v in torch.stack([counts.max(), k.sum()]).tolist())
            if max_n <= 1024:   # native: one CTA per graph sorts its scores; edges filtered + relabelled by a flag / scan / fill pass
                zero = counts.new_zeros(1)
                perm, newid = gnn.topk_pool(score.detach(), torch.cat([zero, torch.cumsum(counts, 0)]), torch.cat([zero, torch.cumsum(k, 0)]), max_n, n_kept)
                ei2, ew2 = gnn.filter_adj(edge_index, edge_attr, newid)
                return x[perm] * score[perm].view(-1, 1), ei2, ew2, batch[perm], perm, score[perm]
        perm, ei2, ew2 = _topk_select_sorted(score, counts, k, edge_index, edge_attr, batch)
        return x[perm] * score[perm].view(-1, 1), ei2, ew2, batch[perm], perm, score[perm]


def _topk_select_sorted(score, counts, k, edge_index, edge_attr, batch):
    """Device-side torch formulation of the selection for the two shapes the sort kernel does not take (a graph above 1024
    nodes -- the reference's graphs stay below ~250 -- or edge weights that carry a gradient): per-graph descending order by a
    stable two-key sort, ties to the lower index, then ``filter_adj``.  ``TopKPooling.forward`` only reaches it with CUDA
    tensors; ``tests/test_networks_glue_cpu.py`` holds it against the restated PyG semantics."""
    order = torch.sort(score, descending=True, stable=True)[1]
    order = order[torch.sort(batch[order], stable=True)[1]]
    start = torch.cumsum(counts, 0) - counts
    rank = torch.arange(order.numel(), device=score.device) - start[batch[order]]
    perm = order[rank < k[batch[order]]]
    mask = perm.new_full((score.size(0),), -1)
    mask[perm] = torch.arange(perm.numel(), device=perm.device)
    row, col = mask[edge_index[0]], mask[edge_index[1]]
    keep = (row >= 0) & (col >= 0)
    return perm, torch.stack([row[keep], col[keep]]), edge_attr[keep]


def _augment_adj(edge_index, edge_weight, num_nodes, batch=None, n_graphs=None):
    """Networks.py:216-225: add_self_loops -> sort -> spspmm(A, A) (coalesced) -> remove_self_loops.  Inference and training
    alike (the edge weights carry no gradient) the block-diagonal batch goes through the native row-accumulator kernels
    (gnn.augment_adj).  CPU tensors raise; ``_augment_adj_sparse`` takes the shapes the kernel does not (see there)."""
    gnn.require_cuda(edge_index, "augment_adj")
    if not edge_weight.requires_grad:
        if batch is None:
            gptr, mx, bt = torch.tensor([0, num_nodes], device=edge_index.device), num_nodes, None
        else:
            G = int(n_graphs) if n_graphs is not None else (int(batch.max()) + 1 if batch.numel() else 0)
            counts = torch.bincount(batch, minlength=G)
            gptr = torch.cat([counts.new_zeros(1), torch.cumsum(counts, 0)])
            mx, bt = (int(counts.max()) if G else 0), batch
        out = gnn.augment_adj(edge_index, edge_weight.float(), num_nodes, bt, gptr, mx) if num_nodes > 0 else None
        if out is not None:
            return out
    return _augment_adj_sparse(edge_index, edge_weight, num_nodes)


def _augment_adj_sparse(edge_index, edge_weight, num_nodes):
    """(A + I)^2 without its diagonal through torch.sparse on the device, for a graph above 1024 nodes (the shared-memory row
    accumulator of ``k_augment_adj`` is sized for that) or edge weights that carry a gradient.  Same structure and order as the
    kernel's output; ``tests/test_networks_glue_cpu.py`` holds it against the restated PyG semantics."""
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    ei = torch.cat([edge_index, torch.stack([loop, loop])], dim=1)
    ew = torch.cat([edge_weight, edge_weight.new_ones(num_nodes)])
    A = torch.sparse_coo_tensor(ei, ew, (num_nodes, num_nodes)).coalesce()
    A2 = torch.sparse.mm(A, A).coalesce()
    ei2, ew2 = A2.indices(), A2.values()
    keep = ei2[0] != ei2[1]
    return ei2[:, keep], ew2[keep]


def _segment_softmax(src, index):
    n = int(index.max()) + 1
    mx = torch.full((n,), -float("inf"), dtype=src.dtype, device=src.device).scatter_reduce(0, index, src, "amax")
    out = (src - mx[index]).exp()
    den = torch.zeros(n, dtype=src.dtype, device=src.device).index_add_(0, index, out)
    return out / (den[index] + 1e-16)


def _global_mean_pool(x, batch):
    n = int(batch.max()) + 1
    s = torch.zeros(n, x.size(1), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    cnt = torch.bincount(batch, minlength=n).clamp(min=1).to(x.dtype)
    return s / cnt.view(-1, 1)


# ======================================================================== GCN family ===
class GCN(torch.nn.Module):
    """Networks.py:12-28 (DQN Q-network): 2 x GCNConv(improved) + ReLU, functional dropout(p=prob), Linear(1000,1)."""
    _out = 1

    def __init__(self):
        super().__init__()
        self.conv1 = GCNConv(5, 1000, improved=True)
        self.conv2 = GCNConv(1000, 1000, improved=True)
        self.fully_con1 = torch.nn.Linear(1000, self._out)

    def _trunk(self, data, p):
        x = data.x
        native_q = (self._out == 1 and not torch.is_grad_enabled() and (p == 0 or p == 0.0) and _PRECISION == "tc3" and x.is_cuda and x.size(1) <= 8
                    and self.conv2.out_channels % 4 == 0 and self.conv2.out_channels <= 1024)
        if native_q and getattr(data, "_dge_structure", None) is None:
            # inference on a batch that arrives as a raw edge list (DeepQ.test): CSRs + normalisation + forward in ONE native call
            return gnn.gcn_q_forward_coo(x, data.edge_index, data.edge_attr, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                                         self.fully_con1.weight[0], self.fully_con1.bias).view(-1, 1)
        gs = _structure(data, x.size(0))
        if native_q:
            # inference: the whole Q-network in one native call (three launches)
            return gnn.gcn_q_forward(x, gs, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                                     self.fully_con1.weight[0], self.fully_con1.bias).view(-1, 1)
        if not torch.is_grad_enabled() and x.size(1) <= 8:   # inference: aggregate the 5 input channels first, transform in the epilogue
            x = gnn.gcn_conv_small_fused(x, self.conv1.weight, self.conv1.bias, gs, improved=True, relu=True)
        else:
            x = self.conv1(x, gs, relu=True)
        fused = (self._out == 1 and not torch.is_grad_enabled() and (p == 0 or p == 0.0))
        if fused:   # inference: conv2 aggregate + ReLU + Linear(1000,1) in one kernel
            # (the head bias is added as a tensor op: reading it on the host would force a sync)
            q = gnn.gcn_aggregate_head(_mm(x, self.conv2.weight), self.conv2.bias, gs, self.fully_con1.weight[0], 0.0)
            return q.view(-1, 1) + self.fully_con1.bias
        x = self.conv2(x, gs, relu=True)
        x = F.dropout(x, p=p)
        return self.fully_con1(x)

    def forward(self, data, prob, batch=None):
        return self._trunk(data, prob)


class PolicyGCN(GCN):
    """Networks.py:31-49"""

    def forward(self, data, mask, batch=None):
        x = self._trunk(data, 0.5)
        x = torch.masked_select(x.view(-1), mask)
        return _segment_softmax(x, torch.masked_select(batch, mask))


class ValueGCN(GCN):
    """Networks.py:52-70"""
    _out = 100

    def forward(self, data, mask, batch=None):
        return _global_mean_pool(self._trunk(data, 0.5), batch).mean(dim=1)


# ======================================================================= GG-NN family ===
class GGNN(torch.nn.Module):
    """Networks.py:73-86"""
    _out = 1

    def __init__(self):
        super().__init__()
        self.gconv1 = GatedGraphConv(1000, 3)
        self.fully_con1 = torch.nn.Linear(1000, self._out)

    def _trunk(self, data, p):
        gs = _structure(data, data.x.size(0))
        x = F.relu(self.gconv1(data.x, gs))
        return self.fully_con1(F.dropout(x, p=p))

    def forward(self, data, prob, batch=None):
        return self._trunk(data, prob)


class PolicyGGNN(GGNN):
    """Networks.py:89-104"""

    def forward(self, data, mask, batch=None):
        x = torch.masked_select(self._trunk(data, 0.5).view(-1), mask)
        return _segment_softmax(x, torch.masked_select(batch, mask))


class ValueGGNN(GGNN):
    """Networks.py:107-122"""
    _out = 100

    def forward(self, data, mask, batch=None):
        return _global_mean_pool(self._trunk(data, 0.5), batch).mean(dim=1)


# ===================================================================== g-U-Net family ===
class GraphUNet(torch.nn.Module):
    """Networks.py:125-230 (built as GraphUNet(5, 1000, 1000, depth=3), train.py:44)."""
    _out = 1

    def __init__(self, in_channels, hidden_channels, out_channels, depth, pool_ratios=0.5, sum_res=True, act=F.relu):
        super().__init__()
        assert depth >= 1 and sum_res
        self.in_channels, self.hidden_channels, self.out_channels, self.depth = in_channels, hidden_channels, out_channels, depth
        self.pool_ratios = [pool_ratios] * depth if not isinstance(pool_ratios, (list, tuple)) else list(pool_ratios)
        self.act, self.sum_res = act, sum_res
        self.down_convs = torch.nn.ModuleList([GCNConv(in_channels, hidden_channels, improved=True)])
        self.pools = torch.nn.ModuleList()
        for i in range(depth):
            self.pools.append(TopKPooling(hidden_channels, self.pool_ratios[i]))
            self.down_convs.append(GCNConv(hidden_channels, hidden_channels, improved=True))
        self.up_convs = torch.nn.ModuleList([GCNConv(hidden_channels, hidden_channels, improved=True) for _ in range(depth - 1)])
        self.up_convs.append(GCNConv(hidden_channels, out_channels, improved=True))
        self.fully_con1 = torch.nn.Linear(out_channels, self._out)

    def reset_parameters(self):
        for m in list(self.down_convs) + list(self.pools) + list(self.up_convs):
            m.reset_parameters()

    def _trunk(self, data, p, batch):
        x, ei, ew = data.x, data.edge_index, data.edge_attr
        if batch is None:
            batch = ei.new_zeros(x.size(0))
        gs = _structure(data, x.size(0))
        x = self.act(self.down_convs[0](x, gs))
        xs, gss, perms = [x], [gs], []
        for i in range(1, self.depth + 1):
            ei, ew = _augment_adj(ei, ew, x.size(0), batch)
            x, ei, ew, batch, perm, _ = self.pools[i - 1](x, ei, ew, batch)
            gs = gnn.GraphStructure(ei, ew, x.size(0))
            x = self.act(self.down_convs[i](x, gs))
            if i < self.depth:
                xs.append(x); gss.append(gs)
            perms.append(perm)
        for i in range(self.depth):
            j = self.depth - 1 - i
            up = torch.zeros_like(xs[j])
            up[perms[j]] = x
            x = self.up_convs[i](xs[j] + up, gss[j])
            if i < self.depth - 1:
                x = self.act(x)
        self.last_perms = perms        # nodes kept by the pooling levels of this pass (kept for inspection / parity tests)
        x = F.dropout(F.relu(x), p=p)
        return self.fully_con1(x)

    def forward(self, data, prob, batch=None):
        return self._trunk(data, prob, batch)

    def __repr__(self):
        return "{}({}, {}, {}, depth={}, pool_ratios={})".format(self.__class__.__name__, self.in_channels, self.hidden_channels,
                                                                 self.out_channels, self.depth, self.pool_ratios)


class PolicyGraphUNet(GraphUNet):
    """Networks.py:233-339"""

    def forward(self, data, mask, batch=None):
        if batch is None:
            batch = data.edge_index.new_zeros(data.x.size(0))
        x = torch.masked_select(self._trunk(data, 0.5, batch).view(-1), mask)
        return _segment_softmax(x, torch.masked_select(batch, mask))


class ValueGraphUNet(GraphUNet):
    """Networks.py:342-449"""
    _out = 100

    def forward(self, data, mask, batch=None):
        if batch is None:
            batch = data.edge_index.new_zeros(data.x.size(0))
        return _global_mean_pool(self._trunk(data, 0.5, batch), batch).mean(dim=1)
