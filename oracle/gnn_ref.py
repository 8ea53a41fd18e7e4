"""Pure-PyTorch restatement of the PyG-1.x layers the reference's ``scripts/Networks.py``
uses (GCNConv(improved=True), GatedGraphConv, TopKPooling, spspmm-based augment_adj,
segment softmax, global_mean_pool) and of the nine network classes built from them.

TEST INFRASTRUCTURE ONLY (the oracle for the GNN rows a16/a17 of SURVEY section 8).
PyTorch-Geometric / torch_scatter / torch_sparse are not installable here and are unpinned
in the reference (no requirements file) => **parity unpinned**; the shipped state dicts
(``data/torch_weights/*/MyModel.pt``) pin parameter names and shapes, and the golden CSVs
(``data/test_result/*_DQN_GCN.csv``) pin the DQN+GCN policy's choices end to end over 6218 rows
(tests/test_oracle_cpu.py); GG-NN and g-U-Net stay unpinned (the A2C result files are stochastic: dropout is always on).  All ops run in whatever dtype ``x`` has (fp32 in the
reference; tests also use fp64 to bound the fp32 error of the CUDA kernels).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------- functional ---
def add_remaining_self_loops(edge_index, edge_weight, fill_value, num_nodes):
    """PyG utils.add_remaining_self_loops: keep existing self-loop weights, add ``fill_value``
    loops for the other nodes; non-loop edges first, then one loop per node."""
    row, col = edge_index
    mask = row != col
    loop_weight = edge_weight.new_full((num_nodes,), fill_value)
    inv = ~mask
    if inv.any():
        loop_weight[row[inv]] = edge_weight[inv]
    loop_index = torch.arange(num_nodes, dtype=row.dtype, device=row.device)
    ei = torch.cat([edge_index[:, mask], torch.stack([loop_index, loop_index])], dim=1)
    ew = torch.cat([edge_weight[mask], loop_weight])
    return ei, ew


def gcn_norm(edge_index, edge_weight, num_nodes, improved=True):
    """GCNConv.norm (PyG 1.4-1.6): deg over ``row``; D^-1/2 (A + fill I) D^-1/2."""
    fill = 2.0 if improved else 1.0
    ei, ew = add_remaining_self_loops(edge_index, edge_weight, fill, num_nodes)
    row, col = ei
    deg = torch.zeros(num_nodes, dtype=ew.dtype, device=ew.device).index_add_(0, row, ew)
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0
    return ei, dis[row] * ew * dis[col]


def gcn_conv(x, edge_index, edge_weight, weight, bias, improved=True):
    """Networks.py:15-16,22-24 -> GCNConv.forward: X W, then sum_j norm_ij (XW)_j at the
    target node (flow source_to_target: message from edge_index[0] aggregated at edge_index[1])."""
    n = x.size(0)
    ei, norm = gcn_norm(edge_index, edge_weight.to(x.dtype), n, improved)
    xw = x @ weight
    out = torch.zeros_like(xw).index_add_(0, ei[1], norm.view(-1, 1) * xw[ei[0]])
    return out + bias


def gated_graph_conv(x, edge_index, edge_weight, weight, w_ih, w_hh, b_ih, b_hh):
    """Networks.py:76,82 -> GatedGraphConv(1000, 3): zero-pad, then per layer
    m = scatter_add(w * (h W_i)[src] -> dst), h = GRUCell(m, h)."""
    c = weight.size(-1)
    h = x if x.size(1) == c else torch.cat([x, x.new_zeros(x.size(0), c - x.size(1))], dim=1)
    for i in range(weight.size(0)):
        m = h @ weight[i]
        m = torch.zeros_like(m).index_add_(0, edge_index[1], edge_weight.to(m.dtype).view(-1, 1) * m[edge_index[0]])
        h = torch.gru_cell(m, h, w_ih, w_hh, b_ih, b_hh)
    return h


def segment_softmax(src, index, num_segments=None):
    """torch_geometric.utils.softmax (1.x): max-shifted exp / (sum + 1e-16)."""
    n = int(index.max()) + 1 if num_segments is None else num_segments
    mx = torch.full((n,), -float("inf"), dtype=src.dtype, device=src.device).scatter_reduce(0, index, src, "amax")
    out = (src - mx[index]).exp()
    den = torch.zeros(n, dtype=src.dtype, device=src.device).index_add_(0, index, out)
    return out / (den[index] + 1e-16)


def global_mean_pool(x, batch, num_graphs=None):
    n = int(batch.max()) + 1 if num_graphs is None else num_graphs
    s = torch.zeros(n, x.size(1), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    cnt = torch.zeros(n, dtype=x.dtype, device=x.device).index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return s / cnt.clamp(min=1).view(-1, 1)


def topk_perm(score, ratio, batch):
    """torch_geometric.nn.pool.topk: per graph keep ceil(ratio*n) nodes, descending score."""
    perms = []
    for g in range(int(batch.max()) + 1):
        idx = (batch == g).nonzero().view(-1)
        k = int(math.ceil(ratio * idx.numel()))
        order = torch.sort(score[idx], descending=True, stable=True)[1][:k]
        perms.append(idx[order])
    return torch.cat(perms)


def filter_adj(edge_index, edge_attr, perm, num_nodes):
    mask = perm.new_full((num_nodes,), -1)
    mask[perm] = torch.arange(perm.numel(), device=perm.device)
    row, col = mask[edge_index[0]], mask[edge_index[1]]
    keep = (row >= 0) & (col >= 0)
    return torch.stack([row[keep], col[keep]]), edge_attr[keep]


def topk_pool(x, edge_index, edge_attr, batch, weight, ratio=0.5, forced_perm=None, record=None):
    """TopKPooling(1000, 0.5).forward: score = tanh(x.w / ||w||).  ``forced_perm`` (parity tests): pool THESE nodes instead of the
    fp64 top-k -- a near-tie of two scores may legitimately be ordered differently in fp32 -- and ``record`` receives
    (score, own perm) so the test can check that the forced selection is a top-k selection up to such ties."""
    score = (x * weight).sum(dim=-1)
    score = torch.tanh(score / weight.norm(p=2, dim=-1))
    perm = topk_perm(score, ratio, batch)
    if record is not None:
        record.append((score.detach(), perm, batch))
    if forced_perm is not None:
        perm = forced_perm
    x = x[perm] * score[perm].view(-1, 1)
    batch = batch[perm]
    ei, ea = filter_adj(edge_index, edge_attr, perm, score.size(0))
    return x, ei, ea, batch, perm


def augment_adj(edge_index, edge_weight, num_nodes):
    """Networks.py:216-225: add_self_loops -> sort -> spspmm (A.A, coalesced) -> remove_self_loops."""
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    ei = torch.cat([edge_index, torch.stack([loop, loop])], dim=1)
    ew = torch.cat([edge_weight, edge_weight.new_ones(num_nodes)])
    A = torch.sparse_coo_tensor(ei, ew, (num_nodes, num_nodes)).coalesce()
    A2 = torch.sparse.mm(A, A).coalesce()
    ei2, ew2 = A2.indices(), A2.values()
    keep = ei2[0] != ei2[1]
    return ei2[:, keep], ew2[keep]


# --------------------------------------------------------------------------- modules ---
class _GCNConvParams(torch.nn.Module):
    """Parameter holder with the PyG-1.x GCNConv layout: ``weight (in,out)``, ``bias (out)``."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(cin, cout))
        self.bias = torch.nn.Parameter(torch.zeros(cout))
        bound = math.sqrt(6.0 / (cin + cout))  # glorot
        torch.nn.init.uniform_(self.weight, -bound, bound)

    def forward(self, x, edge_index, edge_weight):
        return gcn_conv(x, edge_index, edge_weight, self.weight, self.bias, improved=True)


class _GatedParams(torch.nn.Module):
    def __init__(self, c, layers):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(layers, c, c).uniform_(-1 / math.sqrt(c), 1 / math.sqrt(c)))
        self.rnn = torch.nn.GRUCell(c, c)

    def forward(self, x, edge_index, edge_weight):
        return gated_graph_conv(x, edge_index, edge_weight, self.weight, self.rnn.weight_ih, self.rnn.weight_hh,
                                self.rnn.bias_ih, self.rnn.bias_hh)


class _PoolParams(torch.nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(1, c).uniform_(-1 / math.sqrt(c), 1 / math.sqrt(c)))


def _dropout(x, p, mask):
    """F.dropout is functional in the reference => always active (q19).  ``mask`` (same
    shape as x, 0/1 keep mask) makes the draw explicit for parity tests."""
    if mask is not None:
        return x * mask / (1.0 - p)
    return F.dropout(x, p=p)


class GCN(torch.nn.Module):            # Networks.py:12-28
    def __init__(self, out=1):
        super().__init__()
        self.conv1 = _GCNConvParams(5, 1000)
        self.conv2 = _GCNConvParams(1000, 1000)
        self.fully_con1 = torch.nn.Linear(1000, out)

    def trunk(self, data, p, mask=None):
        x, ei, ew = data.x, data.edge_index, data.edge_attr
        x = F.relu(self.conv1(x, ei, ew))
        x = F.relu(self.conv2(x, ei, ew))
        x = _dropout(x, p, mask)
        return self.fully_con1(x)

    def forward(self, data, prob, batch=None, dropout_mask=None):
        return self.trunk(data, prob, dropout_mask)


class PolicyGCN(GCN):                   # Networks.py:31-49
    def forward(self, data, mask, batch=None, dropout_mask=None):
        x = self.trunk(data, 0.5, dropout_mask)
        x = torch.masked_select(x.view(-1), mask)
        return segment_softmax(x, torch.masked_select(batch, mask))


class ValueGCN(GCN):                    # Networks.py:52-70
    def __init__(self):
        super().__init__(out=100)

    def forward(self, data, mask, batch=None, dropout_mask=None):
        return global_mean_pool(self.trunk(data, 0.5, dropout_mask), batch).mean(dim=1)


class GGNN(torch.nn.Module):           # Networks.py:73-86
    def __init__(self, out=1):
        super().__init__()
        self.gconv1 = _GatedParams(1000, 3)
        self.fully_con1 = torch.nn.Linear(1000, out)

    def trunk(self, data, p, mask=None):
        x = F.relu(self.gconv1(data.x, data.edge_index, data.edge_attr))
        return self.fully_con1(_dropout(x, p, mask))

    def forward(self, data, prob, batch=None, dropout_mask=None):
        return self.trunk(data, prob, dropout_mask)


class PolicyGGNN(GGNN):                 # Networks.py:89-104
    def forward(self, data, mask, batch=None, dropout_mask=None):
        x = torch.masked_select(self.trunk(data, 0.5, dropout_mask).view(-1), mask)
        return segment_softmax(x, torch.masked_select(batch, mask))


class ValueGGNN(GGNN):                  # Networks.py:107-122
    def __init__(self):
        super().__init__(out=100)

    def forward(self, data, mask, batch=None, dropout_mask=None):
        return global_mean_pool(self.trunk(data, 0.5, dropout_mask), batch).mean(dim=1)


class GraphUNet(torch.nn.Module):      # Networks.py:125-230
    def __init__(self, in_channels=5, hidden_channels=1000, out_channels=1000, depth=3, pool_ratios=0.5, out=1):
        super().__init__()
        self.depth, self.ratio = depth, pool_ratios
        self.down_convs = torch.nn.ModuleList([_GCNConvParams(in_channels, hidden_channels)])
        self.pools = torch.nn.ModuleList()
        for _ in range(depth):
            self.pools.append(_PoolParams(hidden_channels))
            self.down_convs.append(_GCNConvParams(hidden_channels, hidden_channels))
        self.up_convs = torch.nn.ModuleList([_GCNConvParams(hidden_channels, hidden_channels) for _ in range(depth - 1)])
        self.up_convs.append(_GCNConvParams(hidden_channels, out_channels))
        self.fully_con1 = torch.nn.Linear(out_channels, out)

    def trunk(self, data, p, batch=None, mask=None, forced_perms=None, record=None):
        x, ei, ew = data.x, data.edge_index, data.edge_attr
        if batch is None:
            batch = ei.new_zeros(x.size(0))
        x = F.relu(self.down_convs[0](x, ei, ew))
        xs, eis, ews, perms = [x], [ei], [ew], []
        for i in range(1, self.depth + 1):
            ei, ew = augment_adj(ei, ew, x.size(0))
            x, ei, ew, batch, perm = topk_pool(x, ei, ew, batch, self.pools[i - 1].weight, self.ratio,
                                               None if forced_perms is None else forced_perms[i - 1], record)
            x = F.relu(self.down_convs[i](x, ei, ew))
            if i < self.depth:
                xs.append(x); eis.append(ei); ews.append(ew)
            perms.append(perm)
        for i in range(self.depth):
            j = self.depth - 1 - i
            up = torch.zeros_like(xs[j])
            up[perms[j]] = x
            x = self.up_convs[i](xs[j] + up, eis[j], ews[j])
            if i < self.depth - 1:
                x = F.relu(x)
        x = _dropout(F.relu(x), p, mask)
        return self.fully_con1(x)

    def forward(self, data, prob, batch=None, dropout_mask=None, forced_perms=None, record=None):
        return self.trunk(data, prob, batch, dropout_mask, forced_perms, record)


class PolicyGraphUNet(GraphUNet):       # Networks.py:233-339
    def forward(self, data, mask, batch=None, dropout_mask=None):
        if batch is None:
            batch = data.edge_index.new_zeros(data.x.size(0))
        x = torch.masked_select(self.trunk(data, 0.5, batch, dropout_mask).view(-1), mask)
        return segment_softmax(x, torch.masked_select(batch, mask))


class ValueGraphUNet(GraphUNet):        # Networks.py:342-449
    def __init__(self, *a, **k):
        super().__init__(*a, out=100, **k)

    def forward(self, data, mask, batch=None, dropout_mask=None):
        if batch is None:
            batch = data.edge_index.new_zeros(data.x.size(0))
        return global_mean_pool(self.trunk(data, 0.5, batch, dropout_mask), batch).mean(dim=1)


class Graph:
    """Stand-in for torch_geometric.data.Data (x, edge_index, edge_attr)."""

    def __init__(self, x, edge_index, edge_attr):
        self.x, self.edge_index, self.edge_attr = x, edge_index, edge_attr
