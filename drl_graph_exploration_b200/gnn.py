"""Message-passing primitives of the graph Q-network on the hand-written CUDA kernels of
``csrc/dge_gnn.cu`` (C ABI: include/dge_gnn.h).

``GraphStructure`` turns a PyG-style ``edge_index``/``edge_attr`` pair into deterministic
destination- and source-sorted CSR once per batch; ``gcn_aggregate`` / ``weighted_aggregate``
are autograd functions whose forward *and* backward are the gather kernel (the backward of a
weighted aggregation is the same gather over the transposed CSR, with the ReLU mask fused).
There is no PyTorch fallback: a missing ``libdge.so`` or a CPU tensor raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from .engine import DgeError, load_library

_vp = ctypes.c_void_p
_configured = False


def _lib():
    global _configured
    L = load_library()
    if not _configured:
        L.dge_gnn_csr_build.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]
        L.dge_gcn_norm.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, ctypes.c_float, _vp, _vp, _vp, _vp, _vp]
        L.dge_gnn_aggregate.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp,
                                        ctypes.c_float, _vp, _vp]
        _configured = True
    return L


def _p(t: Optional[torch.Tensor]):
    return None if t is None else _vp(t.data_ptr())


def _st(dev):
    return _vp(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise DgeError(f"{what}: the GNN kernels run on CUDA tensors only (no CPU fallback in the product path)")


launch_count = 0   # kernels launched through this module (bench.py reports it as gpu_launches)


class GraphStructure:
    """CSR views (by destination and by source, rows sorted by edge id) of one batched graph."""

    def __init__(self, edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor], num_nodes: int):
        global launch_count
        _need_cuda(edge_index, "GraphStructure")
        L = _lib()
        dev = edge_index.device
        self.N, self.E = int(num_nodes), int(edge_index.shape[1])
        ei = edge_index.contiguous()
        if ei.dtype != torch.int64:
            ei = ei.long()
        self.src, self.dst = ei[0].contiguous(), ei[1].contiguous()
        self.weight = (torch.ones(self.E, device=dev) if edge_weight is None else edge_weight.contiguous().float())
        self.rowptr_dst, self.perm_dst = self._build_csr(self.dst)
        self._src_csr = None
        self._gcn = {}

    def _build_csr(self, key):
        global launch_count
        L, dev = _lib(), key.device
        i32 = dict(dtype=torch.int32, device=dev)
        rp, pm = torch.empty(self.N + 1, **i32), torch.empty(max(self.E, 1), **i32)
        ws = torch.empty(2 * self.N + self.E, **i32)
        with torch.cuda.device(dev):
            rc = L.dge_gnn_csr_build(self.N, self.E, _p(key), _p(rp), _p(pm), _p(ws), _st(dev))
        if rc:
            raise DgeError(f"dge_gnn_csr_build failed ({rc})")
        launch_count += 5
        return rp, pm

    @classmethod
    def from_csr(cls, edge_index, edge_weight, num_nodes, rowptr_dst, perm_dst, gcn_norm=None, gcn_selfnorm=None):
        """Adopt a destination-sorted CSR (and, optionally, the improved-GCN normalisation) that was
        already built on the device (the engine's graph kernel does) -- zero preprocessing launches."""
        gs = cls.__new__(cls)
        gs.N, gs.E = int(num_nodes), int(edge_index.shape[1])
        gs.src, gs.dst = edge_index[0], edge_index[1]
        gs.weight = edge_weight
        gs.rowptr_dst, gs.perm_dst = rowptr_dst, perm_dst
        gs._src_csr = None
        gs._gcn = {} if gcn_norm is None else {True: (gcn_norm, gcn_selfnorm)}
        return gs

    @property
    def rowptr_src(self):
        return self._source_csr()[0]

    @property
    def perm_src(self):
        return self._source_csr()[1]

    def _source_csr(self):
        """Source-sorted CSR (transposed gather: backward pass, weighted degree) -- built on first use."""
        if self._src_csr is None:
            self._src_csr = self._build_csr(self.src.contiguous())
        return self._src_csr

    def gcn_norm(self, improved: bool = True):
        """(norm[E], selfnorm[N]) of GCNConv.norm -- cached per structure."""
        global launch_count
        if improved not in self._gcn:
            L = _lib()
            dev = self.src.device
            f32 = dict(dtype=torch.float32, device=dev)
            dis, selfw = torch.empty(self.N, **f32), torch.empty(self.N, **f32)
            norm, selfnorm = torch.empty(max(self.E, 1), **f32), torch.empty(self.N, **f32)
            with torch.cuda.device(dev):
                rc = L.dge_gcn_norm(self.N, self.E, _p(self.src), _p(self.dst), _p(self.weight), _p(self.rowptr_src), _p(self.perm_src),
                                    ctypes.c_float(2.0 if improved else 1.0), _p(dis), _p(selfw), _p(norm), _p(selfnorm), _st(dev))
            if rc:
                raise DgeError(f"dge_gcn_norm failed ({rc})")
            launch_count += 3
            self._gcn[improved] = (norm, selfnorm)
        return self._gcn[improved]


def _aggregate(gs: GraphStructure, X, transposed: bool, coef, selfcoef, bias, gate, relu: bool, want_out: bool = True,
               head_w=None, head_b: float = 0.0):
    global launch_count
    _need_cuda(X, "aggregate")
    L = _lib()
    X = X.contiguous()
    N, C = X.shape
    if N != gs.N:
        raise DgeError(f"aggregate: x has {N} rows, graph has {gs.N} nodes")
    out = torch.empty_like(X) if want_out else None
    q = torch.empty(N, dtype=torch.float32, device=X.device) if head_w is not None else None
    rowptr, perm, nbr = (gs.rowptr_src, gs.perm_src, gs.dst) if transposed else (gs.rowptr_dst, gs.perm_dst, gs.src)
    with torch.cuda.device(X.device):
        rc = L.dge_gnn_aggregate(N, C, _p(X), _p(rowptr), _p(perm), _p(nbr), _p(coef), _p(selfcoef), _p(bias),
                                 _p(None if gate is None else gate.contiguous()), int(relu), _p(out),
                                 _p(None if head_w is None else head_w.contiguous()), ctypes.c_float(head_b), _p(q), _st(X.device))
    if rc:
        raise DgeError(f"dge_gnn_aggregate failed ({rc})")
    launch_count += 1
    return out, q


class _AggregateFn(torch.autograd.Function):
    """y = act(bias + selfcoef * x + A_coef x); backward = the same gather over the transposed CSR."""

    @staticmethod
    def forward(ctx, x, bias, gs, coef, selfcoef, relu):
        out, _ = _aggregate(gs, x, False, coef, selfcoef, bias, None, relu)
        ctx.gs, ctx.coef, ctx.selfcoef, ctx.relu, ctx.has_bias = gs, coef, selfcoef, relu, bias is not None
        ctx.save_for_backward(out if relu else None)
        return out

    @staticmethod
    def backward(ctx, gout):
        (out,) = ctx.saved_tensors
        gout = gout.contiguous()
        gx, _ = _aggregate(ctx.gs, gout, True, ctx.coef, ctx.selfcoef, None, out if ctx.relu else None, False)
        gbias = None
        if ctx.has_bias:
            gbias = (torch.where(out > 0, gout, torch.zeros_like(gout)) if ctx.relu else gout).sum(dim=0)
        return gx, gbias, None, None, None, None


def gcn_aggregate(xw: torch.Tensor, bias: Optional[torch.Tensor], gs: GraphStructure, improved: bool = True, relu: bool = False):
    """GCNConv.propagate + bias (+ fused ReLU): D^-1/2 (A + fill I) D^-1/2 (XW) + b."""
    norm, selfnorm = gs.gcn_norm(improved)
    return _AggregateFn.apply(xw, bias, gs, norm, selfnorm, relu)


def gcn_aggregate_head(xw: torch.Tensor, bias, gs: GraphStructure, head_w: torch.Tensor, head_b: float, improved: bool = True):
    """Inference-only fusion of the last GCNConv + ReLU + Linear(C,1): returns q [N] without
    materialising the [N,C] hidden state."""
    norm, selfnorm = gs.gcn_norm(improved)
    _, q = _aggregate(gs, xw, False, norm, selfnorm, bias, None, True, want_out=False, head_w=head_w, head_b=head_b)
    return q


def gcn_conv_small_fused(x: torch.Tensor, weight: torch.Tensor, bias, gs: GraphStructure, improved: bool = True, relu: bool = True):
    """Inference-only first GCN layer: act(bias + (A_hat X) W) with the <= 8 input channels aggregated
    before the transform (one kernel, no GEMM launch, no [N,C] intermediate)."""
    global launch_count
    _need_cuda(x, "gcn_conv_small_fused")
    L = _lib()
    if not hasattr(L, "_conv_small_ready"):
        L.dge_gcn_conv_small.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp]
        L._conv_small_ready = True
    norm, selfnorm = gs.gcn_norm(improved)
    x = x.contiguous().float()
    N, cin = x.shape
    C = weight.shape[1]
    out = torch.empty(N, C, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.dge_gcn_conv_small(N, cin, C, _p(x), _p(gs.rowptr_dst), _p(gs.perm_dst), _p(gs.src), _p(norm), _p(selfnorm),
                                  _p(weight.contiguous()), _p(None if bias is None else bias.contiguous()), int(relu), _p(out), _st(x.device))
    if rc:
        raise DgeError(f"dge_gcn_conv_small failed ({rc})")
    launch_count += 1
    return out


def weighted_aggregate(x: torch.Tensor, gs: GraphStructure):
    """GatedGraphConv.propagate: sum_j w_ji x_j (no self loops, no bias)."""
    return _AggregateFn.apply(x, None, gs, gs.weight, None, False)
