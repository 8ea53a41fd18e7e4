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


require_cuda = _need_cuda   # for the glue in Networks.py


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
            gbias = colsum(torch.where(out > 0, gout, torch.zeros_like(gout)) if ctx.relu else gout)
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


# ------------------------------------------------------------------ tensor-core GEMM (tcgen05, 3xTF32) ---
_gemm_ready = False


def _gemm_lib():
    global _gemm_ready
    L = _lib()
    if not _gemm_ready:
        L.dge_gemm_split_tf32.argtypes = [ctypes.c_int64, _vp, _vp, _vp, _vp]
        L.dge_gemm_prep_weight.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp]
        L.dge_gemm_tf32x3.argtypes = [ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp]
        L.dge_gemm_tf32x3_ex.argtypes = [ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp]
        L.dge_gemm_split_transpose.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp]
        L.dge_gemm_tf32x3_tn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp]
        L.dge_gru_gates.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp]
        L.dge_gru_gates_bwd.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        L.dge_colsum_ws_floats.restype = ctypes.c_int64
        L.dge_colsum_ws_floats.argtypes = [ctypes.c_int]
        L.dge_colsum.argtypes = [ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp, _vp]
        _gemm_ready = True
    return L


def tc_supported(K: int, N: int) -> bool:
    return K % 4 == 0 and N % 4 == 0 and K >= 32


def split_tf32(x: torch.Tensor):
    """x -> (hi, lo) TF32 parts (dge_gemm_split_tf32)."""
    global launch_count
    L = _gemm_lib()
    x = x.contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = L.dge_gemm_split_tf32(x.numel(), _p(x), _p(hi), _p(lo), _st(x.device))
    if rc:
        raise DgeError(f"dge_gemm_split_tf32 failed ({rc})")
    launch_count += 1
    return hi, lo


def _weight_operand(w: torch.Tensor, transposed: bool):
    """(hi, lo) of the K-major B operand: ``transposed`` -> W^T [N,K] of W [K,N] (forward X @ W); else W itself as
    [N', K'] (backward dY @ W^T).  Split once per weight update: the result is cached ON the parameter object (on
    the base tensor for views such as ``weight[i]``), keyed by the autograd version counter."""
    global launch_count
    owner = w._base if w._base is not None else w
    cache = getattr(owner, "_dge_tf32_split", None)
    if cache is None:
        cache = {}
        try:
            owner._dge_tf32_split = cache
        except Exception:
            pass
    key = (w.storage_offset(), tuple(w.shape), tuple(w.stride()), transposed)
    hit = cache.get(key)
    if hit is not None and hit[0] == w._version:
        return hit[1], hit[2]
    L = _gemm_lib()
    wd = w.detach().contiguous()
    if transposed:
        K, N = wd.shape
        hi, lo = torch.empty(N, K, dtype=torch.float32, device=w.device), torch.empty(N, K, dtype=torch.float32, device=w.device)
        with torch.cuda.device(w.device):
            rc = L.dge_gemm_prep_weight(K, N, _p(wd), _p(hi), _p(lo), _st(w.device))
        if rc:
            raise DgeError(f"dge_gemm_prep_weight failed ({rc})")
        launch_count += 1
    else:
        hi, lo = split_tf32(wd)
    cache[key] = (w._version, hi, lo)
    return hi, lo


def _tc_gemm(a: torch.Tensor, bt_hi: torch.Tensor, bt_lo: torch.Tensor, m_dev: Optional[torch.Tensor] = None):
    """C [M,N] = a [M,K] @ Bt[N,K]^T on the tcgen05 kernel."""
    global launch_count
    _need_cuda(a, "tc_gemm")
    L = _gemm_lib()
    M, K = a.shape
    N = bt_hi.shape[0]
    ah, al = split_tf32(a)
    c = torch.empty(M, N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = L.dge_gemm_tf32x3(M, _p(m_dev), N, K, _p(ah), _p(al), _p(bt_hi), _p(bt_lo), _p(c), N, _st(a.device))
    if rc:
        raise DgeError(f"dge_gemm_tf32x3 failed ({rc})")
    launch_count += 1
    return c


def split_transpose(x: torch.Tensor, plain: bool = True):
    """x [M,C] -> ((hi, lo) [M,C] or (None, None), (thi, tlo) [C,Mp]) with Mp = M rounded up to 4 (dge_gemm_split_transpose): the
    row-major split feeds products that contract over C, the transposed one the products that contract over the rows (nodes)."""
    global launch_count
    _need_cuda(x, "split_transpose")
    L = _gemm_lib()
    x = x.contiguous().float()
    M, C = x.shape
    Mp = (M + 3) & ~3
    hi, lo = (torch.empty_like(x), torch.empty_like(x)) if plain else (None, None)
    thi, tlo = torch.empty(C, Mp, dtype=torch.float32, device=x.device), torch.empty(C, Mp, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.dge_gemm_split_transpose(M, C, _p(x), _p(hi), _p(lo), _p(thi), _p(tlo), _st(x.device))
    if rc:
        raise DgeError(f"dge_gemm_split_transpose failed ({rc})")
    launch_count += 1
    return (hi, lo), (thi, tlo)


def _tc_gemm_parts(a_hi: torch.Tensor, a_lo: torch.Tensor, bt_hi: torch.Tensor, bt_lo: torch.Tensor):
    """C [M,N] = A [M,K] @ Bt[N,K]^T from operands that are already split (contraction over the operands' contiguous axis)."""
    global launch_count
    L = _gemm_lib()
    M, K = a_hi.shape
    N = bt_hi.shape[0]
    c = torch.empty(M, N, dtype=torch.float32, device=a_hi.device)
    with torch.cuda.device(a_hi.device):
        rc = L.dge_gemm_tf32x3_ex(M, None, N, K, _p(a_hi), _p(a_lo), 0, _p(bt_hi), _p(bt_lo), 0, _p(c), N, 1, _st(a_hi.device))
    if rc:
        raise DgeError(f"dge_gemm_tf32x3_ex failed ({rc})")
    launch_count += 1
    return c


def _tc_gemm_over_rows(at: tuple, bt: tuple, rows: int):
    """C [Ca,Cb] = A^T B for A [rows,Ca], B [rows,Cb] given as their transposed splits ([Ca,Mp], [Cb,Mp] from ``split_transpose``):
    the weight-gradient shape of autograd's mm backward (few output tiles, K = nodes) -- the library splits K across the SMs and the
    slices add into the zeroed output."""
    global launch_count
    L = _gemm_lib()
    (a_hi, a_lo), (b_hi, b_lo) = at, bt
    Ca, Mp = a_hi.shape
    Cb = b_hi.shape[0]
    c = torch.zeros(Ca, Cb, dtype=torch.float32, device=a_hi.device)
    with torch.cuda.device(a_hi.device):
        rc = L.dge_gemm_tf32x3_ex(Ca, None, Cb, rows, _p(a_hi), _p(a_lo), Mp, _p(b_hi), _p(b_lo), Mp, _p(c), Cb, 0, _st(a_hi.device))
    if rc:
        raise DgeError(f"dge_gemm_tf32x3_ex (K = rows) failed ({rc})")
    launch_count += 2
    return c


def tc_gemm_tn(a: tuple, b: tuple):
    """C [Ca,Cb] = A^T B for A [rows,Ca], B [rows,Cb] given as their plain (hi, lo) splits -- the SAME splits the forward and grad-input
    products use: the tensor core reads both operands MN-major (dge_gemm_tf32x3_tn), no transposed copy is made.  K = rows is split
    across the SMs when the output has few tiles; the slices add into the zeroed output."""
    global launch_count
    L = _gemm_lib()
    (a_hi, a_lo), (b_hi, b_lo) = a, b
    rows, Ca = a_hi.shape
    Cb = b_hi.shape[1]
    c = torch.zeros(Ca, Cb, dtype=torch.float32, device=a_hi.device)
    with torch.cuda.device(a_hi.device):
        rc = L.dge_gemm_tf32x3_tn(Ca, Cb, rows, _p(a_hi), _p(a_lo), Ca, _p(b_hi), _p(b_lo), Cb, _p(c), Cb, 0, _st(a_hi.device))
    if rc:
        raise DgeError(f"dge_gemm_tf32x3_tn failed ({rc})")
    launch_count += 2
    return c


_colsum_ws: dict = {}


def colsum(x: torch.Tensor) -> torch.Tensor:
    """Column sums of x [M,C] (dge_colsum: deterministic two-stage reduction) -- the bias gradients of the dense layers."""
    global launch_count
    _need_cuda(x, "colsum")
    L = _gemm_lib()
    x = x.contiguous().float()
    M, C = x.shape
    key = (x.device, torch.cuda.current_stream(x.device).cuda_stream)
    need = int(L.dge_colsum_ws_floats(C))
    ws = _colsum_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = _colsum_ws[key] = torch.empty(max(need, 296 * 3072), dtype=torch.float32, device=x.device)
    out = torch.empty(C, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.dge_colsum(M, C, _p(x), C, _p(out), _p(ws), _st(x.device))
    if rc:
        raise DgeError(f"dge_colsum failed ({rc})")
    launch_count += 2
    return out


class _TcMatmulFn(torch.autograd.Function):
    """x [M,K] @ w [K,N] with forward, grad-input AND grad-weight on the tcgen05 3xTF32 kernel.  The weight gradient x^T dy contracts
    over the nodes; the tensor core reads the (hi, lo) splits of x and dy MN-major (``tc_gemm_tn``), i.e. the splits the forward and the
    grad-input product make anyway -- nothing is transposed."""

    @staticmethod
    def forward(ctx, x, w):
        hi, lo = _weight_operand(w, True)
        xh, xl = split_tf32(x.float())
        if ctx.needs_input_grad[1]:
            ctx.xs = (xh, xl)
        ctx.save_for_backward(w)
        return _tc_gemm_parts(xh, xl, hi, lo)

    @staticmethod
    def backward(ctx, gy):
        (w,) = ctx.saved_tensors
        gx = gw = None
        gh, gl = split_tf32(gy.contiguous().float())
        if ctx.needs_input_grad[1]:
            gw = tc_gemm_tn(ctx.xs, (gh, gl))
            ctx.xs = None
        if ctx.needs_input_grad[0]:
            hi, lo = _weight_operand(w, False)            # dY [M,N] @ W^T: Bt = W [K,N] as stored
            gx = _tc_gemm_parts(gh, gl, hi, lo)
        return gx, gw


def tc_matmul(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """x [M,K] @ w [K,N], fp32-quality on the 5th-gen tensor cores (csrc/dge_gemm.cu)."""
    if not (torch.is_grad_enabled() and (x.requires_grad or w.requires_grad)):
        # no backward will follow (inference, or the acting forward of a trainer under no_grad): no transposed operand is prepared
        hi, lo = _weight_operand(w, True)
        return _tc_gemm_parts(*split_tf32(x.float()), hi, lo)
    return _TcMatmulFn.apply(x, w)


class _TcLinearFn(torch.autograd.Function):
    """x [M,I] @ w[O,I]^T (``F.linear`` without bias: the GRUCell transforms, whose weights already are the K-major B operand),
    all three products on the tcgen05 3xTF32 kernel."""

    @staticmethod
    def forward(ctx, x, w):
        hi, lo = _weight_operand(w, False)
        xh, xl = split_tf32(x.float())
        if ctx.needs_input_grad[1]:
            ctx.xs = (xh, xl)
        ctx.save_for_backward(w)
        return _tc_gemm_parts(xh, xl, hi, lo)

    @staticmethod
    def backward(ctx, gy):
        (w,) = ctx.saved_tensors
        gx = gw = None
        gh, gl = split_tf32(gy.contiguous().float())
        if ctx.needs_input_grad[1]:
            gw = tc_gemm_tn((gh, gl), ctx.xs)                      # dW [O,I] = dY^T X, both operands as stored
            ctx.xs = None
        if ctx.needs_input_grad[0]:
            hi, lo = _weight_operand(w, True)             # dY [M,O] @ W [O,I]: Bt = W^T [I,O]
            gx = _tc_gemm_parts(gh, gl, hi, lo)
        return gx, gw


def tc_linear(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    if not (torch.is_grad_enabled() and (x.requires_grad or w.requires_grad)):
        hi, lo = _weight_operand(w, False)
        return _tc_gemm_parts(*split_tf32(x.float()), hi, lo)
    return _TcLinearFn.apply(x, w)


class _GruGatesFn(torch.autograd.Function):
    """h' = GRU gates(gi, gh, b_ih, b_hh, h) (dge_gru_gates); backward = dge_gru_gates_bwd + two dge_colsum for the biases."""

    @staticmethod
    def forward(ctx, gi, gh, b_ih, b_hh, h):
        global launch_count
        L = _gemm_lib()
        gi, gh, h = gi.contiguous(), gh.contiguous(), h.contiguous().float()
        b_ih, b_hh = b_ih.detach().contiguous(), b_hh.detach().contiguous()
        N, C = h.shape
        out = torch.empty_like(h)
        with torch.cuda.device(h.device):
            rc = L.dge_gru_gates(N, C, _p(gi), _p(gh), _p(b_ih), _p(b_hh), _p(h), 0, _p(out), _st(h.device))
        if rc:
            raise DgeError(f"dge_gru_gates failed ({rc})")
        launch_count += 1
        ctx.save_for_backward(gi, gh, b_ih, b_hh, h)
        return out

    @staticmethod
    def backward(ctx, gout):
        global launch_count
        L = _gemm_lib()
        gi, gh, b_ih, b_hh, h = ctx.saved_tensors
        gout = gout.contiguous().float()
        N, C = h.shape
        dgi, dgh, dh = torch.empty_like(gi), torch.empty_like(gh), torch.empty_like(h)
        with torch.cuda.device(h.device):
            rc = L.dge_gru_gates_bwd(N, C, _p(gi), _p(gh), _p(b_ih), _p(b_hh), _p(h), _p(gout), _p(dgi), _p(dgh), _p(dh), _st(h.device))
        if rc:
            raise DgeError(f"dge_gru_gates_bwd failed ({rc})")
        launch_count += 1
        return dgi, dgh, (colsum(dgi) if ctx.needs_input_grad[2] else None), (colsum(dgh) if ctx.needs_input_grad[3] else None), dh


def gru_cell_train(m: torch.Tensor, h: torch.Tensor, rnn: torch.nn.GRUCell) -> torch.Tensor:
    """``rnn(m, h)`` of GatedGraphConv UNDER AUTOGRAD (Networks.py:76-82 inside DeepQ.train / A2C.train): both transforms, their
    grad-input and grad-weight products on the tcgen05 3xTF32 GEMM, gate arithmetic and its backward in ``dge_gru_gates(_bwd)``."""
    _need_cuda(m, "gru_cell_train")
    gi, gh = tc_linear(m, rnn.weight_ih), tc_linear(h, rnn.weight_hh)
    return _GruGatesFn.apply(gi, gh, rnn.bias_ih, rnn.bias_hh, h)


# ------------------------------------------------------------- DQN training step of the GCN, one native call ---
_train_ws: dict = {}
_train_seed = [0x5DEECE66D]


def _train_lib():
    L = _gemm_lib()
    if not hasattr(L, "_train_ready"):
        L.dge_gcn_train_ws_floats.restype = ctypes.c_int64
        L.dge_gcn_train_ws_floats.argtypes = [ctypes.c_int, ctypes.c_int]
        L.dge_gcn_train_step.argtypes = [ctypes.c_int] * 3 + [_vp] * 20 + [ctypes.c_float, ctypes.c_float, ctypes.c_uint64] + [_vp] * 10
        L.dge_gemm_tf32x3_ex.argtypes = [ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp]
        L.dge_clamp_adam_step.argtypes = [ctypes.c_int64, _vp, _vp, _vp, _vp, _vp] + [ctypes.c_float] * 6 + [_vp]
        L._train_ready = True
    return L


def gcn_train_eligible(model) -> bool:
    """Networks.GCN itself (the DQN Q-network) on CUDA with the shapes the native step takes."""
    return QForwardPlan.eligible(model) and all(p.grad is not None and p.grad.is_contiguous() for p in model.parameters())


def gcn_train_step(model, x: torch.Tensor, gs: GraphStructure, action, y, inv_batch: float, drop_p: float, seed: Optional[int] = None,
                   return_debug: bool = False):
    """Forward (functional dropout p), DeepQ.cost and the whole backward pass of ``Networks.GCN`` in ONE native call
    (``dge_gcn_train_step``, csrc/dge_train.cu: 15 launches, the three [nodes,1000]x[1000,1000]-shaped products on the tcgen05
    3xTF32 kernel, no autograd graph).  The gradients are WRITTEN into the parameters' ``.grad`` buffers (which must exist and be
    contiguous: ``dist.FlatGradBucket``); returns (loss, q) as device tensors -- nothing is read back.  ``action`` / ``y``: [N] float
    tensors (None: a = 1 / y = 0, i.e. loss = sum(q^2) * inv_batch)."""
    global launch_count
    _need_cuda(x, "gcn_train_step")
    L = _train_lib()
    dev = x.device
    x = x.contiguous().float()
    N, cin = x.shape
    C = model.conv1.weight.shape[1]
    norm, selfnorm = gs.gcn_norm(True)
    w2 = model.conv2.weight
    t_hi, t_lo = _weight_operand(w2, True)
    s_hi, s_lo = _weight_operand(w2, False)
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    need = int(L.dge_gcn_train_ws_floats(N, C))
    ws = _train_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = _train_ws[key] = torch.empty(max(need, 1 << 22), dtype=torch.float32, device=dev)
    q = torch.empty(N, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    if seed is None:
        _train_seed[0] = (_train_seed[0] * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        seed = _train_seed[0]
    a = None if action is None else action.contiguous().float()
    yy = None if y is None else y.contiguous().float()
    m = model
    hw, hb = m.fully_con1.weight, m.fully_con1.bias
    with torch.cuda.device(dev):
        rc = L.dge_gcn_train_step(N, cin, C, _p(x), _p(gs.rowptr_dst), _p(gs.perm_dst), _p(gs.rowptr_src), _p(gs.perm_src), _p(gs.src), _p(gs.dst),
                                  _p(norm), _p(selfnorm), _p(m.conv1.weight.detach()), _p(m.conv1.bias.detach()), _p(t_hi), _p(t_lo), _p(s_hi), _p(s_lo),
                                  _p(m.conv2.bias.detach()), _p(hw.detach()), _p(hb.detach()), _p(a), _p(yy),
                                  ctypes.c_float(inv_batch), ctypes.c_float(drop_p), ctypes.c_uint64(seed),
                                  _p(m.conv1.weight.grad), _p(m.conv1.bias.grad), _p(w2.grad), _p(m.conv2.bias.grad), _p(hw.grad), _p(hb.grad),
                                  _p(loss), _p(q), _p(ws), _st(dev))
    if rc:
        raise DgeError(f"dge_gcn_train_step failed ({rc})")
    launch_count += 15
    if return_debug:   # tests: the activations the backward pass used (h1, dropout(h2)) -- views of the workspace, valid until the next call
        NC = N * C
        return loss[0], q, dict(h1=ws[8 * N:8 * N + NC].view(N, C), d2=ws[8 * N + 4 * NC:8 * N + 5 * NC].view(N, C))
    return loss[0], q


# ------------------------------------------------------------- whole Q-network forward, one native call ---
_q_ws: dict = {}


def _q_forward_lib():
    L = _gemm_lib()
    if not hasattr(L, "_q_forward_ready"):
        L.dge_gcn_q_forward.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int] + [_vp] * 16
        L._q_forward_ready = True
    return L


def _q_forward_model_args(w1: torch.Tensor, b1, w2: torch.Tensor, b2, head_w: torch.Tensor, head_b):
    """The seven pointer arguments of ``dge_gcn_q_forward`` that depend on the model only (w1, b1, W2^T hi / lo, b2, head weight,
    head bias) and the tensors that keep them alive."""
    hi, lo = _weight_operand(w2, True)
    keep = [w1.detach().contiguous(), None if b1 is None else b1.detach().contiguous(), hi, lo, None if b2 is None else b2.detach().contiguous(),
            head_w.detach().contiguous(), None if head_b is None else head_b.detach()]
    return [_p(t) for t in keep], keep


def gcn_q_forward(x: torch.Tensor, gs: GraphStructure, w1: torch.Tensor, b1, w2: torch.Tensor, b2, head_w: torch.Tensor, head_b: torch.Tensor):
    """Networks.GCN.forward at inference (prob = 0) through ``dge_gcn_q_forward``: one ctypes call, three launches
    (fused first layer with the TF32 split in its epilogue -> tcgen05 GEMM -> aggregate + ReLU + head).  Returns q [N]."""
    global launch_count
    _need_cuda(x, "gcn_q_forward")
    L = _q_forward_lib()
    norm, selfnorm = gs.gcn_norm(True)
    margs, keep = _q_forward_model_args(w1, b1, w2, b2, head_w, head_b)
    x = x.contiguous().float()
    N, cin = x.shape
    C = w1.shape[1]
    dev = x.device
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)      # one workspace per stream: calls on different streams may overlap
    ws = _q_ws.get(key)
    if ws is None or ws.numel() < 3 * N * C:      # grow-only workspace: no allocator traffic in the acting loop
        ws = _q_ws[key] = torch.empty(max(3 * N * C, 1 << 22), dtype=torch.float32, device=dev)
    q = torch.empty(N, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.dge_gcn_q_forward(N, cin, C, _p(x), _p(gs.rowptr_dst), _p(gs.perm_dst), _p(gs.src), _p(norm), _p(selfnorm), *margs,
                                 _p(ws), _p(q), _st(dev))
    if rc:
        raise DgeError(f"dge_gcn_q_forward failed ({rc})")
    launch_count += 3
    return q


_coo_ws: dict = {}


def gcn_q_forward_coo(x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor, w1: torch.Tensor, b1, w2: torch.Tensor, b2, head_w: torch.Tensor,
                      head_b: torch.Tensor):
    """``Networks.GCN.forward(data, 0)`` for a batch that arrives as a raw edge list: CSR builds, normalisation and the three forward
    launches in ONE native call (``dge_gcn_q_forward_coo``) on grow-only scratch -- one ctypes round trip and one ``torch.empty`` (q) per pass
    instead of four calls and a dozen allocations.  Returns q [N]."""
    global launch_count
    _need_cuda(x, "gcn_q_forward_coo")
    L = _gemm_lib()
    if not hasattr(L, "_coo_ready"):
        L.dge_gcn_q_forward_coo_iws.restype = ctypes.c_int64
        L.dge_gcn_q_forward_coo_iws.argtypes = [ctypes.c_int, ctypes.c_int]
        L.dge_gcn_q_forward_coo_fws.restype = ctypes.c_int64
        L.dge_gcn_q_forward_coo_fws.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.dge_gcn_q_forward_coo.argtypes = [ctypes.c_int] * 4 + [_vp] * 15
        L._coo_ready = True
    x = x.contiguous().float()
    N, cin = x.shape
    C = w1.shape[1]
    ei = edge_index if edge_index.dtype == torch.int64 else edge_index.long()
    E = int(ei.shape[1])
    src, dst = ei[0].contiguous(), ei[1].contiguous()
    ew = (torch.ones(E, device=x.device) if edge_attr is None else edge_attr.contiguous().float())
    margs, keep = _q_forward_model_args(w1, b1, w2, b2, head_w, head_b)
    dev = x.device
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ni, nf = int(L.dge_gcn_q_forward_coo_iws(N, E)), int(L.dge_gcn_q_forward_coo_fws(N, E, C))
    ws = _coo_ws.get(key)
    if ws is None or ws[0].numel() < ni or ws[1].numel() < nf:
        ws = _coo_ws[key] = (torch.empty(max(ni, 1 << 20), dtype=torch.int32, device=dev), torch.empty(max(nf, 1 << 22), dtype=torch.float32, device=dev))
    q = torch.empty(N, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.dge_gcn_q_forward_coo(N, E, cin, C, _p(x), _p(src), _p(dst), _p(ew), *margs, _p(ws[0]), _p(ws[1]), _p(q), _st(dev))
    if rc:
        raise DgeError(f"dge_gcn_q_forward_coo failed ({rc})")
    launch_count += 16
    return q


class QForwardPlan:
    """``gcn_q_forward`` for the acting loop with its argument list prepared once.  Between two ticks of ``runner.PolicyLoop``
    every pointer of the call is the same -- the weights (inference) and the engine-owned graph batch buffers (x, destination
    CSR, GCN normalisation: ``envs.exploration_env.GraphBatch``) -- and only the node count changes, so the per-tick Python glue
    of the generic route (``GraphBatch.data()`` views, ``GraphStructure.from_csr``, module dispatch, 16 pointer conversions; ~90 us
    of the loop's critical host path per tick, profiles/r01_policyloop_host_profile.txt) shrinks to one ctypes call.  Same entry
    point, same arguments, same three launches as ``Networks.GCN.forward(data, 0.0)`` on ``batch.data()``: identical Q-values
    (tests/test_q_plan_cpu.py compares the argument lists, tests/test_zz_q_plan_gpu.py the values).  The model arguments are
    re-derived when a parameter changes (version counter / storage)."""

    def __init__(self, model, batch):
        self.model, self.batch = model, batch
        self.cin, self.C = (int(v) for v in model.conv1.weight.shape)
        self.dev = batch.x.device
        self.q = torch.empty(batch.node_cap, dtype=torch.float32, device=self.dev)
        self._gargs = [_p(batch.x), _p(batch.csr_rowptr), _p(batch.csr_perm), _p(batch.edge_index), _p(batch.gcn_norm), _p(batch.gcn_selfnorm)]
        self._qarg = _p(self.q)
        self._sig = self._margs = self._keep = self._ws = self._wsarg = None

    @staticmethod
    def eligible(model) -> bool:
        """The conditions under which ``Networks.GCN._trunk`` itself takes the ``dge_gcn_q_forward`` route (precision mode and
        autograd state are the caller's to check per tick)."""
        from . import Networks
        if type(model) is not Networks.GCN:      # the DQN Q-network itself (a subclass may read the trunk differently)
            return False
        try:
            cin, C = model.conv1.weight.shape
            return (getattr(model, "_out", None) == 1 and model.conv1.weight.is_cuda and cin <= 8 and model.conv2.out_channels % 4 == 0
                    and model.conv2.out_channels <= 1024 and tuple(model.conv2.weight.shape) == (C, C) and model.conv1.improved and model.conv2.improved)
        except AttributeError:
            return False

    def _params(self):
        m = self.model
        return (m.conv1.weight, m.conv1.bias, m.conv2.weight, m.conv2.bias, m.fully_con1.weight, m.fully_con1.bias)

    def __call__(self, n_nodes: int, graph_args=None) -> torch.Tensor:
        """Q-values [n_nodes] of the batch currently in the engine's graph buffers (a view of the plan's output buffer).
        ``graph_args``: the six graph pointers (x, rowptr, perm, src, norm, selfnorm) when the batch lives elsewhere -- the host
        loop's packed arena (runner.packed_graph_args)."""
        global launch_count
        n = int(n_nodes)
        gargs = self._gargs if graph_args is None else graph_args
        if n > self.batch.node_cap:
            raise DgeError("QForwardPlan: more nodes than the graph batch holds")
        sig = tuple((t.data_ptr(), t._version) if t is not None else None for t in self._params())
        if sig != self._sig:
            w1, b1, w2, b2, hw, hb = self._params()
            self._margs, self._keep = _q_forward_model_args(w1, b1, w2, b2, hw[0], hb)
            self._sig = sig
        if self._ws is None or self._ws.numel() < 3 * n * self.C:
            self._ws = torch.empty(max(3 * n * self.C, 1 << 22), dtype=torch.float32, device=self.dev)
            self._wsarg = _p(self._ws)
        L = _q_forward_lib()
        if torch.cuda.current_device() == self.dev.index:
            rc = L.dge_gcn_q_forward(n, self.cin, self.C, *gargs, *self._margs, self._wsarg, self._qarg, _st(self.dev))
        else:
            with torch.cuda.device(self.dev):
                rc = L.dge_gcn_q_forward(n, self.cin, self.C, *gargs, *self._margs, self._wsarg, self._qarg, _st(self.dev))
        if rc:
            raise DgeError(f"dge_gcn_q_forward failed ({rc})")
        launch_count += 3
        return self.q[:n]


def gru_cell_inference(m: torch.Tensor, h: torch.Tensor, rnn: torch.nn.GRUCell, relu: bool = False) -> torch.Tensor:
    """``rnn(m, h)`` of GatedGraphConv at inference: both dense transforms on the tcgen05 3xTF32 GEMM (the GRUCell weights
    [3C, C] already are the K-major B operand), gate arithmetic in ``dge_gru_gates``."""
    global launch_count
    _need_cuda(m, "gru_cell_inference")
    L = _gemm_lib()
    if not hasattr(L, "_gru_ready"):
        L.dge_gru_gates.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp]
        L._gru_ready = True
    ih_hi, ih_lo = _weight_operand(rnn.weight_ih, False)
    hh_hi, hh_lo = _weight_operand(rnn.weight_hh, False)
    m, h = m.contiguous().float(), h.contiguous().float()
    gi, gh = _tc_gemm(m, ih_hi, ih_lo), _tc_gemm(h, hh_hi, hh_lo)
    N, C = h.shape
    out = torch.empty_like(h)
    with torch.cuda.device(h.device):
        rc = L.dge_gru_gates(N, C, _p(gi), _p(gh), _p(rnn.bias_ih.detach().contiguous()), _p(rnn.bias_hh.detach().contiguous()), _p(h), int(relu),
                             _p(out), _st(h.device))
    if rc:
        raise DgeError(f"dge_gru_gates failed ({rc})")
    launch_count += 1
    return out


def augment_adj(edge_index: torch.Tensor, edge_weight: torch.Tensor, num_nodes: int, batch: Optional[torch.Tensor], graph_ptr: torch.Tensor,
                max_graph_nodes: int):
    """GraphUNet.augment_adj on the native kernels: remove_self_loops((A + I)^2), entries sorted by (row, col) like
    torch_sparse.spspmm's coalesced output.  The edge list must be coalesced (no duplicate pairs).  Returns None when a graph
    is too large for the shared-memory row accumulator (the caller then takes its generic path)."""
    global launch_count
    _need_cuda(edge_index, "augment_adj")
    L = _lib()
    if not hasattr(L, "_augment_ready"):
        L.dge_gnn_augment_adj_count.argtypes = [ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp]
        L.dge_gnn_augment_adj_fill.argtypes = [ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        L._augment_ready = True
    if max_graph_nodes > 1024:
        return None
    dev, N = edge_index.device, int(num_nodes)
    gs = GraphStructure(edge_index, edge_weight, N)
    rp, pm = gs.rowptr_src, gs.perm_src
    cnt = torch.empty(N, dtype=torch.int32, device=dev)
    outptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
    gp = graph_ptr.long().contiguous()
    bt = None if batch is None else batch.long().contiguous()
    with torch.cuda.device(dev):
        rc = L.dge_gnn_augment_adj_count(N, _p(rp), _p(pm), _p(gs.dst), _p(gs.weight), _p(bt), _p(gp), int(max_graph_nodes), _p(cnt), _p(outptr), _st(dev))
    if rc:
        raise DgeError(f"dge_gnn_augment_adj_count failed ({rc})")
    E2 = int(outptr[N].item())                       # host sync: sizes the output (spspmm synchronises for the same reason)
    row, col = torch.empty(max(E2, 1), dtype=torch.int64, device=dev), torch.empty(max(E2, 1), dtype=torch.int64, device=dev)
    val = torch.empty(max(E2, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.dge_gnn_augment_adj_fill(N, _p(rp), _p(pm), _p(gs.dst), _p(gs.weight), _p(bt), _p(gp), _p(outptr), _p(row), _p(col), _p(val), _st(dev))
    if rc:
        raise DgeError(f"dge_gnn_augment_adj_fill failed ({rc})")
    launch_count += 3
    return torch.stack([row[:E2], col[:E2]]), val[:E2]


def topk_pool(score: torch.Tensor, graph_ptr: torch.Tensor, k_ptr: torch.Tensor, max_graph_nodes: int, n_kept: int):
    """PyG ``topk``: per graph the k best-scored nodes in descending order -> (perm [n_kept], newid [N])."""
    global launch_count
    _need_cuda(score, "topk_pool")
    L = _lib()
    if not hasattr(L, "_topk_ready"):
        L.dge_gnn_topk_pool.argtypes = [ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp]
        L.dge_gnn_filter_adj_count.argtypes = [ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp]
        L.dge_gnn_filter_adj_fill.argtypes = [ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
        L._topk_ready = True
    dev, N = score.device, score.numel()
    perm = torch.empty(max(n_kept, 1), dtype=torch.int64, device=dev)
    newid = torch.empty(max(N, 1), dtype=torch.int64, device=dev)
    G = graph_ptr.numel() - 1
    with torch.cuda.device(dev):
        rc = L.dge_gnn_topk_pool(G, int(max_graph_nodes), _p(score.contiguous().float()), _p(graph_ptr.long().contiguous()),
                                 _p(k_ptr.long().contiguous()), _p(perm), _p(newid), _st(dev))
    if rc:
        raise DgeError(f"dge_gnn_topk_pool failed ({rc})")
    launch_count += 1
    return perm[:n_kept], newid[:N]


def filter_adj(edge_index: torch.Tensor, edge_weight: torch.Tensor, newid: torch.Tensor):
    """PyG ``filter_adj``: edges whose ends both survive the pooling, relabelled, original order."""
    global launch_count
    L = _lib()
    dev, E = edge_index.device, edge_index.size(1)
    if E == 0:
        return edge_index, edge_weight
    ei = edge_index.contiguous()
    src, dst, w = ei[0].contiguous(), ei[1].contiguous(), edge_weight.contiguous().float()
    flag, pos = torch.empty(E, dtype=torch.int32, device=dev), torch.empty(E + 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = L.dge_gnn_filter_adj_count(E, _p(src), _p(dst), _p(newid), _p(flag), _p(pos), _st(dev))
    if rc:
        raise DgeError(f"dge_gnn_filter_adj_count failed ({rc})")
    E2 = int(pos[E].item())                                  # host sync: sizes the pooled edge list
    out = torch.empty(2, max(E2, 1), dtype=torch.int64, device=dev)
    ow = torch.empty(max(E2, 1), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = L.dge_gnn_filter_adj_fill(E, _p(src), _p(dst), _p(w), _p(newid), _p(flag), _p(pos), _p(out[0]), _p(out[1]), _p(ow), _st(dev))
    if rc:
        raise DgeError(f"dge_gnn_filter_adj_fill failed ({rc})")
    launch_count += 3
    return out[:, :E2], ow[:E2]
