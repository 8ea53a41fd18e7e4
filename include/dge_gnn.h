/* ============================================================================
 * dge_gnn.h -- C ABI of the GNN message-passing kernels in libdge.so.
 *
 * Replaces, for the graph Q-network of scripts/Networks.py (GCN: Networks.py:12-28;
 * the conv layers of GGNN / GraphUNet reuse the same aggregation), the PyG /
 * torch_scatter call sites listed in SURVEY section 2.3: add_remaining_self_loops +
 * weighted-degree scatter_add + symmetric norm (GCNConv.norm) and the per-edge
 * gather * scale + scatter_add of GCNConv.propagate / GatedGraphConv.propagate.
 * Plain device pointers; int64 edge indices exactly as torch `edge_index` rows.
 * Return 0 ok, -1 bad argument, -2 CUDA error.  All launches are async on `stream`.
 * ==========================================================================*/
#ifndef DGE_GNN_H_
#define DGE_GNN_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Deterministic CSR of E edges grouped by key[e] in [0,N) (key = edge_index[1] for the
 * forward gather, edge_index[0] for the transposed/backward gather); rows sorted by edge id.
 * rowptr [N+1], perm [E], ws [2N+E] scratch. */
int dge_gnn_csr_build(int N, int E, const int64_t *key, int32_t *rowptr, int32_t *perm, int32_t *ws, void *stream);

/* GCNConv.norm(improved: fill = 2): dis [N] = deg^-1/2, selfw [N] self-loop weight,
 * norm [E] per-edge coefficient (0 on explicit self loops), selfnorm [N]. */
int dge_gcn_norm(int N, int E, const int64_t *src, const int64_t *dst, const float *w, const int32_t *rowptr_src,
                 const int32_t *perm_src, float fill, float *dis, float *selfw, float *norm, float *selfnorm, void *stream);

/* out[i,:] = act(bias + selfcoef[i] X[i,:] + sum_{p in rowptr[i]..rowptr[i+1]} coef[perm[p]] * mask(X[nbr[perm[p]],:]))
 *   X [N,C] f32 row-major (C <= 1024), nbr [E] int64 = the *other* endpoint of each edge,
 *   coef [E], selfcoef [N] nullable, bias [C] nullable, gate [N,C] nullable (rows are
 *   multiplied by gate>0: ReLU backward fused into the gather), relu 0/1, out [N,C] nullable,
 *   head_w [C] nullable: also emit q[i] = dot(out[i,:], head_w) + head_b (Linear(C,1) fused). */
int dge_gnn_aggregate(int N, int C, const float *X, const int32_t *rowptr, const int32_t *perm, const int64_t *nbr,
                      const float *coef, const float *selfcoef, const float *bias, const float *gate, int relu, float *out,
                      const float *head_w, float head_b, float *q, void *stream);

/* First GCN layer fused (GCNConv(5 -> C), Networks.py:15,22): out = act(bias + (A_hat X) W) with the
 * Cin <= 8 input channels aggregated BEFORE the transform; W [Cin,C] row-major.  Inference only. */
int dge_gcn_conv_small(int N, int Cin, int C, const float *X, const int32_t *rowptr, const int32_t *perm, const int64_t *nbr,
                       const float *coef, const float *selfcoef, const float *W, const float *bias, int relu, float *out, void *stream);

/* ---- dense node-MLP GEMM on the tcgen05 tensor cores (csrc/dge_gemm.cu): replaces the cuBLAS SGEMM behind
 * GCNConv's `torch.matmul(x, self.weight)` (Networks.py:22-24 via PyG), GatedGraphConv's `h @ W[i]` and the GRUCell
 * transforms (Networks.py:76-82).  C[M,N] = A[M,K] * Bt[N,K]^T in 3xTF32 (fp32-quality: ~1e-6 relative), fp32
 * accumulation in tensor memory.  Operands are pre-split into (hi, lo) TF32 parts:
 *   dge_gemm_split_tf32   x[n] -> hi[n], lo[n]                    (n % 4 == 0; activations, or a weight used as Bt as is)
 *   dge_gemm_prep_weight  W[K,N] row-major -> Wt_hi, Wt_lo [N,K]   (the Bt operand of X @ W)
 *   dge_gemm_tf32x3       K % 4 == 0, ldc % 4 == 0, 16-byte aligned pointers; M_dev nullable: live row count read on
 *                         the device (<= M), so a batch sized on the device needs no host sync; rows >= *M_dev untouched. */
int dge_gemm_split_tf32(int64_t n, const float *x, float *hi, float *lo, void *stream);
int dge_gemm_prep_weight(int K, int N, const float *W, float *Wt_hi, float *Wt_lo, void *stream);
int dge_gemm_tf32x3(int M, const int32_t *M_dev, int N, int K, const float *A_hi, const float *A_lo, const float *Bt_hi,
                    const float *Bt_lo, float *C, int ldc, void *stream);
/* general form: operand row pitches lda / ldb in floats (multiples of 4, >= K; 0 = K, which then may be any length) and `splits`
 * K slices per output tile (> 1: C must be zeroed by the caller, the slices add their partial products with atomics; 0: chosen
 * by the library -- the weight-gradient shape x^T dy of autograd's mm backward, few output tiles and K = nodes, is split
 * across the SMs).                                                                                                     */
int dge_gemm_tf32x3_ex(int M, const int32_t *M_dev, int N, int K, const float *A_hi, const float *A_lo, int lda, const float *Bt_hi,
                       const float *Bt_lo, int ldb, float *C, int ldc, int splits, void *stream);

/* transposed-A form for the weight gradient x^T dy of autograd's mm backward (torch: `x.t() @ dy`, cuBLAS SGEMM 'TN'): C[M,N] = A^T B with
 * A [K,M] and B [K,N] row-major as stored (row pitches lda >= M, ldb >= N, multiples of 4; 0 = M / N), contraction over the K rows (nodes).
 * Both operands are read MN-major by the tensor core (TMA boxes of 32 columns x 32 rows, major bits of the instruction descriptor), so the
 * (hi, lo) splits of x and dy that the forward and grad-input products already use serve here too -- no transposed copies.            */
int dge_gemm_tf32x3_tn(int M, int N, int K, const float *A_hi, const float *A_lo, int lda, const float *B_hi, const float *B_lo, int ldb,
                       float *C, int ldc, int splits, void *stream);

/* ---- Q-values of the GCN Q-network (Networks.GCN.forward(data, 0): scripts/Networks.py:18-28, called by DeepQ.test, policy.py:255-259) from a raw
 * PyG edge list in one call: both CSRs, the improved-GCN normalisation, fused first layer, tcgen05 GEMM, aggregation + ReLU + head.  src / dst =
 * edge_index rows [E] int64, w = edge_attr [E]; W2t_(hi,lo) = dge_gemm_prep_weight(W2); head_b_dev [1] on the device.  iws / fws: scratch of
 * dge_gcn_q_forward_coo_iws(N, E) int32 / dge_gcn_q_forward_coo_fws(N, E, C) floats (fws 16-byte aligned).  q [N].                          */
int64_t dge_gcn_q_forward_coo_iws(int N, int E);
int64_t dge_gcn_q_forward_coo_fws(int N, int E, int C);
int dge_gcn_q_forward_coo(int N, int E, int Cin, int C, const float *x, const int64_t *src, const int64_t *dst, const float *w,
                          const float *W1, const float *b1, const float *W2t_hi, const float *W2t_lo, const float *b2, const float *head_w,
                          const float *head_b_dev, int32_t *iws, float *fws, float *q, void *stream);

/* ---- one DQN training step of the GCN Q-network without autograd (csrc/dge_train.cu).  Replaces DeepQ.train + DeepQ.cost
 * (scripts/policy.py:234-253: model(data, 0.5) -> sum((Q a - y)^2) / BATCH -> backward) for scripts/Networks.py:12-28:
 * forward with functional dropout drop_p (Philox stream keyed by drop_seed), cost, and the gradients of all six parameters.
 * (rowptr_d, perm_d): destination-sorted CSR (dge_gnn_csr_build on edge_index[1]); (rowptr_s, perm_s): source-sorted;
 * src / dst = edge_index rows; norm / selfnorm from dge_gcn_norm(fill = 2).  W2t_(hi,lo) = dge_gemm_prep_weight(W2),
 * W2_(hi,lo) = dge_gemm_split_tf32(W2).  act / y [N] nullable (a = 1 / y = 0).  Gradients are written to gW1 [Cin,C], gb1 [C],
 * gW2 [C,C], gb2 [C], gWh [C], gbh [1]; loss [1] and q [N] stay on the device.  ws: dge_gcn_train_ws_floats(N, C) floats,
 * 16-byte aligned.  Cin <= 8, C % 4 == 0, C <= 1024.                                                                     */
int64_t dge_gcn_train_ws_floats(int N, int C);
int dge_gcn_train_step(int N, int Cin, int C, const float *x, const int32_t *rowptr_d, const int32_t *perm_d, const int32_t *rowptr_s,
                       const int32_t *perm_s, const int64_t *src, const int64_t *dst, const float *norm, const float *selfnorm,
                       const float *W1, const float *b1, const float *W2t_hi, const float *W2t_lo, const float *W2_hi, const float *W2_lo,
                       const float *b2, const float *Wh, const float *bh, const float *act, const float *y, float inv_batch, float drop_p,
                       uint64_t drop_seed, float *gW1, float *gb1, float *gW2, float *gb2, float *gWh, float *gbh, float *loss, float *q,
                       float *ws, void *stream);
/* The optimizer half of the step (policy.py:251-253: clamp(+-clamp) on every gradient element, then torch.optim.Adam.step) as one
 * kernel over flat buffers of n floats: g <- clamp(g * gscale), m, v, p updated with torch's Adam arithmetic (amsgrad off, no
 * weight decay); step [1] int64 on the device is incremented first.                                                      */
int dge_clamp_adam_step(int64_t n, float *p, float *g, float *m, float *v, int64_t *step, float lr, float beta1, float beta2, float eps,
                        float clamp, float gscale, void *stream);


/* ---- GRU cell of the GG-NN family (torch.nn.GRUCell inside PyG GatedGraphConv, Networks.py:73-86): the gate arithmetic
 * after the two dense transforms gi = m W_ih^T, gh = h W_hh^T [N,3C] (gate order r | z | n), biases [3C], previous state
 * h [N,C] -> h' [N,C] (optionally ReLU'd).  C % 4 == 0, 16-byte aligned pointers.                                    */
int dge_gru_gates(int N, int C, const float *gi, const float *gh, const float *b_ih, const float *b_hh, const float *h, int relu,
                  float *out, void *stream);
/* Backward of the same cell (what autograd does inside torch.nn.GRUCell under A2C.train / DeepQ.train of the GG-NN nets,
 * policy.py:241-253,474-497): from the saved forward inputs and gout = dL/dh' [N,C] (no ReLU on this path) the gates are
 * recomputed and
 *   dgi [N,3C] = (da_r | da_z | da_n),   dgh [N,3C] = (da_r | da_z | da_n * r),   dh [N,C] = gout * z   (the direct path)
 * with da_n = gout (1 - z)(1 - n^2), da_r = da_n (gh_n + b_hn) r (1 - r), da_z = gout (h - n) z (1 - z).  The dense products
 * (dm = dgi W_ih, dh += dgh W_hh, dW_ih = dgi^T m, dW_hh = dgh^T h) are dge_gemm_tf32x3_ex calls, the bias gradients dge_colsum. */
int dge_gru_gates_bwd(int N, int C, const float *gi, const float *gh, const float *b_ih, const float *b_hh, const float *h, const float *gout,
                      float *dgi, float *dgh, float *dh, void *stream);
/* out [C] = column sums of X [N,C] (row pitch ldx floats), deterministic: row slabs -> partials -> slab-order sum.
 * ws: dge_colsum_ws_floats(C) floats.  The bias gradients of the dense layers (sum over nodes of dL/dy). */
int64_t dge_colsum_ws_floats(int C);
int dge_colsum(int N, int C, const float *X, int ldx, float *out, float *ws, void *stream);
/* X [N,C] -> its TF32 (hi, lo) split [N,C] (nullable pair) and the split of X^T as (thi, tlo) [C,Np], Np = (N + 3) & ~3, pad
 * columns zero: both K-major operand forms of dge_gemm_tf32x3_ex -- the transposed one feeds the weight gradient x^T dy (K = nodes). */
int dge_gemm_split_transpose(int N, int C, const float *X, float *hi, float *lo, float *thi, float *tlo, void *stream);

/* ---- g-U-Net: GraphUNet.augment_adj (Networks.py:216-225: add_self_loops -> spspmm(A, A) -> remove_self_loops, coalesced)
 * for a block-diagonal batch.  rowptr_src / perm_src = source-sorted CSR of the edge list (dge_gnn_csr_build on edge_index[0]),
 * dst = edge_index[1], w = edge weights, batch [N] graph of every node (nullable: one graph), graph_ptr [G+1] node ranges.
 * Pass 1 counts the entries of every output row and scans them into outptr [N+1] (outptr[N] = E', which the caller reads to
 * size the outputs); pass 2 writes (row, col, value) sorted by (row, col).  Returns -3 when a graph has more than 1024 nodes. */
int dge_gnn_augment_adj_count(int N, const int32_t *rowptr_src, const int32_t *perm_src, const int64_t *dst, const float *w,
                              const int64_t *batch, const int64_t *graph_ptr, int max_graph_nodes, int32_t *cnt, int32_t *outptr, void *stream);
int dge_gnn_augment_adj_fill(int N, const int32_t *rowptr_src, const int32_t *perm_src, const int64_t *dst, const float *w,
                             const int64_t *batch, const int64_t *graph_ptr, const int32_t *outptr, int64_t *out_row, int64_t *out_col,
                             float *out_val, void *stream);

/* ---- g-U-Net: TopKPooling (Networks.py:150-158 via PyG topk / filter_adj).  score [N] (tanh(x.w/|w|), computed by the caller),
 * graph_ptr / k_ptr [G+1] = node ranges and prefix sums of k = ceil(ratio n) per graph.  perm [sum k]: kept nodes, per graph in
 * descending score order (ties: lower index first); newid [N]: new index of a kept node, -1 otherwise.  Graphs of at most 1024
 * nodes (else -3).  filter_adj keeps the edges whose ends both survive, relabelled, in their original order: ..._count writes
 * flag [E] and pos [E+1] (pos[E] = number of kept edges, read by the caller to size the outputs), ..._fill compacts.          */
int dge_gnn_topk_pool(int G, int max_graph_nodes, const float *score, const int64_t *graph_ptr, const int64_t *k_ptr, int64_t *perm,
                      int64_t *newid, void *stream);
int dge_gnn_filter_adj_count(int E, const int64_t *src, const int64_t *dst, const int64_t *newid, int32_t *flag, int32_t *pos, void *stream);
int dge_gnn_filter_adj_fill(int E, const int64_t *src, const int64_t *dst, const float *w, const int64_t *newid, const int32_t *flag,
                            const int32_t *pos, int64_t *out_src, int64_t *out_dst, float *out_w, void *stream);

/* ---- the whole DQN Q-network forward at inference (Networks.GCN.forward with prob = 0, Networks.py:18-28: two
 * GCNConv(improved) + ReLU and the Linear(C,1) head) in ONE call / three launches: fused first layer with the TF32 split
 * in its epilogue -> dge_gemm_tf32x3 -> aggregate + bias + ReLU + head.  rowptr/perm = destination-sorted CSR, src =
 * edge_index[0], norm/selfnorm from dge_gcn_norm (or the graph kernel), W1 [Cin,C], W2t_hi/lo = dge_gemm_prep_weight of
 * conv2.weight, head_w [C], head_b_dev [1] on the device (nullable), ws >= 3*N*C floats (16-byte aligned), q [N].   */
int dge_gcn_q_forward(int N, int Cin, int C, const float *x, const int32_t *rowptr, const int32_t *perm, const int64_t *src,
                      const float *norm, const float *selfnorm, const float *W1, const float *b1, const float *W2t_hi,
                      const float *W2t_lo, const float *b2, const float *head_w, const float *head_b_dev, float *ws, float *q,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif
