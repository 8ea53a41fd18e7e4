// Training step of the DQN Q-network (Networks.GCN) on hand-written kernels: forward with functional dropout, the DQN cost,
// the whole backward pass and the clamp + Adam update -- no autograd graph, no library GEMM, no eager element-wise kernels.
//
// Replaces DeepQ.train / DeepQ.cost (scripts/policy.py:234-253: model(data, 0.5), sum((Q a - y)^2) / BATCH, backward,
// clamp(+-0.5), Adam step) for the model of scripts/Networks.py:12-28, row a17 of SURVEY section 8:
//     h1 = relu(A (x W1) + b1)          A = D^-1/2 (Adj + 2I) D^-1/2, = (A x) W1: 5 input channels are aggregated first
//     h2 = relu(A (h1 W2) + b2)
//     q  = dropout_p(h2) Wh^T + bh
// Backward, with dq = 2 (q a - y) a / BATCH  (non-zero on the action nodes only):
//     dWh = dq^T d2, dbh = sum dq                              d2 = dropout(h2)                         (k_colreduce<0>)
//     dz2 = (dq Wh) * [d2 != 0] / (1 - p)   never materialised: rows are formed on the fly where dq != 0
//     db2 = sum_n dz2                                                                                   (k_colreduce<0>)
//     g2  = A^T dz2                      gather over the source-sorted CSR, neighbours with dq = 0 skipped   (k_agg_bwd)
//     dW2 = h1^T g2                      tcgen05 3xTF32 GEMM over K = nodes, both operands read as stored (MN-major: dge_gemm_tf32x3_tn), K split across CTAs
//     dh1 = g2 W2^T                      tcgen05 3xTF32 GEMM
//     dz1 = dh1 * [h1 > 0],  db1 = sum_n dz1,  dW1 = (A x)^T dz1                                         (k_colreduce<1>)
// The column reductions are deterministic (row slabs -> partials -> the last CTA adds them in slab order).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "../../include/dge_gnn.h"

int dge_agg_fwd_train_bulk(int N, int C, const float *X, const int32_t *rowptr, const int32_t *perm, const int64_t *nbr, const float *coef,
                           const float *selfcoef, const float *bias, float drop_p, uint64_t drop_seed, const float *head_w, const float *head_b_dev,
                           float *d2, float *q, cudaStream_t st);

namespace {

constexpr int NSLAB = 296;       // row slabs of a column reduction (2 per SM)

__device__ __forceinline__ void tf32_split(float a, float &hi, float &lo) {
  uint32_t hb, lb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(a));
  hi = __uint_as_float(hb);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(a - hi));
  lo = __uint_as_float(lb);
}

// Philox4x32-10 (the engine's generator, csrc/dge_internal.cuh): four 32-bit draws for (key, counter)
__device__ __forceinline__ uint4 philox(uint64_t key, uint64_t ctr) {
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// ---- first layer: ax = A x (Cin <= 8 channels), h1 = relu(ax W1 + b1) ------------------------------------------------------
__global__ void __launch_bounds__(256) k_conv1_train(int N, int Cin, int C, const float *__restrict__ X, const int32_t *__restrict__ rowptr,
                                                     const int32_t *__restrict__ perm, const int64_t *__restrict__ nbr, const float *__restrict__ coef,
                                                     const float *__restrict__ selfcoef, const float *__restrict__ W, const float *__restrict__ bias,
                                                     float *__restrict__ ax /*[N,8]*/, float *__restrict__ h1 /*[N,C]*/) {
  __shared__ float agg[8];
  for (int i = blockIdx.x; i < N; i += gridDim.x) {
    if (threadIdx.x < 32) {   // one warp aggregates the input row: lanes over edges (fixed order per lane, shuffle tree => deterministic)
      float a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = 0.f;
      for (int p = rowptr[i] + threadIdx.x; p < rowptr[i + 1]; p += 32) {
        const int e = perm[p];
        const int n = (int)nbr[e];
        const float cf = coef[e];
#pragma unroll
        for (int k = 0; k < 8; ++k) if (k < Cin) a[k] += cf * X[(size_t)n * Cin + k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
      if (threadIdx.x == 0) {
        const float sc = selfcoef[i];
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float v = (k < Cin) ? a[k] + sc * X[(size_t)i * Cin + k] : 0.f; agg[k] = v; ax[(size_t)i * 8 + k] = v; }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float r = bias ? bias[c] : 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) if (k < Cin) r += agg[k] * W[(size_t)k * C + c];
      h1[(size_t)i * C + c] = fmaxf(r, 0.f);
    }
    __syncthreads();
  }
}

// ---- X [N,C] -> (hi, lo) [N,C] and the transposed (thi, tlo) [C,Np]: both K-major operand forms of the tensor-core GEMM ---------
__global__ void __launch_bounds__(256) k_split_transpose(int N, int C, int Np, const float *__restrict__ X, float *__restrict__ hi, float *__restrict__ lo,
                                                         float *__restrict__ thi, float *__restrict__ tlo) {
  __shared__ float th[32][33], tl[32][33];
  const int n0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, c = c0 + tx;
    float h = 0.f, l = 0.f;
    if (n < N && c < C) {
      tf32_split(X[(size_t)n * C + c], h, l);
      if (hi) { hi[(size_t)n * C + c] = h; lo[(size_t)n * C + c] = l; }
    }
    th[r][tx] = h; tl[r][tx] = l;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, n = n0 + tx;
    if (c < C && n < Np) { thi[(size_t)c * Np + n] = th[tx][r]; tlo[(size_t)c * Np + n] = tl[tx][r]; }   // (pad columns n >= N: zeros)
  }
}

// ---- second layer aggregation + bias + ReLU + dropout + head: d2 = dropout(relu(A t2 + b2)), q = d2 Wh + bh -----------------------
__global__ void __launch_bounds__(256) k_agg_fwd_train(int N, int C, const float *__restrict__ X, const int32_t *__restrict__ rowptr,
                                                       const int32_t *__restrict__ perm, const int64_t *__restrict__ nbr, const float *__restrict__ coef,
                                                       const float *__restrict__ selfcoef, const float *__restrict__ bias, float drop_p, uint64_t seed,
                                                       const float *__restrict__ head_w, const float *__restrict__ head_b, float *__restrict__ d2,
                                                       float *__restrict__ q) {
  const int c0 = threadIdx.x * 4;
  const bool act = c0 < C;
  const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  __shared__ float red[8];
  for (int i = blockIdx.x; i < N; i += gridDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto add_row = [&](int n, float cf) {
      if (!act) return;
      const float4 x = *reinterpret_cast<const float4 *>(X + (size_t)n * C + c0);
      acc.x += cf * x.x; acc.y += cf * x.y; acc.z += cf * x.z; acc.w += cf * x.w;
    };
    const int lo = rowptr[i], hi = rowptr[i + 1];
    int p = lo;
    for (; p + 7 < hi; p += 8) {   // eight neighbour rows in flight: indices and coefficients first, then the row loads back to back
      int nn[8]; float ff[8]; float4 xx[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int e = perm[p + u]; nn[u] = (int)nbr[e]; ff[u] = coef[e]; }
      if (act) {
#pragma unroll
        for (int u = 0; u < 8; ++u) xx[u] = *reinterpret_cast<const float4 *>(X + (size_t)nn[u] * C + c0);
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += ff[u] * xx[u].x; acc.y += ff[u] * xx[u].y; acc.z += ff[u] * xx[u].z; acc.w += ff[u] * xx[u].w; }
      }
    }
    for (; p + 3 < hi; p += 4) {   // four neighbour rows in flight
      const int e0 = perm[p], e1 = perm[p + 1], e2 = perm[p + 2], e3 = perm[p + 3];
      const int n0 = (int)nbr[e0], n1 = (int)nbr[e1], n2 = (int)nbr[e2], n3 = (int)nbr[e3];
      const float f0 = coef[e0], f1 = coef[e1], f2 = coef[e2], f3 = coef[e3];
      add_row(n0, f0); add_row(n1, f1); add_row(n2, f2); add_row(n3, f3);
    }
    for (; p < hi; ++p) { const int e0 = perm[p]; add_row((int)nbr[e0], coef[e0]); }
    add_row(i, selfcoef[i]);
    float hq = 0.f;
    if (act) {
      const float4 b = *reinterpret_cast<const float4 *>(bias + c0);
      float r[4] = {fmaxf(acc.x + b.x, 0.f), fmaxf(acc.y + b.y, 0.f), fmaxf(acc.z + b.z, 0.f), fmaxf(acc.w + b.w, 0.f)};
      if (drop_p > 0.f) {   // functional dropout (Networks.py:26): keep with probability 1 - p, scale by 1 / (1 - p)
        const uint4 u = philox(seed, (uint64_t)i * 256 + threadIdx.x);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int v = 0; v < 4; ++v) r[v] = ((uu[v] >> 8) * (1.0f / 16777216.0f) >= drop_p) ? r[v] * scale : 0.f;
      }
      *reinterpret_cast<float4 *>(d2 + (size_t)i * C + c0) = make_float4(r[0], r[1], r[2], r[3]);
      const float4 w = *reinterpret_cast<const float4 *>(head_w + c0);
      hq = r[0] * w.x + r[1] * w.y + r[2] * w.z + r[3] * w.w;
    }
    for (int o = 16; o > 0; o >>= 1) hq += __shfl_xor_sync(0xffffffffu, hq, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = hq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int wv = 0; wv < 8; ++wv) s += red[wv];
      q[i] = s + head_b[0];
    }
    __syncthreads();
  }
}

// ---- DeepQ.cost (policy.py:234-239): loss = sum((q a - y)^2) * inv_batch, dq = 2 (q a - y) a * inv_batch --------------------------
__global__ void __launch_bounds__(1024) k_dqn_cost(int N, const float *__restrict__ q, const float *__restrict__ a, const float *__restrict__ y, float inv_batch,
                                                   float *__restrict__ dq, float *__restrict__ loss) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < N; i += 1024) {
    const float av = a ? a[i] : 1.f, d = q[i] * av - (y ? y[i] : 0.f);
    s += d * d;
    dq[i] = 2.f * d * av * inv_batch;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = red[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) loss[0] = s * inv_batch;
  }
}

// ---- column reductions over the nodes -------------------------------------------------------------------------------------
// MODE 0 (head):     gWh[c] = sum_n dq[n] d2[n,c];  gbh = sum_n dq[n];  gb2[c] = Wh[c] s sum_n dq[n] [d2[n,c] != 0]
// MODE 1 (layer 1):  gb1[c] = sum_n dz1[n,c];  gW1[k,c] = sum_n ax[n,k] dz1[n,c]   with dz1 = dh1 * [h1 > 0]
// Row slabs (n = slab, slab + NSLAB, ...) -> partials; k_colreduce_final adds the partials in slab order.
template <int MODE>
__global__ void __launch_bounds__(256) k_colreduce(int N, int C, int Cin, const float *__restrict__ dq, const float *__restrict__ A /*d2 | dh1*/,
                                                   const float *__restrict__ Bm /*- | h1_hi*/, const float *__restrict__ ax, const float *__restrict__ Wh, float scale,
                                                   float *__restrict__ part /*[NSLAB][NV][C]*/, unsigned *__restrict__ counter,
                                                   float *__restrict__ o0 /*gWh | gb1*/, float *__restrict__ o1 /*gb2 | gW1 [Cin,C]*/, float *__restrict__ o2 /*gbh | -*/) {
  constexpr int NV = MODE == 0 ? 2 : 9;      // accumulated vectors per column
  const int c0 = threadIdx.x * 4;
  const bool act = c0 < C;
  float acc[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[v][j] = 0.f;
  float sdq = 0.f;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    if (MODE == 0) {
      const float d = dq[n];
      if (d == 0.f) continue;             // DQN: only the action nodes carry a gradient
      sdq += d;
      if (act) {
        const float4 x = *reinterpret_cast<const float4 *>(A + (size_t)n * C + c0);
        acc[0][0] += d * x.x; acc[0][1] += d * x.y; acc[0][2] += d * x.z; acc[0][3] += d * x.w;
        acc[1][0] += x.x != 0.f ? d : 0.f; acc[1][1] += x.y != 0.f ? d : 0.f; acc[1][2] += x.z != 0.f ? d : 0.f; acc[1][3] += x.w != 0.f ? d : 0.f;
      }
    } else if (act) {
      const float4 g = *reinterpret_cast<const float4 *>(A + (size_t)n * C + c0);
      const float4 m = *reinterpret_cast<const float4 *>(Bm + (size_t)n * C + c0);
      const float z[4] = {m.x > 0.f ? g.x : 0.f, m.y > 0.f ? g.y : 0.f, m.z > 0.f ? g.z : 0.f, m.w > 0.f ? g.w : 0.f};
      const float4 a0 = *reinterpret_cast<const float4 *>(ax + (size_t)n * 8), a1 = *reinterpret_cast<const float4 *>(ax + (size_t)n * 8 + 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[0][j] += z[j];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[1 + k][j] += av[k] * z[j];
      }
    }
  }
  float *mine = part + (size_t)blockIdx.x * NV * C;
  if (act)
#pragma unroll
    for (int v = 0; v < NV; ++v) *reinterpret_cast<float4 *>(mine + (size_t)v * C + c0) = make_float4(acc[v][0], acc[v][1], acc[v][2], acc[v][3]);
  if (MODE == 0 && threadIdx.x == 0) part[(size_t)gridDim.x * NV * C + blockIdx.x] = sdq;
}

// second stage: one thread per (vector, column) adds the slabs' partials in slab order (deterministic; coalesced across columns)
template <int MODE>
__global__ void __launch_bounds__(256) k_colreduce_final(int C, int Cin, int nslab, const float *__restrict__ part, const float *__restrict__ Wh, float scale,
                                                         float *__restrict__ o0, float *__restrict__ o1, float *__restrict__ o2) {
  constexpr int NV = MODE == 0 ? 2 : 9;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < NV * C) {
    const int v = idx / C, c = idx - v * C;
    float s = 0.f;
    for (int b = 0; b < nslab; ++b) s += part[((size_t)b * NV + v) * C + c];
    if (MODE == 0) {
      if (v == 0) o0[c] = s; else o1[c] = s * Wh[c] * scale;
    } else {
      if (v == 0) o0[c] = s; else if (v - 1 < Cin) o1[(size_t)(v - 1) * C + c] = s;
    }
  }
  if (MODE == 0 && idx == 0) {
    float s = 0.f;
    for (int b = 0; b < nslab; ++b) s += part[(size_t)nslab * NV * C + b];
    o2[0] = s;
  }
}

// ---- g2 = A^T dz2, dz2[n,c] = dq[n] s Wh[c] [d2[n,c] != 0] formed on the fly; gather over the source-sorted CSR --------------------
__global__ void __launch_bounds__(256) k_agg_bwd_train(int N, int C, const float *__restrict__ d2, const float *__restrict__ dq, const float *__restrict__ Wh,
                                                       float scale, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ perm,
                                                       const int64_t *__restrict__ nbr, const float *__restrict__ coef, const float *__restrict__ selfcoef,
                                                       float *__restrict__ g2) {
  const int c0 = threadIdx.x * 4;
  const bool act = c0 < C;
  float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
  if (act) { w = *reinterpret_cast<const float4 *>(Wh + c0); w.x *= scale; w.y *= scale; w.z *= scale; w.w *= scale; }
  for (int i = blockIdx.x; i < N; i += gridDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto add_row = [&](int n, float cf) {
      const float d = dq[n] * cf;
      if (d == 0.f || !act) return;       // (warp-uniform: dq, cf are per row)
      const float4 x = *reinterpret_cast<const float4 *>(d2 + (size_t)n * C + c0);
      acc.x += x.x != 0.f ? d * w.x : 0.f; acc.y += x.y != 0.f ? d * w.y : 0.f; acc.z += x.z != 0.f ? d * w.z : 0.f; acc.w += x.w != 0.f ? d * w.w : 0.f;
    };
    for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) { const int e = perm[p]; add_row((int)nbr[e], coef[e]); }
    add_row(i, selfcoef[i]);
    if (act) *reinterpret_cast<float4 *>(g2 + (size_t)i * C + c0) = acc;
  }
}

// ---- clamp + Adam on the flat buffers (policy.py:251-253; torch.optim.Adam's arithmetic, amsgrad off, no weight decay) ---------------
__global__ void __launch_bounds__(256) k_clamp_adam(int64_t n, float *__restrict__ p, float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                                                    const int64_t *__restrict__ step, float lr, float b1, float b2, float eps, float clamp, float gscale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = (double)step[0];
  const float bc1 = (float)(1.0 - pow((double)b1, t)), bc2s = (float)sqrt(1.0 - pow((double)b2, t));
  float gi = g[i] * gscale;
  if (clamp > 0.f) gi = fminf(fmaxf(gi, -clamp), clamp);
  g[i] = gi;                                   // the gradient Adam saw stays readable (tests, logging)
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2s + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}
__global__ void k_step_inc(int64_t *step) { step[0] += 1; }

inline int node_grid(int N) { return N < 148 * 16 ? N : 148 * 16; }
#define CKL() do { if (cudaGetLastError() != cudaSuccess) return -2; } while (0)

}  // namespace

extern "C" int64_t dge_gcn_train_ws_floats(int N, int C) {
  const int64_t Np = (N + 3) & ~3, NC = (int64_t)N * C;
  // ax | h1 | h1_hi | h1_lo | t2 (later dh1) | d2 | g2 | g2_hi | g2_lo | h1t_hi | h1t_lo | g2t_hi | g2t_lo | dq | partials | counters
  return 8 * (int64_t)N + 9 * NC + 4 * (int64_t)C * Np + (((int64_t)N + 3) & ~3) + (int64_t)NSLAB * 9 * C + NSLAB + 64;
}

// One training step's forward + backward.  CSRs: (rowptr_d, perm_d) rows = destination (forward gather, neighbour = src),
// (rowptr_s, perm_s) rows = source (backward gather, neighbour = dst).  W2t_(hi,lo) = (W2^T) split [C_out, C_in] (forward operand),
// W2_(hi,lo) = W2 split as stored [C_in, C_out] (grad-input operand).  Gradients are WRITTEN to gW1 [Cin,C], gb1 [C], gW2 [C,C],
// gb2 [C], gWh [C], gbh [1]; loss [1] and q [N] stay on the device.  drop_p = 0 disables the dropout.
extern "C" int dge_gcn_train_step(int N, int Cin, int C, const float *x, const int32_t *rowptr_d, const int32_t *perm_d, const int32_t *rowptr_s,
                                  const int32_t *perm_s, const int64_t *src, const int64_t *dst, const float *norm, const float *selfnorm,
                                  const float *W1, const float *b1, const float *W2t_hi, const float *W2t_lo, const float *W2_hi, const float *W2_lo,
                                  const float *b2, const float *Wh, const float *bh, const float *act, const float *y, float inv_batch, float drop_p,
                                  uint64_t drop_seed, float *gW1, float *gb1, float *gW2, float *gb2, float *gWh, float *gbh, float *loss, float *q,
                                  float *ws, void *stream) {
  if (N <= 0 || Cin <= 0 || Cin > 8 || C <= 0 || C > 1024 || (C & 3) || !x || !rowptr_d || !perm_d || !rowptr_s || !perm_s || !selfnorm || !W1 || !b1 ||
      !W2t_hi || !W2t_lo || !W2_hi || !W2_lo || !b2 || !Wh || !bh || !gW1 || !gb1 || !gW2 || !gb2 || !gWh || !gbh || !loss || !q || !ws || ((uintptr_t)ws & 15) ||
      drop_p < 0.f || drop_p >= 1.f)
    return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t Np = (N + 3) & ~3, NC = (int64_t)N * C;
  float *ax = ws, *h1 = ax + 8 * (int64_t)N, *h1_hi = h1 + NC, *h1_lo = h1_hi + NC, *t2 = h1_lo + NC, *d2 = t2 + NC, *g2 = d2 + NC, *g2_hi = g2 + NC,
        *g2_lo = g2_hi + NC, *h1t_hi = g2_lo + NC, *h1t_lo = h1t_hi + (int64_t)C * Np, *g2t_hi = h1t_lo + (int64_t)C * Np, *g2t_lo = g2t_hi + (int64_t)C * Np,
        *dq = g2t_lo + (int64_t)C * Np, *part = dq + ((N + 3) & ~3);
  unsigned *counter = reinterpret_cast<unsigned *>(part + (int64_t)NSLAB * 9 * C + NSLAB);
  float *dh1 = t2;
  const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  (void)counter;
  // ---- forward
  k_conv1_train<<<node_grid(N), 256, 0, st>>>(N, Cin, C, x, rowptr_d, perm_d, src, norm, selfnorm, W1, b1, ax, h1); CKL();
  int rc = dge_gemm_split_tf32(NC, h1, h1_hi, h1_lo, st);                                             // (one split serves h1 W2 and, as stored, h1^T g2)
  if (rc) return rc;
  rc = dge_gemm_tf32x3_ex(N, nullptr, C, C, h1_hi, h1_lo, 0, W2t_hi, W2t_lo, 0, t2, C, 1, st);
  if (rc) return rc;
  rc = dge_agg_fwd_train_bulk(N, C, t2, rowptr_d, perm_d, src, norm, selfnorm, b2, drop_p, drop_seed, Wh, bh, d2, q, st);   // (opt-in, DGE_AGG_BULK=1)
  if (rc == -1) { k_agg_fwd_train<<<node_grid(N), 256, 0, st>>>(N, C, t2, rowptr_d, perm_d, src, norm, selfnorm, b2, drop_p, drop_seed, Wh, bh, d2, q); CKL(); }
  else if (rc) return rc;
  // ---- cost
  k_dqn_cost<<<1, 1024, 0, st>>>(N, q, act, y, inv_batch, dq, loss); CKL();
  // ---- backward
  k_colreduce<0><<<NSLAB, 256, 0, st>>>(N, C, Cin, dq, d2, nullptr, nullptr, Wh, scale, part, counter, gWh, gb2, gbh); CKL();
  k_colreduce_final<0><<<(2 * C + 255) / 256, 256, 0, st>>>(C, Cin, NSLAB, part, Wh, scale, gWh, gb2, gbh); CKL();
  k_agg_bwd_train<<<node_grid(N), 256, 0, st>>>(N, C, d2, dq, Wh, scale, rowptr_s, perm_s, dst, norm, selfnorm, g2); CKL();
  rc = dge_gemm_split_tf32(NC, g2, g2_hi, g2_lo, st);
  if (rc) return rc;
  rc = dge_gemm_tf32x3_ex(N, nullptr, C, C, g2_hi, g2_lo, 0, W2_hi, W2_lo, 0, dh1, C, 1, st);            // dh1 = g2 W2^T
  if (rc) return rc;
  if (cudaMemsetAsync(gW2, 0, sizeof(float) * (size_t)C * C, st) != cudaSuccess) return -2;
  rc = dge_gemm_tf32x3_tn(C, C, N, h1_hi, h1_lo, C, g2_hi, g2_lo, C, gW2, C, 0, st);                     // dW2 = h1^T g2, K = nodes, operands as stored (MN-major)
  if (rc) return rc;
  k_colreduce<1><<<NSLAB, 256, 0, st>>>(N, C, Cin, nullptr, dh1, h1_hi, ax, nullptr, 1.f, part, counter + 1, gb1, gW1, nullptr); CKL();
  k_colreduce_final<1><<<(9 * C + 255) / 256, 256, 0, st>>>(C, Cin, NSLAB, part, nullptr, 1.f, gb1, gW1, nullptr); CKL();
  return 0;
}

// X [N,C] -> (hi, lo) [N,C] (nullable pair) and (thi, tlo) [C,Np]: the operand forms of the tcgen05 GEMM for autograd's mm backward
// (gnn._TcMatmulFn / _TcLinearFn: every family's dense layers under training, not only the fused GCN step above)
extern "C" int dge_gemm_split_transpose(int N, int C, const float *X, float *hi, float *lo, float *thi, float *tlo, void *stream) {
  if (N <= 0 || C <= 0 || !X || !thi || !tlo || (!hi != !lo)) return -1;
  const int64_t Np = ((int64_t)N + 3) & ~3;
  const dim3 tg((C + 31) / 32, (unsigned)((Np + 31) / 32));
  k_split_transpose<<<tg, 256, 0, static_cast<cudaStream_t>(stream)>>>(N, C, (int)Np, X, hi, lo, thi, tlo);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// p -= Adam(clamp(g * gscale)) on flat buffers of n floats; step [1] int64 on the device is incremented first (t = 1 on the first call)
extern "C" int dge_clamp_adam_step(int64_t n, float *p, float *g, float *m, float *v, int64_t *step, float lr, float beta1, float beta2, float eps,
                                   float clamp, float gscale, void *stream) {
  if (n <= 0 || !p || !g || !m || !v || !step) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_step_inc<<<1, 1, 0, st>>>(step);
  k_clamp_adam<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, p, g, m, v, step, lr, beta1, beta2, eps, clamp, gscale);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
