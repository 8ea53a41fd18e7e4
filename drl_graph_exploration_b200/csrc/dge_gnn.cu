// GNN message-passing kernels for the graph Q-network (scripts/Networks.py, rows a16/a17).
//
// The reference runs PyG's GCNConv: X@W (cuBLAS), add-remaining-self-loops, weighted-degree
// scatter_add, per-edge gather*scale and a scatter_add with atomics over [E+N, 1000] floats
// (torch_scatter).  Here the edge list is turned once per batch into destination-sorted CSR
// (deterministic: rows sorted by edge id) and the aggregation becomes a gather:
// one CTA per destination node, 250 threads x float4 = one 1000-channel row, neighbours'
// rows streamed with 16-byte coalesced loads, self-loop / bias / ReLU (and, for the output
// layer, the Linear(1000,1) head) fused in the epilogue -- no atomics, no [E,1000] temporary.
// The dense X@W / grad GEMMs are tensor-core GEMMs outside this file.
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dge_gnn.h"

namespace {

// ----------------------------------------------------------------- CSR build ---
__global__ void k_count(int E, const int64_t *key, int32_t *cnt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) atomicAdd(&cnt[(int)key[e]], 1);
}

// single-CTA exclusive scan, tile by tile: coalesced loads of 1024 counts, shuffle scan inside the warps, one shared-memory
// pass over the 32 warp totals, running carry in a register; the next tile's load is issued before the current tile's scan
// (N is at most a few 100k nodes: a handful of tiles, latency-bound -- a multi-CTA decoupled look-back would not pay).
__global__ void __launch_bounds__(1024) k_scan(int N, const int32_t *cnt, int32_t *rowptr) {
  __shared__ int32_t wsum[32];
  __shared__ int32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int carry = 0;
  int nxt = (tid < N) ? cnt[tid] : 0;
  for (int base = 0; base < N; base += 1024) {
    const int v = nxt;
    const int nb = base + 1024 + tid;
    nxt = (nb < N) ? cnt[nb] : 0;
    int x = v;                                   // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
      wsum[lane] = w;                            // inclusive totals of the warps
      if (lane == 31) s_carry = w;
    }
    __syncthreads();
    const int before = carry + (warp ? wsum[warp - 1] : 0) + x - v;
    if (base + tid < N) rowptr[base + tid] = before;
    carry += s_carry;
    __syncthreads();
  }
  if (tid == 0) rowptr[N] = carry;
}

__global__ void k_fill(int E, const int64_t *key, const int32_t *rowptr, int32_t *cursor, int32_t *perm) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int r = (int)key[e];
  perm[rowptr[r] + atomicAdd(&cursor[r], 1)] = e;
}

// deterministic summation order: every edge finds its rank inside its row (rows are short,
// degree ~2..T), one thread per edge, and lands at rowptr[row] + rank  => rows sorted by edge id.
__global__ void k_rank_rows(int E, const int64_t *key, const int32_t *rowptr, const int32_t *tmp, int32_t *perm) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int r = (int)key[e];
  const int lo = rowptr[r], hi = rowptr[r + 1];
  int rank = 0;
  for (int q = lo; q < hi; ++q) rank += (tmp[q] < e) ? 1 : 0;
  perm[lo + rank] = e;
}

// ------------------------------------------------------------------ GCN norm ---
// GCNConv.norm (PyG 1.x, improved=True): add *remaining* self loops with weight `fill`,
// deg_i = sum of the weights of the edges whose source is i, norm_e = deg^-1/2[src] w deg^-1/2[dst].
__global__ void __launch_bounds__(256) k_gcn_deg(int N, const int32_t *rowptr_src, const int32_t *perm_src, const int64_t *src, const int64_t *dst,
                                                 const float *w, float fill, float *dis, float *selfw) {
  // one warp per node: lanes stride the node's out-edges (coalesced reads of the row), partial sums combined by a fixed
  // shuffle tree => deterministic whatever the edge count
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  float deg = 0.f, loopw = 0.f;
  int has_loop = 0;
  for (int p = rowptr_src[i] + lane; p < rowptr_src[i + 1]; p += 32) {
    const int e = perm_src[p];
    if ((int)dst[e] == i) { has_loop = 1; loopw = w[e]; continue; }   // existing loops are re-appended after the plain edges
    deg += w[e];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    deg += __shfl_xor_sync(0xffffffffu, deg, o);
    const float lw2 = __shfl_xor_sync(0xffffffffu, loopw, o);
    const int hl2 = __shfl_xor_sync(0xffffffffu, has_loop, o);
    if (hl2 && !has_loop) loopw = lw2;            // (a node has at most one self loop in these graphs; the last one wins like before)
    has_loop |= hl2;
  }
  if (lane == 0) {
    const float lw = has_loop ? loopw : fill;
    deg += lw;
    dis[i] = (deg > 0.f) ? 1.0f / sqrtf(deg) : 0.f;   // deg^-0.5 with inf -> 0
    selfw[i] = lw;
  }
}
__global__ void k_gcn_norm(int E, const int64_t *src, const int64_t *dst, const float *w, const float *dis, float *norm) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int s = (int)src[e], d = (int)dst[e];
  norm[e] = (s == d) ? 0.f : dis[s] * w[e] * dis[d];   // self loops are carried by selfnorm
}
__global__ void k_gcn_selfnorm(int N, const float *dis, const float *selfw, float *selfnorm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) selfnorm[i] = dis[i] * selfw[i] * dis[i];
}

// ----------------------------------------------------------------- aggregate ---
// out[i,:] = act( bias + selfcoef[i]*X[i,:] + sum_{p in row i} coef[perm[p]] * X[nbr[perm[p]],:] )
// optional: multiply gathered rows by (gate[n,:] > 0) (ReLU backward fused into the gather)
// optional head: q[i] = dot(out[i,:], head_w) + head_b   (Linear(1000,1) fused; out may be null)
template <bool V4>
__global__ void __launch_bounds__(256) k_aggregate(int N, int C, const float *__restrict__ X, const int32_t *__restrict__ rowptr,
                                                   const int32_t *__restrict__ perm, const int64_t *__restrict__ nbr,
                                                   const float *__restrict__ coef, const float *__restrict__ selfcoef,
                                                   const float *__restrict__ bias, const float *__restrict__ gate, int relu,
                                                   float *__restrict__ out, const float *__restrict__ head_w, float head_b,
                                                   float *__restrict__ q, const float *__restrict__ head_b_dev = nullptr,
                                                   const int32_t *__restrict__ N_dev = nullptr) {
  constexpr int VEC = 4;
  const int c0 = threadIdx.x * VEC;
  const bool act = c0 < C;
  const int n_live = N_dev ? min(N, *N_dev) : N;     // a batch sized on the device: the grid is a capacity, the loop bound is live
  for (int i = blockIdx.x; i < n_live; i += gridDim.x) {
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  auto add_row = [&](int n, float cf) {
    if (!act) return;
    if (V4) {
      float4 x = *reinterpret_cast<const float4 *>(X + (size_t)n * C + c0);
      if (gate) {
        const float4 g = *reinterpret_cast<const float4 *>(gate + (size_t)n * C + c0);
        x.x = g.x > 0.f ? x.x : 0.f; x.y = g.y > 0.f ? x.y : 0.f; x.z = g.z > 0.f ? x.z : 0.f; x.w = g.w > 0.f ? x.w : 0.f;
      }
      acc[0] += cf * x.x; acc[1] += cf * x.y; acc[2] += cf * x.z; acc[3] += cf * x.w;
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (c0 + v < C) {
          float x = X[(size_t)n * C + c0 + v];
          if (gate && !(gate[(size_t)n * C + c0 + v] > 0.f)) x = 0.f;
          acc[v] += cf * x;
        }
      }
    }
  };
  const int lo = rowptr[i], hi = rowptr[i + 1];
  int p = lo;
  if (V4 && !gate) {
    for (; p + 7 < hi; p += 8) {   // eight neighbour rows in flight: indices and coefficients first, then the row loads back to back
      int nn[8]; float ff[8]; float4 xx[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int e = perm[p + u]; nn[u] = (int)nbr[e]; ff[u] = coef[e]; }
      if (act) {
#pragma unroll
        for (int u = 0; u < 8; ++u) xx[u] = *reinterpret_cast<const float4 *>(X + (size_t)nn[u] * C + c0);
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc[0] += ff[u] * xx[u].x; acc[1] += ff[u] * xx[u].y; acc[2] += ff[u] * xx[u].z; acc[3] += ff[u] * xx[u].w; }
      }
    }
  }
  for (; p + 1 < hi; p += 2) {   // two neighbour rows in flight
    const int e0 = perm[p], e1 = perm[p + 1];
    const int n0 = (int)nbr[e0], n1 = (int)nbr[e1];
    const float f0 = coef[e0], f1 = coef[e1];
    add_row(n0, f0);
    add_row(n1, f1);
  }
  if (p < hi) { const int e0 = perm[p]; add_row((int)nbr[e0], coef[e0]); }
  if (selfcoef) add_row(i, selfcoef[i]);
  float hq = 0.f;
  if (act) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (c0 + v < C) {
        float r = acc[v] + (bias ? bias[c0 + v] : 0.f);
        if (relu) r = fmaxf(r, 0.f);
        acc[v] = r;
        if (head_w) hq += r * head_w[c0 + v];
      }
    }
    if (out) {
      if (V4) *reinterpret_cast<float4 *>(out + (size_t)i * C + c0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      else
        for (int v = 0; v < VEC; ++v) if (c0 + v < C) out[(size_t)i * C + c0 + v] = acc[v];
    }
  }
  if (head_w) {   // block reduction of the head dot product (fixed order => deterministic)
    __shared__ float red[8];
    for (int o = 16; o > 0; o >>= 1) hq += __shfl_xor_sync(0xffffffffu, hq, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = hq;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int wv = 0; wv < 8; ++wv) s += red[wv];
      q[i] = s + head_b + (head_b_dev ? head_b_dev[0] : 0.f);
    }
    __syncthreads();   // `red` is reused by the next node of this CTA
  }
  }
}

// ------------------------------------------------------------ aggregate, bulk-async staged ---
// The same gather (GCNConv.propagate, Networks.py:22-24 via PyG) with the neighbour rows STAGED: a producer warp streams every
// neighbour's feature row (C floats = 4 KB for C = 1000) into a shared-memory ring with one bulk asynchronous copy per row
// (cp.async.bulk global -> shared, completion counted in bytes on the slot's mbarrier); the eight consumer warps wait on the
// slot, accumulate coef * row in registers (float4 per thread) and hand the slot back.  The producer runs ahead across node
// boundaries -- RING rows (32 KB) are in flight per CTA whatever the consumers do, instead of the two loads per thread of the
// direct-load kernel (ncu r01: 28 % of DRAM peak, 17.9 long-scoreboard stalls per issue).  Several destination rows per CTA
// (grid-stride); epilogue as above (self loop = one more staged row, bias, ReLU, optional dropout, optional head).
constexpr int AGG_GROUP = 4;          // rows per ring slot: one mbarrier round trip per AGG_GROUP rows
constexpr int AGG_RING = 4, AGG_CONSUMERS = 256, AGG_THREADS = AGG_CONSUMERS + 32;   // AGG_RING slots x AGG_GROUP rows in flight per CTA
__device__ __forceinline__ uint32_t agg_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void agg_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ uint4 agg_philox(uint64_t key, uint64_t ctr) {   // Philox4x32-10 (as csrc/dge_train.cu)
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
__global__ void __launch_bounds__(AGG_THREADS) k_aggregate_bulk(int N, int C, const float *__restrict__ X, const int32_t *__restrict__ rowptr,
                                                                const int32_t *__restrict__ perm, const int64_t *__restrict__ nbr,
                                                                const float *__restrict__ coef, const float *__restrict__ selfcoef,
                                                                const float *__restrict__ bias, int relu, float drop_p, uint64_t drop_seed,
                                                                float *__restrict__ out, const float *__restrict__ head_w, float head_b,
                                                                float *__restrict__ q, const float *__restrict__ head_b_dev,
                                                                const int32_t *__restrict__ N_dev) {
  extern __shared__ __align__(16) unsigned char agg_smem[];
  float *ring = reinterpret_cast<float *>(agg_smem);                                   // [AGG_RING][AGG_GROUP][C]
  uint64_t *bars = reinterpret_cast<uint64_t *>(agg_smem + (size_t)AGG_RING * AGG_GROUP * C * sizeof(float));   // full[RING], empty[RING]
  __shared__ float red[8];
  const int n_live = N_dev ? min(N, *N_dev) : N;
  const uint32_t full0 = agg_smem_u32(bars), empty0 = full0 + 8 * AGG_RING;
  if (threadIdx.x == 0) {
    for (int s = 0; s < AGG_RING; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * s), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8 * s), "r"(AGG_CONSUMERS / 32) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t row_bytes = (uint32_t)C * sizeof(float);
  if (threadIdx.x >= AGG_CONSUMERS) {
    // ---- producer: one lane issues the bulk copies, in exactly the order the consumers fold the rows
    if (threadIdx.x == AGG_CONSUMERS) {
      uint32_t it = 0;
      for (int i = blockIdx.x; i < n_live; i += gridDim.x) {
        const int lo = rowptr[i], hi = rowptr[i + 1];
        const int nrows = hi - lo + (selfcoef ? 1 : 0);                                // the last one: the node's own row (self loop)
        for (int r0 = 0; r0 < nrows; r0 += AGG_GROUP, ++it) {                          // a slot takes up to AGG_GROUP rows of ONE node
          const int g = min(AGG_GROUP, nrows - r0);
          const uint32_t s = it % AGG_RING, ph = (it / AGG_RING) & 1;
          agg_mbar_wait(empty0 + 8 * s, ph ^ 1);                                       // (free at first use)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8 * s), "r"(row_bytes * (uint32_t)g) : "memory");
          for (int u = 0; u < g; ++u) {
            const int p = lo + r0 + u;
            const int n = p < hi ? (int)nbr[perm[p]] : i;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(agg_smem_u32(ring + ((size_t)s * AGG_GROUP + u) * C)), "l"(X + (size_t)n * C), "r"(row_bytes), "r"(full0 + 8 * s) : "memory");
          }
        }
      }
    }
    return;
  }
  // ---- consumers
  const int c0 = threadIdx.x * 4, lane = threadIdx.x & 31;
  const bool act = c0 < C;
  const float dscale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  uint32_t it = 0;
  for (int i = blockIdx.x; i < n_live; i += gridDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int lo = rowptr[i], hi = rowptr[i + 1];
    const int nrows = hi - lo + (selfcoef ? 1 : 0);
    for (int r0 = 0; r0 < nrows; r0 += AGG_GROUP, ++it) {
      const int g = min(AGG_GROUP, nrows - r0);
      float cf[AGG_GROUP];
#pragma unroll
      for (int u = 0; u < AGG_GROUP; ++u) { const int p = lo + r0 + u; cf[u] = u < g ? (p < hi ? coef[perm[p]] : selfcoef[i]) : 0.f; }
      const uint32_t s = it % AGG_RING, ph = (it / AGG_RING) & 1;
      agg_mbar_wait(full0 + 8 * s, ph);
      if (act) {
#pragma unroll
        for (int u = 0; u < AGG_GROUP; ++u) {
          if (u < g) {
            const float4 x = *reinterpret_cast<const float4 *>(ring + ((size_t)s * AGG_GROUP + u) * C + c0);
            acc.x += cf[u] * x.x; acc.y += cf[u] * x.y; acc.z += cf[u] * x.z; acc.w += cf[u] * x.w;
          }
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + 8 * s) : "memory");
    }
    float hq = 0.f;
    if (act) {
      float r[4] = {acc.x, acc.y, acc.z, acc.w};
      if (bias) { const float4 b = *reinterpret_cast<const float4 *>(bias + c0); r[0] += b.x; r[1] += b.y; r[2] += b.z; r[3] += b.w; }
      if (relu) { r[0] = fmaxf(r[0], 0.f); r[1] = fmaxf(r[1], 0.f); r[2] = fmaxf(r[2], 0.f); r[3] = fmaxf(r[3], 0.f); }
      if (drop_p > 0.f) {   // functional dropout (Networks.py:26): keep with probability 1 - p, scale by 1 / (1 - p)
        const uint4 u = agg_philox(drop_seed, (uint64_t)i * 256 + threadIdx.x);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int v = 0; v < 4; ++v) r[v] = ((uu[v] >> 8) * (1.0f / 16777216.0f) >= drop_p) ? r[v] * dscale : 0.f;
      }
      if (out) *reinterpret_cast<float4 *>(out + (size_t)i * C + c0) = make_float4(r[0], r[1], r[2], r[3]);
      if (head_w) { const float4 w = *reinterpret_cast<const float4 *>(head_w + c0); hq = r[0] * w.x + r[1] * w.y + r[2] * w.z + r[3] * w.w; }
    }
    if (head_w) {   // reduction of the head dot product over the consumer warps (fixed order => deterministic); named barrier: the producer is not part of it
      for (int o = 16; o > 0; o >>= 1) hq += __shfl_xor_sync(0xffffffffu, hq, o);
      if (lane == 0) red[threadIdx.x >> 5] = hq;
      asm volatile("bar.sync 1, %0;" ::"r"(AGG_CONSUMERS) : "memory");
      if (threadIdx.x == 0) {
        float sacc = 0.f;
        for (int wv = 0; wv < 8; ++wv) sacc += red[wv];
        q[i] = sacc + head_b + (head_b_dev ? head_b_dev[0] : 0.f);
      }
      asm volatile("bar.sync 1, %0;" ::"r"(AGG_CONSUMERS) : "memory");
    }
  }
}
inline size_t agg_bulk_smem(int C) { return (size_t)AGG_RING * AGG_GROUP * C * sizeof(float) + 2 * AGG_RING * sizeof(uint64_t); }
// the bulk kernel takes rows of a multiple of 16 bytes at 16-byte aligned addresses (C % 4 == 0), up to 1024 channels
inline bool agg_bulk_ok(int C, const float *X, const float *out, const float *bias, const float *head_w) {
  // opt-in (DGE_AGG_BULK=1): measured SLOWER than the direct-load kernel with eight rows in flight at the 4 KB rows of this network
  // (profiles/r02_aggregate_ab.md) -- a row is one float4 per thread, so the per-slot mbarrier round trip is not amortised
  static const bool on = [] { const char *v = getenv("DGE_AGG_BULK"); return v && v[0] == '1'; }();
  if (!on) return false;
  return C % 4 == 0 && C <= 1024 && ((uintptr_t)X % 16 == 0) && (!out || (uintptr_t)out % 16 == 0) && (!bias || (uintptr_t)bias % 16 == 0) &&
         (!head_w || (uintptr_t)head_w % 16 == 0);
}
inline int agg_bulk_launch(int grid, int N, int C, const float *X, const int32_t *rowptr, const int32_t *perm, const int64_t *nbr, const float *coef,
                           const float *selfcoef, const float *bias, int relu, float drop_p, uint64_t drop_seed, float *out, const float *head_w,
                           float head_b, float *q, const float *head_b_dev, const int32_t *N_dev, cudaStream_t st) {
  const size_t smem = agg_bulk_smem(C);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(k_aggregate_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    configured = smem;
  }
  k_aggregate_bulk<<<grid, AGG_THREADS, smem, st>>>(N, C, X, rowptr, perm, nbr, coef, selfcoef, bias, relu, drop_p, drop_seed, out, head_w, head_b, q, head_b_dev, N_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
#define CK() (cudaGetLastError() == cudaSuccess ? 0 : -2)

}  // namespace

extern "C" int dge_gnn_csr_build(int N, int E, const int64_t *key, int32_t *rowptr, int32_t *perm, int32_t *ws /*[2N+E]*/, void *stream) {
  if (N <= 0 || E < 0 || (E > 0 && !key) || !rowptr || !perm || !ws) return -1;   // an edgeless batch is legal: every row empty
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(ws, 0, sizeof(int32_t) * 2 * (size_t)N, st) != cudaSuccess) return -2;
  if (E > 0) k_count<<<cdiv(E, 256), 256, 0, st>>>(E, key, ws);
  k_scan<<<1, 1024, 0, st>>>(N, ws, rowptr);
  if (E > 0) {
    int32_t *tmp = ws + 2 * (size_t)N;
    k_fill<<<cdiv(E, 256), 256, 0, st>>>(E, key, rowptr, ws + N, tmp);
    k_rank_rows<<<cdiv(E, 256), 256, 0, st>>>(E, key, rowptr, tmp, perm);
  }
  return CK();
}

extern "C" int dge_gcn_norm(int N, int E, const int64_t *src, const int64_t *dst, const float *w, const int32_t *rowptr_src,
                            const int32_t *perm_src, float fill, float *dis, float *selfw, float *norm, float *selfnorm, void *stream) {
  if (N <= 0) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_gcn_deg<<<cdiv(N, 8), 256, 0, st>>>(N, rowptr_src, perm_src, src, dst, w, fill, dis, selfw);
  if (E > 0) k_gcn_norm<<<cdiv(E, 256), 256, 0, st>>>(E, src, dst, w, dis, norm);
  k_gcn_selfnorm<<<cdiv(N, 256), 256, 0, st>>>(N, dis, selfw, selfnorm);
  return CK();
}

extern "C" int dge_gnn_aggregate(int N, int C, const float *X, const int32_t *rowptr, const int32_t *perm, const int64_t *nbr,
                                 const float *coef, const float *selfcoef, const float *bias, const float *gate, int relu, float *out,
                                 const float *head_w, float head_b, float *q, void *stream) {
  if (N <= 0 || C <= 0 || C > 1024 || !X || !rowptr || !perm) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!gate && agg_bulk_ok(C, X, out, bias, head_w))   // neighbour rows staged by bulk asynchronous copies
    return agg_bulk_launch(N < 148 * 4 ? N : 148 * 4, N, C, X, rowptr, perm, nbr, coef, selfcoef, bias, relu, 0.f, 0, out, head_w, head_b, q, nullptr, nullptr, st);
  if (C % 4 == 0 && ((uintptr_t)X % 16 == 0) && (!out || (uintptr_t)out % 16 == 0) && (!gate || (uintptr_t)gate % 16 == 0))
    k_aggregate<true><<<N, 256, 0, st>>>(N, C, X, rowptr, perm, nbr, coef, selfcoef, bias, gate, relu, out, head_w, head_b, q);
  else
    k_aggregate<false><<<N, 256, 0, st>>>(N, C, X, rowptr, perm, nbr, coef, selfcoef, bias, gate, relu, out, head_w, head_b, q);
  return CK();
}

// ------------------------------------------------- first GCN layer, fused -----------------------
// GCNConv(5 -> 1000): the aggregation is linear, so A_hat (X W) = (A_hat X) W.  With only 5 input
// channels it is cheaper to aggregate the 5-vectors first (20 B per neighbour instead of 4 KB) and apply
// W (5 x C, L1-resident) + bias + ReLU in the epilogue: no [N,5]x[5,C] GEMM launch, no [N,C] intermediate.
namespace {
template <int CIN_MAX>
__global__ void __launch_bounds__(256) k_gcn_conv_small(int N, int Cin, int C, const float *__restrict__ X, const int32_t *__restrict__ rowptr,
                                                        const int32_t *__restrict__ perm, const int64_t *__restrict__ nbr,
                                                        const float *__restrict__ coef, const float *__restrict__ selfcoef,
                                                        const float *__restrict__ W, const float *__restrict__ bias, int relu,
                                                        float *__restrict__ out, float *__restrict__ out_hi = nullptr,
                                                        float *__restrict__ out_lo = nullptr, const int32_t *__restrict__ N_dev = nullptr) {
  __shared__ float agg[CIN_MAX];
  const int n_live = N_dev ? min(N, *N_dev) : N;
  for (int i = blockIdx.x; i < n_live; i += gridDim.x) {
  if (threadIdx.x < 32) {   // one warp aggregates the input row: lanes over edges (fixed order per lane, shuffle tree => deterministic)
    float a[CIN_MAX];
#pragma unroll
    for (int k = 0; k < CIN_MAX; ++k) a[k] = 0.f;
    const int lo = rowptr[i], hi = rowptr[i + 1];
    for (int p = lo + threadIdx.x; p < hi; p += 32) {
      const int e = perm[p];
      const int n = (int)nbr[e];
      const float cf = coef[e];
#pragma unroll
      for (int k = 0; k < CIN_MAX; ++k) if (k < Cin) a[k] += cf * X[(size_t)n * Cin + k];
    }
#pragma unroll
    for (int k = 0; k < CIN_MAX; ++k)
      for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (threadIdx.x == 0) {
      const float sc = selfcoef ? selfcoef[i] : 0.f;
#pragma unroll
      for (int k = 0; k < CIN_MAX; ++k) agg[k] = (k < Cin) ? a[k] + sc * X[(size_t)i * Cin + k] : 0.f;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float r = bias ? bias[c] : 0.f;
#pragma unroll
    for (int k = 0; k < CIN_MAX; ++k) if (k < Cin) r += agg[k] * W[(size_t)k * C + c];
    if (relu) r = fmaxf(r, 0.f);
    if (out) out[(size_t)i * C + c] = r;
    if (out_hi) {   // operand of the 3xTF32 tensor-core GEMM that follows: (tf32(r), tf32(r - tf32(r))), as dge_gemm_split_tf32
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(r));
      const float ho = __uint_as_float(hb);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(r - ho));
      out_hi[(size_t)i * C + c] = ho;
      out_lo[(size_t)i * C + c] = __uint_as_float(lb);
    }
  }
  __syncthreads();   // `agg` is reused by the next node of this CTA
  }
}
}  // namespace

extern "C" int dge_gcn_conv_small(int N, int Cin, int C, const float *X, const int32_t *rowptr, const int32_t *perm, const int64_t *nbr,
                                  const float *coef, const float *selfcoef, const float *W, const float *bias, int relu, float *out, void *stream) {
  if (N <= 0 || Cin <= 0 || Cin > 8 || C <= 0 || !X || !rowptr || !perm || !W || !out) return -1;
  k_gcn_conv_small<8><<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(N, Cin, C, X, rowptr, perm, nbr, coef, selfcoef, W, bias, relu, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---------------------------------------------- whole Q-network forward (inference), one call ------------------
// Networks.GCN.forward with prob = 0 (Networks.py:18-28): relu(GCNConv(5,C)) -> relu(GCNConv(C,C)) -> Linear(C,1), as
// three launches: first layer fused (aggregate 5 channels, transform, ReLU, TF32 split in the epilogue) -> tcgen05
// 3xTF32 GEMM -> aggregate + bias + ReLU + head dot product.  ws = 3 * N * C floats (h_hi | h_lo | h W2).
// grid of a node-parallel kernel: one CTA per node up to a few waves of the machine, then grid-stride
static inline int node_grid(int N) { return N < 148 * 16 ? N : 148 * 16; }

// N_dev (nullable): the live node count on the device (<= N, which then is the capacity the launch is sized for) -- the form the
// sync-free acting loop uses (dge_policy_tick, csrc/dge_tick.cu)
int dge_gcn_q_forward_dev(int N, const int32_t *N_dev, int Cin, int C, const float *x, const int32_t *rowptr, const int32_t *perm, const int64_t *src,
                          const float *norm, const float *selfnorm, const float *W1, const float *b1, const float *W2t_hi,
                          const float *W2t_lo, const float *b2, const float *head_w, const float *head_b_dev, float *ws, float *q,
                          cudaStream_t st) {
  if (N <= 0 || Cin <= 0 || Cin > 8 || C <= 0 || C > 1024 || (C & 3) || !x || !rowptr || !perm || !selfnorm || !W1 || !W2t_hi ||
      !W2t_lo || !head_w || !ws || !q || ((uintptr_t)ws & 15))
    return -1;
  float *h_hi = ws, *h_lo = ws + (size_t)N * C, *xw = ws + 2 * (size_t)N * C;
  k_gcn_conv_small<8><<<node_grid(N), 256, 0, st>>>(N, Cin, C, x, rowptr, perm, src, norm, selfnorm, W1, b1, 1, nullptr, h_hi, h_lo, N_dev);
  if (cudaGetLastError() != cudaSuccess) return -2;
  const int rc = dge_gemm_tf32x3(N, N_dev, C, C, h_hi, h_lo, W2t_hi, W2t_lo, xw, C, st);
  if (rc) return rc;
  if (agg_bulk_ok(C, xw, nullptr, b2, head_w))
    return agg_bulk_launch(N < 148 * 4 ? N : 148 * 4, N, C, xw, rowptr, perm, src, norm, selfnorm, b2, 1, 0.f, 0, nullptr, head_w, 0.f, q, head_b_dev, N_dev, st);
  k_aggregate<true><<<node_grid(N), 256, 0, st>>>(N, C, xw, rowptr, perm, src, norm, selfnorm, b2, nullptr, 1, nullptr, head_w, 0.f, q, head_b_dev, N_dev);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int dge_gcn_q_forward(int N, int Cin, int C, const float *x, const int32_t *rowptr, const int32_t *perm, const int64_t *src,
                                 const float *norm, const float *selfnorm, const float *W1, const float *b1, const float *W2t_hi,
                                 const float *W2t_lo, const float *b2, const float *head_w, const float *head_b_dev, float *ws, float *q,
                                 void *stream) {
  return dge_gcn_q_forward_dev(N, nullptr, Cin, C, x, rowptr, perm, src, norm, selfnorm, W1, b1, W2t_hi, W2t_lo, b2, head_w, head_b_dev, ws, q,
                               static_cast<cudaStream_t>(stream));
}

// The same forward from a raw PyG edge list, in ONE call: destination- and source-sorted CSR (dge_gnn_csr_build x 2), the improved-GCN
// normalisation (dge_gcn_norm) and dge_gcn_q_forward -- what Networks.GCN.forward(data, 0) does for a batch that arrives as
// (x, edge_index, edge_attr) (DeepQ.test, policy.py:255-259), without four round trips through the caller's language.
// iws: 4 * (N + 1) + 2 * E + 2 * (2 * N + E) int32 scratch; fws: 3 * N + E + 3 * N * C floats (16-byte aligned).
extern "C" int64_t dge_gcn_q_forward_coo_iws(int N, int E) { return 4 * ((int64_t)N + 1) + 2 * (int64_t)E + 2 * (2 * (int64_t)N + E) + 8; }
extern "C" int64_t dge_gcn_q_forward_coo_fws(int N, int E, int C) { return 4 * (((int64_t)N + 3) & ~3) + (((int64_t)E + 3) & ~3) + 3 * (int64_t)N * C + 8; }
extern "C" int dge_gcn_q_forward_coo(int N, int E, int Cin, int C, const float *x, const int64_t *src, const int64_t *dst, const float *w,
                                     const float *W1, const float *b1, const float *W2t_hi, const float *W2t_lo, const float *b2, const float *head_w,
                                     const float *head_b_dev, int32_t *iws, float *fws, float *q, void *stream) {
  if (N <= 0 || E < 0 || !x || (E > 0 && (!src || !dst || !w)) || !iws || !fws || !q || ((uintptr_t)fws & 15)) return -1;
  int32_t *rowptr_d = iws, *rowptr_s = rowptr_d + (N + 1), *perm_d = rowptr_s + (N + 1), *perm_s = perm_d + (E > 0 ? E : 1), *ws_d = perm_s + (E > 0 ? E : 1),
          *ws_s = ws_d + (2 * (size_t)N + E);
  const size_t Np = ((size_t)N + 3) & ~(size_t)3, Ep = ((size_t)E + 3) & ~(size_t)3;
  float *dis = fws, *selfw = dis + Np, *selfnorm = selfw + Np, *norm = selfnorm + Np, *ws = norm + Ep + Np;   // (one spare Np block keeps ws 16-byte aligned)
  int rc = dge_gnn_csr_build(N, E, dst, rowptr_d, perm_d, ws_d, stream);
  if (rc) return rc;
  rc = dge_gnn_csr_build(N, E, src, rowptr_s, perm_s, ws_s, stream);
  if (rc) return rc;
  rc = dge_gcn_norm(N, E, src, dst, w, rowptr_s, perm_s, 2.0f, dis, selfw, norm, selfnorm, stream);
  if (rc) return rc;
  return dge_gcn_q_forward_dev(N, nullptr, Cin, C, x, rowptr_d, perm_d, src, norm, selfnorm, W1, b1, W2t_hi, W2t_lo, b2, head_w, head_b_dev, ws, q,
                               static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------------ GRU cell gates (GG-NN) ---------------------
// torch.nn.GRUCell after its two dense transforms (GatedGraphConv.forward -> self.rnn(m, h), Networks.py:73-86 via PyG):
//   r = sigmoid(gi_r + b_ir + gh_r + b_hr),  z = sigmoid(gi_z + b_iz + gh_z + b_hz),
//   n = tanh(gi_n + b_in + r * (gh_n + b_hn)),  h' = (1 - z) * n + z * h          (gate order r | z | n along 3C)
// gi = m W_ih^T and gh = h W_hh^T come from the tensor-core GEMM; this kernel is the HBM-bound rest: 7C floats in, C out
// per node, float4 lanes, optional fused ReLU on the output (the activation GGNN applies after the last layer).
namespace {
__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + __expf(-v)); }
__global__ void __launch_bounds__(256) k_gru_gates(int64_t n4, int C4, const float4 *__restrict__ gi, const float4 *__restrict__ gh,
                                                   const float4 *__restrict__ b_ih, const float4 *__restrict__ b_hh,
                                                   const float4 *__restrict__ h, int relu, float4 *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int64_t row = i / C4;
  const int c = (int)(i - row * C4);
  const int64_t g0 = row * 3 * C4 + c;
  auto ld = [](const float4 *p, int64_t k) { return p[k]; };
  const float4 ir = ld(gi, g0), iz = ld(gi, g0 + C4), in_ = ld(gi, g0 + 2 * C4);
  const float4 hr = ld(gh, g0), hz = ld(gh, g0 + C4), hn = ld(gh, g0 + 2 * C4);
  const float4 bir = b_ih[c], biz = b_ih[C4 + c], bin = b_ih[2 * C4 + c];
  const float4 bhr = b_hh[c], bhz = b_hh[C4 + c], bhn = b_hh[2 * C4 + c];
  const float4 hv = h[i];
  auto cell = [&](float ir_, float iz_, float in2, float hr_, float hz_, float hn_, float b0, float b1, float b2, float c0, float c1, float c2, float hp) {
    const float r = sigmoidf_((ir_ + b0) + (hr_ + c0));
    const float z = sigmoidf_((iz_ + b1) + (hz_ + c1));
    const float n = tanhf((in2 + b2) + r * (hn_ + c2));
    const float o = (1.0f - z) * n + z * hp;
    return relu ? fmaxf(o, 0.f) : o;
  };
  float4 o;
  o.x = cell(ir.x, iz.x, in_.x, hr.x, hz.x, hn.x, bir.x, biz.x, bin.x, bhr.x, bhz.x, bhn.x, hv.x);
  o.y = cell(ir.y, iz.y, in_.y, hr.y, hz.y, hn.y, bir.y, biz.y, bin.y, bhr.y, bhz.y, bhn.y, hv.y);
  o.z = cell(ir.z, iz.z, in_.z, hr.z, hz.z, hn.z, bir.z, biz.z, bin.z, bhr.z, bhz.z, bhn.z, hv.z);
  o.w = cell(ir.w, iz.w, in_.w, hr.w, hz.w, hn.w, bir.w, biz.w, bin.w, bhr.w, bhz.w, bhn.w, hv.w);
  out[i] = o;
}
}  // namespace

extern "C" int dge_gru_gates(int N, int C, const float *gi, const float *gh, const float *b_ih, const float *b_hh, const float *h, int relu,
                             float *out, void *stream) {
  if (N <= 0 || C <= 0 || (C & 3) || !gi || !gh || !b_ih || !b_hh || !h || !out) return -1;
  if (((uintptr_t)gi | (uintptr_t)gh | (uintptr_t)b_ih | (uintptr_t)b_hh | (uintptr_t)h | (uintptr_t)out) & 15) return -1;
  const int64_t n4 = (int64_t)N * (C / 4);
  k_gru_gates<<<(unsigned)((n4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n4, C / 4, reinterpret_cast<const float4 *>(gi), reinterpret_cast<const float4 *>(gh), reinterpret_cast<const float4 *>(b_ih),
      reinterpret_cast<const float4 *>(b_hh), reinterpret_cast<const float4 *>(h), relu, reinterpret_cast<float4 *>(out));
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- backward of the cell: gates recomputed from the saved forward inputs (gi, gh, biases, h), one pass, 8C floats in / 7C out per node
namespace {
__global__ void __launch_bounds__(256) k_gru_gates_bwd(int64_t n4, int C4, const float4 *__restrict__ gi, const float4 *__restrict__ gh,
                                                       const float4 *__restrict__ b_ih, const float4 *__restrict__ b_hh, const float4 *__restrict__ h,
                                                       const float4 *__restrict__ gout, float4 *__restrict__ dgi, float4 *__restrict__ dgh,
                                                       float4 *__restrict__ dh) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int64_t row = i / C4;
  const int c = (int)(i - row * C4);
  const int64_t g0 = row * 3 * C4 + c;
  const float4 ir = gi[g0], iz = gi[g0 + C4], in_ = gi[g0 + 2 * C4];
  const float4 hr = gh[g0], hz = gh[g0 + C4], hn = gh[g0 + 2 * C4];
  const float4 bir = b_ih[c], biz = b_ih[C4 + c], bin = b_ih[2 * C4 + c];
  const float4 bhr = b_hh[c], bhz = b_hh[C4 + c], bhn = b_hh[2 * C4 + c];
  const float4 hv = h[i], g = gout[i];
  float ar, az, an, anr, dhd;
  auto cell = [&](float ir_, float iz_, float in2, float hr_, float hz_, float hn_, float b0, float b1, float b2, float c0, float c1, float c2, float hp,
                  float go) {
    const float r = sigmoidf_((ir_ + b0) + (hr_ + c0));
    const float z = sigmoidf_((iz_ + b1) + (hz_ + c1));
    const float q = hn_ + c2;
    const float n = tanhf((in2 + b2) + r * q);
    an = go * (1.0f - z) * (1.0f - n * n);
    ar = an * q * r * (1.0f - r);
    az = go * (hp - n) * z * (1.0f - z);
    anr = an * r;
    dhd = go * z;
  };
  float4 o_r, o_z, o_n, o_nr, o_h;
#define DGE_GRU_LANE(L)                                                                                                  \
  cell(ir.L, iz.L, in_.L, hr.L, hz.L, hn.L, bir.L, biz.L, bin.L, bhr.L, bhz.L, bhn.L, hv.L, g.L);                          \
  o_r.L = ar; o_z.L = az; o_n.L = an; o_nr.L = anr; o_h.L = dhd;
  DGE_GRU_LANE(x) DGE_GRU_LANE(y) DGE_GRU_LANE(z) DGE_GRU_LANE(w)
#undef DGE_GRU_LANE
  dgi[g0] = o_r; dgi[g0 + C4] = o_z; dgi[g0 + 2 * C4] = o_n;
  dgh[g0] = o_r; dgh[g0 + C4] = o_z; dgh[g0 + 2 * C4] = o_nr;
  dh[i] = o_h;
}

// ---- deterministic column sums: slab s adds rows [s * rows, (s + 1) * rows) of its columns, a second launch adds the slabs in order
constexpr int CS_SLABS = 296;
__global__ void __launch_bounds__(256) k_colsum_part(int N, int C, int ldx, const float *__restrict__ X, float *__restrict__ part) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const int rows = (N + CS_SLABS - 1) / CS_SLABS, r0 = blockIdx.y * rows, r1 = min(N, r0 + rows);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 += X[(size_t)r * ldx + c]; s1 += X[(size_t)(r + 1) * ldx + c]; s2 += X[(size_t)(r + 2) * ldx + c]; s3 += X[(size_t)(r + 3) * ldx + c];
  }
  for (; r < r1; ++r) s0 += X[(size_t)r * ldx + c];
  part[(size_t)blockIdx.y * C + c] = (s0 + s1) + (s2 + s3);
}
__global__ void __launch_bounds__(256) k_colsum_final(int C, const float *__restrict__ part, float *__restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int k = 0; k < CS_SLABS; ++k) s += part[(size_t)k * C + c];
  out[c] = s;
}
}  // namespace

extern "C" int dge_gru_gates_bwd(int N, int C, const float *gi, const float *gh, const float *b_ih, const float *b_hh, const float *h, const float *gout,
                                 float *dgi, float *dgh, float *dh, void *stream) {
  if (N <= 0 || C <= 0 || (C & 3) || !gi || !gh || !b_ih || !b_hh || !h || !gout || !dgi || !dgh || !dh) return -1;
  if (((uintptr_t)gi | (uintptr_t)gh | (uintptr_t)b_ih | (uintptr_t)b_hh | (uintptr_t)h | (uintptr_t)gout | (uintptr_t)dgi | (uintptr_t)dgh | (uintptr_t)dh) & 15) return -1;
  const int64_t n4 = (int64_t)N * (C / 4);
  k_gru_gates_bwd<<<(unsigned)((n4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      n4, C / 4, reinterpret_cast<const float4 *>(gi), reinterpret_cast<const float4 *>(gh), reinterpret_cast<const float4 *>(b_ih),
      reinterpret_cast<const float4 *>(b_hh), reinterpret_cast<const float4 *>(h), reinterpret_cast<const float4 *>(gout),
      reinterpret_cast<float4 *>(dgi), reinterpret_cast<float4 *>(dgh), reinterpret_cast<float4 *>(dh));
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int64_t dge_colsum_ws_floats(int C) { return (int64_t)CS_SLABS * (C > 0 ? C : 0); }
extern "C" int dge_colsum(int N, int C, const float *X, int ldx, float *out, float *ws, void *stream) {
  if (N <= 0 || C <= 0 || !X || !out || !ws) return -1;
  if (!ldx) ldx = C;
  if (ldx < C) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_colsum_part<<<dim3((C + 255) / 256, CS_SLABS), 256, 0, st>>>(N, C, ldx, X, ws);
  k_colsum_final<<<(C + 255) / 256, 256, 0, st>>>(C, ws, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ------------------------------------------------------------------ g-U-Net: augmented adjacency ---------------
// GraphUNet.augment_adj (Networks.py:216-225 via PyG): A <- remove_self_loops((A + I)(A + I)), coalesced (sorted by row, col).
// The reference goes through torch_sparse.spspmm (cuSPARSE SpGEMM + sort + coalesce); the batched exploration graphs are
// block diagonal with blocks of at most a few hundred nodes, so one warp per row accumulates its output row DENSELY over the
// columns of its own graph in shared memory: for every k in N(i) + {i} (in edge order: deterministic sums), the lanes add
// a_ik * a_kj for j in N(k) + {k}.  Two passes (count, fill) around the existing exclusive scan; columns come out ascending.
namespace {
constexpr int AUG_MAXG = 1024;   // largest graph (nodes) the shared-memory row accumulator holds
constexpr int AUG_WARPS = 8;

template <bool FILL>
__global__ void __launch_bounds__(AUG_WARPS * 32) k_augment_adj(int N, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ perm,
                                                                const int64_t *__restrict__ dst, const float *__restrict__ w,
                                                                const int64_t *__restrict__ batch, const int64_t *__restrict__ gptr,
                                                                int32_t *__restrict__ cnt, const int32_t *__restrict__ outptr,
                                                                int64_t *__restrict__ orow, int64_t *__restrict__ ocol, float *__restrict__ oval) {
  __shared__ float acc_all[AUG_WARPS][AUG_MAXG];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * AUG_WARPS + warp;
  if (i >= N) return;
  float *acc = acc_all[warp];
  const int g = batch ? (int)batch[i] : 0;
  const int n0 = (int)gptr[g], ng = (int)gptr[g + 1] - n0;
  for (int j = lane; j < ng; j += 32) acc[j] = 0.f;
  __syncwarp();
  const int lo = rowptr[i], hi = rowptr[i + 1];
  for (int p = lo - 1; p < hi; ++p) {              // p == lo - 1: the added self loop (k = i, weight 1)
    const int k = (p < lo) ? i : (int)dst[perm[p]];
    const float aik = (p < lo) ? 1.0f : w[perm[p]];
    const int klo = rowptr[k], khi = rowptr[k + 1];
    for (int q = klo + lane; q < khi; q += 32) {   // distinct columns inside one row of a coalesced input: no lane conflicts
      const int e = perm[q];
      acc[(int)dst[e] - n0] += aik * w[e];
    }
    __syncwarp();
    if (lane == 0) acc[k - n0] += aik;             // a_kk = 1
    __syncwarp();
  }
  // compact the non-zero columns (ascending), dropping the diagonal
  int base = FILL ? outptr[i] : 0, total = 0;
  for (int j0 = 0; j0 < ng; j0 += 32) {
    const int j = j0 + lane;
    const float v = (j < ng) ? acc[j] : 0.f;
    const bool keep = j < ng && v != 0.f && (n0 + j) != i;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (FILL && keep) {
      const int o = base + total + __popc(m & ((1u << lane) - 1));
      orow[o] = i; ocol[o] = n0 + j; oval[o] = v;
    }
    total += __popc(m);
  }
  if (!FILL && lane == 0) cnt[i] = total;
}
}  // namespace

// pass 1: cnt[N] = entries of every output row; the caller scans (dge_gnn_scan) and sizes the outputs; pass 2 fills them.
extern "C" int dge_gnn_augment_adj_count(int N, const int32_t *rowptr_src, const int32_t *perm_src, const int64_t *dst, const float *w,
                                         const int64_t *batch, const int64_t *graph_ptr, int max_graph_nodes, int32_t *cnt, int32_t *outptr, void *stream) {
  if (N <= 0 || !rowptr_src || !perm_src || !graph_ptr || !cnt || !outptr) return -1;
  if (max_graph_nodes > AUG_MAXG) return -3;      // larger graphs: the caller uses its generic path
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_augment_adj<false><<<cdiv(N, AUG_WARPS), AUG_WARPS * 32, 0, st>>>(N, rowptr_src, perm_src, dst, w, batch, graph_ptr, cnt, nullptr, nullptr, nullptr, nullptr);
  k_scan<<<1, 1024, 0, st>>>(N, cnt, outptr);
  return CK();
}
extern "C" int dge_gnn_augment_adj_fill(int N, const int32_t *rowptr_src, const int32_t *perm_src, const int64_t *dst, const float *w,
                                        const int64_t *batch, const int64_t *graph_ptr, const int32_t *outptr, int64_t *out_row, int64_t *out_col,
                                        float *out_val, void *stream) {
  if (N <= 0 || !rowptr_src || !perm_src || !graph_ptr || !outptr || !out_row || !out_col || !out_val) return -1;
  k_augment_adj<true><<<cdiv(N, AUG_WARPS), AUG_WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(N, rowptr_src, perm_src, dst, w, batch, graph_ptr, nullptr,
                                                                                                      outptr, out_row, out_col, out_val);
  return CK();
}

// ------------------------------------------------------------------ g-U-Net: top-k pooling ---------------------
// TopKPooling (Networks.py:150-158 via PyG topk + filter_adj): per graph keep the k = ceil(ratio n) nodes with the highest
// score, in descending score order (ties: lower node index first, like a stable descending sort), relabel, and keep the edges
// whose two ends survive (in their original order).  One CTA per graph: bitonic sort of (score, index) in shared memory.
namespace {
constexpr int TOPK_MAXG = 1024;
__global__ void __launch_bounds__(512) k_topk_pool(const float *__restrict__ score, const int64_t *__restrict__ gptr, const int64_t *__restrict__ kptr,
                                                   int64_t *__restrict__ perm, int64_t *__restrict__ newid) {
  __shared__ float ss[TOPK_MAXG];
  __shared__ int si[TOPK_MAXG];
  const int g = blockIdx.x, tid = threadIdx.x;
  const int n0 = (int)gptr[g], n = (int)gptr[g + 1] - n0;
  const int k0 = (int)kptr[g], k = (int)kptr[g + 1] - k0;
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = tid; i < P; i += blockDim.x) { ss[i] = (i < n) ? score[n0 + i] : -INFINITY; si[i] = (i < n) ? i : 0x7fffffff; }
  __syncthreads();
  // "a before b": higher score first, then lower index (NaN scores do not occur: tanh of finite values)
  auto before = [](float sa, int ia, float sb, int ib) { return sa > sb || (sa == sb && ia < ib); };
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool up = ((lo & size) == 0);                  // ascending position = earlier in the final order
        const float sa = ss[lo], sb = ss[hi];
        const int ia = si[lo], ib = si[hi];
        const bool swap = up ? before(sb, ib, sa, ia) : before(sa, ia, sb, ib);
        if (swap) { ss[lo] = sb; ss[hi] = sa; si[lo] = ib; si[hi] = ia; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n; i += blockDim.x) newid[n0 + si[i]] = (i < k) ? (int64_t)(k0 + i) : (int64_t)-1;
  for (int i = tid; i < k; i += blockDim.x) perm[k0 + i] = n0 + si[i];
}
__global__ void k_edge_keep(int E, const int64_t *__restrict__ src, const int64_t *__restrict__ dst, const int64_t *__restrict__ newid, int32_t *flag) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) flag[e] = (newid[src[e]] >= 0 && newid[dst[e]] >= 0) ? 1 : 0;
}
__global__ void k_edge_compact(int E, const int64_t *__restrict__ src, const int64_t *__restrict__ dst, const float *__restrict__ w,
                               const int64_t *__restrict__ newid, const int32_t *__restrict__ flag, const int32_t *__restrict__ pos,
                               int64_t *__restrict__ osrc, int64_t *__restrict__ odst, float *__restrict__ ow) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E && flag[e]) { const int o = pos[e]; osrc[o] = newid[src[e]]; odst[o] = newid[dst[e]]; ow[o] = w[e]; }
}
}  // namespace

extern "C" int dge_gnn_topk_pool(int G, int max_graph_nodes, const float *score, const int64_t *graph_ptr, const int64_t *k_ptr, int64_t *perm,
                                 int64_t *newid, void *stream) {
  if (G <= 0 || !score || !graph_ptr || !k_ptr || !perm || !newid) return -1;
  if (max_graph_nodes > TOPK_MAXG) return -3;
  k_topk_pool<<<G, 512, 0, static_cast<cudaStream_t>(stream)>>>(score, graph_ptr, k_ptr, perm, newid);
  return CK();
}
// filter_adj: flag + exclusive scan (pos [E+1], pos[E] = E') in the first call, compaction in the second
extern "C" int dge_gnn_filter_adj_count(int E, const int64_t *src, const int64_t *dst, const int64_t *newid, int32_t *flag, int32_t *pos, void *stream) {
  if (E <= 0 || !src || !dst || !newid || !flag || !pos) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_edge_keep<<<cdiv(E, 256), 256, 0, st>>>(E, src, dst, newid, flag);
  k_scan<<<1, 1024, 0, st>>>(E, flag, pos);
  return CK();
}
extern "C" int dge_gnn_filter_adj_fill(int E, const int64_t *src, const int64_t *dst, const float *w, const int64_t *newid, const int32_t *flag,
                                       const int32_t *pos, int64_t *out_src, int64_t *out_dst, float *out_w, void *stream) {
  if (E <= 0 || !src || !dst || !w || !newid || !flag || !pos || !out_src || !out_dst || !out_w) return -1;
  k_edge_compact<<<cdiv(E, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(E, src, dst, w, newid, flag, pos, out_src, out_dst, out_w);
  return CK();
}

// second-layer aggregation of the training forward (csrc/dge_train.cu): d2 = dropout(relu(A t2 + b2)), q = d2 Wh + bh
int dge_agg_fwd_train_bulk(int N, int C, const float *X, const int32_t *rowptr, const int32_t *perm, const int64_t *nbr, const float *coef,
                           const float *selfcoef, const float *bias, float drop_p, uint64_t drop_seed, const float *head_w, const float *head_b_dev,
                           float *d2, float *q, cudaStream_t st) {
  if (!agg_bulk_ok(C, X, d2, bias, head_w)) return -1;
  return agg_bulk_launch(N < 148 * 4 ? N : 148 * 4, N, C, X, rowptr, perm, nbr, coef, selfcoef, bias, 1, drop_p, drop_seed, d2, head_w, 0.f, q, head_b_dev, nullptr, st);
}
