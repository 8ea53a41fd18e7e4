// Dense node-MLP GEMM of the graph Q-network on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the cuBLAS SGEMM behind PyG's GCNConv / GatedGraphConv / GRUCell transforms
// (scripts/Networks.py:22-24,76-82: X @ W with X [nodes,1000] fp32 and W [1000,1000|3000]), rows a16/a17 of
// SURVEY section 8.  The contract on Q-values is 1e-4 relative in fp32, which a single TF32 (10-bit mantissa)
// or BF16 pass does not meet over K = 1000, so the product is evaluated as a 3xTF32 split:
//     x = hi + lo,  hi = tf32(x), lo = tf32(x - hi)          (x - hi is exact in fp32)
//     A B^T ~= Ah Bh^T + Al Bh^T + Ah Bl^T                   (dropped term Al Bl^T ~ 2^-22 relative)
// with fp32 accumulation in tensor memory: error ~1e-6 relative, i.e. fp32-GEMM quality at tensor-core speed.
//
// Kernel shape (persistent: one CTA per SM walks the 128x128 output tiles; 192 threads, warp-specialised; the accumulator is
// double-buffered in tensor memory so the epilogue of a tile overlaps the MMAs of the next one):
//   warp 0   TMA producer: per 32-wide K block four 128x32 fp32 boxes (Ah, Al, Bh, Bl; 128-byte swizzle)
//            into a 3-stage shared-memory ring, completion on an mbarrier (expect_tx);
//   warp 1   allocates 256 TMEM columns, then one elected lane issues tcgen05.mma.kind::tf32 (M128 N128 K8):
//            3 products x 4 K-steps per stage, tcgen05.commit releases the stage / signals the epilogue;
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per warp and pass) -> registers -> 128-byte row
//            segments of C (row / column masked).
// Operands are K-major ([rows, K] with K contiguous): A = activations, B^T = the weight stored [out, in].
// The row count can be read from device memory (M_dev) so that a graph batch whose size is only known on the
// device needs no host synchronisation: the tile loop is bounded by the live rows, surplus CTAs exit immediately.
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dge_gnn.h"

namespace {

constexpr int BM = 128, BK = 32;                 // BK fp32 = 128 bytes = one SWIZZLE_128B row
// Output tile BM x BN_.  BN_ = 128: 64 KB per stage (Ah, Al, Bh, Bl tiles of 16 KB), 3 stages, two 128-column accumulators in tensor memory.
// BN_ = 256 (wide): the A tiles are reused over twice the columns -- 96 KB per stage for twice the MMAs, i.e. 0.75x the shared-memory fill
// traffic per flop (the 3xTF32 product is fed from L2 at ~24 MAC per byte: the fill rate, not the tensor pipe, is what bounds it) -- 2 stages,
// two 256-column accumulators = all 512 columns of tensor memory.
constexpr int A_TILE_BYTES = BM * BK * 4;        // 16 KB
template <int BN_> struct Shape {
  static constexpr int STAGES = BN_ == 256 ? 2 : 3;
  static constexpr int B_TILE_BYTES = BN_ * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN_;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 128 /* barriers */;
  // instruction descriptor (cute::UMMA::InstrDescriptor layout): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN_ >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};
constexpr int GEMM_THREADS = 192;
constexpr int UMMA_K = 8;                        // tf32: 32 bytes of K per instruction


__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must fault (launch error) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > 4000000000ll) __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major tile of 128-byte rows, SWIZZLE_128B,
// 8-row core groups 1024 bytes apart (SBO), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major operand.  For 32-bit elements the tensor core takes ONE shared-memory layout in this mode (cute::UMMA::Layout_MN_SW128_32B_Atom,
// LayoutType::SWIZZLE_128B_BASE32B = 1; canonical form ((T,8,m),(4,k)):((1,T,LBO),(8T,SBO)), T = 4 floats per 16 bytes): a 128-byte row holds
// 32 consecutive M (or N) elements of ONE k, its 32-byte chunks XOR-swizzled with (row & 3) -- the TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
// four k rows form the 512-byte atom, the next four start SBO = 512 bytes on, the next 32 M elements LBO = 4096 bytes on (each 32-column
// chunk of the tile is its own TMA box of BK = 32 rows).
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(4096u >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t IDESC) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {   // arrives on `bar` when every MMA issued so far has retired
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent kernel: gridDim.x CTAs (<= one per SM) walk the live 128x128 tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...
// (column tile fastest, so neighbouring CTAs share an A row block through L2).  The accumulator is double-buffered in tensor
// memory (2 x 128 columns): the epilogue of tile i drains buffer i & 1 while the MMAs of tile i + 1 fill the other one, and the
// TMA ring runs ahead across tile boundaries.  The live row count is read from device memory when M_dev is given.
// TN = true: C = A^T B for A [K, M] and B [K, N] as stored (row-major, the contraction runs over the ROWS: the weight gradient x^T dy of a
// dense layer, K = nodes) -- both operands are read MN-major: every 32-column chunk of a tile is one TMA box of BK rows, the instruction
// descriptor's major bits are set, and an MMA step advances by eight rows (one swizzle atom) instead of 32 bytes.  No transposed copy of
// x or dy is ever made.
template <int BN, bool TN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tf32x3(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
              const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl,
              float *__restrict__ C, int M, const int32_t *__restrict__ M_dev, int N, int K, int ldc, int splits) {
  // splits > 1 (weight-gradient shape: few output tiles, a long K): work item = (tile, K slice); the slices of a tile add their
  // partial products into C with vector atomics (C zeroed by the caller; with two slices the sum is order-independent)
  constexpr int STAGES = Shape<BN>::STAGES, STAGE_BYTES = Shape<BN>::STAGE_BYTES, TMEM_COLS = Shape<BN>::TMEM_COLS;
  constexpr int TILE_BYTES = A_TILE_BYTES, B_TILE = Shape<BN>::B_TILE_BYTES;
  constexpr uint32_t IDESC = Shape<BN>::IDESC | (TN ? ((1u << 15) | (1u << 16)) : 0u);   // bits 15 / 16: A / B are MN-major
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = M_dev ? min(M, *M_dev) : M;
  const int n_tiles = (N + BN - 1) / BN;
  const int total = ((rows + BM - 1) / BM) * n_tiles * splits;
  if ((int)blockIdx.x >= total) return;                    // uniform per CTA, before any barrier / allocation
  extern __shared__ uint8_t smem_raw[];
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B tiles want 1024-byte alignment
  const uint32_t bars = tiles + STAGES * STAGE_BYTES;
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16, slot = tempty0 + 16;
  uint32_t *slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (slot - smem_u32(smem_raw)));
  const int num_kb = (K + BK - 1) / BK;
  auto kb_lo = [&](int w) { return (int)((long long)(w % splits) * num_kb / splits); };
  auto kb_hi = [&](int w) { return (int)((long long)(w % splits + 1) * num_kb / splits); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                                         // one warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;

  if (warp == 0) {
    if (lane == 0) {                                       // ---- TMA producer
      int it = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x) {
        const int t = w / splits;
        const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
        for (int kb = kb_lo(w), ke = kb_hi(w); kb < ke; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty0 + 8 * s, ((it / STAGES) & 1) ^ 1);
          const uint32_t st = tiles + s * STAGE_BYTES, fb = full0 + 8 * s;
          mbar_expect_tx(fb, STAGE_BYTES);
          if constexpr (TN) {   // boxes of 32 columns (inner coordinate) x BK rows: 4 KB each, chunk c of a tile at c * 4096
#pragma unroll
            for (int c = 0; c < BM / 32; ++c) {
              tma_load_2d(st + c * 4096, &mAh, fb, m0 + 32 * c, kb * BK);
              tma_load_2d(st + TILE_BYTES + c * 4096, &mAl, fb, m0 + 32 * c, kb * BK);
            }
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
              tma_load_2d(st + 2 * TILE_BYTES + c * 4096, &mBh, fb, n0 + 32 * c, kb * BK);
              tma_load_2d(st + 2 * TILE_BYTES + B_TILE + c * 4096, &mBl, fb, n0 + 32 * c, kb * BK);
            }
          } else {
            tma_load_2d(st, &mAh, fb, kb * BK, m0);
            tma_load_2d(st + TILE_BYTES, &mAl, fb, kb * BK, m0);
            tma_load_2d(st + 2 * TILE_BYTES, &mBh, fb, kb * BK, n0);
            tma_load_2d(st + 2 * TILE_BYTES + B_TILE, &mBl, fb, kb * BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                       // ---- MMA issuer (one thread drives the tensor core)
      int it = 0, lt = 0;
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++lt) {
        const int acc = lt & 1;
        mbar_wait(tempty0 + 8 * acc, ((lt >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator (free at first use)
        tc_fence_after();
        const uint32_t td = tmem + (uint32_t)(acc * BN);
        const int kb0 = kb_lo(w);
        for (int kb = kb0, ke = kb_hi(w); kb < ke; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t st = tiles + s * STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 32 bytes along K inside the swizzled 128-byte row; MN-major: eight k rows = two 512-byte atoms = 1024 bytes
            const uint32_t off = TN ? k * 1024 : k * UMMA_K * 4;
            const uint64_t ah = TN ? smem_desc_mn(st + off) : smem_desc(st + off), al = TN ? smem_desc_mn(st + TILE_BYTES + off) : smem_desc(st + TILE_BYTES + off);
            const uint64_t bh = TN ? smem_desc_mn(st + 2 * TILE_BYTES + off) : smem_desc(st + 2 * TILE_BYTES + off);
            const uint64_t bl = TN ? smem_desc_mn(st + 2 * TILE_BYTES + B_TILE + off) : smem_desc(st + 2 * TILE_BYTES + B_TILE + off);
            umma_tf32(td, al, bh, (kb != kb0 || k) ? 1u : 0u, IDESC);     // small terms first
            umma_tf32(td, ah, bl, 1u, IDESC);
            umma_tf32(td, ah, bh, 1u, IDESC);
          }
          umma_commit(empty0 + 8 * s);                     // stage free once these MMAs have read it
        }
        umma_commit(tfull0 + 8 * acc);                     // accumulator complete
      }
    }
  } else {                                                 // ---- epilogue: TMEM -> registers -> C
    const int q = warp & 3;                                // a warp reaches TMEM lanes [32 (warp % 4), +32)
    int lt = 0;
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++lt) {
      const int t = w / splits;
      const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
      const int acc = lt & 1;
      mbar_wait(tfull0 + 8 * acc, (lt >> 1) & 1);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      float *crow = C + (size_t)row * ldc + n0;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < rows) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int col = n0 + c * 32 + 4 * i;
            const float4 v = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
            if (col + 3 < N) {
              if (splits > 1) atomicAdd(reinterpret_cast<float4 *>(crow + c * 32 + 4 * i), v);
              else *reinterpret_cast<float4 *>(crow + c * 32 + 4 * i) = v;
            } else {
              for (int j = 0; j < 4; ++j)
                if (col + j < N) { if (splits > 1) atomicAdd(crow + c * 32 + 4 * i + j, __uint_as_float(r[4 * i + j])); else crow[c * 32 + 4 * i + j] = __uint_as_float(r[4 * i + j]); }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);       // 4 epilogue warps -> the accumulator may be overwritten
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

// x -> (tf32(x), tf32(x - tf32(x))), round-to-nearest; 4 elements per thread
__global__ void k_split_tf32(int64_t n4, const float4 *__restrict__ x, float4 *__restrict__ hi, float4 *__restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = x[i];
  float4 h, l;
  auto split = [](float a, float &ho, float &lw) {
    uint32_t hb, lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(a));
    ho = __uint_as_float(hb);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(a - ho));
    lw = __uint_as_float(lb);
  };
  split(v.x, h.x, l.x); split(v.y, h.y, l.y); split(v.z, h.z, l.z); split(v.w, h.w, l.w);
  hi[i] = h; lo[i] = l;
}

// W [K, N] row-major -> (Wt_hi, Wt_lo) [N, K]: the K-major B operand of X @ W
__global__ void __launch_bounds__(256) k_prep_weight_t(int K, int N, const float *__restrict__ W, float *__restrict__ thi, float *__restrict__ tlo) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + tx;
    tile[r][tx] = (k < K && n < N) ? W[(size_t)k * N + n] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    if (n < N && k < K) {
      const float a = tile[tx][r];
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(a));
      const float h = __uint_as_float(hb);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(a - h));
      thi[(size_t)n * K + k] = h; tlo[(size_t)n * K + k] = __uint_as_float(lb);
    }
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// [rows, K] fp32 row-major, box = 128 rows x 32 columns (128 bytes), 128-byte swizzle, out-of-bounds reads give 0
bool make_map(CUtensorMap *m, const float *base, uint64_t rows, uint64_t K, uint64_t pitch = 0, uint32_t box_rows = BM,
              CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[2] = {K, rows};
  const cuuint64_t strides[1] = {(pitch ? pitch : K) * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// K slices per output tile when the caller leaves the choice to the library (splits = 0; the caller zeroes C, the slices add into it):
// (a) as many as fill the SMs when the output has fewer tiles than the machine has SMs and K is long (the weight-gradient shape);
// (b) at least one per 64 K blocks (2048 elements of K): the tensor core's fp32 accumulator truncates, so the error of a product grows
//     with the number of MMAs accumulated in tensor memory (measured: 6e-5 of the largest entry at K = 16896 in one or two slices); partial
//     sums of <= 2048 go through the fp32 atomics' round-to-nearest adds instead.  Two slices add in either order to the same bits; more
//     than two make the last bits depend on the order of arrival.
int auto_splits(long long tiles_cap, int num_kb, int n_sm) {
  int splits = 1;
  if (tiles_cap < n_sm && num_kb >= 64) { splits = (int)(n_sm / tiles_cap); if (splits > num_kb / 16) splits = num_kb / 16; if (splits < 1) splits = 1; }
  const int by_len = (num_kb + 63) / 64;
  return splits > by_len ? splits : by_len;
}

}  // namespace

extern "C" int dge_gemm_split_tf32(int64_t n, const float *x, float *hi, float *lo, void *stream) {
  if (n < 0 || (n & 3) || !x || !hi || !lo) return -1;
  if (n == 0) return 0;
  const int64_t n4 = n / 4;
  k_split_tf32<<<(unsigned)((n4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(n4, reinterpret_cast<const float4 *>(x),
                                                                                              reinterpret_cast<float4 *>(hi), reinterpret_cast<float4 *>(lo));
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int dge_gemm_prep_weight(int K, int N, const float *W, float *Wt_hi, float *Wt_lo, void *stream) {
  if (K <= 0 || N <= 0 || !W || !Wt_hi || !Wt_lo) return -1;
  k_prep_weight_t<<<dim3((N + 31) / 32, (K + 31) / 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(K, N, W, Wt_hi, Wt_lo);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// general form: operand row pitches lda / ldb (floats, multiples of 4, >= K; 0 = K) and `splits` K slices per output tile
// (splits > 1: C must be zeroed by the caller, partial products are added with atomics).  splits = 0: chosen here -- as many
// slices as fill the SMs when the output has fewer tiles than the machine has SMs and K is long (the weight-gradient shape).
extern "C" int dge_gemm_tf32x3_ex(int M, const int32_t *M_dev, int N, int K, const float *A_hi, const float *A_lo, int lda, const float *Bt_hi,
                                  const float *Bt_lo, int ldb, float *C, int ldc, int splits, void *stream) {
  if (M < 0 || N <= 0 || K <= 0 || (ldc & 3) || ldc < N || !A_hi || !A_lo || !Bt_hi || !Bt_lo || !C) return -1;
  if (!lda) lda = K;
  if (!ldb) ldb = K;
  if ((lda & 3) || (ldb & 3) || lda < K || ldb < K) return -1;
  if (((uintptr_t)A_hi | (uintptr_t)A_lo | (uintptr_t)Bt_hi | (uintptr_t)Bt_lo | (uintptr_t)C) & 15) return -1;
  if (M == 0) return 0;
  // tile width: 128 columns; the 256-column shape (DGE_GEMM_BN=256, A/B) measured the same (158 us vs 160 us at M = 16384, N = K = 1000):
  // the kernel already runs each of its three products at the rate of the library's single-pass TF32 GEMM (55 us for one product)
  static const int forced_bn = [] { const char *v = getenv("DGE_GEMM_BN"); return v ? atoi(v) : 0; }();
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  const long long m_tiles = (M + BM - 1) / BM;
  const int bn = (forced_bn == 256 && !M_dev) ? 256 : 128;
  CUtensorMap mAh, mAl, mBh, mBl;
  if (!make_map(&mAh, A_hi, M, K, lda) || !make_map(&mAl, A_lo, M, K, lda) || !make_map(&mBh, Bt_hi, N, K, ldb, bn) || !make_map(&mBl, Bt_lo, N, K, ldb, bn)) return -2;
  static bool configured[2] = {false, false};
  if (!configured[bn == 256]) {
    const cudaError_t ce = bn == 256 ? cudaFuncSetAttribute(k_gemm_tf32x3<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape<256>::SMEM)
                                     : cudaFuncSetAttribute(k_gemm_tf32x3<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape<128>::SMEM);
    if (ce != cudaSuccess) return -2;
    configured[bn == 256] = true;
  }
  const long long tiles_cap = (long long)((N + bn - 1) / bn) * m_tiles;
  const int num_kb = (K + BK - 1) / BK;
  if (splits <= 0) splits = M_dev ? 1 : auto_splits(tiles_cap, num_kb, n_sm);
  if (splits > num_kb) splits = num_kb;
  const long long work = tiles_cap * splits;
  const unsigned grid = (unsigned)(work < n_sm ? work : n_sm);
  if (bn == 256)
    k_gemm_tf32x3<256, false><<<grid, GEMM_THREADS, Shape<256>::SMEM, static_cast<cudaStream_t>(stream)>>>(mAh, mAl, mBh, mBl, C, M, M_dev, N, K, ldc, splits);
  else
    k_gemm_tf32x3<128, false><<<grid, GEMM_THREADS, Shape<128>::SMEM, static_cast<cudaStream_t>(stream)>>>(mAh, mAl, mBh, mBl, C, M, M_dev, N, K, ldc, splits);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// C [M,N] = A^T B with A [K,M] (row pitch lda >= M) and B [K,N] (row pitch ldb >= N) as stored, contraction over the rows: the weight
// gradient of a dense layer (A = x, B = dy, K = nodes) from the SAME (hi, lo) splits the forward / grad-input products use.  splits as in
// dge_gemm_tf32x3_ex (0: chosen here; > 1: C zeroed by the caller).
extern "C" int dge_gemm_tf32x3_tn(int M, int N, int K, const float *A_hi, const float *A_lo, int lda, const float *B_hi, const float *B_lo, int ldb,
                                  float *C, int ldc, int splits, void *stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (ldc & 3) || ldc < N || !A_hi || !A_lo || !B_hi || !B_lo || !C) return -1;
  if (!lda) lda = M;
  if (!ldb) ldb = N;
  if ((lda & 3) || (ldb & 3) || lda < M || ldb < N) return -1;
  if (((uintptr_t)A_hi | (uintptr_t)A_lo | (uintptr_t)B_hi | (uintptr_t)B_lo | (uintptr_t)C) & 15) return -1;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  constexpr int bn = 128;
  // the operands as [K rows, M (or N) columns]: inner dimension = the tile's M / N extent, boxes of 32 columns x BK rows
  CUtensorMap mAh, mAl, mBh, mBl;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;     // (the one layout the tensor core reads MN-major 32-bit operands in)
  if (!make_map(&mAh, A_hi, K, M, lda, BK, sw) || !make_map(&mAl, A_lo, K, M, lda, BK, sw) || !make_map(&mBh, B_hi, K, N, ldb, BK, sw) ||
      !make_map(&mBl, B_lo, K, N, ldb, BK, sw))
    return -2;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(k_gemm_tf32x3<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Shape<128>::SMEM) != cudaSuccess) return -2;
    configured = true;
  }
  const long long tiles_cap = (long long)((N + bn - 1) / bn) * ((M + BM - 1) / BM);
  const int num_kb = (K + BK - 1) / BK;
  if (splits <= 0) splits = auto_splits(tiles_cap, num_kb, n_sm);
  if (splits > num_kb) splits = num_kb;
  const long long work = tiles_cap * splits;
  const unsigned grid = (unsigned)(work < n_sm ? work : n_sm);
  k_gemm_tf32x3<128, true><<<grid, GEMM_THREADS, Shape<128>::SMEM, static_cast<cudaStream_t>(stream)>>>(mAh, mAl, mBh, mBl, C, M, nullptr, N, K, ldc, splits);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int dge_gemm_tf32x3(int M, const int32_t *M_dev, int N, int K, const float *A_hi, const float *A_lo, const float *Bt_hi,
                               const float *Bt_lo, float *C, int ldc, void *stream) {
  if (K & 3) return -1;
  return dge_gemm_tf32x3_ex(M, M_dev, N, K, A_hi, A_lo, 0, Bt_hi, Bt_lo, 0, C, ldc, 1, stream);
}
