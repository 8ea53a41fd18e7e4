// Virtual-map rebuild ("covariance-propagation kernel"): occupancy probability from integer
// visibility counts + per-cell 2x2 information by an ordered covariance-intersection fold
// over the trajectory.  Rows a6+a7(+a8) of SURVEY section 8.
//
// Replaces VirtualMap::updateProbability (VirtualMap.cpp:61-84 -> OccupancyMap::update,
// OccupancyMap.cpp:55-138) and VirtualMap::updateInformation (VirtualMap.cpp:256-316 ->
// predictVirtualLandmark :213-229, covarianceIntersection2D :364-377), which the reference
// runs as an O(T*V) scan with Eigen LLT solves per (pose, cell) pair.
//
// B200 mapping: the fold is sequential per CELL but independent across cells, so one thread
// owns one cell (state in registers: 3 information entries, visibility count, flags) and walks
// the trajectory in order.  A CTA owns a 16x16 tile of cells; warps are 8x4 patches so that the
// 7x7-cell sensor footprint of a pose keeps most lanes of a warp busy.  Poses are pre-digested
// once per env (k_vmap_prep: cos/sin, covariance, validity, per-32-pose bounding boxes) and
// staged through shared memory only for the chunks whose bounding box touches the tile.
// The arithmetic is the closed form of the reference's expression
//   Hl^-1 (R + Hx Sigma Hx^T) Hl^-T = Rot (M^-1 R M^-T + A Sigma A^T) Rot^T
// (Hl = M Rot^T, Hx = M A), evaluated in fp64.
#include "dge_internal.cuh"

namespace {

constexpr int TILE = 16;           // cells per tile side
constexpr int VCH = 32;            // poses per chunk
constexpr int VU = 4;              // poses whose state-independent part is evaluated together (instruction-level parallelism)
constexpr int PREP_W = 12;         // doubles per digested pose: x y c s Sxx Sxy Sxt Syy Syt Stt valid pad

struct VmapCfg {
  double map_min_x, map_min_y, res;
  double max_range, min_range, max_bearing, min_bearing;
  double rb, rr;                   // sigma_b^2, sigma_r^2
  double max_r2_lo, max_r2_hi, min_r2_lo, min_r2_hi;   // guard bands around max_range^2 / min_range^2 (see k_vmap_cells)
  double i0;                       // 1/sigma0^2
  double ptab[6];                  // probability for n_seen = 0..4(+) and for a landmark cell (host-evaluated, q8)
  double wedge_tan;                // tan of the half-width of the rear blind wedge (with guard)
  int fov_wide;
  int rows, cols;
};

// ------------------------------------------------------------------ prep ---
__global__ void __launch_bounds__(128) k_vmap_prep(VmapCfg c, int Tstride, const int32_t *n_poses, int Tfixed,
                                                   const double *pose /*[n,Tstride,3]*/, const double *cov /*[n,Tstride,6]*/,
                                                   const double *info /*nullable [n,Tstride,6]*/, double *prep /*[n,Tstride,PREP_W]*/,
                                                   double *cbox /*[n,nchunk_max,4]*/, int nchunk_max, const uint8_t *mask) {
  const int b = blockIdx.x;
  if (mask && !mask[b]) return;
  const int T = n_poses ? n_poses[b] : Tfixed;
  const double *ps = pose + (size_t)b * Tstride * 3, *cv = cov + (size_t)b * Tstride * 6;
  double *pr = prep + (size_t)b * Tstride * PREP_W;
  for (int k = threadIdx.x; k < T; k += blockDim.x) {
    double s, co;
    sincos(ps[3 * k + 2], &s, &co);
    double *o = pr + (size_t)k * PREP_W;
    o[0] = ps[3 * k]; o[1] = ps[3 * k + 1]; o[2] = co; o[3] = s;
    double S[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { S[i] = cv[6 * k + i]; o[4 + i] = S[i]; }
    double det_info;
    if (info) {
      const double *a = info + ((size_t)b * Tstride + k) * 6;
      det_info = a[0] * (a[3] * a[5] - a[4] * a[4]) - a[1] * (a[1] * a[5] - a[4] * a[2]) + a[2] * (a[1] * a[4] - a[3] * a[2]);
    } else {
      const double dc = S[0] * (S[3] * S[5] - S[4] * S[4]) - S[1] * (S[1] * S[5] - S[4] * S[2]) + S[2] * (S[1] * S[4] - S[3] * S[2]);
      det_info = 1.0 / dc;
    }
    o[10] = (det_info < 1e-10) ? 0.0 : 1.0;   // VirtualMap.cpp:293
    o[11] = 0.0;
  }
  __syncthreads();
  // per-chunk bounding boxes (pose positions inflated by the sensor range)
  const int nch = (T + VCH - 1) / VCH;
  for (int ch = threadIdx.x; ch < nch; ch += blockDim.x) {
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int k = ch * VCH; k < min(T, (ch + 1) * VCH); ++k) {
      const double x = pr[(size_t)k * PREP_W], y = pr[(size_t)k * PREP_W + 1];
      x0 = fmin(x0, x); x1 = fmax(x1, x); y0 = fmin(y0, y); y1 = fmax(y1, y);
    }
    double *o = cbox + ((size_t)b * nchunk_max + ch) * 4;
    o[0] = x0 - c.max_range - 0.01; o[1] = x1 + c.max_range + 0.01; o[2] = y0 - c.max_range - 0.01; o[3] = y1 + c.max_range + 0.01;
  }
}

// ----------------------------------------------------------------- cells ---
__global__ void __launch_bounds__(TILE * TILE, 3) k_vmap_cells(VmapCfg c, int Tstride, const int32_t *n_poses, int Tfixed,
                                                            const double *prep, const double *cbox, int nchunk_max,
                                                            const double *lm /*[n,Lstride,2]*/, const uint8_t *lm_obs /*nullable*/,
                                                            int Lstride, int Lfixed, double *prob /*[n,V]*/, double *vinfo /*[n,V,3]*/,
                                                            int32_t *seen_out /*nullable [n,V]*/, const uint8_t *mask) {
  const int b = blockIdx.y;
  if (mask && !mask[b]) return;
  const int T = n_poses ? n_poses[b] : Tfixed;
  const int tiles_x = (c.cols + TILE - 1) / TILE;
  const int tile_r = (blockIdx.x / tiles_x) * TILE, tile_c = (blockIdx.x % tiles_x) * TILE;
  // warp = 8x4 patch; 8 warps tile the 16x16 block as 2 (x) by 4 (y)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = tile_c + (warp & 1) * 8 + (lane & 7);
  const int row = tile_r + (warp >> 1) * 4 + (lane >> 3);
  const bool valid_cell = row < c.rows && col < c.cols;
  const double cx = c.map_min_x + c.res * (col + 0.5), cy = c.map_min_y + c.res * (row + 0.5);
  // tile bounds (cell centres)
  const double tx0 = c.map_min_x + c.res * (tile_c + 0.5), tx1 = c.map_min_x + c.res * (min(tile_c + TILE, c.cols) - 0.5);
  const double ty0 = c.map_min_y + c.res * (tile_r + 0.5), ty1 = c.map_min_y + c.res * (min(tile_r + TILE, c.rows) - 0.5);

  __shared__ double sp[VCH * PREP_W];
  __shared__ unsigned char s_lmflag[TILE * TILE];

  // landmark cells (OccupancyMap.cpp:126-131): a cell holding a landmark estimate saturates "occupied".
  // One thread per landmark marks the tile-local flag; every cell thread then reads its own flag.
  s_lmflag[threadIdx.x] = 0;
  __syncthreads();
  {
    const double *l = lm + (size_t)b * Lstride * 2;
    for (int j = threadIdx.x; j < Lfixed; j += TILE * TILE) {
      if (lm_obs && !lm_obs[(size_t)b * Lstride + j]) continue;
      const int lr = (int)floor((l[2 * j + 1] - c.map_min_y) / c.res) - tile_r, lc = (int)floor((l[2 * j] - c.map_min_x) / c.res) - tile_c;
      if (lr >= 0 && lr < TILE && lc >= 0 && lc < TILE) s_lmflag[lr * TILE + lc] = 1;
    }
  }
  __syncthreads();
  const bool is_lm = s_lmflag[(row - tile_r) * TILE + (col - tile_c)] != 0;

  double ixx = c.i0, ixy = 0.0, iyy = c.i0;
  bool updated = false;
  int cnt = 0;
  const int nch = (T + VCH - 1) / VCH;
  const double rmax_c2 = (c.max_range + 0.01) * (c.max_range + 0.01);
  for (int ch = 0; ch < nch; ++ch) {
    const double *bx = cbox + ((size_t)b * nchunk_max + ch) * 4;
    if (bx[0] > tx1 || bx[1] < tx0 || bx[2] > ty1 || bx[3] < ty0) continue;   // uniform across the CTA
    const int k0 = ch * VCH, kc = min(VCH, T - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kc * PREP_W; i += TILE * TILE) sp[i] = prep[((size_t)b * Tstride + k0) * PREP_W + i];
    __syncthreads();
    // (cells outside the map stay in the loop -- the warp votes below need every lane -- but are never `near`)
    // The fold over poses is sequential per cell (covariance intersection is order dependent), and a visit is a
    // ~40-deep fp64 dependency chain.  Only the last ~10 operations (the intersection itself) depend on the running
    // state, so the state-independent part (geometry, predicted covariance, its inverse, rotation) is evaluated for
    // VU consecutive poses at once -- straight-line code, VU independent chains in flight -- and then folded in order.
    for (int kb = 0; kb < kc; kb += VU) {
      double dxs[VU], dys[VU], d2s[VU];
      bool near_any = false;
#pragma unroll
      for (int u = 0; u < VU; ++u) {
        const double *p = sp + min(kb + u, kc - 1) * PREP_W;
        dxs[u] = cx - p[0]; dys[u] = cy - p[1];
        d2s[u] = __dadd_rn(__dmul_rn(dxs[u], dxs[u]), __dmul_rn(dys[u], dys[u]));
        near_any |= valid_cell && (kb + u < kc) && d2s[u] < rmax_c2;
      }
      if (!__any_sync(0xffffffffu, near_any)) continue;
      double nxx[VU], nxy[VU], nyy[VU], ndet[VU];
      bool vis[VU], upd[VU];
#pragma unroll
      for (int u = 0; u < VU; ++u) {
        const double *p = sp + min(kb + u, kc - 1) * PREP_W;
        const double dx = dxs[u], dy = dys[u], d2 = d2s[u];
        // range gates  r < max_range (Distance.cpp:86 / checkWithoutMinRange)  and  r > min_range (full check, q10) with
        // r = sqrt(d2): decided on d2 alone outside a relative 1e-9 band around the squared limits, by the reference's
        // own sqrt comparison inside it -- same integer visibility counts, no sqrt on the common path
        bool in_max, out_min;
        if (d2 < c.max_r2_lo) in_max = true; else if (d2 > c.max_r2_hi) in_max = false; else in_max = __dsqrt_rn(d2) < c.max_range;
        if (d2 > c.min_r2_hi) out_min = true; else if (d2 < c.min_r2_lo) out_min = false; else out_min = __dsqrt_rn(d2) > c.min_range;
        const bool inr = valid_cell && (kb + u < kc) && in_max;
        const double co = p[2], si = p[3];
        const double qx = co * dx + si * dy, qy = -si * dx + co * dy;
        // field-of-view gate min_b < atan2(qy,qx) < max_b  (Simulator2D.cpp:100-111)
        bool in_fov;
        if (c.fov_wide && (qx > 0.0 || fabs(qy) > -qx * c.wedge_tan)) in_fov = true;
        else if (inr) { const double bb = atan2(qy, qx); in_fov = bb < c.max_bearing && bb > c.min_bearing; }
        else in_fov = false;
        vis[u] = inr && in_fov;                                                     // occupancy visibility count (q8)
        upd[u] = vis[u] && out_min && p[10] != 0.0;                                 // full check (q10) ; det(info) gate
        // body-frame covariance of the predicted virtual landmark: cb = P + (rr / r^2) [qx qx, qx qy; qx qy, qy qy], with
        // P the bearing-noise + pose-covariance part.  Its inverse (the body-frame information) needs ONE division:
        //   det(cb) r^2 = det(P) r^2 + rr Q,  Q = P00 qy^2 - 2 P01 qx qy + P11 qx^2,   adj(cb) r^2 = adj(P) r^2 + rr [qy qy, -qx qy; ., qx qx]
        const double Sxx = p[4], Sxy = p[5], Sxt = p[6], Syy = p[7], Syt = p[8], Stt = p[9];
        const double qxx = qx * qx, qyy = qy * qy, qxy = qx * qy;
        const double P00 = qyy * (c.rb + Stt) + Sxx - 2.0 * qy * Sxt;
        const double P01 = Sxy - qy * Syt + qx * Sxt - qxy * (c.rb + Stt);
        const double P11 = qxx * (c.rb + Stt) + Syy + 2.0 * qx * Syt;
        const double Q = P00 * qyy - 2.0 * P01 * qxy + P11 * qxx;
        const double inv = 1.0 / ((P00 * P11 - P01 * P01) * d2 + c.rr * Q);
        const double lb00 = (P11 * d2 + c.rr * qyy) * inv, lb01 = -(P01 * d2 + c.rr * qxy) * inv, lb11 = (P00 * d2 + c.rr * qxx) * inv;
        // rotate to the map frame
        const double cc = co * co, ss = si * si, cs = co * si;
        nxx[u] = cc * lb00 - 2.0 * cs * lb01 + ss * lb11;
        nxy[u] = cs * (lb00 - lb11) + (cc - ss) * lb01;
        nyy[u] = ss * lb00 + 2.0 * cs * lb01 + cc * lb11;
        ndet[u] = d2 * inv;                                                          // det of the new information = 1 / det(cb)
      }
#pragma unroll
      for (int u = 0; u < VU; ++u) {
        if (vis[u]) ++cnt;
        if (!upd[u]) continue;
        if (!updated) { ixx = nxx[u]; ixy = nxy[u]; iyy = nyy[u]; updated = true; }
        else {  // covariance intersection on information matrices (VirtualMap.cpp:364-377, q11)
          const double a = ixx * iyy - ixy * ixy, bdet = ndet[u];
          const double cm = iyy * nxx[u] - 2.0 * ixy * nxy[u] + ixx * nyy[u];   // det(m1) * tr(m1^-1 m2)
          const double d = a + bdet - cm;
          double w = 0.5 * (2.0 * bdet - cm) / d;
          if ((w < 0 && d < 0) || (w > 1 && d > 0)) w = 0.0;
          else if ((w < 0 && d > 0) || (w > 1 && d < 0)) w = 1.0;
          ixx = w * ixx + (1.0 - w) * nxx[u]; ixy = w * ixy + (1.0 - w) * nxy[u]; iyy = w * iyy + (1.0 - w) * nyy[u];
        }
      }
    }
  }
  if (!valid_cell) return;
  const size_t cell = (size_t)b * c.rows * c.cols + (size_t)row * c.cols + col;
  prob[cell] = is_lm ? c.ptab[5] : c.ptab[min(cnt, 4)];
  vinfo[cell * 3] = ixx; vinfo[cell * 3 + 1] = ixy; vinfo[cell * 3 + 2] = iyy;
  if (seen_out) seen_out[cell] = is_lm ? -1 : cnt;
}

// --------------------------------------------------------------- metrics ---
// explored fraction (VirtualMap.cpp:47-59), utility(0) = sum trace(cov) (Planner2D.cpp:343-366),
// known-cell count, done flag (exploration_env.py:167-168).
__global__ void __launch_bounds__(256) k_vmap_metrics(dge_config cfg, DgeDims d, const double *prob, const double *vinfo,
                                                      const int32_t *sim_step, const int32_t *status, const double *dist,
                                                      double *metrics, uint8_t *done, const uint8_t *mask,
                                                      const int32_t *n_poses, const int32_t *meas_ptr, unsigned long long *counters, const uint8_t *step_kind) {
  const int b = blockIdx.x;
  if (mask && !mask[b]) return;
  if (threadIdx.x == 0 && counters && step_kind[b]) {   // integer work counters (order-independent): env-steps, sum T, sum M
    const int T = n_poses[b];
    atomicAdd(&counters[0], 1ull);
    atomicAdd(&counters[1], (unsigned long long)T);
    atomicAdd(&counters[2], (unsigned long long)meas_ptr[(size_t)b * (d.Tmax + 1) + T]);
  }
  const double *p = prob + (size_t)b * d.V, *vi = vinfo + (size_t)b * d.V * 3;
  const int extg = 20;
  int n_exp = 0, n_known = 0;
  double tr = 0.0;
  for (int i = threadIdx.x; i < d.V; i += 256) {
    const double x = (i % d.cols + 0.5) * cfg.resolution + cfg.map_min_x, y = (i / d.cols + 0.5) * cfg.resolution + cfg.map_min_y;
    const double pv = p[i];
    if ((pv < 0.49 || pv > 0.6) && cfg.map_min_x + extg <= x && x <= cfg.map_max_x - extg && cfg.map_min_y + extg <= y && y <= cfg.map_max_y - extg) ++n_exp;
    if (pv < cfg.occupancy_threshold) ++n_known;
    const double a = vi[3 * i], bb = vi[3 * i + 1], c = vi[3 * i + 2];
    tr += (a + c) / (a * c - bb * bb);
  }
  __shared__ double s_tr[256];
  __shared__ int s_e[256], s_k[256];
  s_tr[threadIdx.x] = tr; s_e[threadIdx.x] = n_exp; s_k[threadIdx.x] = n_known;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {   // fixed-order tree => run-to-run deterministic
    if (threadIdx.x < o) { s_tr[threadIdx.x] += s_tr[threadIdx.x + o]; s_e[threadIdx.x] += s_e[threadIdx.x + o]; s_k[threadIdx.x] += s_k[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int ce = (d.rows - extg * 2 / (int)cfg.resolution) * (d.cols - extg * 2 / (int)cfg.resolution);
    const double explored = (double)s_e[0] / ce;
    const double pk = (double)s_k[0] / d.V;
    metrics[8 * b + 0] = explored;
    metrics[8 * b + 1] = s_tr[0];                 // utility at distance 0
    metrics[8 * b + 2] = cfg.dist_w0 - (cfg.dist_w0 - cfg.dist_w1) * pk;   // distance weight
    metrics[8 * b + 3] = (double)s_k[0];
    metrics[8 * b + 6] = dist[b];
    done[b] = (sim_step[b] > cfg.max_steps || explored > 0.85 || status[b] == DGE_ECAP) ? 1 : 0;
  }
}

// =============================================================== fused, pose-centric rebuild ===
// k_vmap_env: ONE CTA per environment does the whole rebuild -- pose digest (k_vmap_prep), occupancy + ordered
// covariance-intersection fold (k_vmap_cells) and the map metrics (k_vmap_metrics) -- in one launch.
//
// The cell-centric kernel above tests every (pose, cell) pair of a tile for proximity (1 visit per ~100 tests at
// BASELINE config C4).  Here the work is driven by the poses: a pose can only touch the W x W block of cells around
// it (W = 2 ceil(max_range / res) + 1 = 7), and a visit is split in two:
//   * PREDICT (order-independent, ~3/4 of the arithmetic): for every (pose, cell of its block) pair the gates and the
//     predicted information Lambda_new = (Hl^-1 (R + Hx Sigma Hx^T) Hl^-T)^-1 (VirtualMap.cpp:213-229) -- one thread per
//     pair, every warp of the CTA busy, KC poses (a chunk) at a time into a shared-memory record buffer;
//   * FOLD (order-dependent, VirtualMap.cpp:307-313 / 364-377): one thread per cell of the chunk's bounding box walks the
//     chunk's poses in trajectory order and folds the records that hit its cell into the cell state (2x2 information,
//     visibility count, flags), which lives in shared memory for the whole map (28 bytes per cell).
// The record buffer is double-buffered: the prediction of chunk i+1 is issued in front of the fold of chunk i (one barrier per
// chunk), so the latency-bound fold chains of a few warps run beside the throughput-bound prediction of all of them.
// HBM traffic is the algorithmic minimum: the trajectory is read once (digest written and re-read through L1/L2),
// every output byte is written once.  Arithmetic, predicates and summation orders are those of the kernels above
// (bit-identical results, same parity tests).

// reciprocal for the fused kernel: hardware seed (rcp.approx.ftz.f64, ~2^-23) + two Newton steps = full double accuracy
// up to ~2 ulp, half the dependent latency of the IEEE division sequence (operands here are O(1e-3..1e15): no
// subnormals; x = 0 gives inf like the unguarded division of the reference, quirk q11)
__device__ __forceinline__ double vm_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}
constexpr unsigned ST_UPD = 1u << 30, ST_LM = 1u << 31, ST_CNT = ST_UPD - 1;
constexpr int ENV_THREADS = 256;

struct EnvArgs {
  VmapCfg c;
  int Tstride, Tfixed, Lstride, Lfixed, hw, W;         // hw = ceil(max_range / res), W = 2 hw + 1
  int kc;                                              // poses per chunk
  const int32_t *n_poses;
  const double *pose, *cov, *info;                     // [n,Tstride,3], [n,Tstride,6], nullable [n,Tstride,6]
  double *prep;                                        // [n,Tstride,PREP_W] scratch (L1/L2 resident)
  const double *lm; const uint8_t *lm_obs;             // [n,Lstride,2], nullable [n,Lstride]
  double *prob, *vinfo; int32_t *seen;                 // outputs; seen nullable
  const uint8_t *mask;
  // metrics (engine path only; metrics == nullptr skips them)
  dge_config cfg; DgeDims d;
  const int32_t *sim_step, *status, *meas_ptr; const double *dist; double *metrics; uint8_t *done;
  unsigned long long *counters; const uint8_t *step_kind;
  long long *clocks;                                   // nullable [n,4]: SM clock at start / after digest / after fold / end (thread 0)
  const int32_t *order;                                // nullable [n]: block -> env (cost-ordered placement of the step kernels, dge_slam.cu)
};

template <int WT>   // WT = block width W when known at compile time (7 for the reference's sensor / resolution), 0 = generic
__global__ void __launch_bounds__(ENV_THREADS, 2) k_vmap_env(EnvArgs a) {
  const int b = a.order ? a.order[blockIdx.x] : blockIdx.x;
  if (a.mask && !a.mask[b]) return;
  const VmapCfg &c = a.c;
  const int T = a.n_poses ? a.n_poses[b] : a.Tfixed;
  const int tid = threadIdx.x, nthr = ENV_THREADS;
  const int V = c.rows * c.cols, W = WT ? WT : a.W, WW = W * W, KC = a.kc, hw = a.hw;
  extern __shared__ __align__(16) unsigned char vm_smem[];
  double *sxx = reinterpret_cast<double *>(vm_smem), *sxy = sxx + V, *syy = sxy + V;
  unsigned *sst = reinterpret_cast<unsigned *>(syy + V);       // count | ST_UPD | ST_LM
  unsigned *smk = sst + ((V + 1) & ~1);                        // per cell: which poses of the chunk update it (bit kk; low / high half = buffer 0 / 1)
  int *s_fr = reinterpret_cast<int *>(smk + ((V + 1) & ~1));   // [Tstride] first row of every pose's block
  int *s_fc = s_fr + ((a.Tstride + 1) & ~1);                   // [Tstride] first column
  double *s_red = reinterpret_cast<double *>(s_fc + ((a.Tstride + 1) & ~1));   // [256] + 2 x int[256]
  double *rec0 = s_red + 256 + 256;                            // 2 x { nxx, nxy, nyy, ndet [KC*WW] }
  const int RS = KC * WW;                                      // records per buffer
  int4 *s_box = reinterpret_cast<int4 *>((reinterpret_cast<uintptr_t>(rec0 + 8 * RS) + 15) & ~uintptr_t(15));   // [ceil(Tstride / KC)] cell box [r0, r1) x [c0, c1) of every chunk

  if (a.clocks && tid == 0) a.clocks[4 * b] = clock64();
  // ---- init cell state, digest the trajectory -------------------------------------------------------------
  for (int i = tid; i < V; i += nthr) { sxx[i] = c.i0; sxy[i] = 0.0; syy[i] = c.i0; sst[i] = 0u; smk[i] = 0u; }
  const double *ps = a.pose + (size_t)b * a.Tstride * 3, *cv = a.cov + (size_t)b * a.Tstride * 6;
  double *pr = a.prep + (size_t)b * a.Tstride * PREP_W;
  for (int k = tid; k < T; k += nthr) {
    const double px = ps[3 * k], py = ps[3 * k + 1];
    s_fr[k] = (int)floor((py - c.map_min_y) / c.res) - hw;
    s_fc[k] = (int)floor((px - c.map_min_x) / c.res) - hw;     // (defines the candidate block only)
    double s, co;
    sincos(ps[3 * k + 2], &s, &co);
    double *o = pr + (size_t)k * PREP_W;
    double S[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) S[i] = cv[6 * k + i];
    double det_info;
    if (a.info) {
      const double *q = a.info + ((size_t)b * a.Tstride + k) * 6;
      det_info = q[0] * (q[3] * q[5] - q[4] * q[4]) - q[1] * (q[1] * q[5] - q[4] * q[2]) + q[2] * (q[1] * q[4] - q[3] * q[2]);
    } else {
      const double dc = S[0] * (S[3] * S[5] - S[4] * S[4]) - S[1] * (S[1] * S[5] - S[4] * S[2]) + S[2] * (S[1] * S[4] - S[3] * S[2]);
      det_info = 1.0 / dc;
    }
    o[0] = px; o[1] = py; o[2] = co; o[3] = s;
#pragma unroll
    for (int i = 0; i < 6; ++i) o[4 + i] = S[i];
    o[10] = (det_info < 1e-10) ? 0.0 : 1.0;   // VirtualMap.cpp:293
    o[11] = 0.0;
  }
  __syncthreads();
  const int nch = (T + KC - 1) / KC;
  for (int ch = tid; ch < nch; ch += nthr) {                 // bounding box of the cells a chunk's poses can touch
    int r0 = 0x3fffffff, r1 = -0x3fffffff, c0 = 0x3fffffff, c1 = -0x3fffffff;
    for (int k = ch * KC; k < min(T, (ch + 1) * KC); ++k) {
      const int fr = s_fr[k], fc = s_fc[k];
      r0 = min(r0, fr); r1 = max(r1, fr); c0 = min(c0, fc); c1 = max(c1, fc);
    }
    s_box[ch] = make_int4(max(r0, 0), min(r1 + W, c.rows), max(c0, 0), min(c1 + W, c.cols));
  }
  // landmark cells (OccupancyMap.cpp:126-131)
  {
    const double *l = a.lm + (size_t)b * a.Lstride * 2;
    for (int j = tid; j < a.Lfixed; j += nthr) {
      if (a.lm_obs && !a.lm_obs[(size_t)b * a.Lstride + j]) continue;
      const int lr = (int)floor((l[2 * j + 1] - c.map_min_y) / c.res), lc = (int)floor((l[2 * j] - c.map_min_x) / c.res);
      if (lr >= 0 && lr < c.rows && lc >= 0 && lc < c.cols) atomicOr(&sst[lr * c.cols + lc], ST_LM);
    }
  }
  // this thread's (pose of the chunk, cell of its block) pairs are the same in every chunk: pair p = tid + r * nthr
  constexpr int MAXR = 3;
  int pk[MAXR], pdr[MAXR], pdc[MAXR];
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    const int p = tid + r * nthr;
    pk[r] = p / WW;
    const int j = p - pk[r] * WW;
    pdr[r] = j / W; pdc[r] = j - pdr[r] * W;
  }
  // (the clock reads hang on the barrier's result: BAR.SYNC defers blocking, a bare clock read would run ahead of it)
  const int nb1 = __syncthreads_count(1);
  if (a.clocks && tid == 0 && nb1) a.clocks[4 * b + 1] = clock64();

  // ---- PREDICT: records of chunk ch (poses [ch KC, ch KC + kc)) into buffer ch & 1.  The visibility count is order-independent
  // (integer, quirk q8): it goes straight into the cell state; which poses UPDATE a cell goes into the cell's chunk mask.
  auto predict_chunk = [&](int ch) {
    const int k0 = ch * KC, kc = min(KC, T - k0);
    double *rxx = rec0 + (ch & 1) * 4 * RS, *rxy = rxx + RS, *ryy = rxy + RS, *rdt = ryy + RS;
    const int mshift = (ch & 1) * 16;
#pragma unroll
    for (int r = 0; r < MAXR; ++r) {
      const int p = tid + r * nthr, kk = pk[r];
      if (kk >= kc) break;
      const int k = k0 + kk;
      const int row = s_fr[k] + pdr[r], col = s_fc[k] + pdc[r];
      const bool cell_on = row >= 0 && row < c.rows && col >= 0 && col < c.cols;
      const int idx = row * c.cols + col;
      const double *q = pr + (size_t)k * PREP_W;
      const double2 p01 = *reinterpret_cast<const double2 *>(q), p23 = *reinterpret_cast<const double2 *>(q + 2);
      const double cx = c.map_min_x + c.res * (col + 0.5), cy = c.map_min_y + c.res * (row + 0.5);
      const double dx = cx - p01.x, dy = cy - p01.y;
      const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
      bool in_max, out_min;   // range gates on d2, the reference's sqrt comparison inside a 1e-9 band (see k_vmap_cells)
      if (d2 < c.max_r2_lo) in_max = true; else if (d2 > c.max_r2_hi) in_max = false; else in_max = __dsqrt_rn(d2) < c.max_range;
      if (d2 > c.min_r2_hi) out_min = true; else if (d2 < c.min_r2_lo) out_min = false; else out_min = __dsqrt_rn(d2) > c.min_range;
      const bool inr = cell_on && in_max;
      const double co = p23.x, si = p23.y;
      const double qx = co * dx + si * dy, qy = -si * dx + co * dy;
      bool in_fov;
      if (c.fov_wide && (qx > 0.0 || fabs(qy) > -qx * c.wedge_tan)) in_fov = true;
      else if (inr) { const double bb = atan2(qy, qx); in_fov = bb < c.max_bearing && bb > c.min_bearing; }
      else in_fov = false;
      const bool vis = inr && in_fov;
      if (!vis) continue;
      atomicAdd(&sst[idx], 1u);                          // visibility count (q8)
      if (!(out_min && q[10] != 0.0)) continue;          // full check (q10) ; det(info) gate
      atomicOr(&smk[idx], 1u << (mshift + kk));
      const double2 p45 = *reinterpret_cast<const double2 *>(q + 4), p67 = *reinterpret_cast<const double2 *>(q + 6);
      const double2 p89 = *reinterpret_cast<const double2 *>(q + 8);
      const double Sxx = p45.x, Sxy = p45.y, Sxt = p67.x, Syy = p67.y, Syt = p89.x, Stt = p89.y;
      const double qxx = qx * qx, qyy = qy * qy, qxy = qx * qy;
      const double P00 = qyy * (c.rb + Stt) + Sxx - 2.0 * qy * Sxt;
      const double P01 = Sxy - qy * Syt + qx * Sxt - qxy * (c.rb + Stt);
      const double P11 = qxx * (c.rb + Stt) + Syy + 2.0 * qx * Syt;
      const double Q = P00 * qyy - 2.0 * P01 * qxy + P11 * qxx;
      const double inv = vm_rcp((P00 * P11 - P01 * P01) * d2 + c.rr * Q);
      const double lb00 = (P11 * d2 + c.rr * qyy) * inv, lb01 = -(P01 * d2 + c.rr * qxy) * inv, lb11 = (P00 * d2 + c.rr * qxx) * inv;
      const double cc = co * co, ss = si * si, cs = co * si;
      rxx[p] = cc * lb00 - 2.0 * cs * lb01 + ss * lb11;
      rxy[p] = cs * (lb00 - lb11) + (cc - ss) * lb01;
      ryy[p] = ss * lb00 + 2.0 * cs * lb01 + cc * lb11;
      rdt[p] = d2 * inv;
    }
  };
  // ---- FOLD: the cells of the chunk's bounding box, one thread per cell; a cell folds the poses of its chunk mask in
  // trajectory order (lowest bit first) -----------------------------------------------------------------------------------
  auto fold_chunk = [&](int ch) {
    const int k0 = ch * KC;
    const double *rxx = rec0 + (ch & 1) * 4 * RS, *rxy = rxx + RS, *ryy = rxy + RS, *rdt = ryy + RS;
    const int mshift = (ch & 1) * 16;
    const int4 bx = s_box[ch];
    const int r0 = bx.x, r1 = bx.y, c0 = bx.z, c1 = bx.w;
    const int bw = c1 - c0, nbc = (r1 > r0 && bw > 0) ? (r1 - r0) * bw : 0;
    for (int ci = tid; ci < nbc; ci += nthr) {
      const int rr_ = ci / bw, row = r0 + rr_, col = c0 + ci - rr_ * bw;
      const int idx = row * c.cols + col;
      unsigned m = (smk[idx] >> mshift) & 0xffffu;
      if (!m) continue;
      atomicAnd(&smk[idx], ~(0xffffu << mshift));        // (the other half may be set concurrently by the next chunk's prediction)
      double ixx = sxx[idx], ixy = sxy[idx], iyy = syy[idx];
      bool first = !(sst[idx] & ST_UPD);                 // the first hit overwrites the prior (VirtualMap.cpp:307-309)
      while (m) {
        const int kk = __ffs(m) - 1;
        m &= m - 1;
        const int j = kk * WW + (row - s_fr[k0 + kk]) * W + (col - s_fc[k0 + kk]);
        const double nxx = rxx[j], nxy = rxy[j], nyy = ryy[j], bdet = rdt[j];
        // covariance intersection on information matrices (VirtualMap.cpp:364-377, q11)
        const double aa = ixx * iyy - ixy * ixy;
        const double cm = iyy * nxx - 2.0 * ixy * nxy + ixx * nyy;
        const double d = aa + bdet - cm;
        double w = 0.5 * (2.0 * bdet - cm) * vm_rcp(d);
        w = ((w < 0 && d < 0) || (w > 1 && d > 0)) ? 0.0 : (((w < 0 && d > 0) || (w > 1 && d < 0)) ? 1.0 : w);
        ixx = first ? nxx : w * ixx + (1.0 - w) * nxx;
        ixy = first ? nxy : w * ixy + (1.0 - w) * nxy;
        iyy = first ? nyy : w * iyy + (1.0 - w) * nyy;
        first = false;
      }
      sxx[idx] = ixx; sxy[idx] = ixy; syy[idx] = iyy;
      atomicOr(&sst[idx], ST_UPD);
    }
  };
  if (nch > 0) predict_chunk(0);
  __syncthreads();
  for (int ch = 0; ch < nch; ++ch) {
    if (ch + 1 < nch) predict_chunk(ch + 1);    // issued in front of the fold: independent work for the warps that wait on fold chains
    fold_chunk(ch);
    __syncthreads();
  }
  if (a.clocks && tid == 0) a.clocks[4 * b + 2] = clock64();
  // ---- write the map: every output byte once, coalesced -------------------------------------------------------
  const size_t cell0 = (size_t)b * V;
  for (int i = tid; i < V; i += nthr) {
    const unsigned st = sst[i];
    const int cnt = (int)(st & ST_CNT);
    a.prob[cell0 + i] = (st & ST_LM) ? c.ptab[5] : c.ptab[min(cnt, 4)];
    if (a.seen) a.seen[cell0 + i] = (st & ST_LM) ? -1 : cnt;
  }
  for (int i = tid; i < 3 * V; i += nthr) {
    const int cell = i / 3, j = i - 3 * cell;
    a.vinfo[cell0 * 3 + i] = j == 0 ? sxx[cell] : (j == 1 ? sxy[cell] : syy[cell]);
  }
  if (a.clocks && tid == 0) a.clocks[4 * b + 3] = clock64();
  if (!a.metrics) return;

  // ---- metrics (k_vmap_metrics, same summation order: 256 strided partial sums, fixed tree) ---------------------
  if (tid == 0 && a.counters && a.step_kind[b]) {
    atomicAdd(&a.counters[0], 1ull);
    atomicAdd(&a.counters[1], (unsigned long long)T);
    atomicAdd(&a.counters[2], (unsigned long long)a.meas_ptr[(size_t)b * (a.d.Tmax + 1) + T]);
  }
  int *s_e = reinterpret_cast<int *>(s_red + 256), *s_k = s_e + 256;
  {
    const int extg = 20;
    int n_exp = 0, n_known = 0;
    double tr = 0.0;
    for (int i = tid; i < V; i += 256) {
      const double x = (i % c.cols + 0.5) * a.cfg.resolution + a.cfg.map_min_x, y = (i / c.cols + 0.5) * a.cfg.resolution + a.cfg.map_min_y;
      const unsigned st = sst[i];
      const double pv = (st & ST_LM) ? c.ptab[5] : c.ptab[min((int)(st & ST_CNT), 4)];
      if ((pv < 0.49 || pv > 0.6) && a.cfg.map_min_x + extg <= x && x <= a.cfg.map_max_x - extg && a.cfg.map_min_y + extg <= y && y <= a.cfg.map_max_y - extg) ++n_exp;
      if (pv < a.cfg.occupancy_threshold) ++n_known;
      const double qa = sxx[i], qb = sxy[i], qc = syy[i];
      tr += (qa + qc) / (qa * qc - qb * qb);
    }
    s_red[tid] = tr; s_e[tid] = n_exp; s_k[tid] = n_known;
  }
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { s_red[tid] += s_red[tid + o]; s_e[tid] += s_e[tid + o]; s_k[tid] += s_k[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    const int extg = 20;
    const int ce = (a.d.rows - extg * 2 / (int)a.cfg.resolution) * (a.d.cols - extg * 2 / (int)a.cfg.resolution);
    const double explored = (double)s_e[0] / ce;
    const double pk = (double)s_k[0] / a.d.V;
    a.metrics[8 * b + 0] = explored;
    a.metrics[8 * b + 1] = s_red[0];
    a.metrics[8 * b + 2] = a.cfg.dist_w0 - (a.cfg.dist_w0 - a.cfg.dist_w1) * pk;
    a.metrics[8 * b + 3] = (double)s_k[0];
    a.metrics[8 * b + 6] = a.dist[b];
    a.done[b] = (a.sim_step[b] > a.cfg.max_steps || explored > 0.85 || a.status[b] == DGE_ECAP) ? 1 : 0;
  }
}

VmapCfg make_cfg(const dge_config &g, int rows, int cols) {
  VmapCfg c;
  c.map_min_x = g.map_min_x; c.map_min_y = g.map_min_y; c.res = g.resolution;
  c.max_range = g.max_range; c.min_range = g.min_range; c.max_bearing = g.max_bearing; c.min_bearing = g.min_bearing;
  c.rb = g.bearing_noise * g.bearing_noise; c.rr = g.range_noise * g.range_noise;
  c.max_r2_lo = g.max_range * g.max_range * (1.0 - 1e-9); c.max_r2_hi = g.max_range * g.max_range * (1.0 + 1e-9);
  c.min_r2_lo = g.min_range * g.min_range * (1.0 - 1e-9); c.min_r2_hi = g.min_range * g.min_range * (1.0 + 1e-9);
  c.i0 = 1.0 / (g.sigma0 * g.sigma0);
  // OccupancyMap.h:10-19 evaluated on the host with libm, by repeated addition + clamping exactly
  // like OccupancyMap.cpp:55-62 (incl. MAX_LOGODDS = LOGODDS2PROB(0.95), q7)
  const double LF = log(0.3 / (1.0 - 0.3)), LO = log(0.7 / (1.0 - 0.7)), MINL = log(0.05 / (1.0 - 0.05));
  const double MAXL = exp(0.95) / (1.0 + exp(0.95));
  double l = log(0.5 / (1.0 - 0.5));
  for (int n = 0; n <= 4; ++n) {
    c.ptab[n] = exp(l) / (1.0 + exp(l));
    l = fmin(MAXL, fmax(MINL, l + LF));
  }
  const double lo = fmin(MAXL, fmax(MINL, log(0.5 / (1.0 - 0.5)) + LO));
  c.ptab[5] = exp(lo) / (1.0 + exp(lo));
  c.fov_wide = (g.max_bearing >= DGE_PI / 2 && g.min_bearing <= -DGE_PI / 2) ? 1 : 0;
  const double half = fmax(DGE_PI - g.max_bearing, DGE_PI + g.min_bearing);
  c.wedge_tan = tan(half) * (1.0 + 1e-6) + 1e-12;
  c.rows = rows; c.cols = cols;
  return c;
}

}  // namespace

int dge_vmap_nchunk(int T) { return (T + VCH - 1) / VCH; }
int dge_vmap_prep_width() { return PREP_W; }

namespace {
// geometry of the fused kernel for a map: chunk length and shared memory; false if the map does not fit
size_t env_smem(size_t V, int Tstride, int kc, int W) {
  const size_t rs = (size_t)kc * W * W;
  return V * 3 * sizeof(double) + 2 * ((V + 1) & ~(size_t)1) * sizeof(unsigned) + 2 * (size_t)((Tstride + 1) & ~1) * sizeof(int) +
         256 * (sizeof(double) + 2 * sizeof(int)) + 8 * rs * sizeof(double) + (size_t)((Tstride + kc - 1) / kc) * 16 + 32;
}
bool env_plan(const dge_config &g, int rows, int cols, int Tstride, int *hw, int *W, int *kc, size_t *smem) {
  *hw = (int)ceil(g.max_range / g.resolution);
  *W = 2 * *hw + 1;
  if (*W > 15) return false;
  const size_t V = (size_t)rows * cols;
  // chunk length: (pose, cell) pairs of a chunk fill whole rounds of the CTA's threads (kc W^2 just below a multiple of 256;
  // 15 / 10 / 5 for W = 7), at most 16 poses (chunk mask), the longest that leaves room for two CTAs per SM
  int cand[3], nc = 0;
  for (int r = 3; r >= 1; --r) { const int k = (r * ENV_THREADS) / (*W * *W); if (k >= 1 && k <= 16 && (nc == 0 || cand[nc - 1] != k)) cand[nc++] = k; }
  if (nc == 0) cand[nc++] = 1;
  for (int i = 0; i < nc; ++i) {
    *kc = cand[i]; *smem = env_smem(V, Tstride, cand[i], *W);
    if (*smem <= 113 * 1024) return true;
  }
  *kc = cand[nc - 1]; *smem = env_smem(V, Tstride, *kc, *W);
  return *smem <= 226 * 1024;
}
int env_launch(EnvArgs &a, const dge_config &g, int n, int rows, int cols, cudaStream_t st) {
  size_t smem;
  if (!env_plan(g, rows, cols, a.Tstride, &a.hw, &a.W, &a.kc, &smem)) return 1;   // caller falls back to the cell-centric kernels
  void (*kern)(EnvArgs) = a.W == 7 ? k_vmap_env<7> : k_vmap_env<0>;
  static size_t configured[2] = {0, 0};
  size_t &cf = configured[a.W == 7 ? 1 : 0];
  if (smem > 48 * 1024 && smem > cf) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DGE_ECUDA;
    cf = smem;
  }
  kern<<<n, ENV_THREADS, smem, st>>>(a);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}
}  // namespace

int dge_launch_vmap(dge_engine *e, const uint8_t *mask, cudaStream_t st) {
  {
    EnvArgs a;
    a.c = make_cfg(e->cfg, e->d.rows, e->d.cols);
    a.Tstride = e->d.Tmax; a.Tfixed = 0; a.Lstride = e->d.Lt; a.Lfixed = e->d.Lt;
    a.n_poses = e->n_poses; a.pose = e->est_pose; a.cov = e->pose_cov; a.info = e->pose_info; a.prep = e->vm_prep;
    a.lm = e->est_l; a.lm_obs = e->observed; a.prob = e->prob; a.vinfo = e->vinfo; a.seen = e->seen; a.mask = mask;
    a.cfg = e->cfg; a.d = e->d; a.sim_step = e->sim_step; a.status = e->status; a.meas_ptr = e->meas_ptr; a.dist = e->dist;
    a.metrics = e->metrics; a.done = e->done; a.counters = e->count_steps ? e->counters : nullptr; a.step_kind = e->step_kind;
    a.clocks = nullptr;
    a.order = (e->step_order_live && mask == e->active) ? e->step_order : nullptr;   // same envs, same costs as the SLAM launch just before
    const int rc = env_launch(a, e->cfg, e->d.B, e->d.rows, e->d.cols, st);
    if (rc != 1) return rc;
  }

  const VmapCfg c = make_cfg(e->cfg, e->d.rows, e->d.cols);
  const int nchm = dge_vmap_nchunk(e->d.Tmax);
  k_vmap_prep<<<e->d.B, 128, 0, st>>>(c, e->d.Tmax, e->n_poses, 0, e->est_pose, e->pose_cov, e->pose_info, e->vm_prep, e->vm_cbox, nchm, mask);
  const int tiles = ((e->d.cols + TILE - 1) / TILE) * ((e->d.rows + TILE - 1) / TILE);
  k_vmap_cells<<<dim3(tiles, e->d.B), TILE * TILE, 0, st>>>(c, e->d.Tmax, e->n_poses, 0, e->vm_prep, e->vm_cbox, nchm, e->est_l, e->observed,
                                                             e->d.Lt, e->d.Lt, e->prob, e->vinfo, e->seen, mask);
  k_vmap_metrics<<<e->d.B, 256, 0, st>>>(e->cfg, e->d, e->prob, e->vinfo, e->sim_step, e->status, e->dist, e->metrics, e->done, mask,
                                         e->n_poses, e->meas_ptr, e->count_steps ? e->counters : nullptr, e->step_kind);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}

// stand-alone rebuild on caller-provided belief states (C4 roofline sweep, kernel-level parity)
int dge_vmap_standalone(const dge_config *cfg, int n, int T, const double *pose, const double *cov, int L, const double *lm,
                        double *prob, double *vinfo, int32_t *seen, double *prep_ws, double *cbox_ws, cudaStream_t st) {
  const int cols = (int)floor((cfg->map_max_x - cfg->map_min_x) / cfg->resolution);
  const int rows = (int)floor((cfg->map_max_y - cfg->map_min_y) / cfg->resolution);
  const VmapCfg c = make_cfg(*cfg, rows, cols);
  {
    EnvArgs a;
    a.c = c;
    a.Tstride = T; a.Tfixed = T; a.Lstride = L; a.Lfixed = L;
    a.n_poses = nullptr; a.pose = pose; a.cov = cov; a.info = nullptr; a.prep = prep_ws;
    a.lm = lm; a.lm_obs = nullptr; a.prob = prob; a.vinfo = vinfo; a.seen = seen; a.mask = nullptr;
    a.metrics = nullptr; a.done = nullptr; a.counters = nullptr; a.step_kind = nullptr;
    a.sim_step = nullptr; a.status = nullptr; a.meas_ptr = nullptr; a.dist = nullptr;
    a.cfg = *cfg; a.d = DgeDims{};
    a.clocks = reinterpret_cast<long long *>(cbox_ws);   // the chunk-box scratch is unused by the fused kernel: phase clocks for dev profiling
    a.order = nullptr;
    const int rc = env_launch(a, *cfg, n, rows, cols, st);
    if (rc != 1) return rc;
  }
  const int nchm = dge_vmap_nchunk(T);
  k_vmap_prep<<<n, 128, 0, st>>>(c, T, nullptr, T, pose, cov, nullptr, prep_ws, cbox_ws, nchm, nullptr);
  const int tiles = ((cols + TILE - 1) / TILE) * ((rows + TILE - 1) / TILE);
  k_vmap_cells<<<dim3(tiles, n), TILE * TILE, 0, st>>>(c, T, nullptr, T, prep_ws, cbox_ws, nchm, lm, nullptr, L, L, prob, vinfo, seen, nullptr);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}
