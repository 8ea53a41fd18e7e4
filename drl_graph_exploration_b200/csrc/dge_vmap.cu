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
#include <cstdlib>

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
                                                      const int32_t *n_poses, const int32_t *meas_ptr, unsigned long long *counters, const uint8_t *step_kind,
                                                      const uint8_t *observed) {
  const int b = blockIdx.x;
  if (mask && !mask[b]) return;
  if (threadIdx.x == 0 && counters && step_kind[b] == 1) {   // integer work counters (order-independent): env-steps, sum T, sum M
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
    // exploration_env.py:416-419: a world in which the four forced steps of reset() saw no landmark is regenerated -- here the episode ends
    // at once (uncounted) and the next reset draws the next world of the env's seed sequence
    bool blind = false;
    if (step_kind && step_kind[b] == 2) { int n = 0; for (int j = 0; j < d.Lt; ++j) n += observed[(size_t)b * d.Lt + j]; blind = n == 0; }
    done[b] = (sim_step[b] > cfg.max_steps || explored > 0.85 || status[b] == DGE_ECAP || blind) ? 1 : 0;
  }
}

// =============================================================== fused, pose-centric rebuild ===
// k_vmap_env: ONE CTA per environment does the whole rebuild -- pose digest, occupancy + ordered covariance-intersection
// fold and the map metrics -- in one launch, the map state (2x2 information + visibility count + flags, 28 bytes per cell)
// living in shared memory from the first pose to the write-out.
//
// The work is driven by the poses: a pose can only touch the W x W block of cells around it (W = 2 ceil(max_range / res) + 1
// = 7), and a visit is split in two.  Up front, one thread per pose digests the trajectory (cos / sin, covariance, det gate,
// first row / column of the block) into scratch: the transcendental latency is paid once, in parallel.  Then, per chunk of
// KC <= 16 consecutive poses (digest rows scratch -> registers -> shared memory, one chunk ahead):
//   * PREDICT (order-independent, ~3/4 of the arithmetic): every warp takes a contiguous slice of the chunk's (pose, cell)
//     pairs; a branch-free GATE pass (range / field-of-view predicates, GR pairs per lane side by side; the 1e-9 bands around
//     the range limits and the rear wedge go to a rare exact path with the reference's own sqrt / atan2 comparisons; integer
//     visibility count) ballot-compacts the pairs that update a cell into a warp-private list, and the SPD pass -- the predicted
//     information Lambda_new = (Hl^-1 (R + Hx Sigma Hx^T) Hl^-T)^-1 (VirtualMap.cpp:213-229) -- runs on the dense list only
//     (no lane of an fp64 instruction idles on a pair outside the sensor disc).  Records go to a shared-memory buffer; a pair
//     without update leaves a negative sentinel.
//   * FOLD (order-dependent, VirtualMap.cpp:307-313 / 364-377): one thread per cell of the chunk's box collects the chunk's
//     poses that hit its cell (bit mask from the sentinels), then folds their records in trajectory order, the next record
//     loaded in front of the current fold's dependent chain.
// Two barriers per chunk.  HBM traffic is the algorithmic minimum: the trajectory is read once, every output byte is
// written once -- the information matrices by bulk copies (cp.async.bulk shared -> global) of the cell-interleaved state.

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
constexpr int DIG_W = 12;          // doubles per digested pose in shared memory: x y c s Sxx Sxy Sxt Syy Syt Stt valid (fr|fc)

struct EnvArgs {
  VmapCfg c;
  int Tstride, Tfixed, Lstride, Lfixed, hw, W;         // hw = ceil(max_range / res), W = 2 hw + 1
  int kc;                                              // poses per chunk (<= 16)
  int ppw;                                             // (pose, cell) pairs per warp and chunk = ceil(kc W^2 / warps), rounded up to 8
  const int32_t *n_poses;
  const double *pose, *cov, *info;                     // [n,Tstride,3], [n,Tstride,6], nullable [n,Tstride,6]
  double *prep;                                        // [n,Tstride,PREP_W] digest scratch (written and re-read by the env's CTA: L2)
  const double *lm; const uint8_t *lm_obs;             // [n,Lstride,2], nullable [n,Lstride]
  double *prob, *vinfo; int32_t *seen;                 // outputs; seen nullable
  const uint8_t *mask;
  // metrics (engine path only; metrics == nullptr skips them)
  dge_config cfg; DgeDims d;
  const int32_t *sim_step, *status, *meas_ptr; const double *dist; double *metrics; uint8_t *done;
  unsigned long long *counters; const uint8_t *step_kind;
  long long *clocks;                                   // nullable [n,4]: SM clock at start / after digest / after fold / end (thread 0)
  long long *phase;                                    // nullable [n,8]: cycles summed over the chunks (thread 0): gate, SPD, wait, fold, wait
  const int32_t *order;                                // nullable [n]: block -> env (cost-ordered placement of the step kernels, dge_slam.cu)
};

// SPEC (512 threads): warp-specialised -- warps [0, 8) PREDICT chunk i + 1 into the second record buffer while warps [8, 16) FOLD chunk i
// (the fold is a handful of long dependent chains, the prediction is throughput work: run beside each other, not behind);
// the two groups meet at named barriers (records of a buffer full / free), never at a CTA-wide barrier inside the loop.
template <int WT, int NTHR, bool SPEC>   // WT = block width W when known at compile time (7 for the reference's sensor / resolution), 0 = generic
__global__ void __launch_bounds__(NTHR, 2) k_vmap_env(EnvArgs a) {
  constexpr int NWARP = SPEC ? 8 : NTHR / 32;            // warps that predict
  constexpr int GR = 4;                                  // pairs a lane gates side by side
  const int b = a.order ? a.order[blockIdx.x] : blockIdx.x;
  if (a.mask && !a.mask[b]) return;
  const VmapCfg &c = a.c;
  const int T = a.n_poses ? a.n_poses[b] : a.Tfixed;
  const int tid = threadIdx.x, nthr = NTHR, lane = tid & 31, warp = tid >> 5;
  const int V = c.rows * c.cols, W = WT ? WT : a.W, WW = W * W, KC = a.kc, hw = a.hw, PPW = a.ppw;
  extern __shared__ __align__(16) unsigned char vm_smem[];
  double *sinf = reinterpret_cast<double *>(vm_smem);          // [V][3] xx xy yy, cell-interleaved = the layout of the output
  unsigned *sst = reinterpret_cast<unsigned *>(sinf + 3 * (size_t)V);   // count | ST_UPD | ST_LM
  double *dig = reinterpret_cast<double *>(sst + ((V + 1) & ~1));        // [2][KC][DIG_W]
  const int RS = KC * WW;                                       // records per chunk
  double *rec0 = dig + 2 * KC * DIG_W;                          // records [NB][3][RS]: one buffer (predict and fold separated by barriers) or two (SPEC)
  constexpr int NB = SPEC ? 2 : 1;
  double *s_red = rec0 + NB * 3 * RS;                           // [256] + 2 x int[256]
  int *s_fr = reinterpret_cast<int *>(s_red + 256 + 256);       // [2][16] first row of every pose's block
  int *s_fc = s_fr + 32;                                        // [2][16] first column
  int4 *s_box = reinterpret_cast<int4 *>((reinterpret_cast<uintptr_t>(s_fc + 32) + 15) & ~uintptr_t(15));   // [2] cell box [r0, r1) x [c0, c1) of the chunk
  unsigned short *s_list = reinterpret_cast<unsigned short *>(s_box + 2);   // [NWARP][PPW] compacted pair indices

  if (a.clocks && tid == 0) a.clocks[4 * b] = clock64();
  const double *ps = a.pose + (size_t)b * a.Tstride * 3, *cv = a.cov + (size_t)b * a.Tstride * 6;
  const double *pinfo = a.info ? a.info + (size_t)b * a.Tstride * 6 : nullptr;
  const int nch = (T + KC - 1) / KC;
  double *pr = a.prep + (size_t)b * a.Tstride * PREP_W;

  // ---- init cell state, then (one barrier later: the landmark flags need it) the digest of the whole trajectory, one thread per pose,
  // and the landmark cells -- their cold global loads are in flight together ----------------------------------------------------
  for (int i = tid; i < V; i += nthr) { sinf[3 * i] = c.i0; sinf[3 * i + 1] = 0.0; sinf[3 * i + 2] = c.i0; sst[i] = 0u; }
  __syncthreads();
  for (int k = nthr - 1 - tid; k < T; k += nthr) {   // (highest threads first: the low ones carry the landmark loop)
    const double px = ps[3 * k], py = ps[3 * k + 1];
    const int fr = (int)floor((py - c.map_min_y) / c.res) - hw, fc = (int)floor((px - c.map_min_x) / c.res) - hw;   // (defines the candidate block only)
    double s, co;
    sincos(ps[3 * k + 2], &s, &co);
    double S[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) S[i] = cv[6 * k + i];
    double det_info;
    if (pinfo) {
      const double *q = pinfo + 6 * k;
      det_info = q[0] * (q[3] * q[5] - q[4] * q[4]) - q[1] * (q[1] * q[5] - q[4] * q[2]) + q[2] * (q[1] * q[4] - q[3] * q[2]);
    } else {
      const double dc = S[0] * (S[3] * S[5] - S[4] * S[4]) - S[1] * (S[1] * S[5] - S[4] * S[2]) + S[2] * (S[1] * S[4] - S[3] * S[2]);
      det_info = 1.0 / dc;
    }
    double *o = (k < KC) ? dig + (size_t)k * DIG_W : pr + (size_t)k * PREP_W;   // the first chunk's rows go straight to shared memory
    o[0] = px; o[1] = py; o[2] = co; o[3] = s;
#pragma unroll
    for (int i = 0; i < 6; ++i) o[4 + i] = S[i];
    o[10] = (det_info < 1e-10) ? 0.0 : 1.0;   // VirtualMap.cpp:293
    o[11] = __hiloint2double(fr, fc);
    if (k < KC) { s_fr[k] = fr; s_fc[k] = fc; }
  }
  {  // landmark cells (OccupancyMap.cpp:126-131)
    const double *l = a.lm + (size_t)b * a.Lstride * 2;
    for (int j = tid; j < a.Lfixed; j += nthr) {
      if (a.lm_obs && !a.lm_obs[(size_t)b * a.Lstride + j]) continue;
      const int lr = (int)floor((l[2 * j + 1] - c.map_min_y) / c.res), lc = (int)floor((l[2 * j] - c.map_min_x) / c.res);
      if (lr >= 0 && lr < c.rows && lc >= 0 && lc < c.cols) atomicOr(&sst[lr * c.cols + lc], ST_LM);
    }
  }
  // a chunk's digest rows travel scratch -> registers (in front of the previous chunk's prediction) -> shared memory (behind its fold)
  double drow = 0.0;
  auto load_rows = [&](int ch) {
    const int kk = tid / DIG_W, k = ch * KC + kk;
    if (tid < KC * DIG_W && k < T) drow = pr[(size_t)k * PREP_W + (tid - kk * DIG_W)];
  };
  auto store_rows = [&](int ch) {
    const int kk = tid / DIG_W, i = tid - kk * DIG_W, k = ch * KC + kk, bf = ch & 1;
    if (tid < KC * DIG_W && k < T) {
      dig[((size_t)bf * KC + kk) * DIG_W + i] = drow;
      if (i == 11) { s_fr[bf * 16 + kk] = __double2hiint(drow); s_fc[bf * 16 + kk] = __double2loint(drow); }
    }
  };
  auto chunk_box = [&](int ch) {   // last warp: cell box of the chunk's blocks (read by the fold, behind the next barrier)
    const int bf = ch & 1, kc = min(KC, T - ch * KC);
    int fr = lane < kc ? s_fr[bf * 16 + lane] : 0x3fffffff, fc = lane < kc ? s_fc[bf * 16 + lane] : 0x3fffffff;
    int frm = lane < kc ? fr : -0x3fffffff, fcm = lane < kc ? fc : -0x3fffffff;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      fr = min(fr, __shfl_xor_sync(0xffffffffu, fr, o)); frm = max(frm, __shfl_xor_sync(0xffffffffu, frm, o));
      fc = min(fc, __shfl_xor_sync(0xffffffffu, fc, o)); fcm = max(fcm, __shfl_xor_sync(0xffffffffu, fcm, o));
    }
    if (lane == 0) s_box[bf] = make_int4(max(fr, 0), min(frm + W, c.rows), max(fc, 0), min(fcm + W, c.cols));
  };
  // (the clock reads hang on the barrier's result: BAR.SYNC defers blocking, a bare clock read would run ahead of it)
  const int nb1 = __syncthreads_count(1);
  if (a.clocks && tid == 0 && nb1) a.clocks[4 * b + 1] = clock64();

  unsigned short *mylist = s_list + (warp < NWARP ? warp : 0) * PPW;
  long long ph[5] = {0, 0, 0, 0, 0};
  long long t0 = 0;
  // ---- chunk ch, PREDICT into record buffer rb (the calling warps: [0, NWARP)) -----------------------------------------
  auto predict = [&](const int ch, const int rb) {
    const int k0 = ch * KC, kc = min(KC, T - k0), bf = ch & 1;
    const double *dg = dig + (size_t)bf * KC * DIG_W;
    const int *fr_ = s_fr + bf * 16, *fc_ = s_fc + bf * 16;
    double *rxx = rec0 + (size_t)rb * 3 * RS, *rxy = rxx + RS, *ryy = rxy + RS;
    // ---- PREDICT, gate pass: this warp's slice of the chunk's pairs, GR pairs per lane evaluated side by side
    // (independent straight-line instances: instruction-level parallelism for the latency of the fp64 predicates) -----
    const int np = kc * WW, p_lo = min(np, warp * PPW), p_hi = min(np, p_lo + PPW);
    int n_act = 0;
    for (int p0 = p_lo; p0 < p_hi; p0 += 32 * GR) {
      bool upd[GR], vis[GR], amb[GR];
      int cidx[GR];
#pragma unroll
      for (int r = 0; r < GR; ++r) {   // branch-free fast path: the predicates away from their knife edges
        const int pr_ = p0 + 32 * r + lane, p = min(pr_, p_hi - 1);
        const bool pv = pr_ < p_hi;
        const int kk = p / WW, j = p - kk * WW, dr = j / W, dcl = j - dr * W;
        const int row = fr_[kk] + dr, col = fc_[kk] + dcl;
        const bool cell_on = pv && row >= 0 && row < c.rows && col >= 0 && col < c.cols;
        const double *q = dg + kk * DIG_W;
        const double cx = c.map_min_x + c.res * (col + 0.5), cy = c.map_min_y + c.res * (row + 0.5);
        const double dx = cx - q[0], dy = cy - q[1];
        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        const bool in_max = d2 < c.max_r2_lo, out_min = d2 > c.min_r2_hi;
        const bool band = (d2 >= c.max_r2_lo && d2 <= c.max_r2_hi) || (d2 >= c.min_r2_lo && d2 <= c.min_r2_hi);
        const double co = q[2], si = q[3];
        const double qx = co * dx + si * dy, qy = -si * dx + co * dy;
        const bool fov_fast = c.fov_wide && (qx > 0.0 || fabs(qy) > -qx * c.wedge_tan);
        amb[r] = cell_on && (band || (in_max && !fov_fast));
        vis[r] = cell_on && in_max && fov_fast;
        upd[r] = vis[r] && out_min && q[10] != 0.0;            // full check (q10) ; det(info) gate
        cidx[r] = row * c.cols + col;
      }
#pragma unroll
      for (int r = 0; r < GR; ++r) {
        const int p = p0 + 32 * r + lane;
        if (amb[r]) {   // rare: inside a 1e-9 band around a range limit (the reference's own sqrt comparison decides) or in the rear wedge (atan2)
          const int kk = p / WW, j = p - kk * WW, dr = j / W, dcl = j - dr * W;
          const int row = fr_[kk] + dr, col = fc_[kk] + dcl;
          const double *q = dg + kk * DIG_W;
          const double cx = c.map_min_x + c.res * (col + 0.5), cy = c.map_min_y + c.res * (row + 0.5);
          const double dx = cx - q[0], dy = cy - q[1];
          const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
          bool in_max, out_min;
          if (d2 < c.max_r2_lo) in_max = true; else if (d2 > c.max_r2_hi) in_max = false; else in_max = __dsqrt_rn(d2) < c.max_range;
          if (d2 > c.min_r2_hi) out_min = true; else if (d2 < c.min_r2_lo) out_min = false; else out_min = __dsqrt_rn(d2) > c.min_range;
          const double co = q[2], si = q[3];
          const double qx = co * dx + si * dy, qy = -si * dx + co * dy;
          bool in_fov;
          if (c.fov_wide && (qx > 0.0 || fabs(qy) > -qx * c.wedge_tan)) in_fov = true;
          else if (in_max) { const double bb = atan2(qy, qx); in_fov = bb < c.max_bearing && bb > c.min_bearing; }
          else in_fov = false;
          vis[r] = in_max && in_fov;
          upd[r] = vis[r] && out_min && q[10] != 0.0;
        }
        if (vis[r]) atomicAdd(&sst[cidx[r]], 1u);              // visibility count (q8): order-independent
        if (p < p_hi && !upd[r]) rxx[p] = -1.0;                // sentinel: this pose does not update this cell
        const unsigned m = __ballot_sync(0xffffffffu, upd[r]);
        if (upd[r]) mylist[n_act + __popc(m & ((1u << lane) - 1u))] = (unsigned short)p;
        n_act += __popc(m);
      }
    }
    __syncwarp();
    if (a.phase) { const long long t1 = clock64(); ph[0] += t1 - t0; t0 = t1; }
    // ---- PREDICT, SPD pass on the compacted list, two entries per lane side by side ---------------------------------
    for (int i0 = 0; i0 < n_act; i0 += 64) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = i0 + 32 * r + lane;
        const bool on = i < n_act;
        const int p = mylist[min(i, n_act - 1)];
        const int kk = p / WW, j = p - kk * WW, dr = j / W, dcl = j - dr * W;
        const int row = fr_[kk] + dr, col = fc_[kk] + dcl;
        const double *q = dg + kk * DIG_W;
        const double cx = c.map_min_x + c.res * (col + 0.5), cy = c.map_min_y + c.res * (row + 0.5);
        const double dx = cx - q[0], dy = cy - q[1];
        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        const double co = q[2], si = q[3];
        const double qx = co * dx + si * dy, qy = -si * dx + co * dy;
        // body-frame covariance of the predicted virtual landmark: cb = P + (rr / r^2) [qx qx, qx qy; qx qy, qy qy], with P the
        // bearing-noise + pose-covariance part.  Its inverse (the body-frame information) needs ONE reciprocal:
        //   det(cb) r^2 = det(P) r^2 + rr Q,  Q = P00 qy^2 - 2 P01 qx qy + P11 qx^2,   adj(cb) r^2 = adj(P) r^2 + rr [qy qy, -qx qy; ., qx qx]
        const double Sxx = q[4], Sxy = q[5], Sxt = q[6], Syy = q[7], Syt = q[8], Stt = q[9];
        const double qxx = qx * qx, qyy = qy * qy, qxy = qx * qy;
        const double P00 = qyy * (c.rb + Stt) + Sxx - 2.0 * qy * Sxt;
        const double P01 = Sxy - qy * Syt + qx * Sxt - qxy * (c.rb + Stt);
        const double P11 = qxx * (c.rb + Stt) + Syy + 2.0 * qx * Syt;
        const double Q = P00 * qyy - 2.0 * P01 * qxy + P11 * qxx;
        const double inv = vm_rcp((P00 * P11 - P01 * P01) * d2 + c.rr * Q);
        const double lb00 = (P11 * d2 + c.rr * qyy) * inv, lb01 = -(P01 * d2 + c.rr * qxy) * inv, lb11 = (P00 * d2 + c.rr * qxx) * inv;
        const double cc = co * co, ss = si * si, cs = co * si;   // rotate to the map frame
        if (on) {
          rxx[p] = cc * lb00 - 2.0 * cs * lb01 + ss * lb11;     // (diagonal of an SPD matrix: > 0, never the sentinel)
          rxy[p] = cs * (lb00 - lb11) + (cc - ss) * lb01;
          ryy[p] = ss * lb00 + 2.0 * cs * lb01 + cc * lb11;
        }
      }
    }
    if (a.phase) { const long long t1 = clock64(); ph[1] += t1 - t0; t0 = t1; }
  };
  // ---- chunk ch, FOLD from record buffer rb: thread ftid of fthr folding threads ---------------------------------------
  auto fold = [&](const int ch, const int rb, const int ftid, const int fthr) {
    const int k0 = ch * KC, kc = min(KC, T - k0), bf = ch & 1;
    (void)k0;
    const int *fr_ = s_fr + bf * 16, *fc_ = s_fc + bf * 16;
    const double *rxx = rec0 + (size_t)rb * 3 * RS, *rxy = rxx + RS, *ryy = rxy + RS;
    // ---- FOLD: the cells of the chunk's box, one thread per cell, the chunk's poses in trajectory order ---------
    {
      const int4 bx = s_box[bf];
      const int r0 = bx.x, c0 = bx.z, bw = bx.w - bx.z, nbc = (bx.y > r0 && bw > 0) ? (bx.y - r0) * bw : 0;
      for (int ci = ftid; ci < nbc; ci += fthr) {
        const int rr_ = ci / bw, row = r0 + rr_, col = c0 + ci - rr_ * bw;
        const int idx = row * c.cols + col;
        unsigned m = 0;
#pragma unroll 4
        for (int kk = 0; kk < kc; ++kk) {
          const unsigned dr = (unsigned)(row - fr_[kk]), dcl = (unsigned)(col - fc_[kk]);
          if (dr < (unsigned)W && dcl < (unsigned)W && rxx[kk * WW + dr * W + dcl] >= 0.0) m |= 1u << kk;
        }
        if (!m) continue;
        double *st = sinf + 3 * (size_t)idx;
        double ixx = st[0], ixy = st[1], iyy = st[2];
        bool first = !(sst[idx] & ST_UPD);                 // the first hit overwrites the prior (VirtualMap.cpp:307-309)
        int kk = __ffs(m) - 1;
        m &= m - 1;
        int j = kk * WW + (row - fr_[kk]) * W + (col - fc_[kk]);
        double nxx = rxx[j], nxy = rxy[j], nyy = ryy[j];
        for (;;) {
          double fxx = 0, fxy = 0, fyy = 0;
          const bool more = m != 0;
          if (more) {                                      // next record in flight beside this fold's dependent chain
            kk = __ffs(m) - 1;
            m &= m - 1;
            j = kk * WW + (row - fr_[kk]) * W + (col - fc_[kk]);
            fxx = rxx[j]; fxy = rxy[j]; fyy = ryy[j];
          }
          // covariance intersection on information matrices (VirtualMap.cpp:364-377, q11): a = det m1, b = det m2, w = num / d
          const double bdet = nxx * nyy - nxy * nxy;
          const double aa = ixx * iyy - ixy * ixy;
          const double cm = iyy * nxx - 2.0 * ixy * nxy + ixx * nyy;
          const double d = aa + bdet - cm, num = bdet - 0.5 * cm;
          // the clip of q11 -- (w<0 & d<0) | (w>1 & d>0) -> 0 ; (w<0 & d>0) | (w>1 & d<0) -> 1 -- decided from the signs of num, d
          // and num - d beside the reciprocal (w<0 <=> num, d of opposite sign; w>1 <=> num - d has the sign of d), not behind it
          const bool dneg = d < 0, dpos = d > 0;
          const bool wneg = (num < 0 && dpos) || (num > 0 && dneg), wbig = (num > d && dpos) || (num < d && dneg);
          const bool to0 = first || (wneg && dneg) || (wbig && dpos), to1 = (wneg && dpos) || (wbig && dneg);
          double w = num * vm_rcp(d);
          w = to0 ? 0.0 : (to1 ? 1.0 : w);
          ixx = fma(w, ixx - nxx, nxx);
          ixy = fma(w, ixy - nxy, nxy);
          iyy = fma(w, iyy - nyy, nyy);
          first = false;
          if (!more) break;
          nxx = fxx; nxy = fxy; nyy = fyy;
        }
        st[0] = ixx; st[1] = ixy; st[2] = iyy;
        atomicOr(&sst[idx], ST_UPD);                        // (atomic: the next chunk's gate pass may be counting visibility on this word)
      }
    }
  };
  if (!SPEC) {
    for (int ch = 0; ch < nch; ++ch) {
      t0 = a.phase ? clock64() : 0;
      if (ch + 1 < nch) load_rows(ch + 1);
      if (warp == NWARP - 1) chunk_box(ch);
      predict(ch, 0);
      {
        const int nbw = __syncthreads_count(1);
        if (a.phase && nbw) { const long long t1 = clock64(); ph[2] += t1 - t0; t0 = t1; }
      }
      fold(ch, 0, tid, nthr);
      if (a.phase) { const long long t1 = clock64(); ph[3] += t1 - t0; t0 = t1; }
      if (ch + 1 < nch) store_rows(ch + 1);
      {
        const int nbw = __syncthreads_count(1);
        if (a.phase && nbw) { const long long t1 = clock64(); ph[4] += t1 - t0; }
      }
    }
  } else {
    // named barriers: FULL[b] = 1 + b (records of buffer b written: predict arrives, fold waits), FREE[b] = 3 + b (fold done with buffer b and
    // with the chunk's block origins / box: fold arrives, predict waits), 5 = the predict group, 6 = the fold group
    // (immediate barrier ids: with ids in registers the compiler reserves all 16 named barriers of the CTA)
    auto bar_sync = [](int id, int n) {
      switch (id) {
        case 1: asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); break;
        case 2: asm volatile("bar.sync 2, %0;" ::"r"(n) : "memory"); break;
        case 3: asm volatile("bar.sync 3, %0;" ::"r"(n) : "memory"); break;
        case 4: asm volatile("bar.sync 4, %0;" ::"r"(n) : "memory"); break;
        case 5: asm volatile("bar.sync 5, %0;" ::"r"(n) : "memory"); break;
        default: asm volatile("bar.sync 6, %0;" ::"r"(n) : "memory"); break;
      }
    };
    auto bar_arrive = [](int id, int n) {
      switch (id) {
        case 1: asm volatile("bar.arrive 1, %0;" ::"r"(n) : "memory"); break;
        case 2: asm volatile("bar.arrive 2, %0;" ::"r"(n) : "memory"); break;
        case 3: asm volatile("bar.arrive 3, %0;" ::"r"(n) : "memory"); break;
        default: asm volatile("bar.arrive 4, %0;" ::"r"(n) : "memory"); break;
      }
    };
    constexpr int GRP = 256;
    if (warp < NWARP) {
      for (int ch = 0; ch < nch; ++ch) {
        t0 = a.phase ? clock64() : 0;
        if (ch >= 2) bar_sync(3 + (ch & 1), 2 * GRP);                   // fold(ch - 2) has left buffer ch & 1
        if (ch >= 1) {                                                  // rows of this chunk: registers -> shared memory, then the chunk's box
          store_rows(ch);
          bar_sync(5, GRP);
        }
        if (warp == NWARP - 1) chunk_box(ch);
        if (a.phase) { const long long t1 = clock64(); ph[2] += t1 - t0; t0 = t1; }
        if (ch + 1 < nch) load_rows(ch + 1);
        predict(ch, ch & 1);
        bar_arrive(1 + (ch & 1), 2 * GRP);                              // (release: the records, origins and box of chunk ch)
      }
    } else {
      const int ftid = tid - GRP;
      for (int ch = 0; ch < nch; ++ch) {
        t0 = a.phase ? clock64() : 0;
        bar_sync(1 + (ch & 1), 2 * GRP);
        if (a.phase) { const long long t1 = clock64(); ph[4] += t1 - t0; t0 = t1; }
        fold(ch, ch & 1, ftid, GRP);
        bar_sync(6, GRP);                                               // a cell may change hands between chunks: every fold of chunk ch first
        if (a.phase) { const long long t1 = clock64(); ph[3] += t1 - t0; }
        if (ch + 2 < nch) bar_arrive(3 + (ch & 1), 2 * GRP);
      }
    }
  }
  if (a.phase && tid == 0) for (int i = 0; i < (SPEC ? 3 : 5); ++i) a.phase[8 * b + i] = ph[i];
  if (SPEC && a.phase && tid == 256) { a.phase[8 * b + 3] = ph[3]; a.phase[8 * b + 4] = ph[4]; }
  // ---- write the map: every output byte once ---------------------------------------------------------------------
  const size_t cell0 = (size_t)b * V;
  const bool bulk = (V & 1) == 0;                        // 24 V bytes and the env's offset are multiples of 16
  // the information matrices leave as bulk copies shared -> global of the cell-interleaved state: every writer orders its
  // generic-proxy stores before the async proxy, then one thread issues the copies
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const int nb2 = __syncthreads_count(1);
  if (a.clocks && tid == 0 && nb2) a.clocks[4 * b + 2] = clock64();
  if (bulk && tid == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(sinf);
    char *dst = reinterpret_cast<char *>(a.vinfo + cell0 * 3);
    const uint32_t total = 24u * (uint32_t)V;
    for (uint32_t off = 0; off < total; off += 32768u) {
      const uint32_t n = min(32768u, total - off);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(src + off), "r"(n) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  for (int i = tid; i < V; i += nthr) {
    const unsigned st = sst[i];
    const int cnt = (int)(st & ST_CNT);
    a.prob[cell0 + i] = (st & ST_LM) ? c.ptab[5] : c.ptab[min(cnt, 4)];
    if (a.seen) a.seen[cell0 + i] = (st & ST_LM) ? -1 : cnt;
  }
  if (!bulk) for (int i = tid; i < 3 * V; i += nthr) a.vinfo[cell0 * 3 + i] = sinf[i];
  if (a.clocks && tid == 0) a.clocks[4 * b + 3] = clock64();
  if (a.metrics) {
    // ---- metrics (k_vmap_metrics, same summation order: 256 strided partial sums, fixed tree) ---------------------
    if (tid == 0 && a.counters && a.step_kind[b] == 1) {
      atomicAdd(&a.counters[0], 1ull);
      atomicAdd(&a.counters[1], (unsigned long long)T);
      atomicAdd(&a.counters[2], (unsigned long long)a.meas_ptr[(size_t)b * (a.d.Tmax + 1) + T]);
    }
    int *s_e = reinterpret_cast<int *>(s_red + 256), *s_k = s_e + 256;
    if (tid < 256) {
      const int extg = 20;
      int n_exp = 0, n_known = 0;
      double tr = 0.0;
      for (int i = tid; i < V; i += 256) {
        const double x = (i % c.cols + 0.5) * a.cfg.resolution + a.cfg.map_min_x, y = (i / c.cols + 0.5) * a.cfg.resolution + a.cfg.map_min_y;
        const unsigned st = sst[i];
        const double pv = (st & ST_LM) ? c.ptab[5] : c.ptab[min((int)(st & ST_CNT), 4)];
        if ((pv < 0.49 || pv > 0.6) && a.cfg.map_min_x + extg <= x && x <= a.cfg.map_max_x - extg && a.cfg.map_min_y + extg <= y && y <= a.cfg.map_max_y - extg) ++n_exp;
        if (pv < a.cfg.occupancy_threshold) ++n_known;
        const double qa = sinf[3 * i], qb = sinf[3 * i + 1], qc = sinf[3 * i + 2];
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
      // exploration_env.py:416-419: a world in which the four forced steps of reset() saw no landmark is regenerated -- here the episode
      // ends at once (uncounted) and the next reset draws the next world of the env's seed sequence
      bool blind = false;
      if (a.step_kind[b] == 2 && a.lm_obs) { int n = 0; for (int j = 0; j < a.Lfixed; ++j) n += a.lm_obs[(size_t)b * a.Lstride + j]; blind = n == 0; }
      a.done[b] = (a.sim_step[b] > a.cfg.max_steps || explored > 0.85 || a.status[b] == DGE_ECAP || blind) ? 1 : 0;
    }
  }
  // the shared-memory source of the bulk copies must stay alive until they have been read
  if (bulk && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
int env_threads() {   // DGE_VMAP_THREADS=256: the single-group kernel (A/B); default 512 = warp-specialised predict || fold
  static const int n = [] { const char *v = getenv("DGE_VMAP_THREADS"); return (v && atoi(v) == 256) ? 256 : 512; }();
  return n;
}
size_t env_smem(size_t V, int kc, int W, int nbuf) {
  const size_t rs = (size_t)kc * W * W, nw = 8, ppw = (rs + nw - 1) / nw;
  return V * 3 * sizeof(double) + ((V + 1) & ~(size_t)1) * sizeof(unsigned) + 2 * (size_t)kc * DIG_W * sizeof(double) + nbuf * 3 * rs * sizeof(double) +
         256 * (sizeof(double) + 2 * sizeof(int)) + 64 * sizeof(int) + 16 + 2 * sizeof(int4) + nw * ((ppw + 7) & ~(size_t)7) * sizeof(unsigned short) + 32;
}
bool env_plan(const dge_config &g, int rows, int cols, int nbuf, int *hw, int *W, int *kc, int *ppw, size_t *smem) {
  *hw = (int)ceil(g.max_range / g.resolution);
  *W = 2 * *hw + 1;
  if (*W > 15) return false;
  const size_t V = (size_t)rows * cols, nw = 8;
  // chunk length: up to 16 poses (the fold's hit mask; its digest rows are moved by 16 x 12 threads), shorter only if the record buffer(s)
  // would not leave room for two CTAs per SM
  for (int k = 16; k >= 1; k = k > 8 ? k - 2 : k >> 1) {
    *kc = k; *smem = env_smem(V, k, *W, nbuf);
    *ppw = (int)((((size_t)k * *W * *W + nw - 1) / nw + 7) & ~(size_t)7);
    if (*smem <= 112 * 1024 || (k == 1 && *smem <= 226 * 1024)) return true;
  }
  return false;
}
int env_launch(EnvArgs &a, const dge_config &g, int n, int rows, int cols, cudaStream_t st) {
  size_t smem;
  int nthr = env_threads();
  if (nthr == 512 && !env_plan(g, rows, cols, 2, &a.hw, &a.W, &a.kc, &a.ppw, &smem)) nthr = 256;   // (two record buffers do not fit: single-group kernel)
  if (nthr == 512 && a.kc < 8) nthr = 256;                                                           // (too short chunks for the pipelined schedule)
  if (nthr == 256 && !env_plan(g, rows, cols, 1, &a.hw, &a.W, &a.kc, &a.ppw, &smem)) return 1;       // caller falls back to the cell-centric kernels
  const int vi = (a.W == 7 ? 1 : 0) + (nthr == 512 ? 2 : 0);
  void (*kern)(EnvArgs) = vi == 3 ? k_vmap_env<7, 512, true> : vi == 2 ? k_vmap_env<0, 512, true> : vi == 1 ? k_vmap_env<7, 256, false> : k_vmap_env<0, 256, false>;
  static size_t configured[4] = {0, 0, 0, 0};
  size_t &cf = configured[vi];
  if (smem > 48 * 1024 && smem > cf) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DGE_ECUDA;
    cf = smem;
  }
  kern<<<n, nthr, smem, st>>>(a);
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
    a.clocks = nullptr; a.phase = nullptr;
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
                                         e->n_poses, e->meas_ptr, e->count_steps ? e->counters : nullptr, e->step_kind, e->observed);
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
    // dev profiling: phase clocks in the chunk-box scratch, which the fused kernel does not use ([n,4] clocks, then [n,8] per-chunk phase sums)
    a.clocks = reinterpret_cast<long long *>(cbox_ws);
    a.phase = reinterpret_cast<long long *>(cbox_ws) + (size_t)n * 4;
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
