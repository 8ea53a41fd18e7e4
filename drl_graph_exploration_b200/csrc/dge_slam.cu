// SLAM solve + marginal recovery: one environment per CTA.
//
// Replaces, for a batch of environments, SLAM2D::optimize (SLAM2D.cpp:374-430): gtsam
// ISAM2::update / calculateEstimate plus FastMarginals::marginalCovariance for every pose
// and landmark (FastMarginals.cpp:121-186), i.e. rows a4+a5 of SURVEY section 8.
//
// Design (B200-first, not a translation of the Bayes-tree code):
//   * ISAM2 is emulated by its schedule: fixed linearisation point theta, relinearise every
//     `relin_skip`-th update the variables with |delta|_inf >= relin_thresh, exact solve.
//   * the information matrix is block-tridiagonal in the pose chain (3x3 blocks) with an
//     arrow to the landmarks (2x2 blocks).  Poses are eliminated along the chain; each CTA
//     thread owns ONE landmark column of the border, so the forward elimination and the
//     backward substitution are sync-free per-thread recurrences (3 FMAs-chains in
//     registers) and the O(T n^2) parts (Schur complement, W Sigma W^T) are GEMM-shaped
//     phases parallel over poses.
//   * all arithmetic fp64 (prior information 1/sigma^2 ~ 3e7 next to O(1) blocks).
//   * workspace (D,g,U | Dinv,FU,f per pose; border rows Bt / FB) lives in HBM but is
//     written and re-read by the same CTA within microseconds => L2-resident.
//   * INCREMENTAL between relinearisations (what ISAM2 is for): as long as no linearisation point moves, the
//     forward elimination of the closed poses 0..T-2 cannot change when pose T arrives -- the new factors touch the
//     last two poses only, and a new landmark adds a border column that is zero on every older pose.  The state of the
//     elimination after the closed poses (chain carries, border carries, partial Schur complement, partial landmark
//     rhs) is cached per env; a step then closes ONE pose, eliminates the newest one provisionally, inverts the Schur
//     complement and runs the backward pass + marginals.  A step that relinearises (every relin_skip-th update, if a
//     delta exceeds the threshold) re-eliminates from a CHECKPOINT: the variables that move are, as a rule, the poses added since
//     the last relinearisation and the landmarks first seen from them (their linearisation points are dead-reckoned guesses),
//     i.e. the last ~10 poses, so every rebuild also saves the elimination state three poses before its end; the next one resumes
//     there if no moved pose (minus one: its odometry factor reaches back) and no moved landmark's first observation lies before
//     it, else from pose 0.  Border columns are indexed by the landmark's SLOT
//     (order of first observation since the last rebuild), so that columns never move while the cache lives.
#include <cstdlib>

#include "dge_internal.cuh"

namespace {

constexpr int NT = 256;        // threads per CTA (>= the number of border columns, 2*Lt <= 128)
constexpr int NH = NT / 64;    // row groups per column in the Gauss-Jordan sweep
constexpr int CH = 32;         // poses staged per shared-memory chunk
constexpr int SW = 39;         // doubles per staged pose: D(6) g(3) U(9) gnext(3) | Dinv(6) FU(9) f(3)
constexpr int GK = 8;          // poses per staged chunk of border rows in the Schur-complement GEMM
constexpr int WS_POSE = DGE_WS_POSE;    // doubles per pose in ws_pose: A-data D(6) g(3) U(9) gnext(3) @0 | B-data Dinv(6) FU(9) f(3) @21
constexpr int WIDE_N2C = 96;   // more border columns than this: the backward pass works on half-size chunks (shared-memory budget)
constexpr int WS_MEAS = 14;    // doubles per measurement in ws_meas: C(3) gl(2) | D contribution(6) g contribution(3)

struct SlamArgs {
  dge_config cfg;
  DgeDims d;
  int32_t *n_poses, *update_count, *status;
  const double *prior_pose, *odom, *lm_true;
  double *lin_pose, *est_pose, *delta_pose, *pose_cov, *pose_info;
  const int32_t *meas_ptr, *meas_id, *meas_pose;
  const double *meas_b, *meas_r;
  const uint8_t *observed;
  double *lin_l, *est_l, *delta_l, *land_cov;
  double *ws_pose, *ws_meas, *ws_Bt, *ws_FB;
  int32_t *ws_midx;
  int32_t *lm_slot, *fc_valid;   // [B,Lt] landmark id -> border slot ; [B] number of poses the cached elimination state was saved at
  const int32_t *lm_first;       // [B,Lt] pose at which a landmark was first observed (k_move_measure)
  int32_t *ck_pos;               // [B, DGE_CK_STRIDE] stack of checkpoints: closed poses behind each, physical slot of each, count
  double *ck_state;              // [B, DGE_CK_DEPTH, DGE_FC_WIDTH(Lt)] checkpoints of the elimination state, each laid out like fc_state
  double *fc_state;              // [B, DGE_FC_WIDTH(Lt)] cached state: cD(6) cg(3) pad | cB [N2C][3] | gl [N2C] | S_partial [N2C][N2C]
  int incremental;               // 0: every step eliminates from pose 0 (A/B switch, DGE_SLAM_INCREMENTAL=0)
  const int32_t *order;          // nullable [B]: block -> env (cost-ordered placement, k_step_order)
  double *metrics;
  long long *clocks;   // [B,12] optional: SM clock at the phase boundaries (thread 0), for in-situ phase timing
};

__device__ __forceinline__ void predict_br(const Pose3 &p, double lx, double ly, double &bearing, double &range, double *Hx, double *Hl) {
  double s, c;
  sincos(p.th, &s, &c);
  const double dx = lx - p.x, dy = ly - p.y;
  const double qx = c * dx + s * dy, qy = -s * dx + c * dy;
  range = sqrt(dx * dx + dy * dy);
  bearing = atan2(qy, qx);
  const double r2 = qx * qx + qy * qy;
  const double bx = -qy / r2, by = qx / r2, rx = qx / range, ry = qy / range;
  Hx[0] = -bx; Hx[1] = -by; Hx[2] = bx * qy - by * qx;
  Hx[3] = -rx; Hx[4] = -ry; Hx[5] = rx * qy - ry * qx;
  Hl[0] = bx * c - by * s; Hl[1] = bx * s + by * c;
  Hl[2] = rx * c - ry * s; Hl[3] = rx * s + ry * c;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {   // release: the producer's earlier shared stores are visible to waiters
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// reciprocal of a pivot: hardware seed + two Newton steps (~2 ulp), half the dependent latency of the IEEE division
__device__ __forceinline__ double gj_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(NT, 2) k_slam(SlamArgs a, const uint8_t *mask) {
  const int b = a.order ? a.order[blockIdx.x] : blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (mask && !mask[b]) return;
  const int T = a.n_poses[b];
  const int Lt = a.d.Lt, Tmax = a.d.Tmax, N2C = 2 * Lt;  // N2C = border stride in the workspace
  const int ldS = N2C;                                   // row pitch of S / Sigma_ll in shared memory (= the cache's: copied straight)
  extern __shared__ __align__(16) double smem[];
  // shared layout
  double *S = smem;                               // [N2C*N2C]  Schur complement -> Sigma_ll
  double *stage = S + (size_t)N2C * N2C;          // [CH*SW]   per-pose 3x3 blocks of the current chunk
  double *colp = stage + CH * SW;                 // [N2C]
  double *gl = colp + N2C;                        // [N2C]
  double *dl = gl + N2C;                          // [N2C]
  double *red = dl + N2C;                         // [NT/32 * 2]
  double *gjb = red + 2 * (NT / 32);              // [258] pivot row / column / reciprocal exchange of the Gauss-Jordan sweep
  double *gbuf = gjb + 258;                       // [2*GK*3*N2C] border-row staging (Schur GEMM) / per-warp W_k (phase E)
  double *lastB = gbuf + 2 * GK * 3 * N2C;        // [2][3*N2C] border rows Bt of the pose closed last (k = T-2) and of the newest (open) pose
  double *lastFB = lastB + 6 * N2C;               // [2][3*N2C] ... and their FB rows
  double *openB = lastB + 3 * N2C, *openFB = lastFB + 3 * N2C;
  // FB rows of the current chunk of the backward pass [ce*3*N2C]: own region, or (wide borders: half-size chunks) the upper half of gbuf
  const bool wide = N2C > WIDE_N2C;
  double *FBs = wide ? gbuf + GK * 3 * N2C : lastFB + 6 * N2C;
  int *lidx = (int *)(lastFB + 6 * N2C + (wide ? 0 : 2 * GK * 3 * N2C));    // [Lt]  id -> border slot (-1 unobserved)
  int *lid = lidx + Lt;                           // [Lt]  slot -> id
  __shared__ int s_nl, s_nl_old, s_bad, red_i[NT / 32];
  __shared__ unsigned char s_obs[64];
  __shared__ uint64_t s_bar[CH];   // one mbarrier per pose of the staged chunk: B0 (producer) -> B1 (consumers)

  const double wo[3] = {1.0 / (a.cfg.trans_noise * a.cfg.trans_noise), 1.0 / (a.cfg.trans_noise * a.cfg.trans_noise),
                        1.0 / (a.cfg.rot_noise * a.cfg.rot_noise)};
  const double wm[2] = {1.0 / (a.cfg.bearing_noise * a.cfg.bearing_noise), 1.0 / (a.cfg.range_noise * a.cfg.range_noise)};

  double *lin = a.lin_pose + (size_t)b * Tmax * 3, *est = a.est_pose + (size_t)b * Tmax * 3, *del = a.delta_pose + (size_t)b * Tmax * 3;
  const double *od = a.odom + (size_t)b * Tmax * 3;
  double *linl = a.lin_l + (size_t)b * Lt * 2, *estl = a.est_l + (size_t)b * Lt * 2, *dell = a.delta_l + (size_t)b * Lt * 2;
  const uint8_t *obs = a.observed + (size_t)b * Lt;
  const int32_t *mptr = a.meas_ptr + (size_t)b * (Tmax + 1);
  const int32_t *mid = a.meas_id + (size_t)b * a.d.Mmax;
  const double *mb = a.meas_b + (size_t)b * a.d.Mmax, *mr = a.meas_r + (size_t)b * a.d.Mmax;
  double *wsp = a.ws_pose + (size_t)b * Tmax * WS_POSE;
  double *wsm = a.ws_meas + (size_t)b * a.d.Mmax * WS_MEAS;
  double *wBt = a.ws_Bt + (size_t)b * Tmax * 3 * N2C;
  double *wFB = a.ws_FB + (size_t)b * Tmax * 3 * N2C;
  int32_t *wmi = a.ws_midx + (size_t)b * Tmax * Lt;
  int32_t *slot_g = a.lm_slot + (size_t)b * Lt;
  double *fc = a.fc_state + (size_t)b * DGE_FC_WIDTH(Lt);     // cD(6) cg(3) | cB | gl | S_partial
  double *fc_cB = fc + 16, *fc_gl = fc_cB + 3 * N2C, *fc_S = fc_gl + N2C;
  int32_t *ckp = a.ck_pos + (size_t)b * DGE_CK_STRIDE;        // positions [DEPTH] | physical slots [DEPTH] | count

  // ---------------------------------------------------------------- step 0 ---
  // ISAM2 relinearisation schedule (gtsam ISAM2::update: ++update_count; every
  // relinearizeSkip-th call, variables with max|delta| >= relinearizeThreshold move their
  // linearisation point: theta <- theta (+) delta, delta <- 0).
  const int uc = a.update_count[b] + 1;
  if (tid == 0) {
    s_bad = 0;
    for (int i = 0; i < CH; ++i) mbar_init(&s_bar[i], 1);
  }
  for (int j = tid; j < Lt; j += NT) { s_obs[j] = obs[j]; lidx[j] = slot_g[j]; }   // (thread 0 walks them below: no serial chain of global loads)
  int moved = 0, rmin = 0x7fffffff;   // rmin: first pose whose blocks a moved variable touches
  if (a.cfg.relin_skip > 0 && uc % a.cfg.relin_skip == 0) {
    for (int k = tid; k < T; k += NT) {
      const double d0 = del[3 * k], d1 = del[3 * k + 1], d2 = del[3 * k + 2];
      if (fmax(fabs(d0), fmax(fabs(d1), fabs(d2))) >= a.cfg.relin_thresh) {
        moved = 1; rmin = min(rmin, k - 1);   // (the odometry factor k-1 -> k enters the blocks of pose k-1)
        const Pose3 p = dge_compose(Pose3{lin[3 * k], lin[3 * k + 1], lin[3 * k + 2]}, Pose3{d0, d1, d2});
        lin[3 * k] = p.x; lin[3 * k + 1] = p.y; lin[3 * k + 2] = p.th;
        del[3 * k] = 0; del[3 * k + 1] = 0; del[3 * k + 2] = 0;
      }
    }
    for (int j = tid; j < Lt; j += NT) {
      if (!obs[j]) continue;
      if (fmax(fabs(dell[2 * j]), fabs(dell[2 * j + 1])) >= a.cfg.relin_thresh) {
        moved = 1; rmin = min(rmin, a.lm_first[(size_t)b * Lt + j]);
        linl[2 * j] += dell[2 * j]; linl[2 * j + 1] += dell[2 * j + 1];
        dell[2 * j] = 0; dell[2 * j + 1] = 0;
      }
    }
  }
  // the cached elimination state is usable iff nothing was relinearised and it was saved one pose ago
  const int any_moved = __syncthreads_or(moved);
  const bool valid = !any_moved && a.incremental && T >= 2 && a.fc_valid[b] == T - 1;
  // a rebuild resumes from the NEWEST checkpoint nothing in front of which has moved (block-wide minimum of rmin), else from pose 0; the
  // checkpoints behind that point describe a prefix that no longer exists and are dropped; the rebuild leaves a new one three poses before its end
  int c_use = 0, ck_cnt = 0, ck_src_slot = 0;
  if (!valid && a.incremental) {
    ck_cnt = ckp[2 * DGE_CK_DEPTH];
    int limit = T - 2;
    if (any_moved) {
      rmin = __reduce_min_sync(0xffffffffu, rmin);
      if (lane == 0) red_i[warp] = rmin;
      __syncthreads();
      int r = red_i[0];
      for (int w = 1; w < NT / 32; ++w) r = min(r, red_i[w]);
      limit = min(limit, r);
      __syncthreads();
    }
    while (ck_cnt > 0 && ckp[ck_cnt - 1] > limit) --ck_cnt;
    if (ck_cnt > 0 && ckp[ck_cnt - 1] > 0) { c_use = ckp[ck_cnt - 1]; ck_src_slot = ckp[DGE_CK_DEPTH + ck_cnt - 1]; }
    else ck_cnt = 0;
  }
  const int c_new = (!valid && a.incremental && T - 3 > c_use) ? T - 3 : -1;   // position of the checkpoint this step leaves (-1: none)
  // its storage: the first free physical slot, or (stack full) the oldest checkpoint's, which is dropped
  const int ck_dst_slot = ckp[DGE_CK_DEPTH + (ck_cnt < DGE_CK_DEPTH ? ck_cnt : 0)];
  const double *ck = a.ck_state + ((size_t)b * DGE_CK_DEPTH + ck_src_slot) * DGE_FC_WIDTH(Lt);      // resumed from (c_use > 0)
  const double *ck_S = ck + 16 + 4 * N2C;
  double *ckn = a.ck_state + ((size_t)b * DGE_CK_DEPTH + ck_dst_slot) * DGE_FC_WIDTH(Lt);           // saved to (c_new >= 0)
  double *ckn_cB = ckn + 16, *ckn_gl = ckn_cB + 3 * N2C, *ckn_S = ckn_gl + N2C;
  const int k_lo = valid ? T - 2 : c_use;      // poses [k_lo, T-1) are closed in this step; pose T-1 stays open
  const int kz = valid ? T - 1 : c_use;        // poses whose factors are linearised in this step
  const bool append = valid || c_use > 0;      // border slots are kept and new landmarks appended (a rebuild from pose 0 renumbers them)
  if (tid == 0) {  // border slots: a rebuild numbers the observed landmarks in id order, a light step appends the new ones
    int n = 0;
    if (!append) {
      for (int j = 0; j < Lt; ++j) { if (s_obs[j]) { lidx[j] = n; lid[n] = j; slot_g[j] = n; ++n; } else { lidx[j] = -1; slot_g[j] = -1; } }
    } else {
      for (int j = 0; j < Lt; ++j) { const int sl = s_obs[j] ? lidx[j] : -1; lidx[j] = sl; if (sl >= 0) { lid[sl] = j; n = max(n, sl + 1); } }
      s_nl_old = n;
      for (int j = 0; j < Lt; ++j) if (s_obs[j] && lidx[j] < 0) { lidx[j] = n; lid[n] = j; slot_g[j] = n; ++n; }
    }
    if (!append) s_nl_old = n;
    s_nl = n;
  }
  // zero the sparse border inputs of this step (and, on a rebuild, the cached state)
  for (size_t i = (size_t)kz * 3 * N2C + tid; i < (size_t)T * 3 * N2C; i += NT) wBt[i] = 0.0;
  for (size_t i = (size_t)kz * Lt + tid; i < (size_t)T * Lt; i += NT) wmi[i] = 0;
  // S starts from the cached partial Schur complement (copied asynchronously: it is first touched after phase B) or from zero
  if (append) { const double *src = valid ? fc_S : ck_S; for (int i = tid; i < N2C * N2C / 2; i += NT) cp_async16(S + 2 * i, src + 2 * i); cp_async_commit(); }
  else {
    for (int i = tid; i < N2C * N2C; i += NT) S[i] = 0.0;
    for (int i = tid; i < 16 + 4 * N2C; i += NT) fc[i] = 0.0;
  }
  __syncthreads();
  const int nl = s_nl, n2 = 2 * nl;
  {   // a landmark first seen in a light step opens a border column that is zero on every closed pose (the workspace may hold an older episode's rows)
    const int c_old = 2 * s_nl_old, nc = n2 - c_old;
    for (int i = tid; i < nc * 3 * k_lo; i += NT) wFB[(size_t)(i / nc) * N2C + c_old + (i % nc)] = 0.0;
  }

  if (a.clocks && tid == 0) a.clocks[12 * b +0] = clock64();
  // ---------------------------------------------------------------- phase A ---
  // whitened linearisation of every factor at theta.
  // A1: one thread per bearing-range factor (BearingRangeFactor<Pose2,Point2>): pose-pose, pose-landmark
  //     and landmark-landmark blocks + right-hand sides -> ws_meas / border rows.
  const int M = mptr[T];
  const int32_t *mpose = a.meas_pose + (size_t)b * a.d.Mmax;
  for (int p = mptr[kz] + tid; p < M; p += NT) {
    const int k = mpose[p];
    const Pose3 pk{lin[3 * k], lin[3 * k + 1], lin[3 * k + 2]};
    const int id = mid[p], jr = lidx[id], c0 = 2 * jr;
    double bb, rg, Hx[6], Hl[4];
    predict_br(pk, linl[2 * id], linl[2 * id + 1], bb, rg, Hx, Hl);
    double r0 = bb - mb[p];                      // Rot2 local coordinates: wrap to (-pi, pi]
    if (r0 > DGE_PI) r0 -= 2 * DGE_PI; else if (r0 <= -DGE_PI) r0 += 2 * DGE_PI;
    const double r1 = rg - mr[p];
    double *m = wsm + (size_t)p * WS_MEAS;
    m[0] = Hl[0] * wm[0] * Hl[0] + Hl[2] * wm[1] * Hl[2];
    m[1] = Hl[0] * wm[0] * Hl[1] + Hl[2] * wm[1] * Hl[3];
    m[2] = Hl[1] * wm[0] * Hl[1] + Hl[3] * wm[1] * Hl[3];
    m[3] = -(Hl[0] * wm[0] * r0 + Hl[2] * wm[1] * r1);
    m[4] = -(Hl[1] * wm[0] * r0 + Hl[3] * wm[1] * r1);
    int q = 5;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int jj = i; jj < 3; ++jj, ++q) m[q] = Hx[i] * wm[0] * Hx[jj] + Hx[3 + i] * wm[1] * Hx[3 + jj];
      m[11 + i] = -(Hx[i] * wm[0] * r0 + Hx[3 + i] * wm[1] * r1);
      wBt[((size_t)k * 3 + i) * N2C + c0] = Hx[i] * wm[0] * Hl[0] + Hx[3 + i] * wm[1] * Hl[2];
      wBt[((size_t)k * 3 + i) * N2C + c0 + 1] = Hx[i] * wm[0] * Hl[1] + Hx[3 + i] * wm[1] * Hl[3];
    }
    wmi[(size_t)k * Lt + jr] = p + 1;
  }
  // A2: one thread per pose: prior (k = 0) and the odometry factor k -> k+1 (its contribution to
  //     pose k+1 is parked in gnext and picked up by the chain); measurement terms summed in factor order.
  __syncthreads();
  for (int k = k_lo + tid; k < T; k += NT) {
    const Pose3 pk{lin[3 * k], lin[3 * k + 1], lin[3 * k + 2]};
    double D[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0}, U[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gn[3] = {0, 0, 0};
    double *w = wsp + (size_t)k * WS_POSE;
    const bool fresh = k >= kz;   // the pose's own factors (prior / incoming odometry / measurements) are linearised now; else they are in the workspace
    if (!fresh) {
#pragma unroll
      for (int i = 0; i < 6; ++i) D[i] = w[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) g[i] = w[6 + i];
    } else if (k == 0) {  // PriorFactor<Pose2>: error = -Local(x, prior), H = I
      const Pose3 e = dge_between(pk, Pose3{a.prior_pose[3 * b], a.prior_pose[3 * b + 1], a.prior_pose[3 * b + 2]}, nullptr);
      const double wp[3] = {1.0 / (a.cfg.sigma_x0 * a.cfg.sigma_x0), 1.0 / (a.cfg.sigma_y0 * a.cfg.sigma_y0),
                            1.0 / (a.cfg.sigma_theta0 * a.cfg.sigma_theta0)};
      D[0] += wp[0]; D[3] += wp[1]; D[5] += wp[2];
      g[0] += wp[0] * e.x; g[1] += wp[1] * e.y; g[2] += wp[2] * e.th;   // g -= w * (-e)
    } else {       // odometry factor (k-1 -> k): Jacobian wrt x_k is I (its rhs arrives through gnext of pose k-1)
      D[0] += wo[0]; D[3] += wo[1]; D[5] += wo[2];
    }
    if (k + 1 < T) {  // BetweenFactor<Pose2>(x_k, x_k+1, odom): error = Local(odom, between), J = [H1 | I]
      double H1[9];
      const Pose3 pn{lin[3 * (k + 1)], lin[3 * (k + 1) + 1], lin[3 * (k + 1) + 2]};
      const Pose3 h = dge_between(pk, pn, H1);
      const Pose3 e = dge_between(Pose3{od[3 * k], od[3 * k + 1], od[3 * k + 2]}, h, nullptr);
      const double r[3] = {e.x, e.y, e.th};
      int q = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int jj = i; jj < 3; ++jj, ++q) D[q] += H1[i] * wo[0] * H1[jj] + H1[3 + i] * wo[1] * H1[3 + jj] + H1[6 + i] * wo[2] * H1[6 + jj];
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) U[i * 3 + jj] = H1[jj * 3 + i] * wo[jj];
        g[i] -= H1[i] * wo[0] * r[0] + H1[3 + i] * wo[1] * r[1] + H1[6 + i] * wo[2] * r[2];
        gn[i] = -wo[i] * r[i];
      }
    }
    if (fresh)
      for (int p = mptr[k]; p < mptr[k + 1]; ++p) {
        const double *m = wsm + (size_t)p * WS_MEAS;
#pragma unroll
        for (int i = 0; i < 6; ++i) D[i] += m[5 + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) g[i] += m[11 + i];
      }
#pragma unroll
    for (int i = 0; i < 6; ++i) w[i] = D[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { w[6 + i] = g[i]; w[18 + i] = gn[i]; }
#pragma unroll
    for (int i = 0; i < 9; ++i) w[9 + i] = U[i];
  }
  __syncthreads();

  if (a.clocks && tid == 0) a.clocks[12 * b +1] = clock64();
  // ---------------------------------------------------------------- phase B ---
  // forward elimination along the chain, per 32-pose chunk:
  //   B0  warp 0 runs the 3x3 pose recurrence  D~_k = D_k - U_{k-1}^T D~_{k-1}^-1 U_{k-1}  (fp64 issue is
  //       2 cycles per warp instruction whatever the lane count, so the recurrence is done ONCE, not per thread);
  //   B1  thread <-> border column: Bt = B_k + carry, FB = D~^-1 Bt, carry' = -U^T FB  (registers only,
  //       next pose's border row prefetched), own diagonal block of S and own rhs accumulated on the way.
  // column c lives on thread NT-1-c so that warp 0 carries columns only when n2 > NT - 32.
  const int ccol = NT - 1 - tid;
  const bool colv = ccol < n2;
  double sdB0 = 0, sdB1 = 0;   // landmark-landmark blocks of the factors at c_new and behind (added to S after the checkpoint)
  {
    const int c = ccol, jr = c >> 1, comp = c & 1;
    double cD[6] = {0, 0, 0, 0, 0, 0}, cg[3] = {0, 0, 0}, gp[3] = {0, 0, 0};        // warp-0 carried state
    double cB[3] = {0, 0, 0}, sd0 = 0, sd1 = 0, glc = 0;                             // column state
    double sdA0 = 0, sdA1 = 0, glB = 0;   // checkpoint split of the landmark terms: diagonal blocks of the factors in front of c_new, rhs of those behind
    if (append) {   // resume behind the closed poses (light step) / behind the checkpoint (rebuild): chain carries, the rhs the last closed pose parked for its successor, border carries
      const double *rs = valid ? fc : ck;
      if (warp == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) cD[i] = rs[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) { cg[i] = rs[6 + i]; gp[i] = k_lo > 0 ? wsp[(size_t)(k_lo - 1) * WS_POSE + 18 + i] : 0.0; }
      }
      if (colv) {
#pragma unroll
        for (int i = 0; i < 3; ++i) cB[i] = rs[16 + 3 * c + i];
        glc = rs[16 + 3 * N2C + c];
      }
    }
    if (c_new >= 0 && ccol < N2C && !colv) {   // columns without a landmark yet: zero in the checkpoint (a later slot resumes from zeros)
      ckn_cB[3 * ccol] = 0.0; ckn_cB[3 * ccol + 1] = 0.0; ckn_cB[3 * ccol + 2] = 0.0; ckn_gl[ccol] = 0.0;
    }
    for (int k0 = k_lo; k0 < T; k0 += CH) {
      const int kc = min(CH, T - k0);
      const uint32_t par = ((k0 - k_lo) / CH) & 1;   // every pose barrier completes one phase per chunk
      __syncthreads();
      for (int i = tid; i < kc * 21; i += NT) stage[(i / 21) * SW + (i % 21)] = wsp[(size_t)(k0 + i / 21) * WS_POSE + (i % 21)];
      __syncthreads();
      if (warp == 0) {
        // B0: the 3x3 recurrence, scalar and identical on every lane (no shuffles: a shuffle round trip costs ~40
        // cycles here, the whole inverse ~80); ~100 fp64 instructions = ~200 issue cycles per pose.
        const long long tb0 = (a.clocks && tid == 0) ? clock64() : 0;
        for (int kk = 0; kk < kc; ++kk) {
          double *w = stage + kk * SW;
          if ((k0 + kk == T - 1 || k0 + kk == c_new) && lane == 0) {   // the closed poses end here: the state the next step resumes from (/ the checkpoint)
            double *sv = (k0 + kk == c_new) ? ckn : fc;
#pragma unroll
            for (int i = 0; i < 6; ++i) sv[i] = cD[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) sv[6 + i] = cg[i];
          }
          const double d0 = w[0] + cD[0], d1 = w[1] + cD[1], d2 = w[2] + cD[2], d3 = w[3] + cD[3], d4 = w[4] + cD[4], d5 = w[5] + cD[5];
          const double g0 = w[6] + cg[0] + gp[0], g1 = w[7] + cg[1] + gp[1], g2 = w[8] + cg[2] + gp[2];
          gp[0] = w[18]; gp[1] = w[19]; gp[2] = w[20];
          const double u00 = w[9], u01 = w[10], u02 = w[11], u10 = w[12], u11 = w[13], u12 = w[14], u20 = w[15], u21 = w[16], u22 = w[17];
          // adjugate C of D~ (symmetric); D~^-1 = C / det.  Everything on the loop-carried path is expressed through C
          // so that the reciprocal (~80 cycles) runs beside the two 3x3 products instead of in front of them.
          const double c00 = d3 * d5 - d4 * d4, c01 = d2 * d4 - d1 * d5, c02 = d1 * d4 - d2 * d3;
          const double c11 = d0 * d5 - d2 * d2, c12 = d1 * d2 - d0 * d4, c22 = d0 * d3 - d1 * d1;
          const double det = d0 * c00 + d1 * c01 + d2 * c02;
          if (!(det > 0.0) || !(d0 > 0.0)) s_bad = 1;
          const double rd = 1.0 / det;
          // CU = C U, Cg = C g
          const double a00 = c00 * u00 + c01 * u10 + c02 * u20, a01 = c00 * u01 + c01 * u11 + c02 * u21, a02 = c00 * u02 + c01 * u12 + c02 * u22;
          const double a10 = c01 * u00 + c11 * u10 + c12 * u20, a11 = c01 * u01 + c11 * u11 + c12 * u21, a12 = c01 * u02 + c11 * u12 + c12 * u22;
          const double a20 = c02 * u00 + c12 * u10 + c22 * u20, a21 = c02 * u01 + c12 * u11 + c22 * u21, a22 = c02 * u02 + c12 * u12 + c22 * u22;
          const double e0 = c00 * g0 + c01 * g1 + c02 * g2, e1 = c01 * g0 + c11 * g1 + c12 * g2, e2 = c02 * g0 + c12 * g1 + c22 * g2;
          // carries for the next pose: -U^T D~^-1 U (symmetric), -U^T D~^-1 g
          const double nrd = -rd;
          cD[0] = (u00 * a00 + u10 * a10 + u20 * a20) * nrd; cD[1] = (u00 * a01 + u10 * a11 + u20 * a21) * nrd; cD[2] = (u00 * a02 + u10 * a12 + u20 * a22) * nrd;
          cD[3] = (u01 * a01 + u11 * a11 + u21 * a21) * nrd; cD[4] = (u01 * a02 + u11 * a12 + u21 * a22) * nrd; cD[5] = (u02 * a02 + u12 * a12 + u22 * a22) * nrd;
          cg[0] = (u00 * e0 + u10 * e1 + u20 * e2) * nrd; cg[1] = (u01 * e0 + u11 * e1 + u21 * e2) * nrd; cg[2] = (u02 * e0 + u12 * e1 + u22 * e2) * nrd;
          if (lane == 0) {   // Dinv | FU | f -> shared (read by the column warps and copied to the workspace below)
            w[21] = c00 * rd; w[22] = c01 * rd; w[23] = c02 * rd; w[24] = c11 * rd; w[25] = c12 * rd; w[26] = c22 * rd;
            w[27] = a00 * rd; w[28] = a01 * rd; w[29] = a02 * rd; w[30] = a10 * rd; w[31] = a11 * rd; w[32] = a12 * rd; w[33] = a20 * rd; w[34] = a21 * rd; w[35] = a22 * rd;
            w[36] = e0 * rd; w[37] = e1 * rd; w[38] = e2 * rd;
            mbar_arrive(&s_bar[kk]);
          }
        }
        if (a.clocks && tid == 0) a.clocks[12 * b + 8] = (k0 > k_lo ? a.clocks[12 * b + 8] : 0) + (clock64() - tb0);   // cycles inside the pose recurrence
      } else {
        if (colv) {
          if (k0 == k_lo) {
            // landmark-landmark diagonal block and rhs of this column: a gather over the column's factors linearised in this
            // step, in pose order; independent loads (8 poses in flight), hidden behind the first poses of the chain
            for (int k = kz; k < T; k += 8) {
              int p1[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) p1[u] = (k + u < T) ? wmi[(size_t)(k + u) * Lt + jr] : 0;
              double m0[8], m1[8], m2[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const double *m = wsm + (size_t)max(p1[u] - 1, 0) * WS_MEAS;
                m0[u] = m[comp]; m1[u] = m[1 + comp]; m2[u] = m[3 + comp];
              }
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (p1[u]) {
                  if (k + u < c_new) { sdA0 += m0[u]; sdA1 += m1[u]; glc += m2[u]; }
                  else { sd0 += m0[u]; sd1 += m1[u]; glB += m2[u]; }
                }
            }
            glc += glB;
          }
          // B1: border column c.  Bt = B_k + carry, FB = D~^-1 Bt, carry' = -U^T FB, rhs; border rows prefetched
          // 4 poses ahead (an L2 round trip is longer than one pose of the chain)
          const long long tb1 = (a.clocks && tid == NT - 1) ? clock64() : 0;
          double nb[4][3];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 3; ++i) nb[u][i] = (u < kc) ? wBt[((size_t)(k0 + u) * 3 + i) * N2C + c] : 0.0;
          for (int kb = 0; kb < kc; kb += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int kk = kb + u;
              if (kk < kc) {
                const int k = k0 + kk;
                const bool open = k == T - 1;
                if (open) {   // state behind the closed poses (see B0)
                  fc_cB[3 * c] = cB[0]; fc_cB[3 * c + 1] = cB[1]; fc_cB[3 * c + 2] = cB[2];
                  fc_gl[c] = glc;
                }
                if (k == c_new) {   // checkpoint: carries into pose c_new; rhs without the factors at c_new and behind
                  ckn_cB[3 * c] = cB[0]; ckn_cB[3 * c + 1] = cB[1]; ckn_cB[3 * c + 2] = cB[2];
                  ckn_gl[c] = glc - glB;
                }
                const double b0 = nb[u][0] + cB[0], b1 = nb[u][1] + cB[1], b2 = nb[u][2] + cB[2];
                if (kk + 4 < kc) {
#pragma unroll
                  for (int i = 0; i < 3; ++i) nb[u][i] = wBt[((size_t)(k + 4) * 3 + i) * N2C + c];
                }
                while (!mbar_try_wait(&s_bar[kk], par)) { }
                const double *w = stage + kk * SW;
                const double e0 = w[21], e1 = w[22], e2 = w[23], e3 = w[24], e4 = w[25], e5 = w[26];
                const double fb0 = e0 * b0 + e1 * b1 + e2 * b2, fb1 = e1 * b0 + e3 * b1 + e4 * b2, fb2 = e2 * b0 + e4 * b1 + e5 * b2;
                glc -= b0 * w[36] + b1 * w[37] + b2 * w[38];
#pragma unroll
                for (int i = 0; i < 3; ++i) cB[i] = -(w[9 + i] * fb0 + w[12 + i] * fb1 + w[15 + i] * fb2);
                if (!open) {   // the open pose keeps its raw border row: it is eliminated again, with the same carry, when the next pose closes it
                  wBt[((size_t)k * 3 + 0) * N2C + c] = b0; wBt[((size_t)k * 3 + 1) * N2C + c] = b1; wBt[((size_t)k * 3 + 2) * N2C + c] = b2;
                }
                if (k >= T - 2) {   // the last two poses' rows stay on the SM for the rank-3 updates of S
                  double *lb = lastB + (k - (T - 2)) * 3 * N2C, *lf = lastFB + (k - (T - 2)) * 3 * N2C;
                  lb[c] = b0; lb[N2C + c] = b1; lb[2 * N2C + c] = b2;
                  lf[c] = fb0; lf[N2C + c] = fb1; lf[2 * N2C + c] = fb2;
                }
                wFB[((size_t)k * 3 + 0) * N2C + c] = fb0; wFB[((size_t)k * 3 + 1) * N2C + c] = fb1; wFB[((size_t)k * 3 + 2) * N2C + c] = fb2;
              }
            }
          }
          if (a.clocks && tid == NT - 1) a.clocks[12 * b + 9] = (k0 > k_lo ? a.clocks[12 * b + 9] : 0) + (clock64() - tb1);   // cycles of border column 0 in its recurrence
        }
        // copy the chunk's Dinv | FU | f to the workspace (needed again by the backward pass)
        while (!mbar_try_wait(&s_bar[kc - 1], par)) { }
        for (int i = tid - 32; i < kc * 18; i += NT - 32) wsp[(size_t)(k0 + i / 18) * WS_POSE + 21 + (i % 18)] = stage[(i / 18) * SW + 21 + (i % 18)];
      }
    }
    // S = cached partial Schur complement (zero on a rebuild) + the landmark-landmark blocks gathered above; gl = reduced rhs
    cp_async_wait_all();
    __syncthreads();
    if (colv) {   // (the blocks of the factors at c_new and behind follow after the checkpoint of S is saved)
      S[(2 * jr) * ldS + c] += (c_new >= 0) ? sdA0 : sd0;
      S[(2 * jr + 1) * ldS + c] += (c_new >= 0) ? sdA1 : sd1;
      gl[c] = glc;
    }
    if (c_new < 0) { sd0 = 0; sd1 = 0; }
    sdB0 = sd0; sdB1 = sd1;
  }
  __syncthreads();

  if (a.clocks && tid == 0) a.clocks[12 * b +2] = clock64();
  // ---------------------------------------------------------------- phase S ---
  // Schur complement S -= sum_k Bt_k^T FB_k : an n2 x 3T x n2 GEMM.  Border rows are staged through shared
  // memory GK poses at a time (one linear, coalesced copy per operand); every thread keeps up to two 4x4
  // tiles of the upper triangle in registers across all chunks; the result is mirrored.
  const int Tc = T - 1;                         // closed poses [k_lo, Tc) enter the cached partial sum; the open pose is added after the save
  // Schur-complement update over the closed poses [ka, kb): S -= sum_k Bt_k^T FB_k
  auto schur = [&](const int ka, const int kb) {
    if (n2 <= 0 || kb <= ka) return;
    const int nt = (n2 + 3) / 4;
    const int ntile = nt * (nt + 1) / 2;
    if (ntile <= NT / 2 && N2C <= 64) {
      // (the usual case, up to 60 border columns) two thread groups split the k dimension: group g takes poses
      // [4g, 4g+4) of every staged round of GK = 8 poses, one 4x4 tile per thread, so all 8 warps carry DFMAs;
      // the next round's rows are prefetched into registers while the current one is multiplied.
      const int grp = tid >> 7, lt = tid & 127;
      const bool tv = lt < ntile;
      int tr = 0, rem = tv ? lt : 0;
      while (rem >= nt - tr) { rem -= nt - tr; ++tr; }
      const int tc = tr + rem, r0 = tr * 4, c0 = tc * 4;
      double acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.0;
      double *gA = gbuf, *gB = gbuf + GK * 3 * N2C;
      constexpr int PF = 6;                                     // 2 operands x GK*3*N2C / NT <= 2 * PF  (N2C <= 64)
      double pa[PF], pb[PF];
      const int rows0 = min(GK, kb - ka) * 3;
      {
        const size_t base0 = (size_t)ka * 3 * N2C;
#pragma unroll
        for (int u = 0; u < PF; ++u) {
          const int i = tid + u * NT;
          pa[u] = (i < rows0 * N2C) ? wBt[base0 + i] : 0.0; pb[u] = (i < rows0 * N2C) ? wFB[base0 + i] : 0.0;
        }
      }
      for (int k0 = ka; k0 < kb; k0 += GK) {
        const int rows = min(GK, kb - k0) * 3;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PF; ++u) {
          const int i = tid + u * NT;
          if (i < rows * N2C) { gA[i] = pa[u]; gB[i] = pb[u]; }
        }
        __syncthreads();
        if (k0 + GK < kb) {
          const int rown = min(GK, kb - k0 - GK) * 3;
          const size_t base = (size_t)(k0 + GK) * 3 * N2C;
#pragma unroll
          for (int u = 0; u < PF; ++u) {
            const int i = tid + u * NT;
            pa[u] = (i < rown * N2C) ? wBt[base + i] : 0.0; pb[u] = (i < rown * N2C) ? wFB[base + i] : 0.0;
          }
        }
        if (tv) {
          const int kb = grp * (GK / 2) * 3, ke = min(rows, kb + (GK / 2) * 3);
          const double *ar = gA + kb * N2C + r0, *bc = gB + kb * N2C + c0;
          for (int ki = kb; ki < ke; ++ki, ar += N2C, bc += N2C) {   // columns beyond n2 read zeros/garbage of padded rows but are never stored
            const double2 a01 = *reinterpret_cast<const double2 *>(ar), a23 = *reinterpret_cast<const double2 *>(ar + 2);
            const double2 b01 = *reinterpret_cast<const double2 *>(bc), b23 = *reinterpret_cast<const double2 *>(bc + 2);
            acc[0] += a01.x * b01.x; acc[1] += a01.x * b01.y; acc[2] += a01.x * b23.x; acc[3] += a01.x * b23.y;
            acc[4] += a01.y * b01.x; acc[5] += a01.y * b01.y; acc[6] += a01.y * b23.x; acc[7] += a01.y * b23.y;
            acc[8] += a23.x * b01.x; acc[9] += a23.x * b01.y; acc[10] += a23.x * b23.x; acc[11] += a23.x * b23.y;
            acc[12] += a23.y * b01.x; acc[13] += a23.y * b01.y; acc[14] += a23.y * b23.x; acc[15] += a23.y * b23.y;
          }
        }
      }
      // S -= acc: group 0 first, then group 1 (two fixed-order passes keep the result deterministic)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        __syncthreads();
        if (tv && grp == g) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int r = r0 + i, cc = c0 + j;
              if (r < n2 && cc < n2 && (tr != tc || cc >= r)) {
                const double v = S[r * ldS + cc] - acc[i * 4 + j];
                S[r * ldS + cc] = v;
                if (r != cc) S[cc * ldS + r] = v;
              }
            }
        }
      }
    } else if (ntile <= 2 * NT) {
      int tr[2], tc[2];
      bool tv[2];
      double acc[2][16];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int t = tid + s * NT;
        tv[s] = t < ntile;
        int r_ = 0, rem = tv[s] ? t : 0;
        while (rem >= nt - r_) { rem -= nt - r_; ++r_; }
        tr[s] = r_; tc[s] = r_ + rem;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[s][i] = 0.0;
      }
      double *gA = gbuf, *gB = gbuf + GK * 3 * N2C;
      for (int k0 = ka; k0 < kb; k0 += GK) {
        const int rows = min(GK, kb - k0) * 3;
        __syncthreads();
        for (int i = tid; i < rows * N2C; i += NT) { gA[i] = wBt[(size_t)k0 * 3 * N2C + i]; gB[i] = wFB[(size_t)k0 * 3 * N2C + i]; }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (!tv[s]) continue;
          const int r0 = tr[s] * 4, c0 = tc[s] * 4;   // columns beyond n2 read zeros/garbage of padded rows but are never stored
          for (int ki = 0; ki < rows; ++ki) {
            const double *ar = gA + ki * N2C + r0, *bc = gB + ki * N2C + c0;
            const double a0 = ar[0], a1 = ar[1], a2 = ar[2], a3 = ar[3], b0 = bc[0], b1 = bc[1], b2 = bc[2], b3 = bc[3];
            acc[s][0] += a0 * b0; acc[s][1] += a0 * b1; acc[s][2] += a0 * b2; acc[s][3] += a0 * b3;
            acc[s][4] += a1 * b0; acc[s][5] += a1 * b1; acc[s][6] += a1 * b2; acc[s][7] += a1 * b3;
            acc[s][8] += a2 * b0; acc[s][9] += a2 * b1; acc[s][10] += a2 * b2; acc[s][11] += a2 * b3;
            acc[s][12] += a3 * b0; acc[s][13] += a3 * b1; acc[s][14] += a3 * b2; acc[s][15] += a3 * b3;
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (!tv[s]) continue;
        const int r0 = tr[s] * 4, c0 = tc[s] * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = r0 + i, cc = c0 + j;
            if (r < n2 && cc < n2 && (tr[s] != tc[s] || cc >= r)) {
              const double v = S[r * ldS + cc] - acc[s][i * 4 + j];
              S[r * ldS + cc] = v;
              if (r != cc) S[cc * ldS + r] = v;
            }
          }
      }
    } else {   // very wide borders (more than 88 landmark columns): direct global-memory version
      for (int t = tid; t < ntile; t += NT) {
        int tr = 0, rem = t;
        while (rem >= nt - tr) { rem -= nt - tr; ++tr; }
        const int tc = tr + rem;
        const int r0 = tr * 4, c0 = tc * 4;
        double acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = 0.0;
        for (int ki = 3 * ka; ki < 3 * kb; ++ki) {
          const double *br = wBt + (size_t)ki * N2C + r0, *fc = wFB + (size_t)ki * N2C + c0;
          double av[4], bv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { av[i] = (r0 + i < n2) ? br[i] : 0.0; bv[i] = (c0 + i < n2) ? fc[i] : 0.0; }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i * 4 + j] += av[i] * bv[j];
        }
        for (int i = 0; i < 4; ++i)
          for (int j = 0; j < 4; ++j) {
            const int r = r0 + i, cc = c0 + j;
            if (r < n2 && cc < n2 && (tr != tc || cc >= r)) {
              const double v = S[r * ldS + cc] - acc[i * 4 + j];
              S[r * ldS + cc] = v;
              if (r != cc) S[cc * ldS + r] = v;
            }
          }
      }
    }
    };
  if (valid) {   // a light step closes one pose: its term Bt^T FB (rank 3) comes straight from shared memory
    for (int i = tid; i < n2 * ldS; i += NT) {
      const int r = i / ldS, cc = i - r * ldS;
      if (r <= cc && cc < n2) {
        const double v = S[i] - (lastB[r] * lastFB[cc] + lastB[N2C + r] * lastFB[N2C + cc] + lastB[2 * N2C + r] * lastFB[2 * N2C + cc]);
        S[i] = v;
        S[cc * ldS + r] = v;
      }
    }
  } else if (c_new >= 0) {   // rebuild that leaves a checkpoint: the sum in front of c_new first, snapshot, then the rest
    schur(k_lo, c_new);
    __syncthreads();
    for (int i = tid; i < N2C * ldS; i += NT) ckn_S[i] = S[i];    // (every row: later slots must find zeros)
    __syncthreads();
    if (colv) { S[(2 * (ccol >> 1)) * ldS + ccol] += sdB0; S[(2 * (ccol >> 1) + 1) * ldS + ccol] += sdB1; }
    __syncthreads();
    schur(c_new, Tc);
  } else {
    schur(k_lo, Tc);
  }
  __syncthreads();
  // S now is the partial Schur complement behind the closed poses: cache it, then subtract the open pose's term Bt^T FB (rank 3)
  for (int i = tid; i < (valid ? n2 : N2C) * ldS; i += NT) fc_S[i] = S[i];    // (a rebuild writes every row: later slots must find zeros)
  for (int i = tid; i < n2 * ldS; i += NT) {
    const int r = i / ldS, cc = i - r * ldS;
    if (r <= cc && cc < n2) {
      const double v = S[i] - (openB[r] * openFB[cc] + openB[N2C + r] * openFB[N2C + cc] + openB[2 * N2C + r] * openFB[2 * N2C + cc]);
      S[i] = v;
      S[cc * ldS + r] = v;
    }
  }
  __syncthreads();

  // backward pass, chunk by chunk from the end: the FB rows and Dinv | FU | f of a chunk come in by cp.async while the previous
  // chunk's marginals (or, for the first chunk, the inversion below) are computed
  constexpr int SWD = 27;                          // staged doubles per pose: Dinv(6) FU(9) f(3) | P(6) u(3)
  const int CE = wide ? NT / 32 : 2 * (NT / 32);   // poses per chunk (two per warp; wide borders: Wc and the FB rows share gbuf)
  auto prefetch_chunk = [&](int k1, int buf) {
    const int k0 = max(0, k1 - CE), kc = k1 - k0;
    const double *src = wFB + (size_t)k0 * 3 * N2C;
    for (int i = tid; i < kc * 3 * N2C / 2; i += NT) cp_async16(FBs + 2 * i, src + 2 * i);
    double *sd = stage + buf * CE * SWD;
    for (int i = tid; i < kc * 18; i += NT) cp_async8(sd + (i / 18) * SWD + (i % 18), wsp + (size_t)(k0 + i / 18) * WS_POSE + 21 + (i % 18));
    cp_async_commit();
  };
  prefetch_chunk(T, 0);

  if (a.clocks && tid == 0) a.clocks[12 * b +3] = clock64();
  // ---------------------------------------------------------------- phase C ---
  // Sigma_ll = S^-1 (Gauss-Jordan on the SPD Schur complement), dl = Sigma_ll gl.
  if (n2 <= 64) {
    // register-resident version: thread (h, c) = (tid / 64, tid % 64) keeps column c, rows h + 4q (q < 16) in
    // registers for all n2 pivots.  Per pivot only the pivot row / column and the pivot's reciprocal travel through
    // shared memory (double-buffered by parity => one barrier per pivot), the row loads are warp-uniform broadcasts,
    // and the special cases (row p, column p, publishing row / column p+1) sit behind warp-uniform branches:
    // the sweep is bound by instruction issue (2 CTAs x 8 warps share 4 schedulers), so instructions are what counts.
    const int h = tid >> 6, c = tid & 63;
    double *rowb = gjb, *colb = gjb + 128, *pivb = gjb + 256;   // [2][64], [2][64], [2]
    double v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = (h + 4 * q < n2 && c < n2) ? S[(h + 4 * q) * ldS + c] : 0.0;
    if (h == 0) rowb[c] = v[0];
    if (c == 0) {
#pragma unroll
      for (int q = 0; q < 16; ++q) colb[h * 16 + q] = v[q];   // a thread's 16 rows are contiguous: 128-bit loads
    }
    if (tid == 0 && n2 > 0) { pivb[0] = gj_rcp(v[0]); if (!(v[0] > 0.0)) s_bad = 1; }
    __syncthreads();
    // the pivot loop is unrolled completely (p = 4 qq + j): which register holds row p, and which warps hold row / column p and
    // p + 1, are compile-time facts of every unrolled body -- no select chains over the 16 registers, no dynamic indexing
#pragma unroll
    for (int qq = 0; qq < 16; ++qq) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = 4 * qq + j;
        if (p < n2) {
          const int par = p & 1;
          const double *rb = rowb + 64 * par, *cb = colb + 64 * par;
          const double piv = pivb[par];
          const double rpc = (c == p) ? piv : rb[c] * piv;
          if ((p >> 5) == (warp & 1)) {      // this warp holds column p: clear it (its old content travels in cb)
            if (c == p) {
#pragma unroll
              for (int q = 0; q < 16; ++q) v[q] = 0.0;
            }
          }
#pragma unroll
          for (int q = 0; q < 16; q += 2) {
            const double2 f2 = *reinterpret_cast<const double2 *>(cb + h * 16 + q);
            v[q] = fma(-f2.x, rpc, v[q]); v[q + 1] = fma(-f2.y, rpc, v[q + 1]);
          }
          if (h == j) v[qq] = rpc;           // this warp holds row p
          if (p + 1 < n2) {                  // publish pivot row / column p+1 and its reciprocal
            const int pn = p + 1, qn = pn >> 2;
            if (h == (pn & 3)) {
              const double vn = v[qn < 16 ? qn : 15];
              rowb[64 * (par ^ 1) + c] = vn;
              if (c == pn) { pivb[par ^ 1] = gj_rcp(vn); if (!(vn > 0.0)) s_bad = 1; }
            }
            if ((pn >> 5) == (warp & 1)) {
              if (c == pn) {
#pragma unroll
                for (int q = 0; q < 16; q += 2) *reinterpret_cast<double2 *>(colb + 64 * (par ^ 1) + h * 16 + q) = make_double2(v[q], v[q + 1]);
              }
            }
          }
          __syncthreads();
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q)
      if (h + 4 * q < n2 && c < n2) S[(h + 4 * q) * ldS + c] = v[q];
    __syncthreads();
  } else {
  // shared-memory version for wider borders
  // Thread (c, h): column c = tid % 64 (+64 for a second pass when n2 > 64), rows r = h mod NH.
  for (int p = 0; p < n2; ++p) {
    if (tid < n2) colp[tid] = S[tid * ldS + p];
    __syncthreads();
    const double pv = colp[p];
    if (tid == 0 && !(pv > 0.0)) s_bad = 1;
    const double piv = 1.0 / pv;
    const int h = tid >> 6;   // NH = NT/64 row groups per column
    for (int c = tid & 63; c < n2; c += 64) {
      double *__restrict__ Sc = S + c;
      const double *__restrict__ cp = colp;
      if (c != p) {
        const double rowpc = Sc[p * ldS] * piv;
        int r = h;
        for (; r + 5 * NH < n2; r += 6 * NH) {
          double s0 = Sc[r * ldS], s1 = Sc[(r + NH) * ldS], s2 = Sc[(r + 2 * NH) * ldS], s3 = Sc[(r + 3 * NH) * ldS], s4 = Sc[(r + 4 * NH) * ldS], s5 = Sc[(r + 5 * NH) * ldS];
          const double c0 = cp[r], c1 = cp[r + NH], c2 = cp[r + 2 * NH], c3 = cp[r + 3 * NH], c4 = cp[r + 4 * NH], c5 = cp[r + 5 * NH];
          s0 -= c0 * rowpc; s1 -= c1 * rowpc; s2 -= c2 * rowpc; s3 -= c3 * rowpc; s4 -= c4 * rowpc; s5 -= c5 * rowpc;
          if (r != p) Sc[r * ldS] = s0;
          if (r + NH != p) Sc[(r + NH) * ldS] = s1;
          if (r + 2 * NH != p) Sc[(r + 2 * NH) * ldS] = s2;
          if (r + 3 * NH != p) Sc[(r + 3 * NH) * ldS] = s3;
          if (r + 4 * NH != p) Sc[(r + 4 * NH) * ldS] = s4;
          if (r + 5 * NH != p) Sc[(r + 5 * NH) * ldS] = s5;
        }
        for (; r < n2; r += NH)
          if (r != p) Sc[r * ldS] -= cp[r] * rowpc;
      }
    }
    __syncthreads();
    // row p of the non-pivot columns and the pivot column itself (after every thread has finished reading them)
    for (int c = tid; c < n2; c += NT) {
      if (c != p) S[p * ldS + c] *= piv;
      S[c * ldS + p] = (c == p) ? piv : -colp[c] * piv;
    }
    __syncthreads();
  }
  }
  if (tid < n2) {
    double s = 0;
    for (int c = 0; c < n2; ++c) s += S[tid * ldS + c] * gl[c];
    dl[tid] = s;
  }
  __syncthreads();

  if (a.clocks && tid == 0) a.clocks[12 * b +4] = clock64();
  // ------------------------------------------------------------ phases D + E ---
  // backward substitution and marginal recovery, fused per chunk of CE poses (descending) -- W never leaves the SM:
  //   D0  warp 0: P_k = Dinv_k + FU_k P_{k+1} FU_k^T (= [Lambda_xx^-1]_kk), u_k = f_k - FU_k u_{k+1}   -> shared memory;
  //   D1  column threads: W_k = FB_k - FU_k W_{k+1}  (W = Lambda_xx^-1 Lambda_xl), FB rows prefetched    -> shared memory,
  //       pose pairs interleaved [pair][column][6] as phase E reads them (D1 does not depend on D0: different warps);
  //   E   one warp per pose pair of the chunk: y = W_k Sigma_ll with two columns per lane (per row r of Sigma_ll three
  //       128-bit broadcast loads of the 6 interleaved W rows + 2 loads of Sigma_ll feed 12 DFMAs), q = y W_k^T and
  //       v = W_k dl reduced across the warp; lanes 0 / 1 finish the two poses: Sigma_kk = P_k + q, delta_k = u_k - v,
  //       estimate = theta (+) delta, information = Sigma_kk^-1 (SLAM2D.cpp:400).
  double tmax = -1e300;
  {
    const int c = ccol;
    double Wn[3] = {0, 0, 0};
    double P0 = 0, P1 = 0, P2 = 0, P3 = 0, P4 = 0, P5 = 0, un0 = 0, un1 = 0, un2 = 0;   // warp-0 carried state (P symmetric packed)
    double *Wc = gbuf;                                  // [CE / 2][N2C][6]
    int buf = 0;
    for (int k1 = T; k1 > 0; k1 -= CE, buf ^= 1) {
      const int k0 = max(0, k1 - CE), kc = k1 - k0;
      double *sd = stage + buf * CE * SWD;
      cp_async_wait_all();
      __syncthreads();                                  // this chunk's inputs have landed; the previous chunk's E has finished reading Wc / its stage
      if (warp == 0) {
        // D0 across lanes: lane l < 9 owns entry (i, j) = (l / 3, l % 3) of P_k and (j == 0) component i of u_k.  Per pose a lane forms
        // row i of M = FU P (9 FMAs), its entry of P' = Dinv + M FU^T and its component of u' = f - FU u; the new P / u go round the warp
        // by shuffles (the same ~13 dependent FMAs instead of ~60 scalar ones issued redundantly on every lane).  P stays full 3x3 in
        // every lane; symmetric by construction up to rounding: the packed upper triangle is what is published, like the scalar version.
        const int li = lane < 9 ? lane / 3 : 0, lj = lane < 9 ? lane % 3 : 0;
        const int ta = min(li, lj), tb = max(li, lj), p6 = ta * 3 + tb - ta * (ta + 1) / 2;   // packed upper-triangle index of (i, j)
        for (int kk = kc - 1; kk >= 0; --kk) {
          double *w = sd + kk * SWD;   // Dinv(6) FU(9) f(3) | P(6) u(3)
          const double fi0 = w[6 + 3 * li], fi1 = w[7 + 3 * li], fi2 = w[8 + 3 * li];      // row i of FU
          const double fj0 = w[6 + 3 * lj], fj1 = w[7 + 3 * lj], fj2 = w[8 + 3 * lj];      // row j of FU
          // row i of M = FU P
          const double m0 = fi0 * P0 + fi1 * P1 + fi2 * P2, m1 = fi0 * P1 + fi1 * P3 + fi2 * P4, m2 = fi0 * P2 + fi1 * P4 + fi2 * P5;
          const double pn = w[p6] + m0 * fj0 + m1 * fj1 + m2 * fj2;                          // P'(i, j)
          const double vn = w[15 + li] - (fi0 * un0 + fi1 * un1 + fi2 * un2);                // u'(i)
          P0 = __shfl_sync(0xffffffffu, pn, 0); P1 = __shfl_sync(0xffffffffu, pn, 1); P2 = __shfl_sync(0xffffffffu, pn, 2);
          P3 = __shfl_sync(0xffffffffu, pn, 4); P4 = __shfl_sync(0xffffffffu, pn, 5); P5 = __shfl_sync(0xffffffffu, pn, 8);
          un0 = __shfl_sync(0xffffffffu, vn, 0); un1 = __shfl_sync(0xffffffffu, vn, 3); un2 = __shfl_sync(0xffffffffu, vn, 6);
          if (lane == 0) { w[18] = P0; w[19] = P1; w[20] = P2; w[21] = P3; w[22] = P4; w[23] = P5; w[24] = un0; w[25] = un1; w[26] = un2; }
        }
      }
      if (colv) {
        // D1: W_k = FB_k - FU_k W_{k+1}; everything it reads is in shared memory
        for (int kk = kc - 1; kk >= 0; --kk) {
          const double *fb = FBs + (size_t)kk * 3 * N2C + c;
          const double *w = sd + kk * SWD + 6;   // FU
          const double W0 = fb[0] - (w[0] * Wn[0] + w[1] * Wn[1] + w[2] * Wn[2]);
          const double W1 = fb[N2C] - (w[3] * Wn[0] + w[4] * Wn[1] + w[5] * Wn[2]);
          const double W2 = fb[2 * N2C] - (w[6] * Wn[0] + w[7] * Wn[1] + w[8] * Wn[2]);
          double *wo_ = Wc + ((size_t)(kk >> 1) * N2C + c) * 6 + 3 * (kk & 1);
          wo_[0] = W0; wo_[1] = W1; wo_[2] = W2;
          Wn[0] = W0; Wn[1] = W1; Wn[2] = W2;
        }
      }
      __syncthreads();
      if (k0 > 0) prefetch_chunk(k0, buf ^ 1);          // the next chunk streams in beside this chunk's marginals
      if (2 * warp < kc) {
        const double *wb = Wc + (size_t)warp * N2C * 6;   // [n2][6]: W_ka rows 0..2, W_kb rows 0..2 (kb may be beyond the chunk: its sums are not used)
        double qa[6] = {0, 0, 0, 0, 0, 0}, va[3] = {0, 0, 0}, qb[6] = {0, 0, 0, 0, 0, 0}, vb[3] = {0, 0, 0};
        for (int cb = 0; cb < n2; cb += 64) {
          const int c1 = cb + lane, c2 = cb + 32 + lane;
          const bool v1 = c1 < n2, v2 = c2 < n2;
          const double *sp1 = S + (v1 ? c1 : 0), *sp2 = S + (v2 ? c2 : 0);
          const double *wr = wb;
          double ya0 = 0, ya1 = 0, ya2 = 0, yb0 = 0, yb1 = 0, yb2 = 0, za0 = 0, za1 = 0, za2 = 0, zb0 = 0, zb1 = 0, zb2 = 0;
#pragma unroll 4
          for (int r = 0; r < n2; ++r) {
            const double2 w01 = *reinterpret_cast<const double2 *>(wr), w23 = *reinterpret_cast<const double2 *>(wr + 2), w45 = *reinterpret_cast<const double2 *>(wr + 4);
            const double s1 = *sp1, s2 = *sp2;
            ya0 += w01.x * s1; ya1 += w01.y * s1; ya2 += w23.x * s1; yb0 += w23.y * s1; yb1 += w45.x * s1; yb2 += w45.y * s1;
            za0 += w01.x * s2; za1 += w01.y * s2; za2 += w23.x * s2; zb0 += w23.y * s2; zb1 += w45.x * s2; zb2 += w45.y * s2;
            wr += 6; sp1 += ldS; sp2 += ldS;
          }
          if (v1) {
            const double *w = wb + c1 * 6;
            const double d = dl[c1];
            qa[0] += ya0 * w[0]; qa[1] += ya0 * w[1]; qa[2] += ya0 * w[2]; qa[3] += ya1 * w[1]; qa[4] += ya1 * w[2]; qa[5] += ya2 * w[2];
            qb[0] += yb0 * w[3]; qb[1] += yb0 * w[4]; qb[2] += yb0 * w[5]; qb[3] += yb1 * w[4]; qb[4] += yb1 * w[5]; qb[5] += yb2 * w[5];
            va[0] += w[0] * d; va[1] += w[1] * d; va[2] += w[2] * d; vb[0] += w[3] * d; vb[1] += w[4] * d; vb[2] += w[5] * d;
          }
          if (v2) {
            const double *w = wb + c2 * 6;
            const double d = dl[c2];
            qa[0] += za0 * w[0]; qa[1] += za0 * w[1]; qa[2] += za0 * w[2]; qa[3] += za1 * w[1]; qa[4] += za1 * w[2]; qa[5] += za2 * w[2];
            qb[0] += zb0 * w[3]; qb[1] += zb0 * w[4]; qb[2] += zb0 * w[5]; qb[3] += zb1 * w[4]; qb[4] += zb1 * w[5]; qb[5] += zb2 * w[5];
            va[0] += w[0] * d; va[1] += w[1] * d; va[2] += w[2] * d; vb[0] += w[3] * d; vb[1] += w[4] * d; vb[2] += w[5] * d;
          }
        }
        // reduction over the columns: the two halves of the warp first trade poses (lower half keeps pose a, upper half pose b:
        // 9 exchanges), then 4 butterfly steps inside each half -- 45 shuffled values instead of 90
        double r9[9];
        {
          const bool up = lane >= 16;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const double mine = up ? qb[i] : qa[i], send = up ? qa[i] : qb[i];
            r9[i] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const double mine = up ? vb[i] : va[i], send = up ? va[i] : vb[i];
            r9[6 + i] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < 9; ++i) r9[i] += __shfl_xor_sync(0xffffffffu, r9[i], o);
        }
        const int kk = 2 * warp + (lane >> 4);          // lanes 0 / 16 hold the sums of the pair's two poses
        if ((lane & 15) == 0 && kk < kc) {
          const int k = k0 + kk;
          const double *w = sd + kk * SWD + 18;         // P(6) u(3)
          // Sigma_kk = P_k + q, delta_k = u_k - v; the inverse (information) and the estimate follow in one pass over all poses below
          double *pc = a.pose_cov + ((size_t)b * Tmax + k) * 6;
          double tr = 0;
#pragma unroll
          for (int i = 0; i < 6; ++i) { const double cv = w[i] + r9[i]; pc[i] = cv; if (i == 0 || i == 3 || i == 5) tr += cv; }
          del[3 * k] = w[6] - r9[6]; del[3 * k + 1] = w[7] - r9[7]; del[3 * k + 2] = w[8] - r9[8];
          tmax = fmax(tmax, tr);
        }
      }
    }
  }
  __syncthreads();
  // estimate = theta (+) delta, information = Sigma_kk^-1 (SLAM2D.cpp:400): one thread per pose, all poses side by side
  for (int k = tid; k < T; k += NT) {
    const double *pc = a.pose_cov + ((size_t)b * Tmax + k) * 6;
    double C[6], I[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) C[i] = pc[i];
    dge_sym3_inv(C, I);
    double *pi = a.pose_info + ((size_t)b * Tmax + k) * 6;
#pragma unroll
    for (int i = 0; i < 6; ++i) pi[i] = I[i];
    const Pose3 e = dge_compose(Pose3{lin[3 * k], lin[3 * k + 1], lin[3 * k + 2]}, Pose3{del[3 * k], del[3 * k + 1], del[3 * k + 2]});
    est[3 * k] = e.x; est[3 * k + 1] = e.y; est[3 * k + 2] = e.th;
  }
  if (a.clocks && tid == 0) a.clocks[12 * b +5] = clock64();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
  if (lane == 0) red[warp] = tmax;
  // landmarks: delta, estimate, marginal covariance (SLAM2D.cpp:415-424)
  double lerr = 0.0;
  for (int j = tid; j < Lt; j += NT) {
    if (!obs[j]) { lerr += 1.0; continue; }  // exploration_env.py:175: sigma0 (=1.0 default argument) per unobserved landmark
    const int c0 = 2 * lidx[j];
    dell[2 * j] = dl[c0]; dell[2 * j + 1] = dl[c0 + 1];
    const double ex = linl[2 * j] + dl[c0], ey = linl[2 * j + 1] + dl[c0 + 1];
    estl[2 * j] = ex; estl[2 * j + 1] = ey;
    double *lc = a.land_cov + ((size_t)b * Lt + j) * 3;
    lc[0] = S[c0 * ldS + c0]; lc[1] = S[c0 * ldS + c0 + 1]; lc[2] = S[(c0 + 1) * ldS + c0 + 1];
    const double dx = a.lm_true[((size_t)b * Lt + j) * 2] - ex, dy = a.lm_true[((size_t)b * Lt + j) * 2 + 1] - ey;
    lerr += sqrt(dx * dx + dy * dy);
  }
  lerr = warp_sum(lerr);
  if (lane == 0) red[NT / 32 + warp] = lerr;
  __syncthreads();
  if (tid == 0) {
    double m = red[0], le = 0;
    for (int w = 0; w < NT / 32; ++w) { m = fmax(m, red[w]); le += red[NT / 32 + w]; }
    a.metrics[8 * b + 4] = le / Lt;   // ExplorationEnv.get_landmark_error  exploration_env.py:170-176
    a.metrics[8 * b + 5] = m;         // max_uncertainty_of_trajectory        exploration_env.py:190-194
    a.update_count[b] = uc;
    a.fc_valid[b] = s_bad ? 0 : T;
    if (s_bad) ckp[2 * DGE_CK_DEPTH] = 0;
    else if (!valid && a.incremental) {       // the stack after this rebuild: invalidated checkpoints popped, the new one pushed (oldest dropped when full)
      int cnt = ck_cnt;
      if (c_new >= 0) {
        if (cnt == DGE_CK_DEPTH) {
          const int freed = ckp[DGE_CK_DEPTH];
          for (int i = 0; i + 1 < DGE_CK_DEPTH; ++i) { ckp[i] = ckp[i + 1]; ckp[DGE_CK_DEPTH + i] = ckp[DGE_CK_DEPTH + i + 1]; }
          ckp[2 * DGE_CK_DEPTH - 1] = freed;
          cnt = DGE_CK_DEPTH - 1;
        }
        ckp[cnt] = c_new;
        ++cnt;
      }
      ckp[2 * DGE_CK_DEPTH] = cnt;
    }
    if (a.clocks) { a.clocks[12 * b +6] = clock64(); a.clocks[12 * b +7] = T; a.clocks[12 * b + 10] = valid ? 1 : (c_use > 0 ? 2 : 0); a.clocks[12 * b + 11] = n2; }
    if (s_bad) a.status[b] = 1;
  }
}

// Cost-ordered placement of the envs on the SMs.  A launch lasts as long as its slowest CTA, and the step kernels run one CTA per
// env with two CTAs per SM: blocks [0, n_sm) are dispatched one per SM, blocks [n_sm, B) fill the second slots of SMs 0 .. B-n_sm-1.
// The 2 n_sm - B most expensive envs (long trajectory, rebuild due) go to the blocks that keep an SM to themselves, the rest is
// paired expensive-with-cheap.  cost = trajectory length, doubled when the next update is on the relinearisation schedule.
__global__ void __launch_bounds__(1024) k_step_order(int B, int n_sm, int relin_skip, const uint8_t *mask, const int32_t *n_poses,
                                                     const int32_t *update_count, int32_t *order) {
  extern __shared__ int s_cost[];
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int cst = -1;
    if (!mask || mask[b]) {
      cst = n_poses[b];
      if (relin_skip > 0 && (update_count[b] + 1) % relin_skip == 0) cst *= 2;
    }
    s_cost[b] = cst;
  }
  __syncthreads();
  const int solo = (B > n_sm && B < 2 * n_sm) ? 2 * n_sm - B : 0;   // blocks [B - n_sm, n_sm) share their SM with nobody
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int cst = s_cost[b];
    int rank = 0;
    for (int o = 0; o < B; ++o) { const int co = s_cost[o]; rank += (co > cst || (co == cst && o < b)) ? 1 : 0; }
    int blk;
    if (solo == 0) blk = rank;
    else if (rank < solo) blk = (B - n_sm) + rank;               // the most expensive: alone on an SM
    else {
      const int r = rank - solo, half = B - n_sm;                 // 2 * half envs left, two per SM: r-th most expensive with r-th cheapest
      blk = r < half ? r : n_sm + (2 * half - 1 - r);
    }
    order[blk] = b;
  }
}

}  // namespace

size_t dge_slam_smem_bytes(int Lt) {
  const size_t n2c = 2 * (size_t)Lt;
  return (n2c * n2c + CH * SW + 3 * n2c + 2 * (NT / 32) + 258 + 2 * GK * 3 * n2c + 12 * n2c + (n2c > WIDE_N2C ? 0 : 2 * GK * 3 * n2c)) * sizeof(double) +
         2 * (size_t)Lt * sizeof(int) + 16;
}

int dge_launch_slam(dge_engine *e, const uint8_t *mask, cudaStream_t st) {
  SlamArgs a;
  a.cfg = e->cfg; a.d = e->d;
  a.n_poses = e->n_poses; a.update_count = e->update_count; a.status = e->status;
  a.prior_pose = e->prior_pose; a.odom = e->odom; a.lm_true = e->lm_true;
  a.lin_pose = e->lin_pose; a.est_pose = e->est_pose; a.delta_pose = e->delta_pose; a.pose_cov = e->pose_cov; a.pose_info = e->pose_info;
  a.meas_ptr = e->meas_ptr; a.meas_id = e->meas_id; a.meas_pose = e->meas_pose; a.meas_b = e->meas_b; a.meas_r = e->meas_r;
  a.observed = e->observed; a.lin_l = e->lin_l; a.est_l = e->est_l; a.delta_l = e->delta_l; a.land_cov = e->land_cov;
  a.ws_pose = e->ws_pose; a.ws_meas = e->ws_meas; a.ws_Bt = e->ws_Bt; a.ws_FB = e->ws_FB; a.ws_midx = e->ws_midx;
  a.lm_slot = e->lm_slot; a.fc_valid = e->fc_valid; a.fc_state = e->fc_state;
  a.lm_first = e->lm_first; a.ck_pos = e->ck_pos; a.ck_state = e->ck_state;
  static const int incremental = [] { const char *v = getenv("DGE_SLAM_INCREMENTAL"); return (v && v[0] == '0') ? 0 : 1; }();
  a.incremental = incremental;
  static const int ordered = [] { const char *v = getenv("DGE_STEP_ORDER"); return (v && v[0] == '1') ? 1 : 0; }();   // off by default: measured no gain (r02)
  static int n_sm = 0;
  if (!n_sm && (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, e->device) != cudaSuccess || n_sm <= 0)) n_sm = 148;
  a.order = nullptr;
  if (ordered && e->d.B > n_sm && e->d.B < 2 * n_sm && e->d.B <= 4096) {
    k_step_order<<<1, 1024, e->d.B * sizeof(int), st>>>(e->d.B, n_sm, e->cfg.relin_skip, mask, e->n_poses, e->update_count, e->step_order);
    a.order = e->step_order;
  }
  e->step_order_live = a.order != nullptr;
  a.metrics = e->metrics;
  a.clocks = e->slam_clocks;
  const size_t smem = dge_slam_smem_bytes(e->d.Lt);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(k_slam, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DGE_ECUDA;
    configured = smem;
  }
  k_slam<<<e->d.B, NT, smem, st>>>(a, mask);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}
