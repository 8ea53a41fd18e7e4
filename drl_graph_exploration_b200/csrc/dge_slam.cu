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
//   * workspace (D,g,U | Dinv,FU,f | P,u per pose; border rows Bt / FB->W) lives in HBM but is
//     written and re-read by the same CTA within microseconds => L2-resident.
#include "dge_internal.cuh"

namespace {

constexpr int NT = 128;        // threads per CTA; also the max number of border columns (2*Lt <= 128)
constexpr int CH = 32;         // poses staged per shared-memory chunk
constexpr int WS_POSE = 48;    // doubles per pose in ws_pose
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
  double *metrics;
  long long *clocks;   // [B,8] optional: SM clock at the phase boundaries (thread 0), for in-situ phase timing
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

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(NT) k_slam(SlamArgs a, const uint8_t *mask) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (mask && !mask[b]) return;
  const int T = a.n_poses[b];
  const int Lt = a.d.Lt, Tmax = a.d.Tmax, N2C = 2 * Lt;  // N2C = border stride in the workspace
  extern __shared__ double smem[];
  // shared layout
  double *S = smem;                               // [N2C*N2C]  Schur complement -> Sigma_ll
  double *stage = S + (size_t)N2C * N2C;          // [CH*21]
  double *colp = stage + CH * 21;                 // [N2C]
  double *gl = colp + N2C;                        // [N2C]
  double *dl = gl + N2C;                          // [N2C]
  double *red = dl + N2C;                         // [NT/32 * 2]
  int *lidx = (int *)(red + 2 * (NT / 32));       // [Lt]  id -> compact rank (-1 unobserved)
  int *lid = lidx + Lt;                           // [Lt]  rank -> id
  __shared__ int s_nl, s_bad;

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

  // ---------------------------------------------------------------- step 0 ---
  // ISAM2 relinearisation schedule (gtsam ISAM2::update: ++update_count; every
  // relinearizeSkip-th call, variables with max|delta| >= relinearizeThreshold move their
  // linearisation point: theta <- theta (+) delta, delta <- 0).
  const int uc = a.update_count[b] + 1;
  if (tid == 0) { s_bad = 0; }
  if (a.cfg.relin_skip > 0 && uc % a.cfg.relin_skip == 0) {
    for (int k = tid; k < T; k += NT) {
      const double d0 = del[3 * k], d1 = del[3 * k + 1], d2 = del[3 * k + 2];
      if (fmax(fabs(d0), fmax(fabs(d1), fabs(d2))) >= a.cfg.relin_thresh) {
        const Pose3 p = dge_compose(Pose3{lin[3 * k], lin[3 * k + 1], lin[3 * k + 2]}, Pose3{d0, d1, d2});
        lin[3 * k] = p.x; lin[3 * k + 1] = p.y; lin[3 * k + 2] = p.th;
        del[3 * k] = 0; del[3 * k + 1] = 0; del[3 * k + 2] = 0;
      }
    }
    for (int j = tid; j < Lt; j += NT) {
      if (!obs[j]) continue;
      if (fmax(fabs(dell[2 * j]), fabs(dell[2 * j + 1])) >= a.cfg.relin_thresh) {
        linl[2 * j] += dell[2 * j]; linl[2 * j + 1] += dell[2 * j + 1];
        dell[2 * j] = 0; dell[2 * j + 1] = 0;
      }
    }
  }
  if (tid == 0) {  // compact landmark ranks in id order (== gtsam Symbol order of the 'l' keys)
    int n = 0;
    for (int j = 0; j < Lt; ++j) { if (obs[j]) { lidx[j] = n; lid[n] = j; ++n; } else lidx[j] = -1; }
    s_nl = n;
  }
  // zero the sparse border inputs of this step
  for (size_t i = tid; i < (size_t)T * 3 * N2C; i += NT) wBt[i] = 0.0;
  for (size_t i = tid; i < (size_t)T * Lt; i += NT) wmi[i] = 0;
  __syncthreads();
  const int nl = s_nl, n2 = 2 * nl;

  if (a.clocks && tid == 0) a.clocks[8 * b + 0] = clock64();
  // ---------------------------------------------------------------- phase A ---
  // whitened linearisation of every factor at theta.
  // A1: one thread per bearing-range factor (BearingRangeFactor<Pose2,Point2>): pose-pose, pose-landmark
  //     and landmark-landmark blocks + right-hand sides -> ws_meas / border rows.
  const int M = mptr[T];
  const int32_t *mpose = a.meas_pose + (size_t)b * a.d.Mmax;
  for (int p = tid; p < M; p += NT) {
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
  for (int k = tid; k < T; k += NT) {
    const Pose3 pk{lin[3 * k], lin[3 * k + 1], lin[3 * k + 2]};
    double D[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0}, U[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gn[3] = {0, 0, 0};
    if (k == 0) {  // PriorFactor<Pose2>: error = -Local(x, prior), H = I
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
    for (int p = mptr[k]; p < mptr[k + 1]; ++p) {
      const double *m = wsm + (size_t)p * WS_MEAS;
#pragma unroll
      for (int i = 0; i < 6; ++i) D[i] += m[5 + i];
#pragma unroll
      for (int i = 0; i < 3; ++i) g[i] += m[11 + i];
    }
    double *w = wsp + (size_t)k * WS_POSE;
#pragma unroll
    for (int i = 0; i < 6; ++i) w[i] = D[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) { w[6 + i] = g[i]; w[18 + i] = gn[i]; }
#pragma unroll
    for (int i = 0; i < 9; ++i) w[9 + i] = U[i];
  }
  __syncthreads();

  if (a.clocks && tid == 0) a.clocks[8 * b + 1] = clock64();
  // ---------------------------------------------------------------- phase B ---
  // forward elimination along the chain.  The 3x3 pose recurrence is evaluated redundantly
  // by every thread (no communication); thread c additionally carries border column c.
  {
    const int c = tid;
    const bool colv = c < n2;
    const int jr = c >> 1, comp = c & 1;
    double cD[6] = {0, 0, 0, 0, 0, 0}, cg[3] = {0, 0, 0};  // carries -U^T F from the previous pose
    double cB[3] = {0, 0, 0};
    double sd0 = 0, sd1 = 0, glc = 0;                       // own diagonal-block column of S, own gl
    double gprev[3] = {0, 0, 0};                            // rhs of the odometry factor (k-1 -> k) wrt pose k
    for (int k0 = 0; k0 < T; k0 += CH) {
      const int kc = min(CH, T - k0);
      __syncthreads();
      for (int i = tid; i < kc * 21; i += NT) stage[i] = wsp[(size_t)(k0 + i / 21) * WS_POSE + (i % 21)];
      __syncthreads();
      for (int kk = 0; kk < kc; ++kk) {
        const int k = k0 + kk;
        const double *w = stage + kk * 21;
        double D[6], g[3], U[9], Di[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) D[i] = w[i] + cD[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) { g[i] = w[6 + i] + cg[i] + gprev[i]; gprev[i] = w[18 + i]; }
#pragma unroll
        for (int i = 0; i < 9; ++i) U[i] = w[9 + i];
        double det;
        dge_sym3_inv(D, Di, &det);
        if (!(det > 0.0) || !(D[0] > 0.0)) s_bad = 1;
        // fu = Di U ; f = Di g
        double fu[9], f[3];
        const double Dm[9] = {Di[0], Di[1], Di[2], Di[1], Di[3], Di[4], Di[2], Di[4], Di[5]};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
          for (int j = 0; j < 3; ++j) fu[i * 3 + j] = Dm[i * 3] * U[j] + Dm[i * 3 + 1] * U[3 + j] + Dm[i * 3 + 2] * U[6 + j];
          f[i] = Dm[i * 3] * g[0] + Dm[i * 3 + 1] * g[1] + Dm[i * 3 + 2] * g[2];
        }
        // carries for pose k+1: -U^T fu (symmetric), -U^T f
        cD[0] = -(U[0] * fu[0] + U[3] * fu[3] + U[6] * fu[6]);
        cD[1] = -(U[0] * fu[1] + U[3] * fu[4] + U[6] * fu[7]);
        cD[2] = -(U[0] * fu[2] + U[3] * fu[5] + U[6] * fu[8]);
        cD[3] = -(U[1] * fu[1] + U[4] * fu[4] + U[7] * fu[7]);
        cD[4] = -(U[1] * fu[2] + U[4] * fu[5] + U[7] * fu[8]);
        cD[5] = -(U[2] * fu[2] + U[5] * fu[5] + U[8] * fu[8]);
#pragma unroll
        for (int i = 0; i < 3; ++i) cg[i] = -(U[i] * f[0] + U[3 + i] * f[1] + U[6 + i] * f[2]);
        if (tid == 0) {
          double *wo_ = wsp + (size_t)k * WS_POSE + 21;
#pragma unroll
          for (int i = 0; i < 6; ++i) wo_[i] = Di[i];
#pragma unroll
          for (int i = 0; i < 9; ++i) wo_[6 + i] = fu[i];
#pragma unroll
          for (int i = 0; i < 3; ++i) wo_[15 + i] = f[i];
        }
        if (colv) {
          double Bt[3], fb[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) Bt[i] = wBt[((size_t)k * 3 + i) * N2C + c] + cB[i];
          const int p1 = wmi[(size_t)k * Lt + jr];
          if (p1) {  // pose k observes this column's landmark: landmark-landmark block and rhs
            const double *m = wsm + (size_t)(p1 - 1) * WS_MEAS;
            sd0 += comp ? m[1] : m[0];
            sd1 += comp ? m[2] : m[1];
            glc += m[3 + comp];
          }
#pragma unroll
          for (int i = 0; i < 3; ++i) fb[i] = Dm[i * 3] * Bt[0] + Dm[i * 3 + 1] * Bt[1] + Dm[i * 3 + 2] * Bt[2];
          glc -= Bt[0] * f[0] + Bt[1] * f[1] + Bt[2] * f[2];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            cB[i] = -(U[i] * fb[0] + U[3 + i] * fb[1] + U[6 + i] * fb[2]);
            wBt[((size_t)k * 3 + i) * N2C + c] = Bt[i];
            wFB[((size_t)k * 3 + i) * N2C + c] = fb[i];
          }
        }
      }
    }
    // seed S with the landmark-landmark blocks (block diagonal), gl with the reduced rhs
    __syncthreads();
    for (int i = tid; i < n2 * n2; i += NT) S[i] = 0.0;
    __syncthreads();
    if (colv) {
      S[(2 * jr) * n2 + c] = sd0;
      S[(2 * jr + 1) * n2 + c] = sd1;
      gl[c] = glc;
    }
  }
  __syncthreads();

  if (a.clocks && tid == 0) a.clocks[8 * b + 2] = clock64();
  // ---------------------------------------------------------------- phase S ---
  // Schur complement S -= sum_k Bt_k^T FB_k  (n2 x 3T x n2 GEMM, upper 4x4 tiles, mirrored)
  if (n2 > 0) {
    const int nt = (n2 + 3) / 4;
    const int ntile = nt * (nt + 1) / 2;
    for (int t = tid; t < ntile; t += NT) {
      int tr = 0, rem = t;
      while (rem >= nt - tr) { rem -= nt - tr; ++tr; }
      const int tc = tr + rem;
      const int r0 = tr * 4, c0 = tc * 4;
      double acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.0;
      for (int ki = 0; ki < 3 * T; ++ki) {
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
            const double v = S[r * n2 + cc] - acc[i * 4 + j];
            S[r * n2 + cc] = v;
            if (r != cc) S[cc * n2 + r] = v;
          }
        }
    }
  }
  __syncthreads();

  if (a.clocks && tid == 0) a.clocks[8 * b + 3] = clock64();
  // ---------------------------------------------------------------- phase C ---
  // Sigma_ll = S^-1 (in-place Gauss-Jordan on the SPD Schur complement), dl = Sigma_ll gl
  // Thread (c, h): column c = tid % 64 (+64 for a second pass when n2 > 64), rows r = h mod 2.
  // The inner loop is a chain of shared-memory round trips, so loads are batched 6 rows at a time.
  for (int p = 0; p < n2; ++p) {
    if (tid < n2) colp[tid] = S[tid * n2 + p];
    __syncthreads();
    const double pv = colp[p];
    if (tid == 0 && !(pv > 0.0)) s_bad = 1;
    const double piv = 1.0 / pv;
    const int h = tid >> 6;
    for (int c = tid & 63; c < n2; c += 64) {
      double *__restrict__ Sc = S + c;
      const double *__restrict__ cp = colp;
      if (c != p) {
        const double rowpc = Sc[p * n2] * piv;
        int r = h;
        for (; r + 10 < n2; r += 12) {
          double s0 = Sc[r * n2], s1 = Sc[(r + 2) * n2], s2 = Sc[(r + 4) * n2], s3 = Sc[(r + 6) * n2], s4 = Sc[(r + 8) * n2], s5 = Sc[(r + 10) * n2];
          const double c0 = cp[r], c1 = cp[r + 2], c2 = cp[r + 4], c3 = cp[r + 6], c4 = cp[r + 8], c5 = cp[r + 10];
          s0 -= c0 * rowpc; s1 -= c1 * rowpc; s2 -= c2 * rowpc; s3 -= c3 * rowpc; s4 -= c4 * rowpc; s5 -= c5 * rowpc;
          if (r != p) Sc[r * n2] = s0;
          if (r + 2 != p) Sc[(r + 2) * n2] = s1;
          if (r + 4 != p) Sc[(r + 4) * n2] = s2;
          if (r + 6 != p) Sc[(r + 6) * n2] = s3;
          if (r + 8 != p) Sc[(r + 8) * n2] = s4;
          if (r + 10 != p) Sc[(r + 10) * n2] = s5;
        }
        for (; r < n2; r += 2)
          if (r != p) Sc[r * n2] -= cp[r] * rowpc;
      }
    }
    __syncthreads();
    // row p of the non-pivot columns and the pivot column itself (after every thread has finished reading them)
    for (int c = tid; c < n2; c += NT) {
      if (c != p) S[p * n2 + c] *= piv;
      S[c * n2 + p] = (c == p) ? piv : -colp[c] * piv;
    }
    __syncthreads();
  }
  if (tid < n2) {
    double s = 0;
    for (int c = 0; c < n2; ++c) s += S[tid * n2 + c] * gl[c];
    dl[tid] = s;
  }
  __syncthreads();

  if (a.clocks && tid == 0) a.clocks[8 * b + 4] = clock64();
  // --------------------------------------------------------------- phase D1 ---
  // backward substitution: W_k = FB_k - FU_k W_{k+1} (own column), P_k = Dinv_k + FU_k P_{k+1} FU_k^T,
  // u_k = f_k - FU_k u_{k+1} (redundant 3x3 chain).
  {
    const int c = tid;
    const bool colv = c < n2;
    double Wn[3] = {0, 0, 0}, Pn[6] = {0, 0, 0, 0, 0, 0}, un[3] = {0, 0, 0};
    for (int k1 = T; k1 > 0; k1 -= CH) {
      const int k0 = max(0, k1 - CH), kc = k1 - k0;
      __syncthreads();
      for (int i = tid; i < kc * 18; i += NT) stage[i] = wsp[(size_t)(k0 + i / 18) * WS_POSE + 21 + (i % 18)];
      __syncthreads();
      for (int kk = kc - 1; kk >= 0; --kk) {
        const int k = k0 + kk;
        const double *w = stage + kk * 18;
        double Di[6], fu[9], f[3];
#pragma unroll
        for (int i = 0; i < 6; ++i) Di[i] = w[i];
#pragma unroll
        for (int i = 0; i < 9; ++i) fu[i] = w[6 + i];
#pragma unroll
        for (int i = 0; i < 3; ++i) f[i] = w[15 + i];
        // M = fu * Pn (3x3 full), P = Di + M fu^T
        const double Pm[9] = {Pn[0], Pn[1], Pn[2], Pn[1], Pn[3], Pn[4], Pn[2], Pn[4], Pn[5]};
        double M[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) M[i * 3 + j] = fu[i * 3] * Pm[j] + fu[i * 3 + 1] * Pm[3 + j] + fu[i * 3 + 2] * Pm[6 + j];
        double P[6];
        P[0] = Di[0] + M[0] * fu[0] + M[1] * fu[1] + M[2] * fu[2];
        P[1] = Di[1] + M[0] * fu[3] + M[1] * fu[4] + M[2] * fu[5];
        P[2] = Di[2] + M[0] * fu[6] + M[1] * fu[7] + M[2] * fu[8];
        P[3] = Di[3] + M[3] * fu[3] + M[4] * fu[4] + M[5] * fu[5];
        P[4] = Di[4] + M[3] * fu[6] + M[4] * fu[7] + M[5] * fu[8];
        P[5] = Di[5] + M[6] * fu[6] + M[7] * fu[7] + M[8] * fu[8];
        double u[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) u[i] = f[i] - (fu[i * 3] * un[0] + fu[i * 3 + 1] * un[1] + fu[i * 3 + 2] * un[2]);
        if (tid == 0) {
          double *wo_ = wsp + (size_t)k * WS_POSE + 39;
#pragma unroll
          for (int i = 0; i < 6; ++i) wo_[i] = P[i];
#pragma unroll
          for (int i = 0; i < 3; ++i) wo_[6 + i] = u[i];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) Pn[i] = P[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) un[i] = u[i];
        if (colv) {
          double W[3];
#pragma unroll
          for (int i = 0; i < 3; ++i)
            W[i] = wFB[((size_t)k * 3 + i) * N2C + c] - (fu[i * 3] * Wn[0] + fu[i * 3 + 1] * Wn[1] + fu[i * 3 + 2] * Wn[2]);
#pragma unroll
          for (int i = 0; i < 3; ++i) { wFB[((size_t)k * 3 + i) * N2C + c] = W[i]; Wn[i] = W[i]; }
        }
      }
    }
  }
  __syncthreads();

  if (a.clocks && tid == 0) a.clocks[8 * b + 5] = clock64();
  // ---------------------------------------------------------------- phase E ---
  // per pose (one warp each): Sigma_kk = P_k + W_k Sigma_ll W_k^T, delta_k = u_k - W_k dl,
  // estimate = theta (+) delta, information = Sigma_kk^-1 (SLAM2D.cpp:400).
  double tmax = -1e300;
  for (int k = warp; k < T; k += NT / 32) {
    const double *Wk = wFB + (size_t)k * 3 * N2C;
    double q[6] = {0, 0, 0, 0, 0, 0}, v[3] = {0, 0, 0};
    for (int c = lane; c < n2; c += 32) {
      double y0 = 0, y1 = 0, y2 = 0;
      for (int r = 0; r < n2; ++r) {
        const double s = S[r * n2 + c];
        y0 += Wk[r] * s; y1 += Wk[N2C + r] * s; y2 += Wk[2 * N2C + r] * s;
      }
      const double w0 = Wk[c], w1 = Wk[N2C + c], w2 = Wk[2 * N2C + c], d = dl[c];
      q[0] += y0 * w0; q[1] += y0 * w1; q[2] += y0 * w2; q[3] += y1 * w1; q[4] += y1 * w2; q[5] += y2 * w2;
      v[0] += w0 * d; v[1] += w1 * d; v[2] += w2 * d;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) q[i] = warp_sum(q[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
      const double *w = wsp + (size_t)k * WS_POSE + 39;
      double C[6], I[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) C[i] = w[i] + q[i];
      const double d0 = w[6] - v[0], d1 = w[7] - v[1], d2 = w[8] - v[2];
      dge_sym3_inv(C, I);
      double *pc = a.pose_cov + ((size_t)b * Tmax + k) * 6, *pi = a.pose_info + ((size_t)b * Tmax + k) * 6;
#pragma unroll
      for (int i = 0; i < 6; ++i) { pc[i] = C[i]; pi[i] = I[i]; }
      del[3 * k] = d0; del[3 * k + 1] = d1; del[3 * k + 2] = d2;
      const Pose3 e = dge_compose(Pose3{lin[3 * k], lin[3 * k + 1], lin[3 * k + 2]}, Pose3{d0, d1, d2});
      est[3 * k] = e.x; est[3 * k + 1] = e.y; est[3 * k + 2] = e.th;
      tmax = fmax(tmax, C[0] + C[3] + C[5]);
    }
  }
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
    lc[0] = S[c0 * n2 + c0]; lc[1] = S[c0 * n2 + c0 + 1]; lc[2] = S[(c0 + 1) * n2 + c0 + 1];
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
    if (a.clocks) { a.clocks[8 * b + 6] = clock64(); a.clocks[8 * b + 7] = T; }
    if (s_bad) a.status[b] = 1;
  }
}

}  // namespace

size_t dge_slam_smem_bytes(int Lt) {
  const size_t n2c = 2 * (size_t)Lt;
  return (n2c * n2c + CH * 21 + 3 * n2c + 2 * (NT / 32)) * sizeof(double) + 2 * (size_t)Lt * sizeof(int) + 16;
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
