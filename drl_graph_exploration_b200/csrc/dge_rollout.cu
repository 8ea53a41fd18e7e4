// Look-ahead roll-out rewards as batched environment clones (row a14 of SURVEY section 8).
//
// Replaces EMPlanner2D::simulations_reward (Planner2D.cpp:1416-1468) and
// ExplorationEnv.rewards_all_goals (exploration_env.py:145-162): for every frontier of every
// deciding env the reference deep-copies SLAM2D / VirtualMap / Simulator2D, rebuilds a batch
// ISAM2 (SLAM2D::set_copy_isam, SLAM2D.cpp:490-497), executes the line plan step by step with
// full SLAM + virtual-map updates and scores utility(before) - utility(after).  It is the
// dominant cost of training (F x ~4 full env-steps per decision, sequential, on one core).
//
// Here every (env, frontier) pair becomes one slot of a second engine: one CTA copies the
// source env's state (coalesced), re-bases the linearisation point on the current estimate,
// queues the line plan; the ordinary step kernels then advance ALL clones of ALL envs together.
#include "dge_internal.cuh"

namespace {

struct EngPtrs {
  DgeDims d;
  double *true_pose, *lm_true, *prior_pose;
  int32_t *scan_id;
  uint64_t *seed;
  int32_t *n_poses, *sim_step, *update_count, *status, *fc_valid, *ck_pos, *lm_first;
  double *lin_pose, *est_pose, *delta_pose, *odom, *pose_cov, *pose_info;
  int32_t *meas_ptr, *meas_id, *meas_pose;
  double *meas_b, *meas_r;
  uint8_t *observed;
  double *lin_l, *est_l, *delta_l, *land_cov;
  double *prob, *vinfo, *metrics, *dist, *rdist, *plan;
  int32_t *plan_cursor;
  uint8_t *done, *active;
};

EngPtrs ptrs(dge_engine *e) {
  EngPtrs p;
  p.d = e->d; p.true_pose = e->true_pose; p.lm_true = e->lm_true; p.prior_pose = e->prior_pose; p.scan_id = e->scan_id; p.seed = e->seed;
  p.n_poses = e->n_poses; p.sim_step = e->sim_step; p.update_count = e->update_count; p.status = e->status; p.fc_valid = e->fc_valid; p.ck_pos = e->ck_pos; p.lm_first = e->lm_first;
  p.lin_pose = e->lin_pose; p.est_pose = e->est_pose; p.delta_pose = e->delta_pose; p.odom = e->odom; p.pose_cov = e->pose_cov; p.pose_info = e->pose_info;
  p.meas_ptr = e->meas_ptr; p.meas_id = e->meas_id; p.meas_pose = e->meas_pose; p.meas_b = e->meas_b; p.meas_r = e->meas_r;
  p.observed = e->observed; p.lin_l = e->lin_l; p.est_l = e->est_l; p.delta_l = e->delta_l; p.land_cov = e->land_cov;
  p.prob = e->prob; p.vinfo = e->vinfo; p.metrics = e->metrics; p.dist = e->dist; p.rdist = e->rdist; p.plan = e->plan; p.plan_cursor = e->plan_cursor;
  p.done = e->done; p.active = e->active;
  return p;
}

// clone slots: envs selected by `mask` contribute fro_size clones each, in env order.
// cmap [Bc,2] = (src env, frontier index) or (-1,-1); cbase [B] first slot of env b; totals[0]=n_clones, [1]=overflow
__global__ void __launch_bounds__(1024) k_rollout_map(int B, int Bc, const uint8_t *mask, const int32_t *g_sel, dge_graph_out g,
                                                      int32_t *cmap, int32_t *cbase, int32_t *totals) {
  __shared__ int sc[1024];
  const int tid = threadIdx.x;
  const int per = (B + 1023) / 1024;
  const int lo = min(B, tid * per), hi = min(B, lo + per);
  int n = 0;
  for (int b = lo; b < hi; ++b) { const int gi = g_sel[b]; if (gi >= 0 && !(mask && !mask[b])) n += g.fro_size[gi]; }
  sc[tid] = n;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int v = tid >= off ? sc[tid - off] : 0;
    __syncthreads();
    sc[tid] += v;
    __syncthreads();
  }
  const int total = sc[1023];
  const bool ovf = total > Bc;
  int c = tid ? sc[tid - 1] : 0;
  for (int b = lo; b < hi; ++b) {
    const int gi = g_sel[b];
    const int F = (gi >= 0 && !(mask && !mask[b])) ? g.fro_size[gi] : 0;
    cbase[b] = c;
    for (int f = 0; f < F; ++f, ++c)
      if (c < Bc) { cmap[2 * c] = b; cmap[2 * c + 1] = f; }
  }
  __syncthreads();
  for (int i = total + tid; i < Bc; i += 1024) { cmap[2 * i] = -1; cmap[2 * i + 1] = -1; }
  if (tid == 0) { totals[0] = min(total, Bc); totals[1] = ovf ? 1 : 0; }
}

__device__ void rollout_line_plan(const dge_config &cfg, double rx, double ry, double rth, double gx, double gy, double *pl) {
  double root = rth, goal = atan2(gy - ry, gx - rx);   // Planner2D.cpp:972-1038
  if (root < 0) root = DGE_PI * 2 + root;
  if (goal < 0) goal = DGE_PI * 2 + goal;
  const double dr = 180 * DGE_PI / 180;
  double diff = goal - root, sign;
  if (diff > DGE_PI) { diff = 2 * DGE_PI - diff; sign = -1; }
  else if (diff > -DGE_PI && diff < 0) { diff = fabs(diff); sign = -1; }
  else if (diff <= -DGE_PI) { diff = 2 * DGE_PI - fabs(diff); sign = 1; }
  else sign = 1;
  const int quo = (int)(diff / dr);
  const double rem = diff - dr * quo;
  const double dx = rx - gx, dy = ry - gy;
  const double path = sqrt(dx * dx + dy * dy);
  const int dq = (int)(path / cfg.max_edge_length);
  pl[0] = quo; pl[1] = sign; pl[2] = rem; pl[3] = dq; pl[4] = path - dq * cfg.max_edge_length; pl[5] = quo + 1 + dq + 1;
}

template <typename T>
__device__ __forceinline__ void copy_n(T *dst, const T *src, size_t n) {
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// one CTA per clone slot
__global__ void __launch_bounds__(256) k_rollout_clone(dge_config cfg, EngPtrs D, EngPtrs S, const int32_t *cmap, dge_graph_out g, double *u0) {
  const int c = blockIdx.x;
  const int b = cmap[2 * c], f = cmap[2 * c + 1];
  const int tid = threadIdx.x;
  if (b < 0) {   // idle slot: empty queue, never steps
    if (tid == 0) { for (int i = 0; i < 6; ++i) D.plan[6 * c + i] = 0; D.plan_cursor[c] = 0; D.done[c] = 1; D.n_poses[c] = 1; D.status[c] = 0; }
    return;
  }
  const size_t Tm = S.d.Tmax, Lt = S.d.Lt, V = S.d.V, Mm = S.d.Mmax;
  const size_t Td = D.d.Tmax, Md = D.d.Mmax;      // the clone engine may hold longer trajectories than the source (room for the line plan)
  const int T = S.n_poses[b];
  const int M = S.meas_ptr[(size_t)b * (Tm + 1) + T];
  // SLAM2D(slam) copy + set_copy_isam: theta := calculateBestEstimate(), delta := 0, fresh ISAM2 (update count 1)
  copy_n(D.lin_pose + c * Td * 3, S.est_pose + b * Tm * 3, (size_t)T * 3);
  copy_n(D.est_pose + c * Td * 3, S.est_pose + b * Tm * 3, (size_t)T * 3);
  for (size_t i = tid; i < (size_t)T * 3; i += blockDim.x) D.delta_pose[c * Td * 3 + i] = 0.0;
  copy_n(D.odom + c * Td * 3, S.odom + b * Tm * 3, (size_t)T * 3);
  copy_n(D.pose_cov + c * Td * 6, S.pose_cov + b * Tm * 6, (size_t)T * 6);
  copy_n(D.pose_info + c * Td * 6, S.pose_info + b * Tm * 6, (size_t)T * 6);
  copy_n(D.meas_ptr + c * (Td + 1), S.meas_ptr + b * (Tm + 1), (size_t)T + 1);
  copy_n(D.meas_id + c * Md, S.meas_id + b * Mm, (size_t)M);
  copy_n(D.meas_pose + c * Md, S.meas_pose + b * Mm, (size_t)M);
  copy_n(D.meas_b + c * Md, S.meas_b + b * Mm, (size_t)M);
  copy_n(D.meas_r + c * Md, S.meas_r + b * Mm, (size_t)M);
  copy_n(D.lm_true + c * Lt * 2, S.lm_true + b * Lt * 2, Lt * 2);
  copy_n(D.scan_id + c * Lt, S.scan_id + b * Lt, Lt);
  copy_n(D.observed + c * Lt, S.observed + b * Lt, Lt);
  copy_n(D.lm_first + c * Lt, S.lm_first + b * Lt, Lt);
  copy_n(D.lin_l + c * Lt * 2, S.est_l + b * Lt * 2, Lt * 2);
  copy_n(D.est_l + c * Lt * 2, S.est_l + b * Lt * 2, Lt * 2);
  for (size_t i = tid; i < Lt * 2; i += blockDim.x) D.delta_l[c * Lt * 2 + i] = 0.0;
  copy_n(D.land_cov + c * Lt * 3, S.land_cov + b * Lt * 3, Lt * 3);
  copy_n(D.prob + c * V, S.prob + b * V, V);          // VirtualMap copy (Planner2D.cpp:1419)
  copy_n(D.vinfo + c * V * 3, S.vinfo + b * V * 3, V * 3);
  if (tid < 3) { D.true_pose[3 * c + tid] = S.true_pose[3 * b + tid]; D.prior_pose[3 * c + tid] = S.prior_pose[3 * b + tid]; }
  if (tid < 8) D.metrics[8 * c + tid] = S.metrics[8 * b + tid];
  if (tid == 0) {
    D.seed[c] = S.seed[b];                // Simulator2D copy incl. RNG state (:1420): every roll-out of an env sees the same noise stream
    D.n_poses[c] = T; D.sim_step[c] = S.sim_step[b]; D.update_count[c] = 1; D.status[c] = 0;
    D.fc_valid[c] = 0;                    // the clone's SLAM solve starts from a full elimination (no cached state or checkpoint travels with a clone)
    for (int i = 0; i < DGE_CK_DEPTH; ++i) { D.ck_pos[c * DGE_CK_STRIDE + i] = 0; D.ck_pos[c * DGE_CK_STRIDE + DGE_CK_DEPTH + i] = i; }
    D.ck_pos[c * DGE_CK_STRIDE + 2 * DGE_CK_DEPTH] = 0;
    D.dist[c] = 0; D.rdist[c] = 0; D.done[c] = 0; D.active[c] = 0;
    const double *p = S.est_pose + ((size_t)b * Tm + T - 1) * 3;
    const int gi = 0; (void)gi;
    rollout_line_plan(cfg, p[0], p[1], p[2], g.frontier_xy[((size_t)b * S.d.Fmax + f) * 2], g.frontier_xy[((size_t)b * S.d.Fmax + f) * 2 + 1], D.plan + 6 * c);
    D.plan_cursor[c] = 0;
    u0[c] = S.metrics[8 * b + 1];         // calculateUtility(virtual_map, 0)  (:1429)
  }
}

// reward = U(before) - U(after, distance)  (Planner2D.cpp:1463-1464) and the min-max normalisation of
// exploration_env.py:154-161 per source env; one warp per source env.
__global__ void __launch_bounds__(32) k_rollout_rewards(int Fmax, const uint8_t *mask, const int32_t *g_sel, dge_graph_out g, const int32_t *cbase,
                                                        const double *u0, const double *c_metrics, const double *c_rdist, double *raw, double *norm,
                                                        uint8_t *loop_clo) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int gi = g_sel[b];
  if (gi < 0 || (mask && !mask[b])) return;
  const int F = g.fro_size[gi], c0 = cbase[b];
  double mx = -1e300, mn = 1e300;
  int am = 0x7fffffff;
  for (int f = lane; f < F; f += 32) {
    const int c = c0 + f;
    const double r = u0[c] - (c_metrics[8 * c + 1] + c_rdist[c] * c_metrics[8 * c + 2]);
    raw[(size_t)b * Fmax + f] = r;
    if (am == 0x7fffffff || r > mx) { mx = r; am = f; }
    mn = fmin(mn, r);
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double omx = __shfl_xor_sync(0xffffffffu, mx, o), omn = __shfl_xor_sync(0xffffffffu, mn, o);
    const int oam = __shfl_xor_sync(0xffffffffu, am, o);
    if (oam != 0x7fffffff && (am == 0x7fffffff || omx > mx || (omx == mx && oam < am))) { mx = omx; am = oam; }
    mn = fmin(mn, omn);
  }
  if (F <= 0) return;
  const bool nf = (am == 0);                      // is_nf: frontier 0 is the robot's nearest frontier
  const double lo = -1.0, hi = nf ? 0.0 : 1.0;
  for (int f = lane; f < F; f += 32) {
    const double r = raw[(size_t)b * Fmax + f];
    // np.interp(r, (mn, mx), (lo, hi)); degenerate mn == mx -> fp[-1] (q16)
    norm[(size_t)b * Fmax + f] = (mx > mn) ? lo + (r - mn) * ((hi - lo) / (mx - mn)) : hi;
  }
  if (lane == 0) loop_clo[b] = nf ? 0 : 1;
}

}  // namespace

extern "C" int dge_rollout_prepare(dge_handle dst, dge_handle src, const dge_graph_out *g, const uint8_t *mask, int32_t *totals_dev, void *stream) {
  if (!dst || !src || !g || !totals_dev) return DGE_EINVAL;
  if (dst->d.Tmax < src->d.Tmax || dst->d.Lt != src->d.Lt || dst->d.V != src->d.V || dst->device != src->device) return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dst->park_done = 0;   // clones execute their whole plan even if the map crosses the 'explored' threshold on the way
  k_rollout_map<<<1, 1024, 0, st>>>(src->d.B, dst->d.B, mask, src->g_sel, *g, dst->r_cmap, src->r_cbase, totals_dev);
  k_rollout_clone<<<dst->d.B, 256, 0, st>>>(src->cfg, ptrs(dst), ptrs(src), dst->r_cmap, *g, dst->r_u0);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}

extern "C" int dge_rollout_rewards(dge_handle dst, dge_handle src, const dge_graph_out *g, const uint8_t *mask, double *raw_dev, double *norm_dev,
                                   uint8_t *loop_clo_dev, void *stream) {
  if (!dst || !src || !g || !raw_dev || !norm_dev || !loop_clo_dev) return DGE_EINVAL;
  k_rollout_rewards<<<src->d.B, 32, 0, static_cast<cudaStream_t>(stream)>>>(src->d.Fmax, mask, src->g_sel, *g, src->r_cbase, dst->r_u0, dst->metrics,
                                                                              dst->rdist, raw_dev, norm_dev, loop_clo_dev);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}
