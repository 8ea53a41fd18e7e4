// Simulator kernels: environment reset, motion model, landmark-association scan and
// factor append.  One warp per environment (the scan is over <= Lt true landmarks).
//
// Reference behaviour restated (not translated):
//   reset          pyss2d.py:58-138, Simulator2D::addLandmarks (Simulator2D.cpp:445-465)
//   move           SimpleControlModel::evolve (Simulator2D.cpp:161-182), SLAM2D::addOdometry (SLAM2D.cpp:70-89)
//   association    Simulator2D::measure (Simulator2D.cpp:505-527) -> KDTreeR2::queryRadiusNeighbors
//                  (Distance.cpp:78-97, a linear scan) -> BearingRangeSensorModel::measure/check (:100-132)
//   factor append  SLAM2D::addMeasurement (SLAM2D.cpp:103-124)
#include "dge_internal.cuh"

namespace {

struct SimArgs {
  dge_config cfg;
  DgeDims d;
  double *true_pose, *lm_true, *prior_pose;
  int32_t *scan_id;
  uint64_t *seed;
  int32_t *n_poses, *sim_step, *update_count, *status;
  double *lin_pose, *est_pose, *delta_pose, *odom;
  int32_t *meas_ptr, *meas_id, *meas_pose;
  double *meas_b, *meas_r;
  uint8_t *observed;
  int32_t *lm_first, *ck_pos;
  double *lin_l, *est_l, *delta_l;
  double *prob, *vinfo, *metrics, *dist, *rdist, *plan;
  int32_t *plan_cursor;
  uint8_t *done, *active;
  int32_t *forced;      // [B] forced steps left after an in-pipeline reset; DGE_FRESH_BIT = reset this tick
  uint8_t *step_kind;   // [B] 1 = the last step was a policy step (counted), 0 = forced / fresh, 2 = the LAST forced step of an in-pipeline reset
  double forced_odom[3];
};

// Association scan for the pose that was just appended (index k).  Lanes stride over the
// scan slots; the in-range / gate predicates are evaluated in fp64 without FMA contraction
// so that the emitted (id, slot order) list is reproducible.
__device__ void measure_append(const SimArgs &a, int b, int k, const double *noise /*[2*Lt] or null*/, uint64_t key, uint64_t step_ctr) {
  const int lane = threadIdx.x & 31;
  const int Lt = a.d.Lt;
  const double tx = a.true_pose[3 * b], ty = a.true_pose[3 * b + 1], tth = a.true_pose[3 * b + 2];
  double ts, tc;
  sincos(tth, &ts, &tc);
  const double ex = a.est_pose[((size_t)b * a.d.Tmax + k) * 3], ey = a.est_pose[((size_t)b * a.d.Tmax + k) * 3 + 1];
  double es, ec;
  sincos(a.est_pose[((size_t)b * a.d.Tmax + k) * 3 + 2], &es, &ec);
  int count = a.meas_ptr[(size_t)b * (a.d.Tmax + 1) + k];
  const size_t mbase = (size_t)b * a.d.Mmax;
  for (int s0 = 0; s0 < Lt; s0 += 32) {
    const int s = s0 + lane;
    bool keep = false;
    int id = 0;
    double zb = 0, zr = 0;
    if (s < Lt) {
      id = a.scan_id[(size_t)b * Lt + s];
      const double lx = a.lm_true[((size_t)b * Lt + id) * 2], ly = a.lm_true[((size_t)b * Lt + id) * 2 + 1];
      const double dx = lx - tx, dy = ly - ty;
      const double rng = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      if (rng < a.cfg.max_range) {  // Distance.cpp:86-88
        double nb, nr;
        if (noise) { nb = noise[2 * s]; nr = noise[2 * s + 1]; }
        else { dge_normal2(key, step_ctr, 0x100000000ull + (uint64_t)s, nb, nr); nb *= a.cfg.bearing_noise; nr *= a.cfg.range_noise; }
        const double qx = __dadd_rn(__dmul_rn(tc, dx), __dmul_rn(ts, dy)), qy = __dadd_rn(__dmul_rn(-ts, dx), __dmul_rn(tc, dy));
        zb = atan2(qy, qx) + nb;
        zr = rng + nr;
        keep = zb < a.cfg.max_bearing && zb > a.cfg.min_bearing && zr < a.cfg.max_range && zr > a.cfg.min_range;  // Simulator2D.cpp:100-105
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = count + __popc(bal & ((1u << lane) - 1u));
      a.meas_id[mbase + pos] = id;
      a.meas_pose[mbase + pos] = k;
      a.meas_b[mbase + pos] = zb;
      a.meas_r[mbase + pos] = zr;
      if (!a.observed[(size_t)b * Lt + id]) {  // SLAM2D.cpp:112-123: initial value by transform_from
        double sb, cb;
        sincos(zb, &sb, &cb);
        const double qx = zr * cb, qy = zr * sb;
        const double gx = ex + ec * qx - es * qy, gy = ey + es * qx + ec * qy;
        a.observed[(size_t)b * Lt + id] = 1;
        a.lm_first[(size_t)b * Lt + id] = k;
        const size_t li = ((size_t)b * Lt + id) * 2;
        a.lin_l[li] = gx; a.lin_l[li + 1] = gy;
        a.est_l[li] = gx; a.est_l[li + 1] = gy;
        a.delta_l[li] = 0; a.delta_l[li + 1] = 0;
      }
    }
    count += __popc(bal);
  }
  if (lane == 0) a.meas_ptr[(size_t)b * (a.d.Tmax + 1) + k + 1] = count;
}

__global__ void __launch_bounds__(32) k_reset(SimArgs a, const uint8_t *mask, const uint64_t *seeds, const double *start,
                                              const double *lm, const int32_t *scan, const double *noise, int n_forced, uint64_t seed_stride,
                                              unsigned long long *episodes) {
  const int b = blockIdx.x, lane = threadIdx.x;
  if (mask && !mask[b]) return;
  const int Lt = a.d.Lt;
  // the mask may be the env's own `done` flag (dge_reset_done_queued): it is cleared below, after this read
  const uint64_t key = seeds ? seeds[b] : a.seed[b] + seed_stride;
  __syncwarp();
  if (lane == 0) { a.seed[b] = key; if (episodes) atomicAdd(episodes, 1ull); }
  if (lane < DGE_CK_DEPTH) { a.ck_pos[(size_t)b * DGE_CK_STRIDE + lane] = 0; a.ck_pos[(size_t)b * DGE_CK_STRIDE + DGE_CK_DEPTH + lane] = lane; }   // no checkpoints; slots = identity
  if (lane == 0) a.ck_pos[(size_t)b * DGE_CK_STRIDE + 2 * DGE_CK_DEPTH] = 0;
  // ---- start pose (pyss2d.py:88-95: integer x/y on the *map* half-width, whole-degree heading; q2)
  double sx, sy, sth;
  if (start) { sx = start[3 * b]; sy = start[3 * b + 1]; sth = start[3 * b + 2]; }
  else {
    const uint4 r = dge_philox(key, 0xFFFFFFFF00000000ull, 1);
    const int mx = (int)a.cfg.map_max_x;
    sx = (double)(r.x % (uint32_t)mx) - a.cfg.map_max_x / 2;
    sy = (double)(r.y % (uint32_t)mx) - a.cfg.map_max_x / 2;
    sth = (double)(r.z % 360u) * 0.017453292519943295;
  }
  sth = dge_wrap_pi(sth);
  // ---- true landmarks
  if (lm) {
    for (int i = lane; i < 2 * Lt; i += 32) a.lm_true[(size_t)b * Lt * 2 + i] = lm[(size_t)b * Lt * 2 + i];
  } else if (lane == 0) {  // Simulator2D.cpp:452-463 rejection sampling (sequential by construction)
    uint64_t ctr = 0;
    for (int i = 0; i < Lt;) {
      const uint4 r = dge_philox(key, 0xFFFFFFFF00000001ull, ctr++);
      const double x = (a.cfg.env_max_x - a.cfg.env_min_x) * dge_u01(r.x, r.y) + a.cfg.env_min_x;
      const double y = (a.cfg.env_max_y - a.cfg.env_min_y) * dge_u01(r.z, r.w) + a.cfg.env_min_y;
      const double dx = x - sx, dy = y - sy;
      if (sqrt(dx * dx + dy * dy) < 2.0) continue;
      a.lm_true[((size_t)b * Lt + i) * 2] = x;
      a.lm_true[((size_t)b * Lt + i) * 2 + 1] = y;
      ++i;
    }
  }
  for (int i = lane; i < Lt; i += 32) {
    a.scan_id[(size_t)b * Lt + i] = scan ? scan[(size_t)b * Lt + i] : i;
    a.observed[(size_t)b * Lt + i] = 0;
  }
  // ---- virtual map prior (VirtualMap.cpp:318-340)
  const double i0 = 1.0 / (a.cfg.sigma0 * a.cfg.sigma0);
  for (int i = lane; i < a.d.V; i += 32) {
    a.prob[(size_t)b * a.d.V + i] = 0.5;
    a.vinfo[((size_t)b * a.d.V + i) * 3] = i0;
    a.vinfo[((size_t)b * a.d.V + i) * 3 + 1] = 0;
    a.vinfo[((size_t)b * a.d.V + i) * 3 + 2] = i0;
  }
  if (lane < 8) a.metrics[8 * b + lane] = 0;
  if (lane == 0) {
    a.true_pose[3 * b] = sx; a.true_pose[3 * b + 1] = sy; a.true_pose[3 * b + 2] = sth;
    a.prior_pose[3 * b] = sx; a.prior_pose[3 * b + 1] = sy; a.prior_pose[3 * b + 2] = sth;
    const size_t p0 = (size_t)b * a.d.Tmax * 3;
    a.lin_pose[p0] = sx; a.lin_pose[p0 + 1] = sy; a.lin_pose[p0 + 2] = sth;
    a.est_pose[p0] = sx; a.est_pose[p0 + 1] = sy; a.est_pose[p0 + 2] = sth;
    a.delta_pose[p0] = 0; a.delta_pose[p0 + 1] = 0; a.delta_pose[p0 + 2] = 0;
    a.n_poses[b] = 1; a.sim_step[b] = 1; a.update_count[b] = 0; a.status[b] = 0;
    a.meas_ptr[(size_t)b * (a.d.Tmax + 1)] = 0;
    a.dist[b] = 0; a.rdist[b] = 0; a.done[b] = 0; a.active[b] = 1;
    for (int i = 0; i < 6; ++i) a.plan[6 * b + i] = 0;
    a.plan_cursor[b] = 0;
    // in-pipeline reset: the initial optimize() and the forced steps of ExplorationEnv.reset
    // (exploration_env.py:411-414) ride along the next queued ticks instead of extra launches
    a.forced[b] = n_forced > 0 ? (n_forced | DGE_FRESH_BIT) : 0;
    a.step_kind[b] = 0;
  }
  __syncwarp();
  __threadfence_block();
  // first measure() of SS2D.__init__ (pyss2d.py:134); noise slot = call 1
  measure_append(a, b, 0, noise ? noise + (size_t)b * (3 + 4 * Lt) + 3 + 2 * Lt : nullptr, key, 0);
}

__global__ void __launch_bounds__(32) k_move_measure(SimArgs a, const double *odom_in, const uint8_t *mask, const double *noise, int from_queue) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int Lt = a.d.Lt;
  bool act = !(mask && !mask[b]);
  double ox = 0, oy = 0, oth = 0;
  int fl = 0;
  // from_queue: 0 = explicit odom, 1 = queued line plans (parks done envs), 2 = queued, roll-out clones (run every plan
  // to its end), 3 = explicit odom from the host loop, except for envs an in-pipeline reset left in their forced phase
  const bool queued = from_queue == 1 || from_queue == 2;
  if (act && (from_queue == 1 || from_queue == 3)) {
    fl = a.forced[b];
    if (fl & DGE_FRESH_BIT) {   // reset earlier in this tick: no move, but SLAM (the initial optimize) runs
      if (lane == 0) { a.forced[b] = fl & ~DGE_FRESH_BIT; a.active[b] = 1; a.step_kind[b] = 0; }
      return;
    }
  }
  if (act) {
    if (fl > 0) { ox = a.forced_odom[0]; oy = a.forced_odom[1]; oth = a.forced_odom[2]; }
    else if (queued) {  // expand the compact line plan (Planner2D.cpp:982-1038)
      const double *pl = a.plan + 6 * b;
      const int cur = a.plan_cursor[b], nrot = (int)pl[0], nfwd = (int)pl[3], nact = (int)pl[5];
      if (cur >= nact) act = false;
      else if (cur < nrot) oth = pl[1] * DGE_PI;
      else if (cur == nrot) oth = pl[1] * pl[2];
      else if (cur < nrot + 1 + nfwd) ox = a.cfg.max_edge_length;
      else ox = pl[4];
    } else { ox = odom_in[3 * b]; oy = odom_in[3 * b + 1]; oth = odom_in[3 * b + 2]; }
  }
  // pyss2d.py:173-176 (q3): bounds test on the odom vector itself
  if (act && (!(a.cfg.map_min_x < ox && ox < a.cfg.map_max_x) || !(a.cfg.map_min_y < oy && oy < a.cfg.map_max_y))) act = false;
  const int T = a.n_poses[b];
  if (act && T >= a.d.Tmax) { act = false; if (lane == 0) { a.status[b] = DGE_ECAP; a.done[b] = 1; } }
  // (like the reference, a finished episode can still be stepped explicitly; only the queued
  //  mode parks `done` envs until the caller resets them)
  if (act && from_queue == 1 && a.done[b]) act = false;   // from_queue == 2: roll-out clones run their whole plan
  if (lane == 0) { a.active[b] = act ? 1 : 0; a.step_kind[b] = (act && fl == 0) ? 1 : ((act && fl == 1) ? 2 : 0); }
  if (!act) return;
  const uint64_t key = a.seed[b];
  const uint64_t step_ctr = (uint64_t)a.sim_step[b] | ((uint64_t)a.update_count[b] << 32);
  const double *nz = noise ? noise + (size_t)b * (3 + 4 * Lt) : nullptr;
  if (lane == 0) {
    double nx, ny, nt, unused;
    if (nz) { nx = nz[0]; ny = nz[1]; nt = nz[2]; }
    else {
      dge_normal2(key, step_ctr, 0, nx, ny);
      dge_normal2(key, step_ctr, 1, nt, unused);
      nx *= a.cfg.trans_noise; ny *= a.cfg.trans_noise; nt *= a.cfg.rot_noise;
    }
    Pose3 tp{a.true_pose[3 * b], a.true_pose[3 * b + 1], a.true_pose[3 * b + 2]};
    tp = dge_compose(dge_compose(tp, Pose3{ox, oy, oth}), Pose3{nx, ny, nt});  // Simulator2D.cpp:171-173
    a.true_pose[3 * b] = tp.x; a.true_pose[3 * b + 1] = tp.y; a.true_pose[3 * b + 2] = tp.th;
    const size_t pb = (size_t)b * a.d.Tmax * 3;
    a.odom[pb + 3 * (T - 1)] = ox; a.odom[pb + 3 * (T - 1) + 1] = oy; a.odom[pb + 3 * (T - 1) + 2] = oth;
    const Pose3 prev{a.est_pose[pb + 3 * (T - 1)], a.est_pose[pb + 3 * (T - 1) + 1], a.est_pose[pb + 3 * (T - 1) + 2]};
    const Pose3 p2 = dge_compose(prev, Pose3{ox, oy, oth});  // SLAM2D.cpp:80-88
    a.lin_pose[pb + 3 * T] = p2.x; a.lin_pose[pb + 3 * T + 1] = p2.y; a.lin_pose[pb + 3 * T + 2] = p2.th;
    a.est_pose[pb + 3 * T] = p2.x; a.est_pose[pb + 3 * T + 1] = p2.y; a.est_pose[pb + 3 * T + 2] = p2.th;
    a.delta_pose[pb + 3 * T] = 0; a.delta_pose[pb + 3 * T + 1] = 0; a.delta_pose[pb + 3 * T + 2] = 0;
    a.n_poses[b] = T + 1;
    a.sim_step[b] += 1;
    a.dist[b] += sqrt(ox * ox + oy * oy);  // exploration_env.py:103
    a.rdist[b] += sqrt(ox * ox + oy * oy + a.cfg.angle_weight * oth * oth);   // Planner2D.cpp:1440 (roll-out distance)
    if (fl > 0) a.forced[b] = fl - 1;
    else if (queued) a.plan_cursor[b] += 1;
  }
  __syncwarp();
  __threadfence_block();
  // the obstacle-probe measure() of pyss2d.py:182 only burns RNG draws in the reference (q4):
  // nothing to do with a counter-based generator / explicit noise.
  measure_append(a, b, T, nz ? nz + 3 + 2 * Lt : nullptr, key, step_ctr);
}

SimArgs make_args(dge_engine *e, uint8_t *active) {
  SimArgs a;
  a.cfg = e->cfg; a.d = e->d;
  a.true_pose = e->true_pose; a.lm_true = e->lm_true; a.prior_pose = e->prior_pose; a.scan_id = e->scan_id; a.seed = e->seed;
  a.n_poses = e->n_poses; a.sim_step = e->sim_step; a.update_count = e->update_count; a.status = e->status;
  a.lin_pose = e->lin_pose; a.est_pose = e->est_pose; a.delta_pose = e->delta_pose; a.odom = e->odom;
  a.meas_ptr = e->meas_ptr; a.meas_id = e->meas_id; a.meas_pose = e->meas_pose; a.meas_b = e->meas_b; a.meas_r = e->meas_r;
  a.observed = e->observed; a.lm_first = e->lm_first; a.ck_pos = e->ck_pos; a.lin_l = e->lin_l; a.est_l = e->est_l; a.delta_l = e->delta_l;
  a.prob = e->prob; a.vinfo = e->vinfo; a.metrics = e->metrics; a.dist = e->dist; a.rdist = e->rdist; a.plan = e->plan; a.plan_cursor = e->plan_cursor;
  a.done = e->done; a.active = active;
  a.forced = e->forced; a.step_kind = e->step_kind;
  for (int i = 0; i < 3; ++i) a.forced_odom[i] = e->forced_odom[i];
  return a;
}

}  // namespace

int dge_launch_reset(dge_engine *e, const uint8_t *mask, const uint64_t *seeds, const double *start, const double *lm,
                     const int32_t *scan, const double *noise, int n_forced, uint64_t seed_stride, cudaStream_t st) {
  k_reset<<<e->d.B, 32, 0, st>>>(make_args(e, e->active), mask, seeds, start, lm, scan, noise, n_forced, seed_stride,
                                 seed_stride ? e->counters + 3 : nullptr);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}

int dge_launch_move_measure(dge_engine *e, const double *odom, const uint8_t *mask, const double *noise, int from_queue, cudaStream_t st) {
  k_move_measure<<<e->d.B, 32, 0, st>>>(make_args(e, e->active), odom, mask, noise, from_queue);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}
