// Internal definitions of the batched exploration-graph engine (sm_100a only).
// HBM layout: structure-of-arrays, env-major; every per-env array is contiguous so the
// CTA that owns an env streams it with fully coalesced 8-byte loads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dge.h"

#define DGE_PI 3.14159265358979323846
#define DGE_WS_POSE 40
#define DGE_FC_WIDTH(Lt) (16 + 4 * (2 * (Lt)) + (2 * (Lt)) * (2 * (Lt)))
#define DGE_CK_DEPTH 4                       /* checkpoints of the elimination state kept per env (a stack, oldest dropped) */
#define DGE_CK_STRIDE (2 * DGE_CK_DEPTH + 1)  /* ints per env in ck_pos: positions [DEPTH] (stack order), physical slots [DEPTH] (a permutation), count */

struct DgeDims {
  int B, Tmax, Lt, rows, cols, V, Mmax;
  int Ncap, Ecap, Fmax;   // per-env graph capacities
};

struct dge_engine {
  dge_config cfg;
  DgeDims d;
  int device;
  // ---- ground truth / simulator ----
  double *true_pose;     // [B,3]
  double *lm_true;       // [B,Lt,2] by id
  int32_t *scan_id;      // [B,Lt]   scan slot -> id  (std::unordered_map order in the reference, q6)
  uint64_t *seed;        // [B] Philox key
  // ---- SLAM state ----
  int32_t *n_poses, *sim_step, *update_count, *status;
  double *prior_pose;    // [B,3]
  double *lin_pose, *est_pose, *delta_pose, *odom;   // [B,Tmax,3]
  double *pose_cov, *pose_info;                      // [B,Tmax,6]
  int32_t *meas_ptr;     // [B,Tmax+1]
  int32_t *meas_id;      // [B,Mmax]
  int32_t *meas_pose;    // [B,Mmax]  pose index of each measurement factor
  double *meas_b, *meas_r;
  uint8_t *observed;     // [B,Lt]
  double *lin_l, *est_l, *delta_l;   // [B,Lt,2]
  double *land_cov;      // [B,Lt,3]
  // ---- solver workspace (L2-resident, streamed once forward, once backward) ----
  double *ws_pose;       // [B,Tmax,DGE_WS_POSE]: D(6) g(3) U(9) gnext(3) | @21 Dinv(6) FU(9) f(3)
  double *ws_meas;       // [B,Mmax,5]: landmark-landmark block (3) + landmark rhs (2) per measurement
  double *ws_Bt;         // [B,Tmax,3,2Lt]  border rows (pose-landmark blocks, then eliminated rows)
  double *ws_FB;         // [B,Tmax,3,2Lt]  Dinv*Bt (becomes W = Lambda_xx^-1 Lambda_xl in the backward pass)
  int32_t *ws_midx;      // [B,Tmax,Lt]     measurement index+1 of (pose, landmark slot), 0 = none
  int32_t *lm_slot;      // [B,Lt]          landmark id -> border slot of the SLAM solve (-1 unobserved)
  int32_t *fc_valid;     // [B]             number of poses at which the cached forward-elimination state was saved (0 = none)
  int32_t *step_order;   // [B] block -> env of the last SLAM launch (cost-ordered placement); step_order_live: valid for the virtual-map launch that follows
  int step_order_live;
  double *fc_state;      // [B,DGE_FC_WIDTH(Lt)] cached forward-elimination state behind the closed poses (dge_slam.cu)
  double *ck_state;      // [B,DGE_CK_DEPTH,DGE_FC_WIDTH(Lt)] checkpoints of that state: every rebuild leaves one three poses before its end, the next rebuild
                         //                 resumes from the newest one nothing in front of which has moved
  int32_t *ck_pos;       // [B,DGE_CK_STRIDE] the stack: closed poses behind each checkpoint, its physical slot, the count
  int32_t *lm_first;     // [B,Lt]          pose index of a landmark's first observation
  double *vm_prep;       // [B,Tmax,12]     digested poses for the virtual-map kernel
  double *vm_cbox;       // [B,nchunk,4]    per-32-pose bounding boxes
  int32_t *seen;         // [B,V]           integer visibility counts (-1 = landmark cell)
  uint8_t *active;       // [B]             envs that actually stepped in the current dge_step
  // ---- virtual map ----
  double *prob;          // [B,V]
  double *vinfo;         // [B,V,3]
  double *metrics;       // [B,8]
  double *dist;          // [B]
  uint8_t *done;         // [B]
  // ---- action queues ----
  double *plan;          // [B,6]
  int32_t *plan_cursor;  // [B]
  // ---- scratch for host-buffer calls ----
  double *odom_dev_scratch;  // [B,3]
  uint8_t *mask_dev_scratch; // [B]
  uint8_t *mask_dev_scratch2; // [B]   policy-side host calls (dge_graph_host / dge_line_plan_host) -- may run beside dge_step_host
  double *goal_dev_scratch;  // [B,2]
  double *plan_dev_scratch;  // [B,6]
  // graph scratch
  int32_t *g_counts;     // [B,4] N,E,K,F per env
  int32_t *g_frontier;   // [B,Fmax] frontier cell index
  int32_t *g_fassoc;     // [B,Lt+1] node -> frontier association (-1 none): slot 0 robot, 1+i landmark rank i
  int32_t *g_sel;        // [B] position of the env among the selected graphs (-1 = not selected)
  int32_t *g_cnt;        // [B,Ncap] in-degree / row start of each node (CSR build scratch)
  int32_t *g_cur;        // [B,Ncap] fill cursor
  float *g_dis;          // [B,Ncap] deg^-1/2
  int32_t *g_tmp;        // [B,Ecap] unsorted row contents
  double *rdist;         // [B] roll-out distance: sum of sqrt(x^2 + y^2 + angle_weight*theta^2) over executed actions
  int32_t *r_cmap;       // [B,2] (as roll-out engine) source env / frontier of each clone slot
  int32_t *r_cbase;      // [B]   (as source engine) first clone slot of each env
  double *r_u0;          // [B]   (as roll-out engine) utility before the roll-out
  unsigned long long *counters;   // [8] work counters: policy env-steps, sum of T, sum of M, episodes restarted by dge_reset_done_queued,
                                  //     graphs built for a decision, their nodes, their edges, graph batches
  long long *slam_clocks; // [B,12] phase-boundary clocks of the last k_slam launch (+ T in slot 7, sub-phase cycles in 8..11)
  int32_t *forced;       // [B] forced steps left after an in-pipeline reset (| DGE_FRESH_BIT while the initial optimize is pending)
  uint8_t *step_kind;    // [B] 1 = the env's last step was a policy step, 2 = the last forced step of an in-pipeline reset, 0 = forced / fresh
  uint8_t *pending;      // [B] envs that need a decision (dge_mark_pending)
  int64_t *pack_hdr_host; // [16] pinned: header of the last packed graph batch (dge_graph_host_packed_*)
  double *hp_odom;        // [B,3] pinned staging: actions expanded from host plans (dge_step_host_plans_async)
  double *hp_goal;        // [B,2] pinned staging: goals chosen on the host (dge_select_plan_host)
  uint8_t *hp_mask;       // [B]   pinned staging: mask of dge_select_plan_host (value 2 = no frontier left)
  double forced_odom[3]; // host copy of the forced action (exploration_env.py:411-414)
  int count_steps;       // host flag: 1 while stepping on behalf of the policy (reset steps are not counted)
  int park_done;         // host flag: queued stepping skips `done` envs (1, default) or runs every plan to its end (0, roll-out engines)
  // ---- dge_policy_tick (csrc/dge_tick.cu): second stream + fork/join events, capture stream, the captured tick and what it was captured for
  cudaStream_t tick_stream, tick_cap_stream;
  cudaEvent_t ev_fork, ev_move, ev_join;
  cudaStream_t tick_stream2;          // the light group's SLAM -> virtual map chain of a split tick (dge_tick.cu)
  cudaEvent_t ev_split, ev_light;
  uint8_t *group_heavy, *group_light;  // [B] masks of the two groups of a split tick (k_split_groups)
  cudaGraphExec_t tick_exec;
  void *tick_key;
};

// ------------------------------------------------------------- device math ---
__device__ __forceinline__ double dge_wrap_pi(double t) {
  double s, c;
  sincos(t, &s, &c);
  return atan2(s, c);
}

struct Pose3 { double x, y, th; };

__device__ __forceinline__ Pose3 dge_compose(const Pose3 &a, const Pose3 &b) {
  double s, c;
  sincos(a.th, &s, &c);
  Pose3 r;
  r.x = a.x + c * b.x - s * b.y;
  r.y = a.y + s * b.x + c * b.y;
  r.th = dge_wrap_pi(a.th + b.th);
  return r;
}

// between(p1,p2) and (optionally) H1 = d between / d p1  (gtsam Pose2::between)
__device__ __forceinline__ Pose3 dge_between(const Pose3 &p1, const Pose3 &p2, double *H1) {
  double s1, c1, s2, c2;
  sincos(p1.th, &s1, &c1);
  sincos(p2.th, &s2, &c2);
  const double c = c1 * c2 + s1 * s2, s = -s1 * c2 + c1 * s2;
  const double x = p2.x - p1.x, y = p2.y - p1.y;
  Pose3 r;
  r.x = c1 * x + s1 * y;
  r.y = -s1 * x + c1 * y;
  r.th = atan2(s, c);
  if (H1) {
    H1[0] = -c; H1[1] = -s; H1[2] = -s2 * x + c2 * y;
    H1[3] = s;  H1[4] = -c; H1[5] = -c2 * x - s2 * y;
    H1[6] = 0;  H1[7] = 0;  H1[8] = -1;
  }
  return r;
}

// symmetric 3x3 stored as 6: [xx, xy, xt, yy, yt, tt]
__device__ __forceinline__ void dge_sym3_inv(const double *a, double *o, double *det_out = nullptr) {
  const double c00 = a[3] * a[5] - a[4] * a[4];
  const double c01 = a[2] * a[4] - a[1] * a[5];
  const double c02 = a[1] * a[4] - a[2] * a[3];
  const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  const double id = 1.0 / det;
  o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
  o[3] = (a[0] * a[5] - a[2] * a[2]) * id;
  o[4] = (a[1] * a[2] - a[0] * a[4]) * id;
  o[5] = (a[0] * a[3] - a[1] * a[1]) * id;
  if (det_out) *det_out = det;
}

// full 3x3 (row-major 9) inverse of an SPD matrix via its symmetric part
__device__ __forceinline__ void dge_spd3_inv9(const double *A, double *O) {
  const double a[6] = {A[0], 0.5 * (A[1] + A[3]), 0.5 * (A[2] + A[6]), A[4], 0.5 * (A[5] + A[7]), A[8]};
  double o[6];
  dge_sym3_inv(a, o);
  O[0] = o[0]; O[1] = o[1]; O[2] = o[2];
  O[3] = o[1]; O[4] = o[3]; O[5] = o[4];
  O[6] = o[2]; O[7] = o[4]; O[8] = o[5];
}

// ------------------------------------------------------------------ Philox ---
// Philox4x32-10 counter-based generator (perf-mode noise; the reference's libstdc++
// mt19937 streams are reproduced on the host side only, see DESIGN.md "RNG").
__device__ __forceinline__ uint4 dge_philox(uint64_t key, uint64_t ctr_lo, uint64_t ctr_hi) {
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ double dge_u01(uint32_t a, uint32_t b) {  // (0,1), 53 bits
  const uint64_t v = (((uint64_t)a << 32) | b) >> 11;
  return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}
// two independent N(0,1) from one Philox block (Box-Muller)
__device__ __forceinline__ void dge_normal2(uint64_t key, uint64_t ctr_lo, uint64_t ctr_hi, double &n0, double &n1) {
  const uint4 r = dge_philox(key, ctr_lo, ctr_hi);
  const double u0 = dge_u01(r.x, r.y), u1 = dge_u01(r.z, r.w);
  const double rad = sqrt(-2.0 * log(u0));
  double s, c;
  sincospi(2.0 * u1, &s, &c);
  n0 = rad * c; n1 = rad * s;
}

#define DGE_FRESH_BIT 0x40000000

// launch entry points implemented in the .cu files
int dge_launch_reset(dge_engine *e, const uint8_t *mask, const uint64_t *seeds, const double *start, const double *lm,
                     const int32_t *scan, const double *noise, int n_forced, uint64_t seed_stride, cudaStream_t st);
int dge_launch_move_measure(dge_engine *e, const double *odom, const uint8_t *mask, const double *noise, int from_queue, cudaStream_t st);
int dge_launch_slam(dge_engine *e, const uint8_t *mask, cudaStream_t st);
int dge_launch_vmap(dge_engine *e, const uint8_t *mask, cudaStream_t st);
int dge_vmap_standalone(const dge_config *cfg, int n, int T, const double *pose, const double *cov, int L, const double *lm,
                        double *prob, double *vinfo, int32_t *seen, double *prep_ws, double *cbox_ws, cudaStream_t st);
int dge_launch_graph(dge_engine *e, const uint8_t *mask, const dge_graph_out *out, cudaStream_t st);
int dge_launch_mark_pending(dge_engine *e, cudaStream_t st);
int dge_launch_line_plan(dge_engine *e, const double *goal, const uint8_t *mask, double *plan_out, cudaStream_t st);
int dge_launch_select_plan(dge_engine *e, const dge_graph_out *g, const float *q, const uint8_t *mask, int32_t *choice, cudaStream_t st);
void dge_tick_release(dge_engine *e);
int dge_fail(int code, const char *what);   // records the message dge_last_error returns
int dge_vmap_nchunk(int T);
int dge_vmap_prep_width();
size_t dge_slam_smem_bytes(int Lt);

// ------------------------------------------------------------- line planner ---
// closed form of EMPlanner2D::line_planner (Planner2D.cpp:972-1038); the wasted
// initialize() (batch ISAM2 rebuild, :938) is dropped.
__device__ inline void line_plan(const dge_config &cfg, double rx, double ry, double rth, double gx, double gy, double *pl) {
  double root = rth, goal = atan2(gy - ry, gx - rx);
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
  const double drem = path - dq * cfg.max_edge_length;
  pl[0] = quo; pl[1] = sign; pl[2] = rem; pl[3] = dq; pl[4] = drem; pl[5] = quo + 1 + dq + 1;
}


