/* ============================================================================
 * dge.h -- C ABI of the B200-native batched exploration-graph engine (libdge.so)
 *
 * Drop-in boundary for the hot path of RobustFieldAutonomyLab/DRL_graph_exploration.
 * The reference crosses Python<->C++ through two pybind11 modules, one object and
 * one env at a time (src/SS2D.cpp:18-258 `ss2d`, src/Planner2D.cpp:9-106
 * `planner2d`).  A batched GPU engine cannot sit behind an object-per-call API, so
 * every entry point below replaces a *group* of those bindings for B independent
 * environments at once; the reference interface each one replaces is cited.
 * The Python side (`drl_graph_exploration_b200.envs`) re-creates the reference's
 * `ExplorationEnv` / `SS2D` method surface on top of these calls.
 *
 * Conventions: plain pointers + sizes, no torch types.  Pointers suffixed `_dev`
 * are device pointers, `_host` are host pointers (pinned for best speed); `stream`
 * is a `cudaStream_t` passed as void*.  Every call returns 0 on success or a
 * negative DGE_E* code -- it never aborts (the reference `assert`s / throws).
 * All launches are asynchronous on `stream`; host-buffer variants synchronise the
 * stream before returning.  One handle = one device, one stream at a time; one
 * device per PROCESS (the deployment model is one process per GPU: the opt-in
 * shared-memory sizes of the kernels are set once per process, and the caller
 * keeps the handle's device current around every call).
 * ==========================================================================*/
#ifndef DGE_H_
#define DGE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGE_OK 0
#define DGE_EINVAL (-1)   /* bad argument / capacity exceeded at create time        */
#define DGE_ECUDA (-2)    /* CUDA runtime error (see dge_last_error)                */
#define DGE_ENOMEM (-3)
#define DGE_ECAP (-4)     /* a per-env capacity (poses / measurements) overflowed   */

/* Parameters: scripts/envs/exploration_env.ini as read by pyss2d.py:10-55 and
 * overridden by ExplorationEnv.reset (exploration_env.py:399-407).  Angles in
 * radians, already normalised like the reference setters (Simulation2D.h:52-57). */
typedef struct dge_config {
  double env_min_x, env_max_x, env_min_y, env_max_y;
  double map_min_x, map_max_x, map_min_y, map_max_y;
  double resolution, sigma0;
  double bearing_noise, range_noise, min_bearing, max_bearing, min_range, max_range;
  double trans_noise, rot_noise;
  double sigma_x0, sigma_y0, sigma_theta0;
  double angle_weight, dist_w0, dist_w1, max_edge_length, occupancy_threshold;
  double max_steps;
  double relin_thresh;     /* gtsam::ISAM2Params::relinearizeThreshold (SLAM2D.cpp:10) */
  int32_t relin_skip;      /* gtsam::ISAM2Params::relinearizeSkip                      */
  int32_t num_landmarks;   /* [Simulator] num                                         */
} dge_config;

typedef struct dge_engine *dge_handle;

/* Capacities fixed at creation. max_poses bounds the trajectory length per env
 * (an env whose trajectory is full reports done). */
int dge_create(const dge_config *cfg, int n_envs, int max_poses, int device, dge_handle *out);
int dge_destroy(dge_handle h);
const char *dge_last_error(void);
/* out[0..8) = n_envs, max_poses, n_true_landmarks, rows, cols, max_meas_per_env,
 *             max_graph_nodes_per_env, max_graph_edges_per_env */
int dge_dims(dge_handle h, int32_t *out);

/* ---- reset: replaces SS2D.__init__ (pyss2d.py:58-138): Simulator2D(seed),
 * initialize_vehicle, random_landmarks, SLAM2D.add_prior, first measure+optimize.
 *   mask_dev   [B] u8, nullable (null = all envs)
 *   seeds_dev  [B] u64 Philox keys (perf mode RNG)
 *   start_dev  [B,3] f64 nullable: explicit start poses (parity mode; else drawn on device
 *              like pyss2d.py:88-95: integer x/y, whole-degree heading)
 *   lm_dev     [B,Lt,2] f64 nullable: explicit true landmarks by id (else drawn on device
 *              like Simulator2D::addLandmarks, Simulator2D.cpp:445-465)
 *   scan_dev   [B,Lt] i32 nullable: landmark scan order (slot -> id); null = identity
 *   noise_dev  [B, 3+4*Lt] f64 nullable: explicit measurement noise (layout below)   */
int dge_reset(dge_handle h, const uint8_t *mask_dev, const uint64_t *seeds_dev, const double *start_dev,
              const double *lm_dev, const int32_t *scan_dev, const double *noise_dev, void *stream);

/* in-pipeline reset for the queued loop: same world generation as dge_reset, but the initial
 * optimize() (pyss2d.py:135) and the n_forced forced actions of ExplorationEnv.reset
 * (exploration_env.py:411-414, `forced_odom_host` = (1, 1, pi/2)) are executed by the next
 * 1 + n_forced calls of dge_step_queued together with every other env's step -- a reset costs one small
 * launch instead of 1 + 5*n_forced.  Forced steps are not counted in `counters` and the env asks for no decision
 * (dge_state_view.forced != 0) until they are done.                                          */
int dge_reset_queued(dge_handle h, const uint8_t *mask_dev, const uint64_t *seeds_dev, const double *start_dev,
                     const double *lm_dev, const int32_t *scan_dev, const double *forced_odom_host, int n_forced, void *stream);
/* the same for every env whose `done` flag is set, entirely device-side (no mask to build, no host sync):
 * the env's Philox key advances by seed_stride (> 0), `done` is cleared, counters[3] counts the restart. */
int dge_reset_done_queued(dge_handle h, uint64_t seed_stride, const double *forced_odom_host, int n_forced, void *stream);

/* ---- step: replaces ExplorationEnv.step -> SS2D.simulate (exploration_env.py:98-105,
 * pyss2d.py:171-206): Simulator2D.move + SLAM2D.add_odometry, Simulator2D.measure (x2),
 * SLAM2D.add_measurement, SLAM2D.optimize(update_covariance=True),
 * VirtualMap.update_probability + update_information.
 *   odom_dev   [B,3] f64 (x, y, theta) body-frame command
 *   mask_dev   [B] u8 nullable: envs to step
 *   noise_dev  [B, 3+4*Lt] f64 nullable.  Parity mode: [0..3) move noise (x,y,theta);
 *              [3 + c*2*Lt + 2*s + {0,1}] (bearing, range) noise of measure() call c
 *              (0 = obstacle probe, unused; 1 = real) for scan slot s.  Null: Philox.  */
int dge_step(dge_handle h, const double *odom_dev, const uint8_t *mask_dev, const double *noise_dev, void *stream);
/* same, but every env executes the next action of its own queued line plan (filled by
 * dge_select_and_plan); envs whose queue is empty do not step.                       */
int dge_step_queued(dge_handle h, void *stream);
int dge_step_queued_noise(dge_handle h, const double *noise_dev /* [B,3+4*Lt] explicit noise (parity) */, void *stream);
/* the three stages of dge_step, separately launchable (profiling / tests); after the first
 * stage `active` (see dge_state_view) flags the envs that actually moved.             */
int dge_move_measure_queued(dge_handle h, void *stream);
int dge_move_measure(dge_handle h, const double *odom_dev, const uint8_t *mask_dev, const double *noise_dev, void *stream);
int dge_slam_optimize(dge_handle h, const uint8_t *mask_dev, void *stream);   /* SLAM2D::optimize  SLAM2D.cpp:374-430 */
int dge_virtual_map(dge_handle h, const uint8_t *mask_dev, void *stream);     /* VirtualMap.cpp:61-84,256-316         */

/* host-buffer variant of dge_step (the call a ctypes/pybind caller makes): copies odom
 * H2D, steps, copies done flags (and the occupancy maps when obs_host != NULL) D2H.  */
int dge_step_host(dge_handle h, const double *odom_host, const uint8_t *mask_host, uint8_t *done_host,
                  double *obs_host /* [B,rows,cols] nullable */, void *stream);

/* the same for a host-driven acting loop (test.py:100-143 with B envs): flags select
 *   DGE_STEP_NO_SYNC       do not synchronise -- the caller synchronises `stream` before reading the host buffers (lets the
 *                          policy-side host calls below run on another stream meanwhile);
 *   DGE_STEP_HONOR_FORCED  envs that dge_reset_queued / dge_reset_done_queued left in their reset phase (initial optimize +
 *                          forced steps, exploration_env.py:411-414) execute that instead of the supplied odom (they must be
 *                          selected by mask_host);
 * metrics_host [B,8] nullable: dge_state_view.metrics after the step (ExplorationEnv.status / get_landmark_error /
 * max_uncertainty_of_trajectory, exploration_env.py:164-194).                          */
#define DGE_STEP_NO_SYNC 1
#define DGE_STEP_HONOR_FORCED 2
int dge_step_host_async(dge_handle h, const double *odom_host, const uint8_t *mask_host, uint8_t *done_host, double *obs_host,
                        double *metrics_host, int flags, void *stream);

/* the same with the action lists held by the caller in the compact form of dge_line_plan: plans_host [B,6], cursor_host [B] =
 * index of the action to execute ("for act in actions: env.step(act)", test.py:119-120 / policy.py:117-118, for B envs).  */
int dge_step_host_plans_async(dge_handle h, const double *plans_host, const int64_t *cursor_host, const uint8_t *mask_host,
                              uint8_t *done_host, double *obs_host, double *metrics_host, int flags, void *stream);

/* ---- stand-alone virtual-map rebuild on caller-provided belief states (a6+a7;
 * VirtualMap::updateProbability + updateInformation).  n problems, T poses each.
 *   pose_dev [n,T,3], cov_dev [n,T,6] (upper triangle xx,xy,xt,yy,yt,tt of the pose
 *   covariance = inverse information), lm_dev [n,L,2];  outputs prob_dev [n,V],
 *   vinfo_dev [n,V,3] (xx,xy,yy).                                                  */
int dge_virtual_map_rebuild(const dge_config *cfg, int n, int T, const double *pose_dev, const double *cov_dev,
                            int L, const double *lm_dev, double *prob_dev, double *vinfo_dev,
                            int32_t *seen_dev /* [n,V] nullable: integer visibility counts, -1 = landmark cell */,
                            double *ws_dev /* dge_virtual_map_rebuild_ws_doubles(n,T) doubles */, void *stream);
int64_t dge_virtual_map_rebuild_ws_doubles(int n, int T);

/* ---- state views (device pointers owned by the engine; valid until dge_destroy).
 * Replaces the per-object getters of ss2d: SLAM2D.map.iter_trajectory / iter_landmarks,
 * VirtualMap.to_array / to_cov_trace / explored, Simulator2D.vehicle (SS2D.cpp:141-257). */
typedef struct dge_state_view {
  const int32_t *n_poses;        /* [B]                      */
  const int32_t *sim_step;       /* [B] pyss2d self.step      */
  const int32_t *update_count;   /* [B]                      */
  const double *true_pose;       /* [B,3]                    */
  const double *est_pose;        /* [B,Tmax,3]               */
  const double *lin_pose;        /* [B,Tmax,3]               */
  const double *delta_pose;      /* [B,Tmax,3]               */
  const double *pose_cov;        /* [B,Tmax,6] upper triangle */
  const double *pose_info;       /* [B,Tmax,6]               */
  const double *odom;            /* [B,Tmax,3]               */
  const int32_t *meas_ptr;       /* [B,Tmax+1]               */
  const int32_t *meas_id;        /* [B,Mmax]                 */
  const double *meas_bearing;    /* [B,Mmax]                 */
  const double *meas_range;      /* [B,Mmax]                 */
  const double *lm_true;         /* [B,Lt,2] by id           */
  const int32_t *scan_id;        /* [B,Lt]                   */
  const uint8_t *observed;       /* [B,Lt]                   */
  const double *est_l;           /* [B,Lt,2]                 */
  const double *lin_l;           /* [B,Lt,2]                 */
  const double *land_cov;        /* [B,Lt,3] xx,xy,yy        */
  const double *prob;            /* [B,V]                    */
  const double *vinfo;           /* [B,V,3] xx,xy,yy         */
  const int32_t *seen;           /* [B,V] visibility counts, -1 = landmark cell */
  const double *metrics;         /* [B,8]: explored, utility(0), sum cov-trace, #p<thresh,
                                    landmark_error, max pose-cov trace, dist, reserved */
  const uint8_t *done;           /* [B]                      */
  const uint8_t *active;         /* [B] envs that moved in the last (queued) step        */
  const int32_t *status;         /* [B] 0 ok, DGE_ECAP, or 1 = solver breakdown       */
  const double *plan;            /* [B,6] queued line plan (see dge_line_plan)        */
  const int32_t *plan_cursor;    /* [B] next action of the plan                       */
  const int64_t *slam_clocks;    /* [B,12] SM clock at the 7 phase boundaries of the last SLAM launch, [.,7] = T, [.,8..9] = cycles in the pose / border recurrences */
  const int64_t *counters;       /* [8] work counters: env-steps, sum of trajectory lengths, sum of
                                    measurement counts over those steps, episodes restarted
                                    by dge_reset_done_queued, graphs built by dge_graph, their
                                    nodes, their edges, graph batches                          */
  const int32_t *forced;         /* [B] forced steps still queued by dge_reset_queued (bit 30 = initial optimize pending) */
  const int64_t *seed;           /* [B] Philox key of the env's current episode                                         */
  const uint8_t *pending;        /* [B] written by dge_mark_pending                                                     */
} dge_state_view;
int dge_get_state(dge_handle h, dge_state_view *out);
/* steps launched while counting is off (e.g. the 4 forced steps of a reset) do not touch `counters` */
int dge_set_counting(dge_handle h, int on);

/* ---- exploration graph: replaces ExplorationEnv.graph_matrix + frontier
 * (exploration_env.py:196-358), SLAM2D.adjacency_degree_get / key_size / get_key_points
 * (SLAM2D.cpp:141-273) and DeepQ.data_process (policy.py:211-232).  Emits one batched
 * graph (PyG DataLoader layout) for the envs selected by mask, directly on device.
 * Caller-owned output buffers with the capacities reported by dge_dims.             */
typedef struct dge_graph_out {
  float *x;                /* [Ncap,5]  node features f32                              */
  int64_t *edge_index;     /* [2,Ecap]  (row 0 = src, row 1 = dst), stride = edge_cap   */
  float *edge_attr;        /* [Ecap]                                                    */
  int64_t *batch;          /* [Ncap]    graph id (position among selected envs)         */
  int32_t *node_ptr;       /* [B+1]     per selected env                                */
  int32_t *edge_ptr;       /* [B+1]                                                     */
  int32_t *key_size;       /* [B]       K = L + T                                       */
  int32_t *fro_size;       /* [B]       F                                               */
  double *frontier_xy;     /* [B,Fmax,2] goal coordinates of frontier f (indexed by ENV, not by
                              graph position), Fmax = Lt+1                              */
  int32_t *totals;         /* [8]: n_graphs, N_tot, E_tot, overflow flag, #envs done, 0, 0, 0   */
  int64_t node_cap, edge_cap;
  /* optional (all four or none): destination-sorted CSR of the batch + GCNConv(improved) normalisation,
   * built inside the graph kernel so that the GNN needs no preprocessing launches (dge_gnn.h). */
  int32_t *csr_rowptr;     /* [Ncap+1]  incoming edges of node n: csr_perm[csr_rowptr[n] .. csr_rowptr[n+1]) */
  int32_t *csr_perm;       /* [Ecap]    edge ids, ascending inside a row (deterministic summation order)   */
  float *gcn_norm;         /* [Ecap]    deg^-1/2[src] * w * deg^-1/2[dst], deg includes the self loop of 2 */
  float *gcn_selfnorm;     /* [Ncap]    2 / deg                                                            */
} dge_graph_out;
int dge_graph(dge_handle h, const uint8_t *mask_dev, const dge_graph_out *out, void *stream);
/* host-buffer variant of dge_graph: what ExplorationEnv.graph_matrix + DeepQ.data_process hand to the caller, for the
 * envs selected by mask_host, as ONE batch in PyG DataLoader layout.  `dev` = device staging buffers as for dge_graph;
 * the valid prefixes are copied into the (pinned) host buffers: x [N,5], edge_index [2,E] CONTIGUOUS (row 1 starts at
 * element E), edge_attr [E], node_ptr / edge_ptr [G+1], key_size / fro_size [G], frontier_xy [B,Fmax,2] (by env),
 * totals [8] (G, N, E, overflow, #done).  Synchronises `stream`.                                               */
typedef struct dge_graph_host_out {
  float *x; int64_t *edge_index; float *edge_attr;
  int32_t *node_ptr, *edge_ptr, *key_size, *fro_size;
  double *frontier_xy;
  int32_t *totals;
  /* optional (all four or none, and only if `dev` carries them): the destination-sorted CSR and the GCNConv(improved)
   * normalisation of the batch -- csr_rowptr [N+1], csr_perm [E], gcn_norm [E], gcn_selfnorm [N] -- so that a caller
   * who sends the batch back to the device does not have to rebuild them (PyG recomputes them in every GCNConv.forward) */
  int32_t *csr_rowptr, *csr_perm;
  float *gcn_norm, *gcn_selfnorm;
} dge_graph_host_out;
int dge_graph_host(dge_handle h, const uint8_t *mask_host, const dge_graph_out *dev, const dge_graph_host_out *host, void *stream);

/* packed variant: the same batch behind ONE transfer per direction.  `begin` launches the graph kernels and a pack kernel
 * that gathers the valid prefix of every array into `arena_dev` (caller-owned device bytes, >= dge_graph_packed_capacity)
 * and returns without synchronising, so the caller can launch other work (the step of the other envs) meanwhile; `end`
 * synchronises, copies header + payload into `arena_host` (pinned) in one D2H and reports the layout: byte offsets of
 * the sections (16-byte aligned) -- x [N,5] f32, edge_index [2,E] i64 contiguous, edge_attr [E] f32, node_ptr / edge_ptr
 * [G+1] i32, key_size / fro_size [G] i32, frontier_xy [G,Fmax,2] f64 (by GRAPH ordinal, unlike dge_graph_out), csr_rowptr
 * [N+1] i32, csr_perm [E] i32, gcn_norm [E] f32, gcn_selfnorm [N] f32.  Sending `arena_host[:total_bytes]` back to a device
 * buffer gives the GNN its inputs with the same offsets (what DeepQ.test's data.to(device) does, policy.py:255-259).  */
typedef struct dge_graph_packed {
  int64_t total_bytes;
  int64_t x, edge_index, edge_attr, node_ptr, edge_ptr, key_size, fro_size, frontier_xy, csr_rowptr, csr_perm, gcn_norm, gcn_selfnorm;
  int64_t frontier_plan;   /* [G,Fmax,6] f64: the line plan (dge_line_plan's compact form) of EVERY frontier of every graph -- what
                              ExplorationEnv.actions_all_goals returns for the frontier nodes (exploration_env.py:131-143) */
  int32_t n_graphs, n_nodes, n_edges, n_done;
} dge_graph_packed;
int64_t dge_graph_packed_capacity(dge_handle h, const dge_graph_out *dev);
int dge_graph_host_packed_begin(dge_handle h, const uint8_t *mask_host, const dge_graph_out *dev, void *arena_dev, int64_t arena_cap, void *stream);
int dge_graph_host_packed_end(dge_handle h, const void *arena_dev, void *arena_host, int64_t arena_cap, dge_graph_packed *out, void *stream);
/* optional, between ..._begin and ..._end: queue the first `bytes` of the arena for the host right behind the pack kernel (a guess
 * of the batch size, e.g. the previous batch's total_bytes plus a margin); ..._end_prefetched(prefetched = bytes) then skips the
 * second copy and its synchronisation whenever the batch fits in the guess, and copies only the remainder otherwise.       */
int dge_graph_host_packed_prefetch(dge_handle h, const void *arena_dev, void *arena_host, int64_t bytes, void *stream);
int dge_graph_host_packed_end_prefetched(dge_handle h, const void *arena_dev, void *arena_host, int64_t arena_cap, int64_t prefetched,
                                         dge_graph_packed *out, void *stream);
/* host-side policy read-out on such a batch: per selected env (mask_host as given to ..._begin) the first arg-max of q_host over
 * the graph's last fro_size nodes (np.argmax(readout_t[-fro_size:]), test.py:112 / policy.py:109), the chosen frontier as goal and
 * its line plan (actions_all_goals()[key_size + action_index]) into plan_host [B,6]; choice_host [B] nullable (-1: not selected
 * or no frontier left -- that env's episode is declared over, quirk q15).  The plans are taken from the batch's frontier_plan
 * section (no device work, no synchronisation); only when an env has no frontier left the goals go to the device
 * (dge_line_plan_host: that env's done flag is set there) and `stream` is synchronised.                                   */
int dge_select_plan_host(dge_handle h, const void *arena_host, const dge_graph_packed *layout, const float *q_host, const uint8_t *mask_host,
                         double *plan_host, int32_t *choice_host, void *stream);

/* writes dge_state_view.pending: 1 for the envs that need a decision right now (action queue empty, episode
 * running, no forced reset steps outstanding) -- the selection the acting loop of policy.py:236-306 makes,
 * evaluated on the device; pass it as mask_dev to dge_graph.                                            */
int dge_mark_pending(dge_handle h, void *stream);

/* ---- line planner: replaces EMPlanner2D.line_planner (Planner2D.cpp:937-1041) for one
 * goal per env.  goal_dev [B,2]; plan_dev [B,6] = (n_rot_pi, rot_sign, rot_remainder,
 * n_fwd_full, fwd_remainder, n_actions); action i < n_actions is: i < n_rot_pi: (0,0,sign*pi); i == n_rot_pi:
 * (0,0,sign*rot_remainder); then n_fwd_full x (max_edge_length,0,0) and one (fwd_remainder,0,0).  mask nullable;
 * mask value 2 = "this env has no frontier left" (ExplorationEnv.frontier would crash, quirk q15): the episode
 * is declared over (done flag set, empty plan).                                       */
int dge_line_plan(dge_handle h, const double *goal_dev, const uint8_t *mask_dev, double *plan_dev, void *stream);

/* host-buffer variant (EMExplorer.line_plan, pyplanner2d.py:76-78): goal_host [B,2], mask_host [B] nullable, plan_host [B,6]. */
int dge_line_plan_host(dge_handle h, const double *goal_host, const uint8_t *mask_host, double *plan_host, void *stream);

/* ---- policy head on device: per selected graph, argmax of q over its last fro_size
 * nodes (policy.py:109 / test.py:112), goal = that frontier, line plan into the env's
 * action queue.  q_dev [N_tot] f32 in the node order of `g`.                         */
int dge_select_and_plan(dge_handle h, const dge_graph_out *g, const float *q_dev, const uint8_t *mask_dev,
                        int32_t *choice_dev /* [B] nullable: chosen frontier index */, void *stream);

/* ---- one tick of the acting loop on the device: the body of the reference's test loop for B envs at once (test.py:100-143:
 * graph_matrix -> data_process -> model(data, 0) -> np.argmax over the last fro_size nodes -> actions_all_goals -> env.step per
 * action, exploration_env.py:98-105; an episode that ends is re-created like exploration_env.py:389-422).  Per call: every env
 * whose action list ran empty gets its graph built, scored by the GCN Q-network (Networks.GCN, Networks.py:12-28, prob = 0) and
 * a new line plan queued; every env with a queued action executes it (one simulator step); finished episodes restart
 * (dge_reset_done_queued).  No host synchronisation: the size of the decision batch stays on the device (launches are sized by
 * capacity).  `pol` = the network's parameters in the form dge_gcn_q_forward takes them (dge_gnn.h) + workspace:
 *   ws >= 3 * node_cap * C floats (16-byte aligned), q [node_cap] Q-values of the batch's nodes, choice [B] nullable.
 * flags: DGE_TICK_GRAPH       the launch sequence is captured into a CUDA graph at the first call (again whenever an argument
 *                             changes) and replayed with one cudaGraphLaunch on `stream`;
 *        DGE_TICK_ONE_STREAM  step and policy pipelines on `stream` one after the other (default: the step pipeline runs on an
 *                             engine-owned second stream, forked and joined with events -- inside the captured graph too).   */
typedef struct dge_gcn_policy {
  const float *W1, *b1;             /* conv1: [Cin,C], [C] nullable                          */
  const float *W2t_hi, *W2t_lo;     /* conv2 weight as dge_gemm_prep_weight leaves it [C,C]  */
  const float *b2;                  /* [C] nullable                                          */
  const float *head_w, *head_b_dev; /* Linear(C,1): [C], [1] nullable                        */
  float *ws, *q;
  int32_t *choice;
  int64_t node_cap;
  int32_t Cin, C;
} dge_gcn_policy;
#define DGE_TICK_GRAPH 1
#define DGE_TICK_ONE_STREAM 2
int dge_policy_tick(dge_handle h, const dge_graph_out *g, const dge_gcn_policy *pol, uint64_t seed_stride,
                    const double *forced_odom_host, int n_forced, int flags, void *stream);

/* ---- the same tick driven from the HOST with host buffers: the batched form of test.py:100-143 as the reference runs it -- every
 * observation crosses to the host, every action comes from it -- in ONE native call (the host-side loop body of
 * runner.HostPolicyLoop.tick: what the reference's Python does between `env.step` and `model(data)`, for B envs).  Per call:
 *   host     need / has_action / in_reset from (plans, cursor, phase);
 *   `stream`       envs that need a decision: dge_graph_host_packed_begin (graph kernels + pack) ... dge_graph_host_packed_end
 *                  (the batch in arena_host) -> arena_host H2D into arena_dev (DeepQ.test's data.to(device)) -> dge_gcn_q_forward
 *                  -> Q-values D2H into q_host -> dge_select_plan_host (host arg-max, line plans back in plan_host) -> plans / cursor;
 *   `stream_step`  dge_reset_done_queued + dge_step_host_plans_async for the envs with a queued action or in their reset phase
 *                  (done flags, metrics, occupancy maps D2H), joined at the end of the call.
 * The host state (plans [B,6] in dge_line_plan's compact form, cursor [B], phase [B] = ticks of reset work left) is the caller's
 * and is updated in place; all other host pointers are pinned buffers.  The out fields count this tick's work and traffic. */
typedef struct dge_host_loop {
  double *plans; int64_t *cursor; int64_t *phase;            /* host state, updated in place                                  */
  uint8_t *mask, *done, *need;                               /* [B] pinned: step mask / done flags out / decision mask         */
  double *obs; int64_t obs_bytes;                            /* [B,rows,cols] pinned, nullable: occupancy maps after the step  */
  double *metrics;                                           /* [B,8] pinned                                                   */
  void *arena_host; float *q_host; double *plan_host; int32_t *choice_host;   /* pinned: packed batch, Q [node_cap], plans, choices */
  void *arena_pack, *arena_dev; int64_t arena_cap;           /* device arenas: pack target, the policy's copy                  */
  int64_t prefetch_guess;                                    /* in/out: bytes of the batch sent to the host with its header (0: none) */
  int64_t n_stepped, n_graphs, n_nodes, h2d_bytes, d2h_bytes, launches;        /* out: this tick                                */
} dge_host_loop;
int dge_host_policy_tick(dge_handle h, const dge_graph_out *g, const dge_gcn_policy *pol, dge_host_loop *hl, uint64_t seed_stride,
                         const double *forced_odom_host, int n_forced, void *stream, void *stream_step);

/* ---- look-ahead roll-out rewards: replaces EMPlanner2D.simulations_reward
 * (Planner2D.cpp:1416-1468, `planner2d` binding Planner2D.cpp:90) and
 * ExplorationEnv.rewards_all_goals (exploration_env.py:145-162).  `dst` is a second engine with the same
 * map / landmark count and at least the pose capacity of `src` (give it max_poses + the longest line plan, so
 * that a clone of an env near its capacity still runs its whole plan) whose slots receive one clone per (env, frontier) of the envs selected by
 * mask in the graph `g` of `src` (SLAM2D/VirtualMap/Simulator2D copies + set_copy_isam + queued
 * line plan).  The caller then advances `dst` with dge_step_queued until every queue is empty
 * (at most 2 + floor(diagonal / max_edge_length) + 1 calls) and collects the rewards.
 *   totals_dev [2]: number of clones, overflow flag (more clones than dst has slots)
 *   raw_dev / norm_dev [B,Fmax]: U(before) - U(after) per frontier / min-max normalised to
 *   [-1,0] or [-1,1];  loop_clo_dev [B]: the reference's loop-closure flag.            */
int dge_rollout_prepare(dge_handle dst, dge_handle src, const dge_graph_out *g, const uint8_t *mask_dev, int32_t *totals_dev, void *stream);
int dge_rollout_rewards(dge_handle dst, dge_handle src, const dge_graph_out *g, const uint8_t *mask_dev, double *raw_dev, double *norm_dev,
                        uint8_t *loop_clo_dev, void *stream);

/* ---- GNN building blocks (scripts/Networks.py GCN: GCNConv(improved=True) aggregate,
 * bias, ReLU; Linear head) -- hand-written edge/dst-parallel kernels, see dge_gnn.h  */

/* ---- the reference's random streams for `test=True` worlds (host code, csrc/dge_refworld.cu).
 * ExplorationEnv(map_size, env_index, True) of the reference draws its landmarks, control noise and sensor noise from three
 * std::mt19937(env_index) (pyss2d.py:89-119, RNG.h:47-126, Simulator2D.cpp:161-173,436-463,505-527) and visits the landmarks in
 * the iteration order of an std::unordered_map (Simulation2D.h:269).  dge_refworld_* replays those streams on the host and hands
 * out what dge_reset / dge_step take as explicit inputs: landmarks [Lt,2], scan order [Lt], noise rows [3 + 4 Lt].
 * start [3] = the start pose (pyss2d.py:88-95).  One object per episode.                                              */
typedef struct dge_refworld dge_refworld;
dge_refworld *dge_refworld_create(const dge_config *cfg, uint32_t seed, const double *start);
void dge_refworld_destroy(dge_refworld *w);
int dge_refworld_world(const dge_refworld *w, double *landmarks, int32_t *scan, double *init_noise);
int dge_refworld_step(dge_refworld *w, const double *odom, double *noise);

#ifdef __cplusplus
}
#endif
#endif /* DGE_H_ */
