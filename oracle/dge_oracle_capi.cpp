// ============================================================================
// dge_oracle_capi.cpp -- flat C API over the CPU ORACLE for ctypes.
// TEST INFRASTRUCTURE ONLY (see dge_oracle.hpp).  Loaded by oracle/oracle.py.
// ============================================================================
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "dge_oracle.hpp"

using orc::Env;

extern "C" {

void *orc_create(const orc::Config *cfg) { return new Env(*cfg); }
void orc_destroy(void *h) { delete static_cast<Env *>(h); }
void *orc_clone(void *h) { return new Env(*static_cast<Env *>(h)); }
void orc_set_knife(double dx, double dy) { orc::g_knife_dx = dx; orc::g_knife_dy = dy; }   // analysis knob, see dge_oracle.cpp
void orc_set_dense(void *h, int dense) { static_cast<Env *>(h)->use_dense_solver = dense != 0; }

int orc_init(void *h, uint32_t seed, double sx, double sy, double sth, double *noise_rec) {
  Env *e = static_cast<Env *>(h);
  orc::StepNoise rec;
  try { e->init(seed, orc::Pose{sx, sy, sth}, noise_rec ? &rec : nullptr); } catch (...) { return 1; }
  if (noise_rec) std::memcpy(noise_rec, rec.v.data(), sizeof(double) * rec.v.size());
  return 0;
}

int orc_init_landmarks(void *h, uint32_t seed, double sx, double sy, double sth, const double *xy, int L, double *noise_rec) {
  Env *e = static_cast<Env *>(h);
  orc::StepNoise rec;
  std::vector<double> v(xy, xy + 2 * L);
  try { e->init_with_landmarks(seed, orc::Pose{sx, sy, sth}, v, noise_rec ? &rec : nullptr); } catch (...) { return 1; }
  if (noise_rec) std::memcpy(noise_rec, rec.v.data(), sizeof(double) * rec.v.size());
  return 0;
}

// returns 0 ok, 1 solver failure, 2 bounds check fired (no step taken)
int orc_step(void *h, const double *odom, double *noise_rec) {
  Env *e = static_cast<Env *>(h);
  orc::StepNoise rec;
  const int t0 = e->T;
  try { e->step(odom, noise_rec ? &rec : nullptr); } catch (...) { return 1; }
  if (noise_rec && !rec.v.empty()) std::memcpy(noise_rec, rec.v.data(), sizeof(double) * rec.v.size());
  return e->T == t0 ? 2 : 0;
}

// out[0..8) = T, Lt, n_observed, rows, cols, M, sim_step, update_count
void orc_sizes(void *h, int32_t *out) {
  Env *e = static_cast<Env *>(h);
  out[0] = e->T; out[1] = e->Lt; out[2] = e->n_observed(); out[3] = e->rows; out[4] = e->cols;
  out[5] = static_cast<int32_t>(e->meas.size()); out[6] = e->sim_step; out[7] = e->update_count;
}

void orc_get_poses(void *h, double *est, double *lin, double *delta, double *cov, double *info, double *true_pose) {
  Env *e = static_cast<Env *>(h);
  for (int k = 0; k < e->T; ++k) {
    if (est) { est[3 * k] = e->est_pose[k].x; est[3 * k + 1] = e->est_pose[k].y; est[3 * k + 2] = e->est_pose[k].th; }
    if (lin) { lin[3 * k] = e->lin_pose[k].x; lin[3 * k + 1] = e->lin_pose[k].y; lin[3 * k + 2] = e->lin_pose[k].th; }
    if (cov) std::memcpy(cov + 9 * k, e->pose_cov[k].a, 72);
    if (info) std::memcpy(info + 9 * k, e->pose_info[k].a, 72);
  }
  if (delta) std::memcpy(delta, e->delta_pose.data(), sizeof(double) * 3 * e->T);
  if (true_pose) { true_pose[0] = e->true_pose.x; true_pose[1] = e->true_pose.y; true_pose[2] = e->true_pose.th; }
}

void orc_get_landmarks(void *h, uint8_t *observed, double *est, double *lin, double *cov, double *info, double *true_xy, uint32_t *scan_id) {
  Env *e = static_cast<Env *>(h);
  for (int j = 0; j < e->Lt; ++j) {
    if (observed) observed[j] = e->observed[j];
    if (est) { est[2 * j] = e->est_l[2 * j]; est[2 * j + 1] = e->est_l[2 * j + 1]; }
    if (lin) { lin[2 * j] = e->lin_l[2 * j]; lin[2 * j + 1] = e->lin_l[2 * j + 1]; }
    if (cov) std::memcpy(cov + 4 * j, e->land_cov[j].a, 32);
    if (info) std::memcpy(info + 4 * j, e->land_info[j].a, 32);
    if (true_xy) { true_xy[2 * j] = e->lm_x[j]; true_xy[2 * j + 1] = e->lm_y[j]; }
    if (scan_id) scan_id[j] = e->scan_id[j];
  }
}

void orc_get_factors(void *h, double *odom, int32_t *meas_ptr, int32_t *meas_id, double *meas_b, double *meas_r) {
  Env *e = static_cast<Env *>(h);
  if (odom) std::memcpy(odom, e->odom.data(), sizeof(double) * e->odom.size());
  if (meas_ptr) std::memcpy(meas_ptr, e->meas_ptr.data(), sizeof(int32_t) * e->meas_ptr.size());
  for (size_t p = 0; p < e->meas.size(); ++p) {
    if (meas_id) meas_id[p] = e->meas[p].id;
    if (meas_b) meas_b[p] = e->meas[p].bearing;
    if (meas_r) meas_r[p] = e->meas[p].range;
  }
}

void orc_get_vmap(void *h, double *prob, double *vinfo, int32_t *seen, double *trace) {
  Env *e = static_cast<Env *>(h);
  const int V = e->rows * e->cols;
  if (prob) std::memcpy(prob, e->prob.data(), sizeof(double) * V);
  if (vinfo) std::memcpy(vinfo, e->vinfo.data(), sizeof(double) * 4 * V);
  if (seen) std::memcpy(seen, e->seen_count.data(), sizeof(int32_t) * V);
  if (trace) { std::vector<double> t; e->cov_trace(t); std::memcpy(trace, t.data(), sizeof(double) * V); }
}

// out[0..6) = explored, utility(0), landmark_error, max_traj_uncertainty, done, dist
void orc_metrics(void *h, double *out) {
  Env *e = static_cast<Env *>(h);
  out[0] = e->explored(); out[1] = e->utility(0.0); out[2] = e->landmark_error(1.0);
  out[3] = e->max_traj_uncertainty(); out[4] = e->done() ? 1.0 : 0.0; out[5] = e->dist;
}

double orc_utility(void *h, double distance) { return static_cast<Env *>(h)->utility(distance); }

// two-phase graph export: sizes[0..6) = N, K, L, F, E, n_all_frontier_cells
static thread_local orc::GraphOut g_graph;
int orc_graph_build(void *h, int32_t *sizes) {
  Env *e = static_cast<Env *>(h);
  e->graph(g_graph);
  sizes[0] = g_graph.n_nodes; sizes[1] = g_graph.key_size; sizes[2] = g_graph.land_size; sizes[3] = g_graph.fro_size;
  sizes[4] = static_cast<int32_t>(g_graph.edge_src.size()); sizes[5] = static_cast<int32_t>(g_graph.all_frontier_cells.size());
  return 0;
}
void orc_graph_fetch(double *features, int64_t *edge_index /*[2,E]*/, double *edge_w, double *frontier_xy, int32_t *all_cells) {
  const size_t E = g_graph.edge_src.size();
  if (features) std::memcpy(features, g_graph.features.data(), sizeof(double) * g_graph.features.size());
  if (edge_index) { std::memcpy(edge_index, g_graph.edge_src.data(), 8 * E); std::memcpy(edge_index + E, g_graph.edge_dst.data(), 8 * E); }
  if (edge_w) std::memcpy(edge_w, g_graph.edge_w.data(), 8 * E);
  if (frontier_xy) std::memcpy(frontier_xy, g_graph.frontier_xy.data(), 8 * g_graph.frontier_xy.size());
  if (all_cells) std::memcpy(all_cells, g_graph.all_frontier_cells.data(), 4 * g_graph.all_frontier_cells.size());
}

int orc_line_plan(void *h, double gx, double gy, double *out, int max_actions) {
  const std::vector<orc::Pose> a = static_cast<Env *>(h)->line_plan(gx, gy);
  const int n = static_cast<int>(a.size());
  for (int i = 0; i < n && i < max_actions; ++i) { out[3 * i] = a[i].x; out[3 * i + 1] = a[i].y; out[3 * i + 2] = a[i].th; }
  return n;
}

double orc_sim_reward(void *h, const double *actions, int n) {
  std::vector<orc::Pose> a(n);
  for (int i = 0; i < n; ++i) a[i] = orc::Pose{actions[3 * i], actions[3 * i + 1], actions[3 * i + 2]};
  try { return static_cast<Env *>(h)->simulations_reward(a); } catch (...) { return std::nan(""); }
}
// same with explicit noise [n, 3+4*Lt] (kernel parity tests)
double orc_sim_reward_noise(void *h, const double *actions, int n, const double *noise) {
  std::vector<orc::Pose> a(n);
  for (int i = 0; i < n; ++i) a[i] = orc::Pose{actions[3 * i], actions[3 * i + 1], actions[3 * i + 2]};
  try { return static_cast<Env *>(h)->simulations_reward(a, noise); } catch (...) { return std::nan(""); }
}

void orc_virtual_map_rebuild(const orc::Config *cfg, int T, const double *pose, const double *info, int L, const double *lm,
                             int rows, int cols, double *prob, double *vinfo, int32_t *seen) {
  orc::virtual_map_rebuild(*cfg, T, pose, info, L, lm, rows, cols, prob, vinfo, seen);
}

// batched virtual-map rebuild over `n` independent problems on `threads` host threads
// (CPU baseline leg of the C4 sweep).  All arrays are [n, ...] contiguous.
void orc_virtual_map_rebuild_batch(const orc::Config *cfg, int n, int threads, int T, const double *pose, const double *info, int L,
                                   const double *lm, int rows, int cols, double *prob, double *vinfo) {
  const int V = rows * cols;
  auto work = [&](int t) {
    for (int i = t; i < n; i += threads)
      orc::virtual_map_rebuild(*cfg, T, pose + static_cast<size_t>(i) * 3 * T, info + static_cast<size_t>(i) * 9 * T, L,
                               lm + static_cast<size_t>(i) * 2 * L, rows, cols, prob + static_cast<size_t>(i) * V,
                               vinfo + static_cast<size_t>(i) * 4 * V, nullptr);
  };
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
  for (auto &t : th) t.join();
}

int orc_sizeof_config() { return static_cast<int>(sizeof(orc::Config)); }

}  // extern "C"


namespace {
// persistent worker pool (libgomp is not available in this image)
class Pool {
 public:
  explicit Pool(int n) { for (int i = 0; i < n; ++i) th_.emplace_back([this] { work(); }); }
  ~Pool() { { std::lock_guard<std::mutex> l(m_); stop_ = true; ++gen_; } cv_.notify_all(); for (auto &t : th_) t.join(); }
  void run(int n, const std::function<void(int)> &f) {
    std::unique_lock<std::mutex> l(m_);
    fn_ = &f; n_ = n; next_.store(0); pending_ = static_cast<int>(th_.size()); ++gen_;
    cv_.notify_all();
    done_.wait(l, [this] { return pending_ == 0; });
  }
  int size() const { return static_cast<int>(th_.size()); }
 private:
  void work() {
    int seen = 0;
    for (;;) {
      { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return gen_ != seen; }); seen = gen_; if (stop_) return; }
      for (int i; (i = next_.fetch_add(1)) < n_;) (*fn_)(i);
      { std::lock_guard<std::mutex> l(m_); if (--pending_ == 0) done_.notify_one(); }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)> *fn_ = nullptr;
  std::atomic<int> next_{0};
  int n_ = 0, gen_ = 0, pending_ = 0;
  bool stop_ = false;
};
Pool &pool(int threads) {
  static Pool *p = nullptr;
  if (!p || p->size() != threads) { delete p; p = new Pool(threads); }
  return *p;
}
}  // namespace
// ---------------------------------------------------------------------------------------
// Batched drivers for the CPU baseline (bench.py cpu_baseline / --impl reference): each env is
// stepped single-threaded exactly like the reference; independent envs are spread over host
// threads with a persistent worker pool so that Python is off the per-env path.
extern "C" {

// odoms [n,3]; out_done [n] (1 = episode finished), out_T [n] trajectory length
void orc_batch_step(void **hs, const double *odoms, int n, int threads, uint8_t *out_done, int32_t *out_T) {
  pool(threads).run(n, [&](int i) {
    Env *e = static_cast<Env *>(hs[i]);
    try { e->step(odoms + 3 * i, nullptr); } catch (...) { out_done[i] = 1; out_T[i] = e->T; return; }
    out_done[i] = e->done() ? 1 : 0;
    out_T[i] = e->T;
  });
}

static std::vector<orc::GraphOut> g_batch_graphs;
// builds the graphs of n envs in parallel; sizes [n,4] = N, K, F, E per env
void orc_batch_graph_build(void **hs, int n, int threads, int32_t *sizes) {
  g_batch_graphs.resize(n);
  pool(threads).run(n, [&](int i) {
    static_cast<Env *>(hs[i])->graph(g_batch_graphs[i]);
    sizes[4 * i] = g_batch_graphs[i].n_nodes; sizes[4 * i + 1] = g_batch_graphs[i].key_size;
    sizes[4 * i + 2] = g_batch_graphs[i].fro_size; sizes[4 * i + 3] = static_cast<int32_t>(g_batch_graphs[i].edge_src.size());
  });
}
// concatenated (PyG DataLoader layout): x [Ntot,5] f32, edge_index [2,Etot] i64 (offset), edge_attr [Etot] f32
void orc_batch_graph_fetch(int n, float *x, int64_t *edge_index, float *edge_attr, int64_t Etot) {
  int64_t noff = 0, eoff = 0;
  for (int i = 0; i < n; ++i) {
    const orc::GraphOut &g = g_batch_graphs[i];
    for (size_t k = 0; k < g.features.size(); ++k) x[noff * 5 + k] = static_cast<float>(g.features[k]);
    for (size_t k = 0; k < g.edge_src.size(); ++k) {
      edge_index[eoff + k] = g.edge_src[k] + noff; edge_index[Etot + eoff + k] = g.edge_dst[k] + noff;
      edge_attr[eoff + k] = static_cast<float>(g.edge_w[k]);
    }
    noff += g.n_nodes; eoff += static_cast<int64_t>(g.edge_src.size());
  }
}
// line plans towards frontier `choice[i]` of the graphs built by orc_batch_graph_build:
// out [n,max_actions,3], counts [n]
void orc_batch_line_plan(void **hs, int n, const int32_t *choice, int max_actions, double *out, int32_t *counts) {
  for (int i = 0; i < n; ++i) {
    const orc::GraphOut &g = g_batch_graphs[i];
    if (g.fro_size <= 0 || choice[i] < 0) { counts[i] = 0; continue; }
    const std::vector<orc::Pose> a = static_cast<Env *>(hs[i])->line_plan(g.frontier_xy[2 * choice[i]], g.frontier_xy[2 * choice[i] + 1]);
    counts[i] = static_cast<int32_t>(std::min<size_t>(a.size(), max_actions));
    for (int k = 0; k < counts[i]; ++k) { out[(static_cast<size_t>(i) * max_actions + k) * 3] = a[k].x; out[(static_cast<size_t>(i) * max_actions + k) * 3 + 1] = a[k].y; out[(static_cast<size_t>(i) * max_actions + k) * 3 + 2] = a[k].th; }
  }
}
// fresh envs (SS2D.__init__ + the 4 forced steps of ExplorationEnv.reset) in parallel; starts [n,3]
void orc_batch_fresh(void **hs, const uint32_t *seeds, const double *starts, int n, int threads) {
  const double o[3] = {1.0, 1.0, 1.57079632679489661923};
  pool(threads).run(n, [&](int i) {
    Env *e = static_cast<Env *>(hs[i]);
    try {
      e->init(seeds[i], orc::Pose{starts[3 * i], starts[3 * i + 1], starts[3 * i + 2]}, nullptr);
      for (int k = 0; k < 4; ++k) e->step(o, nullptr);
    } catch (...) {}
  });
}

}  // extern "C"
