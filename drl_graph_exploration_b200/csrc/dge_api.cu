// C-ABI of libdge.so (see include/dge.h).  Owns the engine's HBM state; every launch goes to
// the caller's stream so the Python side can order it with torch work or capture it in a
// CUDA graph.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "dge_internal.cuh"

static thread_local std::string g_err;
static int fail(int code, const char *what) {
  cudaError_t ce = cudaGetLastError();
  g_err = std::string(what) + (ce != cudaSuccess ? std::string(": ") + cudaGetErrorString(ce) : std::string());
  return code;
}

int dge_fail(int code, const char *what) { return fail(code, what); }

extern "C" const char *dge_last_error(void) { return g_err.c_str(); }

namespace {
struct Alloc {
  std::vector<void *> ptrs;
  bool ok = true;
  template <typename T>
  T *get(size_t n) {
    void *p = nullptr;
    if (!ok) return nullptr;
    if (cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) { ok = false; return nullptr; }
    cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T));
    ptrs.push_back(p);
    return static_cast<T *>(p);
  }
};
struct EngineBox {
  dge_engine e;
  Alloc al;
};
}  // namespace

extern "C" int dge_create(const dge_config *cfg, int n_envs, int max_poses, int device, dge_handle *out) {
  if (!cfg || !out || n_envs <= 0 || max_poses < 8) return fail(DGE_EINVAL, "dge_create: bad arguments");
  if (cfg->num_landmarks < 1 || 2 * cfg->num_landmarks > 128) return fail(DGE_EINVAL, "dge_create: num_landmarks must be in [1,64]");
  if (cudaSetDevice(device) != cudaSuccess) return fail(DGE_ECUDA, "cudaSetDevice");
  EngineBox *bx = new EngineBox();
  dge_engine &e = bx->e;
  memset(&e, 0, sizeof(e));
  e.cfg = *cfg;
  e.device = device;
  DgeDims &d = e.d;
  d.B = n_envs; d.Tmax = max_poses; d.Lt = cfg->num_landmarks;
  d.cols = (int)floor((cfg->map_max_x - cfg->map_min_x) / cfg->resolution);   // VirtualMap.cpp:319-322
  d.rows = (int)floor((cfg->map_max_y - cfg->map_min_y) / cfg->resolution);
  d.V = d.rows * d.cols;
  d.Mmax = d.Tmax * d.Lt;
  d.Fmax = d.Lt + 1;
  d.Ncap = d.Lt + d.Tmax + d.Fmax;
  d.Ecap = 2 * (d.Mmax + d.Tmax + d.Fmax + d.Lt);
  if (dge_slam_smem_bytes(d.Lt) > 227 * 1024) { delete bx; return fail(DGE_EINVAL, "dge_create: too many landmarks for the SLAM kernel's shared memory"); }
  Alloc &al = bx->al;
  const size_t B = d.B, T = d.Tmax, L = d.Lt, V = d.V, M = d.Mmax;
  e.true_pose = al.get<double>(B * 3); e.lm_true = al.get<double>(B * L * 2); e.scan_id = al.get<int32_t>(B * L); e.seed = al.get<uint64_t>(B);
  e.n_poses = al.get<int32_t>(B); e.sim_step = al.get<int32_t>(B); e.update_count = al.get<int32_t>(B); e.status = al.get<int32_t>(B);
  e.prior_pose = al.get<double>(B * 3);
  e.lin_pose = al.get<double>(B * T * 3); e.est_pose = al.get<double>(B * T * 3); e.delta_pose = al.get<double>(B * T * 3); e.odom = al.get<double>(B * T * 3);
  e.pose_cov = al.get<double>(B * T * 6); e.pose_info = al.get<double>(B * T * 6);
  e.meas_ptr = al.get<int32_t>(B * (T + 1)); e.meas_id = al.get<int32_t>(B * M); e.meas_pose = al.get<int32_t>(B * M); e.meas_b = al.get<double>(B * M); e.meas_r = al.get<double>(B * M);
  e.observed = al.get<uint8_t>(B * L); e.lin_l = al.get<double>(B * L * 2); e.est_l = al.get<double>(B * L * 2); e.delta_l = al.get<double>(B * L * 2);
  e.land_cov = al.get<double>(B * L * 3);
  e.ws_pose = al.get<double>(B * T * DGE_WS_POSE); e.ws_meas = al.get<double>(B * M * 14);
  e.ws_Bt = al.get<double>(B * T * 3 * 2 * L); e.ws_FB = al.get<double>(B * T * 3 * 2 * L);
  e.ws_midx = al.get<int32_t>(B * T * L);
  e.lm_slot = al.get<int32_t>(B * L); e.fc_valid = al.get<int32_t>(B); e.step_order = al.get<int32_t>(B); e.fc_state = al.get<double>(B * (size_t)DGE_FC_WIDTH(L));
  e.ck_state = al.get<double>(B * (size_t)DGE_CK_DEPTH * DGE_FC_WIDTH(L)); e.ck_pos = al.get<int32_t>(B * (size_t)DGE_CK_STRIDE); e.lm_first = al.get<int32_t>(B * L);
  e.vm_prep = al.get<double>(B * T * dge_vmap_prep_width()); e.vm_cbox = al.get<double>(B * (size_t)dge_vmap_nchunk(d.Tmax) * 4);
  e.seen = al.get<int32_t>(B * V); e.active = al.get<uint8_t>(B);
  e.prob = al.get<double>(B * V); e.vinfo = al.get<double>(B * V * 3); e.metrics = al.get<double>(B * 8); e.dist = al.get<double>(B);
  e.done = al.get<uint8_t>(B);
  e.plan = al.get<double>(B * 6); e.plan_cursor = al.get<int32_t>(B);
  e.odom_dev_scratch = al.get<double>(B * 3); e.mask_dev_scratch = al.get<uint8_t>(B);
  e.mask_dev_scratch2 = al.get<uint8_t>(B); e.goal_dev_scratch = al.get<double>(B * 2); e.plan_dev_scratch = al.get<double>(B * 6);
  e.g_counts = al.get<int32_t>(B * 4); e.g_frontier = al.get<int32_t>(B * d.Fmax); e.g_fassoc = al.get<int32_t>(B * (L + 1)); e.g_sel = al.get<int32_t>(B);
  e.g_cnt = al.get<int32_t>(B * (size_t)d.Ncap); e.g_cur = al.get<int32_t>(B * (size_t)d.Ncap); e.g_dis = al.get<float>(B * (size_t)d.Ncap);
  e.g_tmp = al.get<int32_t>(B * (size_t)d.Ecap);
  e.slam_clocks = al.get<long long>(B * 12);
  e.forced = al.get<int32_t>(B); e.step_kind = al.get<uint8_t>(B); e.pending = al.get<uint8_t>(B);
  e.group_heavy = al.get<uint8_t>(B); e.group_light = al.get<uint8_t>(B);
  e.counters = al.get<unsigned long long>(8); e.count_steps = 1; e.park_done = 1;
  e.rdist = al.get<double>(B); e.r_cmap = al.get<int32_t>(B * 2); e.r_cbase = al.get<int32_t>(B); e.r_u0 = al.get<double>(B);
  if (al.ok && cudaMallocHost(reinterpret_cast<void **>(&e.pack_hdr_host), 16 * sizeof(int64_t)) != cudaSuccess) al.ok = false;
  if (al.ok && cudaMallocHost(reinterpret_cast<void **>(&e.hp_odom), B * 3 * sizeof(double)) != cudaSuccess) al.ok = false;
  if (al.ok && cudaMallocHost(reinterpret_cast<void **>(&e.hp_goal), B * 2 * sizeof(double)) != cudaSuccess) al.ok = false;
  if (al.ok && cudaMallocHost(reinterpret_cast<void **>(&e.hp_mask), B) != cudaSuccess) al.ok = false;
  if (!al.ok) {
    for (void *p : al.ptrs) cudaFree(p);
    if (e.pack_hdr_host) cudaFreeHost(e.pack_hdr_host);
    if (e.hp_odom) cudaFreeHost(e.hp_odom);
    if (e.hp_goal) cudaFreeHost(e.hp_goal);
    if (e.hp_mask) cudaFreeHost(e.hp_mask);
    delete bx;
    return fail(DGE_ENOMEM, "dge_create: cudaMalloc failed");
  }
  *out = &bx->e;
  return DGE_OK;
}

extern "C" int dge_destroy(dge_handle h) {
  if (!h) return DGE_EINVAL;
  EngineBox *bx = reinterpret_cast<EngineBox *>(h);   // e is the first member
  cudaSetDevice(h->device);
  dge_tick_release(h);
  for (void *p : bx->al.ptrs) cudaFree(p);
  if (h->pack_hdr_host) cudaFreeHost(h->pack_hdr_host);
  if (h->hp_odom) cudaFreeHost(h->hp_odom);
  if (h->hp_goal) cudaFreeHost(h->hp_goal);
  if (h->hp_mask) cudaFreeHost(h->hp_mask);
  delete bx;
  return DGE_OK;
}

extern "C" int dge_dims(dge_handle h, int32_t *out) {
  if (!h || !out) return DGE_EINVAL;
  const DgeDims &d = h->d;
  out[0] = d.B; out[1] = d.Tmax; out[2] = d.Lt; out[3] = d.rows; out[4] = d.cols; out[5] = d.Mmax; out[6] = d.Ncap; out[7] = d.Ecap;
  return DGE_OK;
}

extern "C" int dge_reset(dge_handle h, const uint8_t *mask, const uint64_t *seeds, const double *start, const double *lm,
                         const int32_t *scan, const double *noise, void *stream) {
  if (!h) return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = dge_launch_reset(h, mask, seeds, start, lm, scan, noise, 0, 0, st);
  if (rc) return fail(rc, "dge_reset: k_reset");
  rc = dge_launch_slam(h, mask, st);   // pyss2d.py:135 self.optimize()
  if (rc) return fail(rc, "dge_reset: k_slam");
  return DGE_OK;
}

extern "C" int dge_reset_queued(dge_handle h, const uint8_t *mask, const uint64_t *seeds, const double *start, const double *lm,
                                const int32_t *scan, const double *forced_odom_host, int n_forced, void *stream) {
  if (!h || !forced_odom_host || n_forced < 1 || n_forced >= DGE_FRESH_BIT) return fail(DGE_EINVAL, "dge_reset_queued: bad arguments");
  for (int i = 0; i < 3; ++i) h->forced_odom[i] = forced_odom_host[i];
  const int rc = dge_launch_reset(h, mask, seeds, start, lm, scan, nullptr, n_forced, 0, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_reset_queued: k_reset") : DGE_OK;
}

extern "C" int dge_reset_done_queued(dge_handle h, uint64_t seed_stride, const double *forced_odom_host, int n_forced, void *stream) {
  if (!h || !forced_odom_host || seed_stride == 0 || n_forced < 1 || n_forced >= DGE_FRESH_BIT) return fail(DGE_EINVAL, "dge_reset_done_queued: bad arguments");
  for (int i = 0; i < 3; ++i) h->forced_odom[i] = forced_odom_host[i];
  const int rc = dge_launch_reset(h, h->done, nullptr, nullptr, nullptr, nullptr, nullptr, n_forced, seed_stride, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_reset_done_queued: k_reset") : DGE_OK;
}

extern "C" int dge_move_measure(dge_handle h, const double *odom, const uint8_t *mask, const double *noise, void *stream) {
  if (!h || !odom) return DGE_EINVAL;
  const int rc = dge_launch_move_measure(h, odom, mask, noise, 0, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_move_measure") : DGE_OK;
}
extern "C" int dge_move_measure_queued(dge_handle h, void *stream) {
  if (!h) return DGE_EINVAL;
  const int rc = dge_launch_move_measure(h, nullptr, nullptr, nullptr, 1, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_move_measure_queued") : DGE_OK;
}
extern "C" int dge_slam_optimize(dge_handle h, const uint8_t *mask, void *stream) {
  if (!h) return DGE_EINVAL;
  const int rc = dge_launch_slam(h, mask, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_slam_optimize") : DGE_OK;
}
extern "C" int dge_virtual_map(dge_handle h, const uint8_t *mask, void *stream) {
  if (!h) return DGE_EINVAL;
  const int rc = dge_launch_vmap(h, mask, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_virtual_map") : DGE_OK;
}

extern "C" int dge_step(dge_handle h, const double *odom, const uint8_t *mask, const double *noise, void *stream) {
  if (!h || !odom) return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = dge_launch_move_measure(h, odom, mask, noise, 0, st);
  if (rc) return fail(rc, "dge_step: move_measure");
  if ((rc = dge_launch_slam(h, h->active, st))) return fail(rc, "dge_step: slam");
  if ((rc = dge_launch_vmap(h, h->active, st))) return fail(rc, "dge_step: vmap");
  return DGE_OK;
}

extern "C" int dge_step_queued_noise(dge_handle h, const double *noise, void *stream) {
  if (!h) return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = dge_launch_move_measure(h, nullptr, nullptr, noise, h->park_done ? 1 : 2, st);
  if (rc) return fail(rc, "dge_step_queued_noise: move_measure");
  if ((rc = dge_launch_slam(h, h->active, st))) return fail(rc, "dge_step_queued_noise: slam");
  if ((rc = dge_launch_vmap(h, h->active, st))) return fail(rc, "dge_step_queued_noise: vmap");
  return DGE_OK;
}

extern "C" int dge_step_queued(dge_handle h, void *stream) {
  if (!h) return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = dge_launch_move_measure(h, nullptr, nullptr, nullptr, 1, st);
  if (rc) return fail(rc, "dge_step_queued: move_measure");
  if ((rc = dge_launch_slam(h, h->active, st))) return fail(rc, "dge_step_queued: slam");
  if ((rc = dge_launch_vmap(h, h->active, st))) return fail(rc, "dge_step_queued: vmap");
  return DGE_OK;
}

extern "C" int dge_step_host(dge_handle h, const double *odom_host, const uint8_t *mask_host, uint8_t *done_host, double *obs_host, void *stream) {
  if (!h || !odom_host) return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t B = h->d.B;
  if (cudaMemcpyAsync(h->odom_dev_scratch, odom_host, B * 3 * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host: H2D odom");
  const uint8_t *mask = nullptr;
  if (mask_host) {
    if (cudaMemcpyAsync(h->mask_dev_scratch, mask_host, B, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host: H2D mask");
    mask = h->mask_dev_scratch;
  }
  const int rc = dge_step(h, h->odom_dev_scratch, mask, nullptr, stream);
  if (rc) return rc;
  if (done_host && cudaMemcpyAsync(done_host, h->done, B, cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host: D2H done");
  if (obs_host && cudaMemcpyAsync(obs_host, h->prob, B * h->d.V * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host: D2H obs");
  if (cudaStreamSynchronize(st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host: sync");
  return DGE_OK;
}

extern "C" int dge_step_host_async(dge_handle h, const double *odom_host, const uint8_t *mask_host, uint8_t *done_host, double *obs_host,
                                   double *metrics_host, int flags, void *stream) {
  if (!h || !odom_host) return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t B = h->d.B;
  if (cudaMemcpyAsync(h->odom_dev_scratch, odom_host, B * 3 * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host_async: H2D odom");
  const uint8_t *mask = nullptr;
  if (mask_host) {
    if (cudaMemcpyAsync(h->mask_dev_scratch, mask_host, B, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host_async: H2D mask");
    mask = h->mask_dev_scratch;
  }
  int rc = dge_launch_move_measure(h, h->odom_dev_scratch, mask, nullptr, (flags & DGE_STEP_HONOR_FORCED) ? 3 : 0, st);
  if (rc) return fail(rc, "dge_step_host_async: move_measure");
  if ((rc = dge_launch_slam(h, h->active, st))) return fail(rc, "dge_step_host_async: slam");
  if ((rc = dge_launch_vmap(h, h->active, st))) return fail(rc, "dge_step_host_async: vmap");
  if (done_host && cudaMemcpyAsync(done_host, h->done, B, cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host_async: D2H done");
  if (metrics_host && cudaMemcpyAsync(metrics_host, h->metrics, B * 8 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host_async: D2H metrics");
  if (obs_host && cudaMemcpyAsync(obs_host, h->prob, B * h->d.V * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host_async: D2H obs");
  if (!(flags & DGE_STEP_NO_SYNC) && cudaStreamSynchronize(st) != cudaSuccess) return fail(DGE_ECUDA, "dge_step_host_async: sync");
  return DGE_OK;
}

extern "C" int dge_graph_host(dge_handle h, const uint8_t *mask_host, const dge_graph_out *dev, const dge_graph_host_out *host, void *stream) {
  if (!h || !dev || !host || !host->x || !host->edge_index || !host->edge_attr || !host->node_ptr || !host->edge_ptr || !host->key_size ||
      !host->fro_size || !host->frontier_xy || !host->totals)
    return fail(DGE_EINVAL, "dge_graph_host: null buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t B = h->d.B;
  const uint8_t *mask = nullptr;
  if (mask_host) {
    if (cudaMemcpyAsync(h->mask_dev_scratch2, mask_host, B, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_graph_host: H2D mask");
    mask = h->mask_dev_scratch2;
  }
  int rc = dge_graph(h, mask, dev, stream);
  if (rc) return rc;
  if (cudaMemcpyAsync(host->totals, dev->totals, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_graph_host: D2H totals");
  if (cudaStreamSynchronize(st) != cudaSuccess) return fail(DGE_ECUDA, "dge_graph_host: sync");
  const size_t G = host->totals[0], N = host->totals[1], E = host->totals[2];
  if (host->totals[3]) return fail(DGE_ECAP, "dge_graph_host: graph batch capacity exceeded");
  if (G == 0) return DGE_OK;
  bool ok = true;
  auto d2h = [&](void *dst, const void *src, size_t bytes) { if (bytes) ok = ok && cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) == cudaSuccess; };
  d2h(host->x, dev->x, N * 5 * sizeof(float));
  d2h(host->edge_index, dev->edge_index, E * sizeof(int64_t));                    // row 0 (sources)
  d2h(host->edge_index + E, dev->edge_index + dev->edge_cap, E * sizeof(int64_t)); // row 1 (destinations): [2,E] contiguous on the host
  d2h(host->edge_attr, dev->edge_attr, E * sizeof(float));
  d2h(host->node_ptr, dev->node_ptr, (G + 1) * sizeof(int32_t));
  d2h(host->edge_ptr, dev->edge_ptr, (G + 1) * sizeof(int32_t));
  d2h(host->key_size, dev->key_size, G * sizeof(int32_t));
  d2h(host->fro_size, dev->fro_size, G * sizeof(int32_t));
  d2h(host->frontier_xy, dev->frontier_xy, B * (size_t)h->d.Fmax * 2 * sizeof(double));
  if (host->csr_rowptr && host->csr_perm && host->gcn_norm && host->gcn_selfnorm && dev->csr_rowptr) {
    d2h(host->csr_rowptr, dev->csr_rowptr, (N + 1) * sizeof(int32_t));
    d2h(host->csr_perm, dev->csr_perm, E * sizeof(int32_t));
    d2h(host->gcn_norm, dev->gcn_norm, E * sizeof(float));
    d2h(host->gcn_selfnorm, dev->gcn_selfnorm, N * sizeof(float));
  }
  if (!ok) return fail(DGE_ECUDA, "dge_graph_host: D2H graph");
  if (cudaStreamSynchronize(st) != cudaSuccess) return fail(DGE_ECUDA, "dge_graph_host: sync");
  return DGE_OK;
}

// ---- host-held action lists: "for act in actions: env.step(act)" (test.py:119-120, policy.py:117-118) for B envs.  The caller keeps
// every env's line plan in the compact form of dge_line_plan and a cursor; action `cursor` is expanded here (Planner2D.cpp:982-1038).
extern "C" int dge_step_host_plans_async(dge_handle h, const double *plans_host, const int64_t *cursor_host, const uint8_t *mask_host,
                                         uint8_t *done_host, double *obs_host, double *metrics_host, int flags, void *stream) {
  if (!h || !plans_host || !cursor_host) return DGE_EINVAL;
  const int B = h->d.B;
  const double pi = DGE_PI, edge = h->cfg.max_edge_length;
  for (int b = 0; b < B; ++b) {
    const double *pl = plans_host + 6 * (size_t)b;
    const int64_t cur = cursor_host[b], nrot = (int64_t)pl[0], nfwd = (int64_t)pl[3];
    double *od = h->hp_odom + 3 * (size_t)b;
    od[0] = od[1] = od[2] = 0.0;
    if (cur < nrot) od[2] = pl[1] * pi;
    else if (cur == nrot) od[2] = pl[1] * pl[2];
    else od[0] = (cur < nrot + 1 + nfwd) ? edge : pl[4];
  }
  return dge_step_host_async(h, h->hp_odom, mask_host, done_host, obs_host, metrics_host, flags, stream);
}

// ---- host-side policy read-out on a packed batch: np.argmax(readout_t[-fro_size:]) (test.py:112, policy.py:109) per graph, the chosen
// frontier's coordinates as goal, line plan from the env's current estimate (actions_all_goals()[key_size + action_index]).
extern "C" int dge_select_plan_host(dge_handle h, const void *arena_host, const dge_graph_packed *lay, const float *q_host, const uint8_t *mask_host,
                                    double *plan_host, int32_t *choice_host, void *stream) {
  if (!h || !arena_host || !lay || !q_host || !mask_host || !plan_host) return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = h->d.B, Fmax = h->d.Fmax;
  const unsigned char *ar = static_cast<const unsigned char *>(arena_host);
  const int32_t *nptr = reinterpret_cast<const int32_t *>(ar + lay->node_ptr), *ks = reinterpret_cast<const int32_t *>(ar + lay->key_size);
  const int32_t *fs = reinterpret_cast<const int32_t *>(ar + lay->fro_size);
  const double *fxy = reinterpret_cast<const double *>(ar + lay->frontier_xy);
  if (lay->frontier_plan > 0) {
    // the batch carries the line plan of every frontier (actions_all_goals): the host picks actions[key_size + action_index] (test.py:113)
    // itself -- no device round trip.  An env without a frontier (q15) needs its done flag set on the device: the general path below.
    bool simple = true;
    int g0 = 0;
    for (int b = 0; b < B && simple; ++b) {
      if (!mask_host[b]) continue;
      if (g0 >= lay->n_graphs) return fail(DGE_EINVAL, "dge_select_plan_host: mask selects more envs than the batch holds");
      if (fs[g0] <= 0) simple = false;
      ++g0;
    }
    if (simple) {
      const double *fpl = reinterpret_cast<const double *>(ar + lay->frontier_plan);
      int g = 0;
      for (int b = 0; b < B; ++b) {
        if (choice_host) choice_host[b] = -1;
        if (!mask_host[b]) continue;
        const int F = fs[g], n0 = nptr[g] + ks[g];
        int best = 0;
        for (int f = 1; f < F; ++f) if (q_host[n0 + f] > q_host[n0 + best]) best = f;   // first maximum, like np.argmax
        for (int i = 0; i < 6; ++i) plan_host[6 * b + i] = fpl[((size_t)g * Fmax + best) * 6 + i];
        if (choice_host) choice_host[b] = best;
        ++g;
      }
      return DGE_OK;
    }
  }
  int g = 0;
  for (int b = 0; b < B; ++b) {
    h->hp_mask[b] = 0;
    if (choice_host) choice_host[b] = -1;
    if (!mask_host[b]) continue;
    if (g >= lay->n_graphs) return fail(DGE_EINVAL, "dge_select_plan_host: mask selects more envs than the batch holds");
    const int F = fs[g], n0 = nptr[g] + ks[g];
    if (F <= 0) { h->hp_mask[b] = 2; ++g; continue; }       // no frontier left (q15): the episode is over
    int best = 0;
    for (int f = 1; f < F; ++f) if (q_host[n0 + f] > q_host[n0 + best]) best = f;   // first maximum, like np.argmax
    h->hp_goal[2 * b] = fxy[((size_t)g * Fmax + best) * 2]; h->hp_goal[2 * b + 1] = fxy[((size_t)g * Fmax + best) * 2 + 1];
    h->hp_mask[b] = 1;
    if (choice_host) choice_host[b] = best;
    ++g;
  }
  return dge_line_plan_host(h, h->hp_goal, h->hp_mask, plan_host, stream);
}

// ---- packed host transfer of a graph batch ---------------------------------------------------------------------
// One arena holds the valid prefix of every array of the batch behind a 128-byte header, so that the batch crosses the
// bus in ONE copy per direction (13 small D2H copies + 7 H2D copies otherwise) and the host / device views are plain
// offsets into the same bytes.  Section order and alignment (16 B) are ABI: see dge_graph_packed in dge.h.
namespace {
__host__ __device__ inline int64_t pk_align(int64_t v) { return (v + 15) & ~int64_t(15); }
constexpr int PK_SECTIONS = 13;
struct PackLayout { int64_t off[PK_SECTIONS]; int64_t total; };
__host__ __device__ inline PackLayout pack_layout(int64_t G, int64_t N, int64_t E, int64_t Fmax) {
  PackLayout L;
  int64_t o = 128;
  // (the small per-graph sections the host reads for its decision come last but one; the last one holds the line plan of EVERY frontier:
  //  ExplorationEnv.actions_all_goals, exploration_env.py:131-143 -- the host then picks actions[key_size + action_index] like test.py:113)
  const int64_t sz[PK_SECTIONS] = {N * 20, E * 16, E * 4, (G + 1) * 4, (G + 1) * 4, G * 4, G * 4, G * Fmax * 16, (N + 1) * 4, E * 4, E * 4, N * 4, G * Fmax * 48};
  for (int i = 0; i < PK_SECTIONS; ++i) { L.off[i] = o; o = pk_align(o + sz[i]); }
  L.total = o;
  return L;
}
__global__ void __launch_bounds__(256) k_graph_pack(dge_graph_out g, const int32_t *g_sel, int B, int Fmax, unsigned char *arena, int64_t cap, dge_config cfg,
                                                    const int32_t *n_poses, const double *est_pose, int Tmax) {
  const int64_t G = g.totals[0], N = g.totals[1], E = g.totals[2];
  const PackLayout L = pack_layout(G, N, E, Fmax);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  int64_t *hdr = reinterpret_cast<int64_t *>(arena);
  if (tid == 0) {
    hdr[0] = G; hdr[1] = N; hdr[2] = E; hdr[3] = g.totals[3] | (L.total > cap ? 2 : 0); hdr[4] = g.totals[4]; hdr[5] = L.total;
  }
  if (L.total > cap || g.totals[3]) return;
  auto cp32 = [&](int64_t off, const void *src, int64_t words) {
    uint32_t *d = reinterpret_cast<uint32_t *>(arena + off);
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
    for (int64_t i = tid; i < words; i += nth) d[i] = s[i];
  };
  cp32(L.off[0], g.x, N * 5);
  cp32(L.off[1], g.edge_index, E * 2);
  cp32(L.off[1] + E * 8, g.edge_index + g.edge_cap, E * 2);
  cp32(L.off[2], g.edge_attr, E);
  cp32(L.off[3], g.node_ptr, G + 1);
  cp32(L.off[4], g.edge_ptr, G + 1);
  cp32(L.off[5], g.key_size, G);
  cp32(L.off[6], g.fro_size, G);
  {   // frontier coordinates by graph ordinal (the device array is indexed by env)
    uint32_t *d = reinterpret_cast<uint32_t *>(arena + L.off[7]);
    const int64_t per = (int64_t)Fmax * 4;     // words per env
    for (int64_t i = tid; i < (int64_t)B * per; i += nth) {
      const int b = (int)(i / per), gi = g_sel[b];
      if (gi >= 0) d[(int64_t)gi * per + (i - (int64_t)b * per)] = reinterpret_cast<const uint32_t *>(g.frontier_xy)[i];
    }
  }
  {   // line plan of every frontier of every selected env, by graph ordinal (same closed form and the same estimate as k_line_plan / k_select_plan)
    double *d = reinterpret_cast<double *>(arena + L.off[12]);
    for (int64_t i = tid; i < (int64_t)B * Fmax; i += nth) {
      const int b = (int)(i / Fmax), f = (int)(i - (int64_t)b * Fmax), gi = g_sel[b];
      if (gi < 0) continue;
      double *pl = d + ((int64_t)gi * Fmax + f) * 6;
      if (f < g.fro_size[gi]) {
        const double *p = est_pose + ((size_t)b * Tmax + n_poses[b] - 1) * 3;
        line_plan(cfg, p[0], p[1], p[2], g.frontier_xy[((size_t)b * Fmax + f) * 2], g.frontier_xy[((size_t)b * Fmax + f) * 2 + 1], pl);
      } else {
        for (int k = 0; k < 6; ++k) pl[k] = 0.0;
      }
    }
  }
  if (g.csr_rowptr) {
    cp32(L.off[8], g.csr_rowptr, N + 1);
    cp32(L.off[9], g.csr_perm, E);
    cp32(L.off[10], g.gcn_norm, E);
    cp32(L.off[11], g.gcn_selfnorm, N);
  }
}
}  // namespace

extern "C" int64_t dge_graph_packed_capacity(dge_handle h, const dge_graph_out *dev) {
  if (!h || !dev) return -1;
  return pack_layout(h->d.B, dev->node_cap, dev->edge_cap, h->d.Fmax).total;
}

extern "C" int dge_graph_host_packed_begin(dge_handle h, const uint8_t *mask_host, const dge_graph_out *dev, void *arena_dev, int64_t arena_cap, void *stream) {
  if (!h || !dev || !arena_dev || arena_cap < 128) return fail(DGE_EINVAL, "dge_graph_host_packed_begin: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t *mask = nullptr;
  if (mask_host) {
    if (cudaMemcpyAsync(h->mask_dev_scratch2, mask_host, h->d.B, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_graph_host_packed_begin: H2D mask");
    mask = h->mask_dev_scratch2;
  }
  const int rc = dge_graph(h, mask, dev, stream);
  if (rc) return rc;
  k_graph_pack<<<148, 256, 0, st>>>(*dev, h->g_sel, h->d.B, h->d.Fmax, static_cast<unsigned char *>(arena_dev), arena_cap, h->cfg, h->n_poses, h->est_pose, h->d.Tmax);
  if (cudaGetLastError() != cudaSuccess) return fail(DGE_ECUDA, "dge_graph_host_packed_begin: k_graph_pack");
  if (cudaMemcpyAsync(h->pack_hdr_host, arena_dev, 128, cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_graph_host_packed_begin: D2H header");
  return DGE_OK;
}

// `prefetched` bytes of the arena were already queued for the host behind ..._begin on the same stream (dge_graph_host_packed_prefetch): when the
// batch fits in them the second copy and its synchronisation are not needed
static int packed_end(dge_handle h, const void *arena_dev, void *arena_host, int64_t arena_cap, int64_t prefetched, dge_graph_packed *out, void *stream);
extern "C" int dge_graph_host_packed_end(dge_handle h, const void *arena_dev, void *arena_host, int64_t arena_cap, dge_graph_packed *out, void *stream) {
  return packed_end(h, arena_dev, arena_host, arena_cap, 0, out, stream);
}
extern "C" int dge_graph_host_packed_prefetch(dge_handle h, const void *arena_dev, void *arena_host, int64_t bytes, void *stream) {
  if (!h || !arena_dev || !arena_host || bytes <= 128) return fail(DGE_EINVAL, "dge_graph_host_packed_prefetch: bad arguments");
  if (cudaMemcpyAsync(static_cast<unsigned char *>(arena_host) + 128, static_cast<const unsigned char *>(arena_dev) + 128, (size_t)(bytes - 128), cudaMemcpyDeviceToHost,
                      static_cast<cudaStream_t>(stream)) != cudaSuccess)
    return fail(DGE_ECUDA, "dge_graph_host_packed_prefetch: D2H");
  return DGE_OK;
}
extern "C" int dge_graph_host_packed_end_prefetched(dge_handle h, const void *arena_dev, void *arena_host, int64_t arena_cap, int64_t prefetched,
                                                    dge_graph_packed *out, void *stream) {
  return packed_end(h, arena_dev, arena_host, arena_cap, prefetched < 128 ? 0 : prefetched, out, stream);
}
static int packed_end(dge_handle h, const void *arena_dev, void *arena_host, int64_t arena_cap, int64_t prefetched, dge_graph_packed *out, void *stream) {
  if (!h || !arena_dev || !arena_host || !out) return fail(DGE_EINVAL, "dge_graph_host_packed_end: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaStreamSynchronize(st) != cudaSuccess) return fail(DGE_ECUDA, "dge_graph_host_packed_end: sync");
  const int64_t *hdr = h->pack_hdr_host;
  const int64_t G = hdr[0], N = hdr[1], E = hdr[2];
  if (hdr[3]) return fail(DGE_ECAP, "dge_graph_host_packed_end: graph batch or arena capacity exceeded");
  const PackLayout L = pack_layout(G, N, E, h->d.Fmax);
  out->n_graphs = (int32_t)G; out->n_nodes = (int32_t)N; out->n_edges = (int32_t)E; out->n_done = (int32_t)hdr[4];
  out->total_bytes = G > 0 ? L.total : 128;
  int64_t *o = &out->x;
  for (int i = 0; i < PK_SECTIONS; ++i) o[i] = L.off[i];
  memcpy(arena_host, hdr, 128);
  if (G == 0) return DGE_OK;
  if (L.total > arena_cap) return fail(DGE_ECAP, "dge_graph_host_packed_end: arena too small");
  const int64_t have = prefetched > 128 ? prefetched : 128;
  if (L.total <= have) return DGE_OK;            // the whole batch came with the prefetch
  if (cudaMemcpyAsync(static_cast<unsigned char *>(arena_host) + have, static_cast<const unsigned char *>(arena_dev) + have, L.total - have, cudaMemcpyDeviceToHost, st) != cudaSuccess)
    return fail(DGE_ECUDA, "dge_graph_host_packed_end: D2H arena");
  if (cudaStreamSynchronize(st) != cudaSuccess) return fail(DGE_ECUDA, "dge_graph_host_packed_end: sync");
  return DGE_OK;
}

extern "C" int dge_line_plan_host(dge_handle h, const double *goal_host, const uint8_t *mask_host, double *plan_host, void *stream) {
  if (!h || !goal_host || !plan_host) return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t B = h->d.B;
  if (cudaMemcpyAsync(h->goal_dev_scratch, goal_host, B * 2 * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_line_plan_host: H2D goals");
  const uint8_t *mask = nullptr;
  if (mask_host) {
    if (cudaMemcpyAsync(h->mask_dev_scratch2, mask_host, B, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_line_plan_host: H2D mask");
    mask = h->mask_dev_scratch2;
  }
  const int rc = dge_launch_line_plan(h, h->goal_dev_scratch, mask, h->plan_dev_scratch, st);
  if (rc) return fail(rc, "dge_line_plan_host");
  if (cudaMemcpyAsync(plan_host, h->plan_dev_scratch, B * 6 * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail(DGE_ECUDA, "dge_line_plan_host: D2H plans");
  if (cudaStreamSynchronize(st) != cudaSuccess) return fail(DGE_ECUDA, "dge_line_plan_host: sync");
  return DGE_OK;
}

extern "C" int dge_virtual_map_rebuild(const dge_config *cfg, int n, int T, const double *pose, const double *cov, int L, const double *lm,
                                       double *prob, double *vinfo, int32_t *seen, double *ws /* dge_virtual_map_rebuild_ws_doubles(n, T) */, void *stream) {
  if (!cfg || n <= 0 || T <= 0 || !pose || !cov || !prob || !vinfo || !ws) return DGE_EINVAL;
  double *prep = ws, *cbox = ws + (size_t)n * T * dge_vmap_prep_width();
  const int rc = dge_vmap_standalone(cfg, n, T, pose, cov, L, lm, prob, vinfo, seen, prep, cbox, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_virtual_map_rebuild") : DGE_OK;
}
extern "C" int64_t dge_virtual_map_rebuild_ws_doubles(int n, int T) {
  return (int64_t)n * T * dge_vmap_prep_width() + (int64_t)n * (dge_vmap_nchunk(T) * 4 + 12);
}

extern "C" int dge_get_state(dge_handle h, dge_state_view *o) {
  if (!h || !o) return DGE_EINVAL;
  o->n_poses = h->n_poses; o->sim_step = h->sim_step; o->update_count = h->update_count;
  o->true_pose = h->true_pose; o->est_pose = h->est_pose; o->lin_pose = h->lin_pose; o->delta_pose = h->delta_pose;
  o->pose_cov = h->pose_cov; o->pose_info = h->pose_info; o->odom = h->odom;
  o->meas_ptr = h->meas_ptr; o->meas_id = h->meas_id; o->meas_bearing = h->meas_b; o->meas_range = h->meas_r;
  o->lm_true = h->lm_true; o->scan_id = h->scan_id; o->observed = h->observed; o->est_l = h->est_l; o->lin_l = h->lin_l;
  o->land_cov = h->land_cov; o->prob = h->prob; o->vinfo = h->vinfo; o->seen = h->seen; o->metrics = h->metrics; o->done = h->done; o->active = h->active; o->status = h->status;
  o->pending = h->pending; o->forced = h->forced; o->seed = reinterpret_cast<const int64_t *>(h->seed); o->plan = h->plan; o->plan_cursor = h->plan_cursor; o->counters = reinterpret_cast<const int64_t *>(h->counters); o->slam_clocks = reinterpret_cast<const int64_t *>(h->slam_clocks);
  return DGE_OK;
}

extern "C" int dge_graph(dge_handle h, const uint8_t *mask, const dge_graph_out *out, void *stream) {
  if (!h || !out || !out->x || !out->edge_index || !out->edge_attr || !out->batch || !out->node_ptr || !out->edge_ptr ||
      !out->key_size || !out->fro_size || !out->frontier_xy || !out->totals)
    return fail(DGE_EINVAL, "dge_graph: null output buffer");
  const int rc = dge_launch_graph(h, mask, out, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_graph") : DGE_OK;
}

extern "C" int dge_mark_pending(dge_handle h, void *stream) {
  if (!h) return DGE_EINVAL;
  const int rc = dge_launch_mark_pending(h, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_mark_pending") : DGE_OK;
}

extern "C" int dge_line_plan(dge_handle h, const double *goal, const uint8_t *mask, double *plan, void *stream) {
  if (!h || !goal || !plan) return DGE_EINVAL;
  const int rc = dge_launch_line_plan(h, goal, mask, plan, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_line_plan") : DGE_OK;
}

extern "C" int dge_select_and_plan(dge_handle h, const dge_graph_out *g, const float *q, const uint8_t *mask, int32_t *choice, void *stream) {
  if (!h || !g || !q) return DGE_EINVAL;
  const int rc = dge_launch_select_plan(h, g, q, mask, choice, static_cast<cudaStream_t>(stream));
  return rc ? fail(rc, "dge_select_and_plan") : DGE_OK;
}

extern "C" int dge_set_counting(dge_handle h, int on) {
  if (!h) return DGE_EINVAL;
  h->count_steps = on ? 1 : 0;
  return DGE_OK;
}
