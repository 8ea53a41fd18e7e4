// Exploration-graph construction on device: frontier detection, frontier association,
// node features, COO edge list in the reference's data_process order, line planner and the
// policy read-out (arg-max over frontier nodes).  Rows a9-a13 of SURVEY section 8.
//
// Replaces ExplorationEnv.frontier / graph_matrix (exploration_env.py:196-358),
// SLAM2D::adjacency_degree_get / key_size / get_key_points (SLAM2D.cpp:141-273),
// DeepQ.data_process (policy.py:211-232) and EMPlanner2D::line_planner
// (Planner2D.cpp:937-1041), which the reference runs as Python loops over a dense N x N
// matrix with ~4K pybind calls per graph.
//
// The dense adjacency never exists here.  Its row-major upper-triangle scan order is
// reproduced analytically: landmark rows (observing poses in time order, then the associated
// frontier), pose rows (next pose; the current pose also links frontier 0), frontier rows (none).
// Integer topology is bit-exact: distance comparisons use unfused fp64 mul/add/sqrt exactly
// like NumPy, ties resolve to the first row-major frontier cell.
#include "dge_internal.cuh"

namespace {

constexpr int GT = 256;

struct GraphArgs {
  dge_config cfg;
  DgeDims d;
  const int32_t *n_poses;
  const double *est_pose, *pose_cov, *odom;
  const int32_t *meas_ptr, *meas_id, *meas_pose;
  const double *meas_r;
  const uint8_t *observed;
  const double *est_l, *land_cov, *prob, *vinfo;
  int32_t *g_counts, *g_frontier, *g_fassoc, *g_sel;
  int32_t *g_cnt, *g_cur, *g_tmp;
  float *g_dis;
};

// envs that need a decision: action queue empty, episode not finished, no forced reset steps outstanding
__global__ void k_mark_pending(int B, const double *plan, const int32_t *cursor, const int32_t *forced, const uint8_t *done, uint8_t *pending) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) pending[b] = (cursor[b] >= (int)plan[6 * b + 5] && !done[b] && forced[b] == 0) ? 1 : 0;
}

__device__ __forceinline__ double dist_np(double ax, double ay, double bx, double by) {  // exploration_env.py:374-376
  const double dx = ax - bx, dy = ay - by;
  return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// ------------------------------------------------------------ pass 1: count ---
__global__ void __launch_bounds__(GT) k_graph_count(GraphArgs a, const uint8_t *mask) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t *cnt = a.g_counts + 4 * b;
  if (mask && !mask[b]) { if (tid == 0) { cnt[0] = 0; cnt[1] = 0; cnt[2] = 0; cnt[3] = 0; } return; }
  const int rows = a.d.rows, cols = a.d.cols, V = a.d.V, Lt = a.d.Lt, T = a.n_poses[b];
  extern __shared__ unsigned char sm_raw[];
  uint8_t *fro = sm_raw;                                   // [V] frontier flag
  int *near_cell = (int *)(sm_raw + ((V + 15) / 16) * 16);  // [Lt+1]
  int *lrank_id = near_cell + Lt + 1;                      // [Lt] rank -> id
  __shared__ int s_L, s_nf;
  const double *p = a.prob + (size_t)b * V;
  const double res = a.cfg.resolution, ext = 20.0;
  // frontier cells  (exploration_env.py:292-325)
  for (int i = tid; i < V; i += GT) {
    const int r = i / cols, c = i % cols;
    uint8_t f = 0;
    if (p[i] < 0.45) {
      int count = 0;
      for (int rr = max(r - 1, 0); rr <= min(r + 1, rows - 1); ++rr)
        for (int cc = max(c - 1, 0); cc <= min(c + 1, cols - 1); ++cc) {
          const double q = p[rr * cols + cc];
          if (0.49 < q && q < 0.51) ++count;
        }
      if (count >= 2) {
        const double x = (c + 0.5) * res + a.cfg.map_min_x, y = (r + 0.5) * res + a.cfg.map_min_y;
        if (a.cfg.map_min_x + ext <= x && x <= a.cfg.map_max_x - ext && a.cfg.map_min_y + ext <= y && y <= a.cfg.map_max_y - ext) f = 1;
      }
    }
    fro[i] = f;
  }
  if (tid == 0) {
    int n = 0;
    for (int j = 0; j < Lt; ++j) if (a.observed[(size_t)b * Lt + j]) lrank_id[n++] = j;
    s_L = n;
  }
  __syncthreads();
  const int L = s_L;
  // nearest frontier of the robot (query 0) and of every landmark in key order (query 1+rank):
  // one warp per query, lexicographic (distance, row-major index) minimum == first strict minimum.
  for (int q = warp; q <= L; q += GT / 32) {
    double px, py;
    if (q == 0) { px = a.est_pose[((size_t)b * a.d.Tmax + T - 1) * 3]; py = a.est_pose[((size_t)b * a.d.Tmax + T - 1) * 3 + 1]; }
    else { const int id = lrank_id[q - 1]; px = a.est_l[((size_t)b * Lt + id) * 2]; py = a.est_l[((size_t)b * Lt + id) * 2 + 1]; }
    double best = 1e300;
    int bi = 0x7fffffff;
    for (int i = lane; i < V; i += 32) {
      if (!fro[i]) continue;
      const double x = (i % cols + 0.5) * res + a.cfg.map_min_x, y = (i / cols + 0.5) * res + a.cfg.map_min_y;
      const double dd = dist_np(px, py, x, y);
      if (dd < best) { best = dd; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) near_cell[q] = (bi == 0x7fffffff) ? -1 : bi;
  }
  __syncthreads();
  if (tid == 0) {
    // frontier list with de-duplication (exploration_env.py:327-338)
    int32_t *fl = a.g_frontier + (size_t)b * a.d.Fmax, *as = a.g_fassoc + (size_t)b * (Lt + 1);
    int F = 0;
    if (near_cell[0] >= 0) {   // q15: no frontier cell at all -> F = 0
      fl[F++] = near_cell[0];
      as[0] = 0;
      for (int ip = 0; ip < L; ++ip) {
        const int c = near_cell[ip + 1];
        int f = -1;
        for (int k = 0; k < F; ++k) if (fl[k] == c) { f = k; break; }
        if (f < 0) { fl[F] = c; f = F++; }
        as[ip + 1] = f;
      }
    } else { for (int i = 0; i <= L; ++i) as[i] = -1; }
    s_nf = F;
  }
  __syncthreads();
  const int F = s_nf, K = L + T;
  // undirected edge count; zero-weight edges are not edges (policy.py:219)
  int ne = 0;
  for (int k = tid; k < T - 1; k += GT) ne += 1;   // odometry weight = |odom_xy| + 0.001 > 0
  const int M = a.meas_ptr[(size_t)b * (a.d.Tmax + 1) + T];
  for (int m = tid; m < M; m += GT) ne += (a.meas_r[(size_t)b * a.d.Mmax + m] != 0.0) ? 1 : 0;
  if (F > 0) {
    for (int q = tid; q <= L; q += GT) {
      const int f = a.g_fassoc[(size_t)b * (Lt + 1) + q];
      const int cell = a.g_frontier[(size_t)b * a.d.Fmax + f];
      const double fx = (cell % cols + 0.5) * res + a.cfg.map_min_x, fy = (cell / cols + 0.5) * res + a.cfg.map_min_y;
      double px, py;
      if (q == 0) { px = a.est_pose[((size_t)b * a.d.Tmax + T - 1) * 3]; py = a.est_pose[((size_t)b * a.d.Tmax + T - 1) * 3 + 1]; }
      else { const int id = lrank_id[q - 1]; px = a.est_l[((size_t)b * Lt + id) * 2]; py = a.est_l[((size_t)b * Lt + id) * 2 + 1]; }
      ne += (dist_np(fx, fy, px, py) != 0.0) ? 1 : 0;
    }
  }
  __shared__ int s_red[GT];
  s_red[tid] = ne;
  __syncthreads();
  for (int o = GT / 2; o > 0; o >>= 1) { if (tid < o) s_red[tid] += s_red[tid + o]; __syncthreads(); }
  if (tid == 0) { cnt[0] = K + F; cnt[1] = 2 * s_red[0]; cnt[2] = K; cnt[3] = F; }
}

// -------------------------------------------------------------- pass 2: scan ---
// single CTA, 1024 threads: three exclusive scans (graph position, node offset, edge offset) over
// the selected envs; each thread owns a contiguous chunk of envs, chunk totals by Hillis-Steele.
__global__ void __launch_bounds__(1024) k_graph_scan(int B, const uint8_t *mask, const int32_t *g_counts, int32_t *g_sel, dge_graph_out o,
                                                     const uint8_t *done, unsigned long long *counters) {
  __shared__ int sg[1024], sn[1024], se[1024], sd[1024];
  const int tid = threadIdx.x;
  const int per = (B + 1023) / 1024;
  const int lo = min(B, tid * per), hi = min(B, lo + per);
  int g = 0, n = 0, e = 0, nd = 0;
  for (int b = lo; b < hi; ++b) {
    nd += done[b] ? 1 : 0;
    if (!(mask && !mask[b])) { ++g; n += g_counts[4 * b]; e += g_counts[4 * b + 1]; }
  }
  sg[tid] = g; sn[tid] = n; se[tid] = e; sd[tid] = nd;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int vg = tid >= off ? sg[tid - off] : 0, vn = tid >= off ? sn[tid - off] : 0, ve = tid >= off ? se[tid - off] : 0,
              vd = tid >= off ? sd[tid - off] : 0;
    __syncthreads();
    sg[tid] += vg; sn[tid] += vn; se[tid] += ve; sd[tid] += vd;
    __syncthreads();
  }
  g = tid ? sg[tid - 1] : 0; n = tid ? sn[tid - 1] : 0; e = tid ? se[tid - 1] : 0;
  for (int b = lo; b < hi; ++b) {
    if (mask && !mask[b]) { g_sel[b] = -1; continue; }
    g_sel[b] = g;
    o.node_ptr[g] = n; o.edge_ptr[g] = e;
    o.key_size[g] = g_counts[4 * b + 2]; o.fro_size[g] = g_counts[4 * b + 3];
    n += g_counts[4 * b]; e += g_counts[4 * b + 1];
    ++g;
  }
  if (tid == 1023) {
    const int G = sg[1023], N = sn[1023], E = se[1023];
    o.node_ptr[G] = N; o.edge_ptr[G] = E;
    o.totals[0] = G; o.totals[1] = N; o.totals[2] = E;
    o.totals[3] = (N > o.node_cap || E > o.edge_cap) ? 1 : 0;
    o.totals[4] = sd[1023];   // envs whose episode is over (lets the host fold the reset check into the same D2H)
    if (o.csr_rowptr && !o.totals[3]) o.csr_rowptr[N] = E;
    if (counters) { counters[4] += G; counters[5] += N; counters[6] += E; counters[7] += 1; }   // single writer: this thread
  }
}

// -------------------------------------------------------------- pass 3: fill ---
__global__ void __launch_bounds__(GT) k_graph_fill(GraphArgs a, dge_graph_out o) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = a.g_sel[b];
  if (g < 0 || o.totals[3]) return;
  const int rows = a.d.rows, cols = a.d.cols, Lt = a.d.Lt, Tmax = a.d.Tmax, T = a.n_poses[b];
  (void)rows;
  const int N = a.g_counts[4 * b], K = a.g_counts[4 * b + 2], F = a.g_counts[4 * b + 3], L = K - T;
  const int n0 = o.node_ptr[g], e0 = o.edge_ptr[g];
  const double res = a.cfg.resolution;
  extern __shared__ unsigned char sm_raw[];
  int *lrank_id = (int *)sm_raw;          // [Lt]
  int *lidx = lrank_id + Lt;              // [Lt] id -> rank
  int *lcnt = lidx + Lt;                  // [Lt+1] observations per landmark rank
  int *rowoff = lcnt + Lt + 1;            // [Lt+1] directed-edge offset of each landmark row
  __shared__ int s_pose_off;
  const int32_t *fl = a.g_frontier + (size_t)b * a.d.Fmax, *as = a.g_fassoc + (size_t)b * (Lt + 1);
  const double *estp = a.est_pose + (size_t)b * Tmax * 3, *estl = a.est_l + (size_t)b * Lt * 2;
  const double rx = estp[3 * (T - 1)], ry = estp[3 * (T - 1) + 1], rth = estp[3 * (T - 1) + 2];
  if (tid == 0) {
    int n = 0;
    for (int j = 0; j < Lt; ++j) { if (a.observed[(size_t)b * Lt + j]) { lidx[j] = n; lrank_id[n++] = j; } else lidx[j] = -1; }
  }
  for (int i = tid; i <= Lt; i += GT) lcnt[i] = 0;
  __syncthreads();
  const int M = a.meas_ptr[(size_t)b * (Tmax + 1) + T];
  const int32_t *mid = a.meas_id + (size_t)b * a.d.Mmax, *mpose = a.meas_pose + (size_t)b * a.d.Mmax;
  const double *mr = a.meas_r + (size_t)b * a.d.Mmax;
  for (int m = tid; m < M; m += GT) if (mr[m] != 0.0) atomicAdd(&lcnt[lidx[mid[m]]], 1);
  __syncthreads();
  auto frontier_w = [&](int q, double &w) {   // weight of the (node q, its frontier) edge; q = 0 robot, 1+rank landmark
    const int f = as[q];
    const int cell = fl[f];
    const double fx = (cell % cols + 0.5) * res + a.cfg.map_min_x, fy = (cell / cols + 0.5) * res + a.cfg.map_min_y;
    double px, py;
    if (q == 0) { px = rx; py = ry; } else { px = estl[2 * lrank_id[q - 1]]; py = estl[2 * lrank_id[q - 1] + 1]; }
    w = dist_np(fx, fy, px, py);
    return f;
  };
  if (tid == 0) {
    int off = 0;
    for (int r = 0; r < L; ++r) {
      rowoff[r] = off;
      off += 2 * lcnt[r];
      if (F > 0) { double w; frontier_w(r + 1, w); if (w != 0.0) off += 2; }
    }
    s_pose_off = off;
  }
  __syncthreads();
  int64_t *src = o.edge_index + e0, *dst = o.edge_index + o.edge_cap + e0;
  float *ew = o.edge_attr + e0;
  auto emit = [&](int pos, int i, int j, double w) {   // policy.py:221-225: (i,j) then (j,i)
    src[pos] = n0 + i; dst[pos] = n0 + j; ew[pos] = (float)w;
    src[pos + 1] = n0 + j; dst[pos + 1] = n0 + i; ew[pos + 1] = (float)w;
  };
  // landmark rows: one warp per landmark, measurements in time order (stable ballot compaction)
  for (int r = warp; r < L; r += GT / 32) {
    const int id = lrank_id[r];
    int base = rowoff[r];
    for (int m0 = 0; m0 < M; m0 += 32) {
      const int m = m0 + lane;
      const bool hit = m < M && mid[m] == id && mr[m] != 0.0;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (hit) emit(base + 2 * __popc(bal & ((1u << lane) - 1u)), r, L + mpose[m], mr[m]);   // SLAM2D.cpp:256
      base += 2 * __popc(bal);
    }
    if (lane == 0 && F > 0) { double w; const int f = frontier_w(r + 1, w); if (w != 0.0) emit(base, r, K + f, w); }
  }
  // pose rows: odometry edge to the next pose (SLAM2D.cpp:239); the current pose links frontier 0
  const double *od = a.odom + (size_t)b * Tmax * 3;
  for (int k = tid; k < T - 1; k += GT) {
    const double w = sqrt(od[3 * k] * od[3 * k] + od[3 * k + 1] * od[3 * k + 1]) + 0.001;
    emit(s_pose_off + 2 * k, L + k, L + k + 1, w);
  }
  if (tid == 0 && F > 0) { double w; frontier_w(0, w); if (w != 0.0) emit(s_pose_off + 2 * (T - 1), K - 1, K, w); }
  // node features (exploration_env.py:226-274), f64 -> f32 like policy.py:229
  for (int i = tid; i < N; i += GT) {
    double x, y, f0;
    if (i < L) { const int id = lrank_id[i]; x = estl[2 * id]; y = estl[2 * id + 1]; const double *c = a.land_cov + ((size_t)b * Lt + id) * 3; f0 = c[0] + c[2]; }
    else if (i < K) { const int k = i - L; x = estp[3 * k]; y = estp[3 * k + 1]; const double *c = a.pose_cov + ((size_t)b * Tmax + k) * 6; f0 = c[0] + c[3] + c[5]; }
    else {
      const int cell = fl[i - K];
      x = (cell % cols + 0.5) * res + a.cfg.map_min_x; y = (cell / cols + 0.5) * res + a.cfg.map_min_y;
      const double *v = a.vinfo + ((size_t)b * a.d.V + cell) * 3;
      f0 = (v[0] + v[2]) / (v[0] * v[2] - v[1] * v[1]);   // VirtualMap::toCovTrace
      o.frontier_xy[((size_t)b * a.d.Fmax + (i - K)) * 2] = x;
      o.frontier_xy[((size_t)b * a.d.Fmax + (i - K)) * 2 + 1] = y;
    }
    const int mj = (int)rint((x - a.cfg.map_min_x) / res - 0.5), mi = (int)rint((y - a.cfg.map_min_y) / res - 0.5);   // coor2index, round-half-even
    double goal = atan2(y - ry, x - rx), root = rth;   // diff_theta  exploration_env.py:378-387
    if (goal < 0) goal = DGE_PI * 2 + goal;
    if (root < 0) root = DGE_PI * 2 + root;
    double diff = goal - root;
    if (diff < 0) diff = DGE_PI * 2 + diff;
    float *xo = o.x + (size_t)(n0 + i) * 5;
    xo[0] = (float)f0;
    xo[1] = (float)dist_np(x, y, rx, ry);
    xo[2] = (float)diff;
    xo[3] = (float)a.prob[(size_t)b * a.d.V + mi * cols + mj];
    xo[4] = (i < K - 1) ? -1.f : (i == K - 1 ? 0.f : 1.f);
    o.batch[n0 + i] = g;
  }
  // ---- destination-sorted CSR + GCNConv(improved=True) normalisation of this graph, in place of the
  // GNN's separate preprocessing launches (count / scan / fill / rank / degree / norm).  Rows sorted by
  // edge id => deterministic aggregation order; the graph is symmetric, so in- and out-weights coincide.
  if (!o.csr_rowptr) return;
  const int E = a.g_counts[4 * b + 1];
  int32_t *cnt = a.g_cnt + (size_t)b * a.d.Ncap, *cur = a.g_cur + (size_t)b * a.d.Ncap, *tmp = a.g_tmp + (size_t)b * a.d.Ecap;
  float *dis = a.g_dis + (size_t)b * a.d.Ncap;
  for (int i = tid; i < N; i += GT) { cnt[i] = 0; cur[i] = 0; }
  __syncthreads();   // also orders the edge writes above before the reads below (same CTA)
  for (int e = tid; e < E; e += GT) atomicAdd(&cnt[(int)(dst[e] - n0)], 1);
  __syncthreads();
  if (warp == 0) {   // exclusive scan of the in-degrees (N <= a few hundred): warp-serial chunks
    int run = 0;
    for (int i0 = 0; i0 < N; i0 += 32) {
      const int i = i0 + lane;
      const int v = i < N ? cnt[i] : 0;
      int inc = v;
      for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += t; }
      if (i < N) { cnt[i] = run + inc - v; o.csr_rowptr[n0 + i] = e0 + run + inc - v; }
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  for (int e = tid; e < E; e += GT) { const int d = (int)(dst[e] - n0); tmp[cnt[d] + atomicAdd(&cur[d], 1)] = e; }
  __syncthreads();
  for (int e = tid; e < E; e += GT) {   // rank inside the row = position in ascending edge-id order
    const int d = (int)(dst[e] - n0);
    const int lo = cnt[d], hi = lo + cur[d];
    int rank = 0;
    for (int q = lo; q < hi; ++q) rank += (tmp[q] < e) ? 1 : 0;
    o.csr_perm[e0 + lo + rank] = e0 + e;
  }
  __syncthreads();
  for (int i = tid; i < N; i += GT) {   // weighted degree in row order + the improved self loop (weight 2)
    float deg = 0.f;
    for (int q = cnt[i]; q < cnt[i] + cur[i]; ++q) deg += ew[o.csr_perm[e0 + q] - e0];
    deg += 2.0f;
    const float ds = 1.0f / sqrtf(deg);
    dis[i] = ds;
    o.gcn_selfnorm[n0 + i] = ds * 2.0f * ds;
  }
  __syncthreads();
  for (int e = tid; e < E; e += GT) o.gcn_norm[e0 + e] = dis[(int)(src[e] - n0)] * ew[e] * dis[(int)(dst[e] - n0)];
}

// ------------------------------------------------------------- line planner ---
// (line_plan: dge_internal.cuh -- the packed host transfer evaluates it for every frontier too)
__global__ void k_line_plan(dge_config cfg, DgeDims d, const int32_t *n_poses, const double *est_pose, const double *goal,
                            const uint8_t *mask, double *plan_out, uint8_t *done) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= d.B || (mask && !mask[b])) return;
  if (mask && mask[b] == 2) {   // no frontier left (q15): the episode is over, like k_select_plan's F <= 0 branch
    for (int i = 0; i < 6; ++i) plan_out[6 * b + i] = 0;
    if (done) done[b] = 1;
    return;
  }
  const int T = n_poses[b];
  const double *p = est_pose + ((size_t)b * d.Tmax + T - 1) * 3;
  line_plan(cfg, p[0], p[1], p[2], goal[2 * b], goal[2 * b + 1], plan_out + 6 * b);
}

// policy read-out: first arg-max of q over the graph's last fro_size nodes (np.argmax,
// policy.py:109 / test.py:112), goal = that frontier, plan into the env's queue.
__global__ void __launch_bounds__(32) k_select_plan(dge_config cfg, DgeDims d, const int32_t *n_poses, const double *est_pose,
                                                    const int32_t *g_sel, dge_graph_out o, const float *q, const uint8_t *mask,
                                                    double *plan, int32_t *cursor, int32_t *choice, uint8_t *done) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int g = g_sel[b];
  if (g < 0 || (mask && !mask[b])) return;
  const int F = o.fro_size[g], K = o.key_size[g], n0 = o.node_ptr[g];
  if (F <= 0) { if (lane == 0) { for (int i = 0; i < 6; ++i) plan[6 * b + i] = 0; cursor[b] = 0; if (choice) choice[b] = -1; done[b] = 1; /* q15: no frontier left = episode over */ } return; }
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int f = lane; f < F; f += 32) {   // ascending f per lane: strict > keeps the first maximum
    const float v = q[n0 + K + f];
    if (bi == 0x7fffffff || v > best) { best = v; bi = f; }
  }
  for (int s = 16; s > 0; s >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, s);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
  }
  if (lane == 0) {
    const int T = n_poses[b];
    const double *p = est_pose + ((size_t)b * d.Tmax + T - 1) * 3;
    line_plan(cfg, p[0], p[1], p[2], o.frontier_xy[((size_t)b * d.Fmax + bi) * 2], o.frontier_xy[((size_t)b * d.Fmax + bi) * 2 + 1], plan + 6 * b);
    cursor[b] = 0;
    if (choice) choice[b] = bi;
  }
}

GraphArgs make_gargs(dge_engine *e) {
  GraphArgs a;
  a.cfg = e->cfg; a.d = e->d; a.n_poses = e->n_poses; a.est_pose = e->est_pose; a.pose_cov = e->pose_cov; a.odom = e->odom;
  a.meas_ptr = e->meas_ptr; a.meas_id = e->meas_id; a.meas_pose = e->meas_pose; a.meas_r = e->meas_r; a.observed = e->observed;
  a.est_l = e->est_l; a.land_cov = e->land_cov; a.prob = e->prob; a.vinfo = e->vinfo;
  a.g_counts = e->g_counts; a.g_frontier = e->g_frontier; a.g_fassoc = e->g_fassoc; a.g_sel = e->g_sel;
  a.g_cnt = e->g_cnt; a.g_cur = e->g_cur; a.g_tmp = e->g_tmp; a.g_dis = e->g_dis;
  return a;
}

}  // namespace

int dge_launch_mark_pending(dge_engine *e, cudaStream_t st) {
  k_mark_pending<<<(e->d.B + 255) / 256, 256, 0, st>>>(e->d.B, e->plan, e->plan_cursor, e->forced, e->done, e->pending);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}

int dge_launch_graph(dge_engine *e, const uint8_t *mask, const dge_graph_out *out, cudaStream_t st) {
  const GraphArgs a = make_gargs(e);
  const size_t sm1 = ((e->d.V + 15) / 16) * 16 + (2 * e->d.Lt + 2) * sizeof(int);
  k_graph_count<<<e->d.B, GT, sm1, st>>>(a, mask);
  k_graph_scan<<<1, 1024, 0, st>>>(e->d.B, mask, e->g_counts, e->g_sel, *out, e->done, e->count_steps ? e->counters : nullptr);
  const size_t sm3 = (4 * e->d.Lt + 4) * sizeof(int);
  k_graph_fill<<<e->d.B, GT, sm3, st>>>(a, *out);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}

int dge_launch_line_plan(dge_engine *e, const double *goal, const uint8_t *mask, double *plan_out, cudaStream_t st) {
  k_line_plan<<<(e->d.B + 127) / 128, 128, 0, st>>>(e->cfg, e->d, e->n_poses, e->est_pose, goal, mask, plan_out, e->done);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}

int dge_launch_select_plan(dge_engine *e, const dge_graph_out *g, const float *q, const uint8_t *mask, int32_t *choice, cudaStream_t st) {
  k_select_plan<<<e->d.B, 32, 0, st>>>(e->cfg, e->d, e->n_poses, e->est_pose, e->g_sel, *g, q, mask, e->plan, e->plan_cursor, choice, e->done);
  return cudaGetLastError() == cudaSuccess ? DGE_OK : DGE_ECUDA;
}
