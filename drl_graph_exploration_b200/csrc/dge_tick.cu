// One tick of the acting loop for B environments, issued without any host synchronisation and replayable as a CUDA graph.
//
// Replaces the body of the reference's test loop (scripts/test.py:100-143: graph_matrix -> data_process -> model ->
// arg-max over the frontier nodes -> actions_all_goals -> env.step per action; exploration_env.py:98-105 is the step) and
// the reset at the end of an episode (exploration_env.py:389-422), for B envs per call:
//
//   stream            mark_pending -> graph (count / scan / fill, mask = pending) -> GCN Q-network (fused first layer ->
//                     tcgen05 GEMM -> aggregate + head; node count read from the graph batch's `totals` on the device,
//                     launches sized by capacity) -> [wait: this tick's move kernel has read the plans] -> select_and_plan
//   engine's stream   reset_done_queued -> move_measure_queued -> SLAM -> virtual map         (forked after mark_pending)
//
// The two pipelines touch disjoint env sets (an env either has a queued action or needs a decision).  Everything the host
// would have to know to size a launch (how many envs decide, how many nodes their graphs have) stays on the device, so the
// whole tick is a fixed launch sequence: it is captured ONCE into a CUDA graph (fork / join included) and replayed with
// one cudaGraphLaunch per tick.
#include <cstdlib>
#include <cstring>

#include "dge_internal.cuh"
#include "../../include/dge_gnn.h"

int dge_gcn_q_forward_dev(int N, const int32_t *N_dev, int Cin, int C, const float *x, const int32_t *rowptr, const int32_t *perm, const int64_t *src,
                          const float *norm, const float *selfnorm, const float *W1, const float *b1, const float *W2t_hi,
                          const float *W2t_lo, const float *b2, const float *head_w, const float *head_b_dev, float *ws, float *q,
                          cudaStream_t st);

namespace {

struct TickKey {          // everything that is baked into a captured tick
  dge_graph_out g;
  dge_gcn_policy pol;
  uint64_t seed_stride;
  double forced_odom[3];
  int n_forced, flags, count_steps, park_done;
};

int ensure_streams(dge_engine *e) {
  if (e->tick_stream) return DGE_OK;
  if (cudaStreamCreateWithFlags(&e->tick_stream, cudaStreamNonBlocking) != cudaSuccess) return DGE_ECUDA;
  if (cudaStreamCreateWithFlags(&e->tick_stream2, cudaStreamNonBlocking) != cudaSuccess) return DGE_ECUDA;
  if (cudaStreamCreateWithFlags(&e->tick_cap_stream, cudaStreamNonBlocking) != cudaSuccess) return DGE_ECUDA;
  for (cudaEvent_t *ev : {&e->ev_fork, &e->ev_move, &e->ev_join, &e->ev_split, &e->ev_light})
    if (cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) return DGE_ECUDA;
  return DGE_OK;
}

// ---- split of the stepping envs into a HEAVY and a LIGHT group (one CTA) --------------------------------------------------
// A SLAM launch lasts as long as its slowest env (a rebuild that reaches far back, or a long trajectory: up to ~2.5x the mean), and the
// virtual-map launch behind it waits for that one CTA although the other envs' estimates have been final for tens of microseconds.  The
// tick therefore runs two SLAM -> virtual-map chains side by side: the n_heavy envs with the largest expected cost (trajectory length,
// doubled when the update is on the relinearisation schedule -- the measure of k_step_order) on the step stream, all others on a second
// stream.  The light group's map rebuild then overlaps the heavy group's SLAM tail; per env nothing changes (both kernels are per-env).
__global__ void __launch_bounds__(1024) k_split_groups(int B, int n_heavy, int relin_skip, const uint8_t *active, const int32_t *n_poses,
                                                       const int32_t *update_count, uint8_t *heavy, uint8_t *light) {
  extern __shared__ int s_cost[];
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int cst = -1;
    if (active[b]) {
      cst = n_poses[b];
      if (relin_skip > 0 && (update_count[b] + 1) % relin_skip == 0) cst *= 2;
    }
    s_cost[b] = cst;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int cst = s_cost[b];
    int rank = 0;
    for (int o = 0; o < B; ++o) { const int co = s_cost[o]; rank += (co > cst || (co == cst && o < b)) ? 1 : 0; }
    const bool act = cst >= 0, hv = act && rank < n_heavy;
    heavy[b] = hv ? 1 : 0;
    light[b] = (act && !hv) ? 1 : 0;
  }
}

// the fixed launch sequence of a tick on `st` (+ the engine's second stream unless DGE_TICK_ONE_STREAM)
int issue_tick(dge_engine *e, const dge_graph_out *g, const dge_gcn_policy *pol, uint64_t seed_stride, int n_forced, int flags, cudaStream_t st) {
  const bool fork = !(flags & DGE_TICK_ONE_STREAM);
  cudaStream_t s1 = fork ? e->tick_stream : st;
  int rc;
  if ((rc = dge_launch_mark_pending(e, st))) return rc;          // before the step pipeline moves the queues
  if (fork) {
    if (cudaEventRecord(e->ev_fork, st) != cudaSuccess || cudaStreamWaitEvent(s1, e->ev_fork, 0) != cudaSuccess) return DGE_ECUDA;
  }
  // ---- step pipeline
  if ((rc = dge_launch_reset(e, e->done, nullptr, nullptr, nullptr, nullptr, nullptr, n_forced, seed_stride, s1))) return rc;
  if ((rc = dge_launch_move_measure(e, nullptr, nullptr, nullptr, 1, s1))) return rc;
  if (fork && cudaEventRecord(e->ev_move, s1) != cudaSuccess) return DGE_ECUDA;
  // opt-in (DGE_TICK_HEAVY=n): measured on B200 at 256 envs, n = 16 / 32 / 64: tick 0.187 ms against 0.182 ms with one SLAM / map launch for all
  // envs (profiles/r02_tick_split_ab.md) -- the light group still contains rebuilds the cost measure does not predict, and the two extra
  // launches + fork / join cost what the overlap wins; off by default
  static const int n_heavy = [] { const char *v = getenv("DGE_TICK_HEAVY"); return v ? atoi(v) : 0; }();
  if (fork && n_heavy > 0 && e->d.B > 2 * n_heavy && e->d.B <= 4096) {
    cudaStream_t s2 = e->tick_stream2;
    k_split_groups<<<1, 1024, e->d.B * sizeof(int), s1>>>(e->d.B, n_heavy, e->cfg.relin_skip, e->active, e->n_poses, e->update_count, e->group_heavy, e->group_light);
    if (cudaGetLastError() != cudaSuccess) return DGE_ECUDA;
    if (cudaEventRecord(e->ev_split, s1) != cudaSuccess || cudaStreamWaitEvent(s2, e->ev_split, 0) != cudaSuccess) return DGE_ECUDA;
    if ((rc = dge_launch_slam(e, e->group_heavy, s1))) return rc;      // (issued first: its CTAs take their SMs before the light ones fill the rest)
    if ((rc = dge_launch_slam(e, e->group_light, s2))) return rc;
    if ((rc = dge_launch_vmap(e, e->group_light, s2))) return rc;
    if ((rc = dge_launch_vmap(e, e->group_heavy, s1))) return rc;
    if (cudaEventRecord(e->ev_light, s2) != cudaSuccess || cudaStreamWaitEvent(s1, e->ev_light, 0) != cudaSuccess) return DGE_ECUDA;
  } else {
    if ((rc = dge_launch_slam(e, e->active, s1))) return rc;
    if ((rc = dge_launch_vmap(e, e->active, s1))) return rc;
  }
  if (fork && cudaEventRecord(e->ev_join, s1) != cudaSuccess) return DGE_ECUDA;
  // ---- policy pipeline
  if ((rc = dge_launch_graph(e, e->pending, g, st))) return rc;
  const int ncap = (int)(g->node_cap < pol->node_cap ? g->node_cap : pol->node_cap);
  rc = dge_gcn_q_forward_dev(ncap, g->totals + 1, pol->Cin, pol->C, g->x, g->csr_rowptr, g->csr_perm, g->edge_index, g->gcn_norm, g->gcn_selfnorm,
                             pol->W1, pol->b1, pol->W2t_hi, pol->W2t_lo, pol->b2, pol->head_w, pol->head_b_dev, pol->ws, pol->q, st);
  if (rc) return rc == -1 ? DGE_EINVAL : DGE_ECUDA;
  if (fork && cudaStreamWaitEvent(st, e->ev_move, 0) != cudaSuccess) return DGE_ECUDA;   // plans are rewritten only after the move kernel has read them
  if ((rc = dge_launch_select_plan(e, g, pol->q, nullptr, pol->choice, st))) return rc;
  if (fork && cudaStreamWaitEvent(st, e->ev_join, 0) != cudaSuccess) return DGE_ECUDA;   // join: the tick ends when both pipelines are done
  return DGE_OK;
}

}  // namespace

void dge_tick_release(dge_engine *e) {
  if (e->tick_exec) { cudaGraphExecDestroy(e->tick_exec); e->tick_exec = nullptr; }
  if (e->tick_key) { free(e->tick_key); e->tick_key = nullptr; }
  if (e->tick_stream) { cudaStreamDestroy(e->tick_stream); e->tick_stream = nullptr; }
  if (e->tick_stream2) { cudaStreamDestroy(e->tick_stream2); e->tick_stream2 = nullptr; }
  if (e->tick_cap_stream) { cudaStreamDestroy(e->tick_cap_stream); e->tick_cap_stream = nullptr; }
  for (cudaEvent_t *ev : {&e->ev_fork, &e->ev_move, &e->ev_join, &e->ev_split, &e->ev_light})
    if (*ev) { cudaEventDestroy(*ev); *ev = nullptr; }
}

extern "C" int dge_policy_tick(dge_handle h, const dge_graph_out *g, const dge_gcn_policy *pol, uint64_t seed_stride,
                               const double *forced_odom_host, int n_forced, int flags, void *stream) {
  if (!h || !g || !pol || !forced_odom_host || seed_stride == 0 || n_forced < 1 || n_forced >= DGE_FRESH_BIT) return DGE_EINVAL;
  if (!g->csr_rowptr || !g->csr_perm || !g->gcn_norm || !g->gcn_selfnorm || !pol->W1 || !pol->W2t_hi || !pol->W2t_lo || !pol->head_w ||
      !pol->ws || !pol->q || pol->node_cap <= 0)
    return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ensure_streams(h)) return DGE_ECUDA;
  for (int i = 0; i < 3; ++i) h->forced_odom[i] = forced_odom_host[i];
  if (!(flags & DGE_TICK_GRAPH)) return issue_tick(h, g, pol, seed_stride, n_forced, flags, st);

  TickKey key;
  memset(&key, 0, sizeof(key));
  key.g = *g; key.pol = *pol; key.seed_stride = seed_stride; key.n_forced = n_forced; key.flags = flags;
  key.count_steps = h->count_steps; key.park_done = h->park_done;
  for (int i = 0; i < 3; ++i) key.forced_odom[i] = forced_odom_host[i];
  if (!h->tick_exec || !h->tick_key || memcmp(h->tick_key, &key, sizeof(key)) != 0) {
    if (h->tick_exec) { cudaGraphExecDestroy(h->tick_exec); h->tick_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(h->tick_cap_stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) return DGE_ECUDA;
    const int rc = issue_tick(h, g, pol, seed_stride, n_forced, flags, h->tick_cap_stream);
    const cudaError_t ce = cudaStreamEndCapture(h->tick_cap_stream, &graph);
    if (rc || ce != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); return rc ? rc : DGE_ECUDA; }
    const cudaError_t ci = cudaGraphInstantiate(&h->tick_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ci != cudaSuccess) { h->tick_exec = nullptr; return DGE_ECUDA; }
    if (!h->tick_key) h->tick_key = malloc(sizeof(TickKey));
    if (!h->tick_key) return DGE_ENOMEM;
    memcpy(h->tick_key, &key, sizeof(key));
  }
  return cudaGraphLaunch(h->tick_exec, st) == cudaSuccess ? DGE_OK : DGE_ECUDA;
}

// ---- the host-driven tick (runner.HostPolicyLoop.tick) as one native call: same sequence of C-ABI calls the Python loop makes, the
// NumPy bookkeeping between them done here.  An env is in exactly one of three states: in its reset phase (phase > 0: the initial
// optimize + forced steps queued by dge_reset_done_queued run, one per tick), holding a queued action (cursor < plan length), or in
// need of a decision.
extern "C" int dge_host_policy_tick(dge_handle h, const dge_graph_out *g, const dge_gcn_policy *pol, dge_host_loop *hl, uint64_t seed_stride,
                                    const double *forced_odom_host, int n_forced, void *stream, void *stream_step) {
  if (!h || !g || !pol || !hl || !forced_odom_host || seed_stride == 0) return DGE_EINVAL;
  if (!hl->plans || !hl->cursor || !hl->phase || !hl->mask || !hl->done || !hl->need || !hl->metrics || !hl->arena_host || !hl->q_host ||
      !hl->plan_host || !hl->choice_host || !hl->arena_pack || !hl->arena_dev || !pol->ws || !pol->q)
    return DGE_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream), s1 = stream_step ? static_cast<cudaStream_t>(stream_step) : st;
  if (ensure_streams(h)) return DGE_ECUDA;
  const int B = h->d.B;
  hl->n_stepped = hl->n_graphs = hl->n_nodes = hl->h2d_bytes = hl->d2h_bytes = hl->launches = 0;
  int n_need = 0, n_step = 0;
  for (int b = 0; b < B; ++b) {
    const bool in_reset = hl->phase[b] > 0;
    const bool has_act = !in_reset && hl->cursor[b] < (int64_t)hl->plans[6 * b + 5];
    const bool need = !in_reset && !has_act;
    hl->need[b] = need ? 1 : 0;
    hl->mask[b] = (has_act || in_reset) ? 1 : 0;
    n_need += need; n_step += has_act;
  }
  int rc;
  if (s1 != st) {   // the step pipeline starts after everything the previous tick left on `stream` (its select / plan)
    if (cudaEventRecord(h->ev_fork, st) != cudaSuccess || cudaStreamWaitEvent(s1, h->ev_fork, 0) != cudaSuccess) return DGE_ECUDA;
  }
  // ---- policy pipeline, part 1 (async): graph kernels + pack; the batch follows its header to the host at once, sized by the previous
  // batch (+ margin): one synchronisation instead of two
  auto issue_graph = [&]() -> int {
    if (!n_need) return DGE_OK;
    int r = dge_graph_host_packed_begin(h, hl->need, g, hl->arena_pack, hl->arena_cap, stream);
    if (r) return r;
    hl->launches += 5; hl->h2d_bytes += B;
    if (hl->prefetch_guess > hl->arena_cap) hl->prefetch_guess = hl->arena_cap;
    if (hl->prefetch_guess > 128) r = dge_graph_host_packed_prefetch(h, hl->arena_pack, hl->arena_host, hl->prefetch_guess, stream);
    return r;
  };
  // ---- step pipeline (async on s1): restart finished episodes, one simulator step from the host's action lists
  auto issue_step = [&]() -> int {
    int r = dge_reset_done_queued(h, seed_stride, forced_odom_host, n_forced, s1);
    if (r) return r;
    if ((r = dge_step_host_plans_async(h, hl->plans, hl->cursor, hl->mask, hl->done, hl->obs, hl->metrics, DGE_STEP_NO_SYNC | DGE_STEP_HONOR_FORCED, s1))) return r;
    hl->launches += 6;
    hl->h2d_bytes += (int64_t)B * 3 * sizeof(double) + B;
    hl->d2h_bytes += B + (int64_t)B * 8 * sizeof(double) + (hl->obs ? hl->obs_bytes : 0);
    return DGE_OK;
  };
  // Which of the two is issued first: the step's chain (H2D actions, four kernels, 2 MB of maps D2H: ~0.25 ms) is as long as the policy's, and it
  // is the one the tick ends on (profiles/r02_hostloop_host_profile_v1.txt: 38 us of join wait), so it goes first: 0.277 vs 0.284 ms per tick
  // (DGE_HOST_TICK_STEP_FIRST=0 for the A/B).  Both orders give the same results, the env sets are disjoint.
  static const int step_first = [] { const char *v = getenv("DGE_HOST_TICK_STEP_FIRST"); return (v && v[0] == '0') ? 0 : 1; }();
  // (an error after the first asynchronous launch: both streams are drained before the call returns, so that no copy into the caller's
  //  pinned buffers is still in flight when the caller unwinds)
  auto bail = [&](int code) -> int { cudaStreamSynchronize(st); if (s1 != st) cudaStreamSynchronize(s1); return code; };
  if ((rc = step_first ? issue_step() : issue_graph())) return bail(rc);
  if ((rc = step_first ? issue_graph() : issue_step())) return bail(rc);
  for (int b = 0; b < B; ++b) {
    if (hl->phase[b] > 0) hl->phase[b] -= 1;
    else if (!hl->need[b]) hl->cursor[b] += 1;
  }
  hl->n_stepped = n_step;
  // ---- policy pipeline, part 2: the batch crosses to the host and back, Q-values come to the host, the host picks the frontiers
  if (n_need) {
    dge_graph_packed pk;
    if ((rc = dge_graph_host_packed_end_prefetched(h, hl->arena_pack, hl->arena_host, hl->arena_cap, hl->prefetch_guess, &pk, stream))) return bail(rc);
    hl->d2h_bytes += pk.total_bytes > hl->prefetch_guess ? pk.total_bytes : (hl->prefetch_guess > 128 ? hl->prefetch_guess : pk.total_bytes);   // bytes that crossed the bus
    if (hl->prefetch_guess > 0) hl->prefetch_guess = pk.total_bytes + pk.total_bytes / 4 + 16384;   // (0 = the caller switched the prefetch off)
    const int n = pk.n_nodes;
    if (pk.n_graphs > 0) {
      if (n > pol->node_cap) return bail(DGE_ECAP);
      if (cudaMemcpyAsync(hl->arena_dev, hl->arena_host, (size_t)pk.total_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return bail(DGE_ECUDA);
      hl->h2d_bytes += pk.total_bytes;
      const unsigned char *ad = static_cast<const unsigned char *>(hl->arena_dev);
      const int grc = dge_gcn_q_forward(n, pol->Cin, pol->C, reinterpret_cast<const float *>(ad + pk.x), reinterpret_cast<const int32_t *>(ad + pk.csr_rowptr),
                                        reinterpret_cast<const int32_t *>(ad + pk.csr_perm), reinterpret_cast<const int64_t *>(ad + pk.edge_index),
                                        reinterpret_cast<const float *>(ad + pk.gcn_norm), reinterpret_cast<const float *>(ad + pk.gcn_selfnorm), pol->W1, pol->b1,
                                        pol->W2t_hi, pol->W2t_lo, pol->b2, pol->head_w, pol->head_b_dev, pol->ws, pol->q, stream);
      if (grc) return bail(grc == -1 ? DGE_EINVAL : DGE_ECUDA);
      if (cudaMemcpyAsync(hl->q_host, pol->q, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) return bail(DGE_ECUDA);
      if (cudaStreamSynchronize(st) != cudaSuccess) return bail(DGE_ECUDA);
      hl->launches += 3; hl->d2h_bytes += (int64_t)n * sizeof(float);
      // host arg-max; the chosen frontier's plan is in the batch (frontier_plan): no device round trip (the rare no-frontier case excepted)
      if ((rc = dge_select_plan_host(h, hl->arena_host, &pk, hl->q_host, hl->need, hl->plan_host, hl->choice_host, stream))) return bail(rc);
      for (int b = 0; b < B; ++b) {
        if (!hl->need[b]) continue;
        if (hl->choice_host[b] < 0) hl->phase[b] = n_forced + 1;        // no frontier left (q15): the episode is over, its reset takes n_forced + 1 ticks
        for (int i = 0; i < 6; ++i) hl->plans[6 * b + i] = hl->plan_host[6 * b + i];
        hl->cursor[b] = 0;
      }
      hl->n_graphs = pk.n_graphs; hl->n_nodes = n;
    }
  }
  // ---- join: the step's host buffers are valid after this
  if (cudaStreamSynchronize(s1) != cudaSuccess) return DGE_ECUDA;
  for (int b = 0; b < B; ++b) {
    if (hl->done[b]) { hl->phase[b] = n_forced + 1; hl->plans[6 * b + 5] = 0.0; hl->cursor[b] = 0; }
  }
  return DGE_OK;
}
