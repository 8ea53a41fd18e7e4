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
#include <cstring>

#include "dge_internal.cuh"

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
  if (cudaStreamCreateWithFlags(&e->tick_cap_stream, cudaStreamNonBlocking) != cudaSuccess) return DGE_ECUDA;
  for (cudaEvent_t *ev : {&e->ev_fork, &e->ev_move, &e->ev_join})
    if (cudaEventCreateWithFlags(ev, cudaEventDisableTiming) != cudaSuccess) return DGE_ECUDA;
  return DGE_OK;
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
  if ((rc = dge_launch_slam(e, e->active, s1))) return rc;
  if ((rc = dge_launch_vmap(e, e->active, s1))) return rc;
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
  if (e->tick_cap_stream) { cudaStreamDestroy(e->tick_cap_stream); e->tick_cap_stream = nullptr; }
  for (cudaEvent_t *ev : {&e->ev_fork, &e->ev_move, &e->ev_join})
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
