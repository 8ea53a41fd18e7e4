"""HostPolicyLoop (host-buffer C ABI: dge_step_host_async / dge_graph_host / dge_line_plan_host) against the
device-resident PolicyLoop: same seeds, same policy => every env goes through the same sequence of operations
(test.py:100-143 per env), so after K ticks the engine states must agree.  The host path re-uploads the graph and
builds its own CSR, so Q-values may differ in the last bit; states are compared exactly for the integer fields and
to 1e-9 for fp64 (a flipped arg-max would show up as a different trajectory length)."""
import numpy as np
import pytest
import torch

from drl_graph_exploration_b200.config import EnvConfig

pytestmark = pytest.mark.gpu


def _mk(n, seed0=300):
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
    env = VecExplorationEnv(n, cfg=EnvConfig(map_size=20, num_landmarks=30), max_poses=96, seed0=seed0)
    env.reset()
    return env


@pytest.mark.parametrize("overlap,packed,native", [(False, True, None), (True, True, None), (True, True, False), (True, False, False)])
def test_host_loop_matches_device_loop(overlap, packed, native):
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.runner import HostPolicyLoop, PolicyLoop
    a, b = _mk(24), _mk(24)
    torch.manual_seed(0)
    model = Networks.GCN().to(a.device).eval()
    dev_loop = PolicyLoop(a, model, overlap=overlap)
    host_loop = HostPolicyLoop(b, model, overlap=overlap, native=native)   # None: the tick is ONE native call (dge_host_policy_tick); False: issued from Python
    host_loop.packed = packed                      # one packed transfer per direction (default) / one copy per array
    K = 120                                        # long enough for several episodes to end and restart in-pipeline
    for _ in range(K):
        dev_loop.tick()
        host_loop.tick()
    torch.cuda.synchronize()
    sa, sb = a.eng.state, b.eng.state
    assert int(sa["counters"][3]) > 0              # some episodes restarted
    for f in ("n_poses", "sim_step", "update_count", "meas_ptr", "observed", "seed", "seen"):
        assert torch.equal(sa[f], sb[f]), f
    assert int(sa["counters"][0]) == int(sb["counters"][0]) == host_loop.steps
    T = sa["n_poses"].cpu().numpy()
    ea, eb = sa["est_pose"].cpu().numpy(), sb["est_pose"].cpu().numpy()
    for i in range(a.B):
        np.testing.assert_allclose(ea[i, :T[i]], eb[i, :T[i]], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(sa["prob"].cpu().numpy(), sb["prob"].cpu().numpy(), rtol=0, atol=0)
    # the host buffers hold what the step returned: done flags, metrics, occupancy maps
    np.testing.assert_array_equal(host_loop.done, sb["done"].cpu().numpy())
    np.testing.assert_array_equal(host_loop.t_obs.numpy(), sb["prob"].cpu().numpy())
    np.testing.assert_array_equal(host_loop.metrics[:, 0], sb["metrics"][:, 0].cpu().numpy())
    assert host_loop.h2d > 0 and host_loop.d2h > 0 and host_loop.graphs == dev_loop.graphs
    assert (host_loop._native is not None) == (native is None and packed)      # the route that was meant to run did run
    a.close(); b.close()


def test_graph_host_equals_device_graph():
    """dge_graph_host: the host batch is the valid prefix of the device batch, edge_index contiguous [2,E]."""
    import ctypes
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.runner import HostPolicyLoop
    env = _mk(8, seed0=7)
    loop = HostPolicyLoop(env, Networks.GCN().to(env.device).eval(), overlap=False)
    loop.need[:] = 1
    loop.need[3] = 0
    mp = ctypes.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
    rc = loop._L.dge_graph_host(env.eng._h, loop.t_need.data_ptr(), ctypes.byref(env.graph.c), ctypes.byref(loop._ho), mp)
    assert rc == 0
    ng, n, e = (int(v) for v in loop.t_tot[:3])
    assert ng == 7 and n > 0 and e > 0
    g = env.graph
    assert torch.equal(loop.t_x[:n], g.x[:n].cpu())
    assert torch.equal(loop.t_ei[:2 * e].view(2, e), g.edge_index[:, :e].cpu())
    assert torch.equal(loop.t_ea[:e], g.edge_attr[:e].cpu())
    assert torch.equal(loop.t_nptr[:ng + 1], g.node_ptr[:ng + 1].cpu())
    assert torch.equal(loop.t_ks[:ng], g.key_size[:ng].cpu()) and torch.equal(loop.t_fs[:ng], g.fro_size[:ng].cpu())
    assert torch.equal(loop.t_fxy, g.frontier_xy.cpu())
    env.close()


def test_packed_graph_transfer_equals_device_graph():
    """dge_graph_host_packed_begin/end: every section of the arena is the valid prefix of the device batch (frontier
    coordinates by graph ordinal), and the header reports the totals."""
    import ctypes
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.runner import HostPolicyLoop
    env = _mk(8, seed0=11)
    loop = HostPolicyLoop(env, Networks.GCN().to(env.device).eval(), overlap=False)
    loop.need[:] = 1
    loop.need[5] = 0
    mp = ctypes.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
    L, pk, g = loop._L, loop._pk, env.graph
    assert L.dge_graph_host_packed_begin(env.eng._h, loop.t_need.data_ptr(), ctypes.byref(g.c), loop.a_pack.data_ptr(), loop.arena_cap, mp) == 0
    assert L.dge_graph_host_packed_end(env.eng._h, loop.a_pack.data_ptr(), loop.a_host.data_ptr(), loop.arena_cap, ctypes.byref(pk), mp) == 0
    ng, n, e = pk.n_graphs, pk.n_nodes, pk.n_edges
    assert ng == 7 and n > 0 and e > 0 and pk.total_bytes <= loop.arena_cap
    hv = lambda off, cnt, dt, sz: loop.a_host[off:off + cnt * sz].view(dt)
    f32, i32, i64, f64 = torch.float32, torch.int32, torch.int64, torch.float64
    assert torch.equal(hv(pk.x, n * 5, f32, 4).view(n, 5), g.x[:n].cpu())
    assert torch.equal(hv(pk.edge_index, 2 * e, i64, 8).view(2, e), g.edge_index[:, :e].cpu())
    assert torch.equal(hv(pk.edge_attr, e, f32, 4), g.edge_attr[:e].cpu())
    assert torch.equal(hv(pk.node_ptr, ng + 1, i32, 4), g.node_ptr[:ng + 1].cpu()) and torch.equal(hv(pk.edge_ptr, ng + 1, i32, 4), g.edge_ptr[:ng + 1].cpu())
    assert torch.equal(hv(pk.key_size, ng, i32, 4), g.key_size[:ng].cpu()) and torch.equal(hv(pk.fro_size, ng, i32, 4), g.fro_size[:ng].cpu())
    envs = [b for b in range(8) if b != 5]
    assert torch.equal(hv(pk.frontier_xy, ng * (env.eng.Lt + 1) * 2, f64, 8).view(ng, env.eng.Lt + 1, 2), g.frontier_xy[envs].cpu())
    assert torch.equal(hv(pk.csr_rowptr, n + 1, i32, 4), g.csr_rowptr[:n + 1].cpu()) and torch.equal(hv(pk.csr_perm, e, i32, 4), g.csr_perm[:e].cpu())
    assert torch.equal(hv(pk.gcn_norm, e, f32, 4), g.gcn_norm[:e].cpu()) and torch.equal(hv(pk.gcn_selfnorm, n, f32, 4), g.gcn_selfnorm[:n].cpu())
    # the plan of every frontier travels with the batch: frontier f = 0 of every graph against dge_line_plan on the same goal
    F = env.eng.Lt + 1
    plans = hv(pk.frontier_plan, ng * F * 6, f64, 8).view(ng, F, 6)
    ref_plan = env.line_plan(g.frontier_xy[:, 0].contiguous()).cpu()
    fro = hv(pk.fro_size, ng, i32, 4)
    for gi, bi in enumerate(envs):
        if int(fro[gi]) > 0:
            assert torch.equal(plans[gi, 0], ref_plan[bi]), gi
        assert bool((plans[gi, int(fro[gi]):] == 0).all())
    hdr = loop.a_host[:48].view(i64)
    assert hdr[:3].tolist() == [ng, n, e] and int(hdr[5]) == pk.total_bytes
    env.close()


@pytest.mark.parametrize("prefetch", [False, True])
def test_native_host_tick_equals_the_python_issued_tick(prefetch):
    """dge_host_policy_tick against the same sequence of C-ABI calls issued from Python: identical host state (plans, cursors, reset
    phases), identical traffic and launch counts, identical engine state, tick by tick.  With the prefetch (the batch follows its
    header to the host at once, sized by the previous batch) the D2H count may only be larger."""
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.runner import HostPolicyLoop
    a, b = _mk(16, seed0=40), _mk(16, seed0=40)
    torch.manual_seed(0)
    model = Networks.GCN().to(a.device).eval()
    nat, py = HostPolicyLoop(a, model, overlap=True), HostPolicyLoop(b, model, overlap=True, native=False)
    nat.prefetch = prefetch
    for t in range(90):
        na, nb = nat.tick(), py.tick()
        assert na == nb, t
        np.testing.assert_array_equal(nat.plans, py.plans); np.testing.assert_array_equal(nat.cursor, py.cursor)
        np.testing.assert_array_equal(nat.phase, py.phase)
    assert (nat.steps, nat.graphs, nat.h2d, nat.launches) == (py.steps, py.graphs, py.h2d, py.launches)
    assert nat.d2h == py.d2h if not prefetch else py.d2h <= nat.d2h <= 3 * py.d2h
    torch.cuda.synchronize()
    sa, sb = a.eng.state, b.eng.state
    for f in ("n_poses", "sim_step", "meas_ptr", "observed", "seed", "seen", "est_pose", "prob"):
        assert torch.equal(sa[f], sb[f]), f
    a.close(); b.close()
