"""The graph kernel's in-place CSR + GCN normalisation (used by the fast policy path) must give the same
Q-values as the generic edge-list preprocessing of gnn.GraphStructure, and the fused first layer the same
as the unfused one."""
import numpy as np
import pytest
import torch

from drl_graph_exploration_b200.config import EnvConfig

pytestmark = pytest.mark.gpu


def test_engine_csr_equals_generic_preprocessing():
    from drl_graph_exploration_b200 import Networks, gnn
    from drl_graph_exploration_b200.data import Data
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv

    env = VecExplorationEnv(32, cfg=EnvConfig(map_size=20, num_landmarks=30), max_poses=96)
    env.reset()
    torch.manual_seed(0)
    model = Networks.GCN().to(env.device).eval()
    # fused vs unfused is a statement about the fusion: both sides use the same GEMM (the unfused path runs under autograd,
    # whose GEMM is the tcgen05 one too -- the default training mode)
    Networks.set_matmul_precision("tc3", train="tc3")
    try:
        _run(env, model, Data)
    finally:
        env.close()


def _run(env, model, Data):
    for it in range(12):
        need = env.needs_decision()
        g = env.build_graph(need)
        ng, n, e = g.sync_sizes()
        if g.n_done:
            env.reset_done()
        if ng:
            d_fast = g.data()
            from drl_graph_exploration_b200 import Networks
            d_ref = Data(d_fast.x.clone(), d_fast.edge_index.clone(), d_fast.edge_attr.clone())
            Networks._structure(d_ref, d_ref.x.size(0))        # the generic preprocessing (CSR builds + normalisation kernels), cached on the data
            d_coo = Data(d_fast.x.clone(), d_fast.edge_index.clone(), d_fast.edge_attr.clone())   # a raw edge list: the single-call route
            with torch.no_grad():
                q_fast = model(d_fast, 0.0)
                q_ref = model(d_ref, 0.0)
                q_coo = model(d_coo, 0.0)
                with torch.enable_grad():
                    q_unfused = model(Data(d_fast.x.clone(), d_fast.edge_index.clone(), d_fast.edge_attr.clone()), 0.0).detach()
            assert getattr(d_coo, "_dge_structure", None) is None and torch.equal(q_coo, q_ref)      # same kernels, same order: identical
            gs_f, gs_r = d_fast._dge_structure, d_ref._dge_structure
            assert torch.equal(gs_f.rowptr_dst, gs_r.rowptr_dst) and torch.equal(gs_f.perm_dst[:e], gs_r.perm_dst[:e])
            nf, sf = gs_f.gcn_norm(True); nr, sr = gs_r.gcn_norm(True)
            assert torch.allclose(nf[:e], nr[:e], rtol=2e-6, atol=0) and torch.allclose(sf, sr, rtol=2e-6, atol=0)
            scale = q_ref.abs().max()
            assert (q_fast - q_ref).abs().max() <= 2e-5 * scale
            assert (q_fast - q_unfused).abs().max() <= 2e-5 * scale
            env.select_and_plan(q_fast, need)
        env.step_queued()
