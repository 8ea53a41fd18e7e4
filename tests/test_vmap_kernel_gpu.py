"""Kernel-level parity of the covariance-propagation kernel (a6+a7) on synthetic belief states
(BASELINE config C4 shape at reduced batch): CUDA `dge_virtual_map_rebuild` vs the CPU oracle's
restatement of VirtualMap::updateProbability / updateInformation."""
import numpy as np
import pytest
import torch

from helpers import borderline_cells, sym3_to_full
from synth import synth_states
from drl_graph_exploration_b200.config import EnvConfig
from oracle.oracle import virtual_map_rebuild as oracle_rebuild

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("map_size,T,L,n", [(60, 64, 200, 6), (60, 300, 200, 3), (20, 40, 30, 6), (100, 100, 50, 2)])
def test_rebuild_matches_oracle(map_size, T, L, n):
    from drl_graph_exploration_b200.engine import virtual_map_rebuild

    cfg = EnvConfig(map_size=map_size, num_landmarks=L)
    pose, cov, cov6, info, lm = synth_states(cfg, n, T, L, seed=T)
    dev = torch.device("cuda")
    prob, vinfo, seen = virtual_map_rebuild(cfg, torch.as_tensor(pose, device=dev), torch.as_tensor(cov6, device=dev),
                                            torch.as_tensor(lm, device=dev), want_seen=True)
    torch.cuda.synchronize()
    prob, vinfo, seen = prob.cpu().numpy(), vinfo.cpu().numpy(), seen.cpu().numpy()
    nb = 0
    for i in range(n):
        p_ref, i_ref, s_ref = oracle_rebuild(cfg, pose[i], info[i].reshape(T, 9), lm[i])
        ok = ~borderline_cells(cfg, pose[i], tol=1e-9)
        nb += int((~ok).sum())
        assert np.array_equal(seen[i][ok], s_ref[ok]), "integer visibility counts"
        assert np.array_equal(prob[i][ok], p_ref[ok]), "occupancy probabilities (bit-exact, closed value set)"
        a, b = sym3_to_full(vinfo[i])[ok], i_ref[ok]
        err = np.abs(a - b) / (1e-9 + np.maximum(np.abs(a), np.abs(b)))
        assert err.max() < 1e-6, err.max()
    assert nb <= 4 * n


def test_rebuild_properties_full_size():
    """C4 full size (1024 envs would be ~100 MB of output; 256 here): size-independent properties --
    SPD cells, untouched cells keep the prior, idempotence, permutation of the batch."""
    from drl_graph_exploration_b200.engine import virtual_map_rebuild

    cfg = EnvConfig(map_size=60, num_landmarks=200)
    n, T, L = 256, 128, 200
    pose, cov, cov6, info, lm = synth_states(cfg, n, T, L, seed=9)
    dev = torch.device("cuda")
    tp, tc, tl = (torch.as_tensor(a, device=dev) for a in (pose, cov6, lm))
    prob, vinfo, seen = virtual_map_rebuild(cfg, tp, tc, tl, want_seen=True)
    prob2, vinfo2, _ = virtual_map_rebuild(cfg, tp, tc, tl)
    assert torch.equal(prob, prob2) and torch.equal(vinfo, vinfo2)   # deterministic / idempotent
    perm = torch.randperm(n, device=dev)
    prob3, vinfo3, _ = virtual_map_rebuild(cfg, tp[perm].contiguous(), tc[perm].contiguous(), tl[perm].contiguous())
    assert torch.equal(prob3, prob[perm]) and torch.equal(vinfo3, vinfo[perm])
    det = vinfo[..., 0] * vinfo[..., 2] - vinfo[..., 1] ** 2
    assert (det > 0).all() and (vinfo[..., 0] > 0).all()
    untouched = seen == 0
    assert (prob[untouched] == 0.5).all()
    assert ((seen > 0) | (seen == -1) | untouched).all()


@pytest.mark.parametrize("map_size,T,L,n", [(60, 48, 200, 96), (100, 40, 50, 70)])
def test_band_group_split_equals_whole_env_kernel_and_oracle(map_size, T, L, n):
    """Throughput regime (n x bands >= 8 x 148): every env is split into band groups of 4 warps.  Same arithmetic, same order
    per cell => bit-identical to the whole-env kernel (which a small batch selects), and oracle parity on a sample."""
    from drl_graph_exploration_b200.engine import virtual_map_rebuild

    cfg = EnvConfig(map_size=map_size, num_landmarks=L)
    pose, cov, cov6, info, lm = synth_states(cfg, n, T, L, seed=1000 + T)
    dev = torch.device("cuda")
    tp, tc, tl = (torch.as_tensor(a, device=dev) for a in (pose, cov6, lm))
    prob, vinfo, seen = virtual_map_rebuild(cfg, tp, tc, tl, want_seen=True)                       # split
    k = 5
    prob_w, vinfo_w, seen_w = virtual_map_rebuild(cfg, tp[:k].contiguous(), tc[:k].contiguous(), tl[:k].contiguous(), want_seen=True)   # whole-env CTAs
    assert torch.equal(prob[:k], prob_w) and torch.equal(vinfo[:k], vinfo_w) and torch.equal(seen[:k], seen_w)
    prob, vinfo, seen = prob.cpu().numpy(), vinfo.cpu().numpy(), seen.cpu().numpy()
    for i in (0, n // 2, n - 1):
        p_ref, i_ref, s_ref = oracle_rebuild(cfg, pose[i], info[i].reshape(T, 9), lm[i])
        ok = ~borderline_cells(cfg, pose[i], tol=1e-9)
        assert np.array_equal(seen[i][ok], s_ref[ok]) and np.array_equal(prob[i][ok], p_ref[ok])
        a, b = sym3_to_full(vinfo[i])[ok], i_ref[ok]
        assert (np.abs(a - b) / (1e-9 + np.maximum(np.abs(a), np.abs(b)))).max() < 1e-6
