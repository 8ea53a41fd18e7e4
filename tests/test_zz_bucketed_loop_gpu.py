"""runner.BucketedPolicyLoop (EXPERIMENTAL, opt-in: DGE_EXPERIMENTAL=1) against runner.PolicyLoop.

The bucketed loop changes only WHEN an env takes its ticks (short trajectories tick more often than long ones, on their own
stream), never what a tick does: after n ticks of its bucket an env must be, bit for bit, in the state the lock-step loop
leaves it in after n ticks.  The file sorts last and is skipped by default: the loop was written after the last GPU session of
round 1 and has not run on a GPU yet."""
import os

import pytest
import torch

from drl_graph_exploration_b200.config import EnvConfig

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("DGE_EXPERIMENTAL") != "1", reason="opt-in: DGE_EXPERIMENTAL=1")]


def _signature(st):
    """Per-env state digest [B, 9] f64: counters, last pose estimate, map and metric sums (all deterministic per env)."""
    n = st["n_poses"].long()
    B = n.numel()
    last = st["est_pose"][torch.arange(B, device=n.device), (n - 1).clamp(min=0)]
    return torch.cat([n.double().view(B, 1), st["sim_step"].double().view(B, 1), st["seed"].double().view(B, 1), last.double(),
                      st["prob"].double().sum(dim=1, keepdim=True), st["metrics"].double().sum(dim=1, keepdim=True),
                      st["plan_cursor"].double().view(B, 1)], dim=1)


def test_bucketed_ticks_leave_every_env_where_lockstep_ticks_do():
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
    from drl_graph_exploration_b200.runner import BucketedPolicyLoop, PolicyLoop

    B, rounds, short_ticks = 24, 10, 3
    cfg = EnvConfig(map_size=20, num_landmarks=12)
    torch.manual_seed(0)
    model = Networks.GCN().cuda().eval()

    env = VecExplorationEnv(B, cfg=cfg, max_poses=96, device=0, seed0=0)
    env.reset()
    lock = PolicyLoop(env, model)
    sigs = [_signature(env.eng.state).clone()]
    for _ in range(rounds * short_ticks):
        lock.tick()
        torch.cuda.synchronize()
        sigs.append(_signature(env.eng.state).clone())
    sigs = torch.stack(sigs)                                   # [ticks + 1, B, 9]
    steps_lock = int(env.eng.state["counters"][0])
    env.close()

    env = VecExplorationEnv(B, cfg=cfg, max_poses=96, device=0, seed0=0)
    env.reset()
    bl = BucketedPolicyLoop(env, model, short_ticks=short_ticks, short_fraction=0.5)
    for _ in range(rounds):
        bl.round()
    bl.join()
    torch.cuda.synchronize()
    n = bl.ticks_of
    assert int(n.min()) >= rounds and int(n.max()) <= rounds * short_ticks and int(n.min()) < int(n.max())
    got = _signature(env.eng.state)
    want = sigs[n, torch.arange(B, device=n.device)]
    bad = (got != want).any(dim=1).nonzero().view(-1).tolist()
    assert not bad, (bad, n[bad].tolist(), got[bad].tolist(), want[bad].tolist())
    steps = int(env.eng.state["counters"][0])
    print(f"bucketed: {steps} env-steps in {rounds} rounds (lock step: {steps_lock} in {rounds * short_ticks} ticks), ticks per env {int(n.min())}..{int(n.max())}")
    env.close()
