"""gnn.QForwardPlan on the GPU: the prepared Q forward of the acting loop returns the very Q-values of Networks.GCN.forward on
the same graph batch, and runner.PolicyLoop leaves the same engine state with and without it.  (Written after the last GPU
session of round 1 -- tests/test_q_plan_cpu.py holds the argument lists equal on CPU -- hence sorted behind the suites that have
been seen green on a B200.)"""
import pytest
import torch

from drl_graph_exploration_b200.config import EnvConfig

pytestmark = pytest.mark.gpu


def test_plan_returns_the_q_values_of_the_module_and_the_loop_is_unchanged():
    from drl_graph_exploration_b200 import Networks, gnn
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
    from drl_graph_exploration_b200.runner import PolicyLoop

    cfg = EnvConfig(map_size=20, num_landmarks=12)
    torch.manual_seed(0)
    model = Networks.GCN().cuda().eval()
    assert gnn.QForwardPlan.eligible(model)

    def run(use_plan, ticks=25):
        env = VecExplorationEnv(16, cfg=cfg, max_poses=96, device=0, seed0=0)
        env.reset()
        loop = PolicyLoop(env, model, device_tick=False)      # the per-launch route (the device tick has its own test)
        assert loop._plan is not None
        if not use_plan:
            loop._plan = None
        for _ in range(ticks):
            loop.tick()
        torch.cuda.synchronize()
        assert (loop._plan is not None) == use_plan          # the prepared route really ran (it disables itself, loudly, on a glue error)
        st = env.eng.state
        out = {k: st[k].clone() for k in ("n_poses", "sim_step", "plan", "plan_cursor", "est_pose", "prob", "metrics")}
        # one more decision round by hand: plan vs module on the same batch
        with torch.no_grad():
            g = env.build_graph(env.mark_pending().clone())
            ng, n, _ = g.sync_sizes()
            q_mod = model(g.data(), 0.0).view(-1).clone() if ng else None
            q_plan = gnn.QForwardPlan(model, g)(n).clone() if ng else None
        env.close()
        return out, q_mod, q_plan

    a, q_mod, q_plan = run(True)
    b, _, _ = run(False)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    if q_mod is not None:
        assert torch.equal(q_mod, q_plan)
