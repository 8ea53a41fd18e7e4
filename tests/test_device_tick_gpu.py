"""dge_policy_tick (include/dge.h): the sync-free tick of the acting loop -- one native call, replayed as a CUDA graph -- leaves
exactly the engine state of the per-launch schedule (runner.PolicyLoop with device_tick=False: ~20 ctypes calls and one size sync
per tick), through episode ends and in-pipeline restarts.  Both replace the loop body of the reference's test.py:100-143."""
import pytest
import torch

from drl_graph_exploration_b200.config import EnvConfig

pytestmark = pytest.mark.gpu

KEYS = ("n_poses", "sim_step", "true_pose", "est_pose", "prob", "seen", "counters", "seed", "plan", "plan_cursor", "forced", "done", "metrics")


def _run(route, ticks, n_envs=48, landmarks=12):
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
    from drl_graph_exploration_b200.runner import PolicyLoop
    cfg = EnvConfig(map_size=20, num_landmarks=landmarks)
    env = VecExplorationEnv(n_envs, cfg=cfg, max_poses=96, device=0, seed0=0)
    env.reset()
    torch.manual_seed(0)
    model = Networks.GCN().to(env.device).eval()
    kw = {"launches": dict(device_tick=False), "eager": dict(device_tick=True, cuda_graph=False), "graph": dict(device_tick=True, cuda_graph=True),
          "graph_one_stream": dict(device_tick=True, cuda_graph=True, overlap=False)}[route]
    loop = PolicyLoop(env, model, **kw)
    assert (loop.device is not None) == (route != "launches")
    for _ in range(ticks):
        loop.tick()
    torch.cuda.synchronize()
    st = env.eng.state
    out = {k: st[k].clone() for k in KEYS}
    graphs = loop.graphs
    env.close()
    return out, graphs


def test_device_tick_equals_the_per_launch_schedule():
    ref, g_ref = _run("launches", 150)
    assert int(ref["counters"][3]) >= 10 and int(ref["counters"][0]) > 48 * 100      # restarts happened, most ticks were policy steps
    assert g_ref == int(ref["counters"][4]) > 0
    for route in ("eager", "graph", "graph_one_stream"):
        out, graphs = _run(route, 150)
        for k in KEYS:
            assert torch.equal(ref[k], out[k]), (route, k)
        assert graphs == g_ref


def test_graph_replay_follows_a_weight_update():
    """The captured tick holds pointers, not values: an in-place weight update is seen by the next replay; a re-allocated
    parameter re-derives the arguments (and re-captures)."""
    from drl_graph_exploration_b200 import Networks
    from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
    from drl_graph_exploration_b200.runner import PolicyLoop
    cfg = EnvConfig(map_size=20, num_landmarks=12)
    outs = []
    for route in ("graph", "launches"):
        env = VecExplorationEnv(16, cfg=cfg, max_poses=96, device=0, seed0=3)
        env.reset()
        torch.manual_seed(0)
        model = Networks.GCN().to(env.device).eval()
        loop = PolicyLoop(env, model, device_tick=(route == "graph"))
        for i in range(40):
            if i == 15:
                with torch.no_grad():
                    model.conv2.weight.mul_(-1.0)          # in place: version counter moves, storage stays
            if i == 28:
                model.fully_con1.weight = torch.nn.Parameter(-model.fully_con1.weight.detach().clone())   # new storage
            loop.tick()
        torch.cuda.synchronize()
        outs.append({k: env.eng.state[k].clone() for k in ("n_poses", "plan", "plan_cursor", "prob")})
        env.close()
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k
