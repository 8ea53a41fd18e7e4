"""dev: lock-step HostPolicyLoop vs PolicyLoop, report the first tick where an env diverges."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from drl_graph_exploration_b200 import Networks
from drl_graph_exploration_b200.config import EnvConfig
from drl_graph_exploration_b200.envs.exploration_env import VecExplorationEnv
from drl_graph_exploration_b200.runner import HostPolicyLoop, PolicyLoop

def mk(n, seed0=300):
    env = VecExplorationEnv(n, cfg=EnvConfig(map_size=20, num_landmarks=30), max_poses=96, seed0=seed0); env.reset(); return env
a, b = mk(24), mk(24)
torch.manual_seed(0)
model = Networks.GCN().to(a.device).eval()
dl, hl = PolicyLoop(a, model, overlap=False), HostPolicyLoop(b, model, overlap=False)
for t in range(200):
    dl.tick(); hl.tick(); torch.cuda.synchronize()
    ca = a._choice.cpu().numpy()
    na, nb = a.eng.state["n_poses"].cpu().numpy(), b.eng.state["n_poses"].cpu().numpy()
    if hasattr(hl, "last_choice"):
        envs, ch = hl.last_choice
        bad = [(int(e), int(c), int(ca[e])) for e, c in zip(envs, ch) if ca[e] != c and ca[e] >= 0]
        if bad:
            print("tick", t, "choice mismatch (env, host, dev):", bad)
            ng, n, e_ = (int(v) for v in hl.t_tot[:3])
            for (e, c, d) in bad:
                gi = list(envs).index(e)
                k, f, n0 = int(hl.ks[gi]), int(hl.fs[gi]), int(hl.nptr[gi])
                print("  host q frontier:", hl.q[n0 + k:n0 + k + f])
            break
    if not np.array_equal(na, nb):
        d = np.nonzero(na != nb)[0]
        print("tick", t, "n_poses differ at envs", d, na[d], nb[d], "host phase", hl.phase[d], "cursor", hl.cursor[d], "plans", hl.plans[d],
              "dev forced", a.eng.state["forced"][d].cpu().numpy(), b.eng.state["forced"][d].cpu().numpy(),
              "done", a.eng.state["done"][d].cpu().numpy(), b.eng.state["done"][d].cpu().numpy())
        break
else:
    print("no divergence in 200 ticks; restarts", int(a.eng.state["counters"][3]), int(b.eng.state["counters"][3]))
