"""dev: one C4-shaped launch of the covariance-propagation kernel (for ncu): n envs, T poses."""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
from synth import synth_states
from drl_graph_exploration_b200.config import EnvConfig
from drl_graph_exploration_b200.engine import load_library, _ptr, _stream_ptr
n, T = int(sys.argv[1]), int(sys.argv[2])
cfg = EnvConfig(map_size=60, num_landmarks=200); cs = cfg.to_struct(); L, V = 200, cfg.rows * cfg.cols
dev = torch.device("cuda"); lib = load_library()
pose, cov, cov6, info, lm = synth_states(cfg, n, T, L, seed=T)
tp, tc, tl = (torch.as_tensor(a, device=dev).contiguous() for a in (pose, cov6, lm))
prob = torch.empty(n, V, dtype=torch.float64, device=dev); vinfo = torch.empty(n, V, 3, dtype=torch.float64, device=dev)
ws = torch.zeros(lib.dge_virtual_map_rebuild_ws_doubles(n, T), dtype=torch.float64, device=dev)
for _ in range(3):
    lib.dge_virtual_map_rebuild(ctypes.byref(cs), n, T, _ptr(tp), _ptr(tc), L, _ptr(tl), _ptr(prob), _ptr(vinfo), None, _ptr(ws), _stream_ptr(dev))
torch.cuda.synchronize()
