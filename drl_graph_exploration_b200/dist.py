"""Multi-GPU plumbing: envs are independent, so the path shards by env index
across ranks with no data-path collective; DQN training adds exactly ONE all-reduce of the flat
fp32 gradient bucket per step (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_envs(n_envs_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous env-index range [lo, hi) of `rank` (env e -> rank e // ceil(B/W))."""
    per = (n_envs_total + world - 1) // world
    lo = min(n_envs_total, rank * per)
    return lo, min(n_envs_total, lo + per)


def env_seeds(lo: int, hi: int, episode: int = 0, stride: int = 1 << 20) -> torch.Tensor:
    """Disjoint Philox keys per (env, episode)."""
    return torch.arange(lo, hi, dtype=torch.int64) + episode * stride


class FlatGradBucket:
    """One contiguous fp32 buffer holding every parameter's gradient (views), so a train step
    issues a single all-reduce (GCN: 1.008 M params = 4.0 MB -- latency-bound on NVLink)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        """sum over ranks, then x 1/W; the reference's element-wise clamp (policy.py:251-252) is
        applied AFTER this so world_size = 1 keeps the single-GPU semantics."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / dist.get_world_size())

    def clamp_(self, lim: float):
        self.flat.clamp_(-lim, lim)


class NativeAdam:
    """``torch.optim.Adam(params, lr)`` (amsgrad off, no weight decay) with the reference's element-wise gradient clamp
    (policy.py:251-253) fused in, as ONE kernel over flat buffers (``dge_clamp_adam_step``, csrc/dge_train.cu): the parameters
    are re-bound as views of one contiguous buffer, their gradients live in a ``FlatGradBucket`` (``.bucket``), first / second
    moments are flat too.  ``step(clamp=, gscale=)``: p -= Adam(clamp(g * gscale)); ``gscale = 1 / world`` folds the averaging
    of the all-reduced gradient sum into the same pass.  The step counter lives on the device (no host state per step)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("NativeAdam runs on CUDA parameters only")
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                off += p.numel()
        self.bucket = FlatGradBucket(self.params)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.step_t = torch.zeros(1, dtype=torch.int64, device=dev)
        self.param_groups = [dict(lr=float(lr), betas=tuple(betas), eps=float(eps), params=self.params)]

    def zero_grad(self, set_to_none: bool = False):
        self.bucket.zero_()

    def step(self, clamp: float = 0.0, gscale: float = 1.0):
        import ctypes
        from .gnn import _train_lib, _p, _st
        g = self.param_groups[0]
        L = _train_lib()
        dev = self.flat.device
        with torch.cuda.device(dev):
            rc = L.dge_clamp_adam_step(self.flat.numel(), _p(self.flat), _p(self.bucket.flat), _p(self.m), _p(self.v), _p(self.step_t),
                                       ctypes.c_float(g["lr"]), ctypes.c_float(g["betas"][0]), ctypes.c_float(g["betas"][1]), ctypes.c_float(g["eps"]),
                                       ctypes.c_float(clamp), ctypes.c_float(gscale), _st(dev))
        if rc:
            raise RuntimeError(f"dge_clamp_adam_step failed ({rc})")
        self._bump_versions()

    def _bump_versions(self):
        """The kernel wrote through raw pointers: advance autograd's version counters by hand (the cache of the weights' TF32
        operand forms, ``gnn._weight_operand``, keys on them)."""
        setv = getattr(torch._C._autograd, "_unsafe_set_version_counter", None)
        if setv is not None:
            setv(self.params, [p._version + 1 for p in self.params])
        else:                      # older torch: an in-place no-op per parameter does the same (at the price of six tiny launches)
            with torch.no_grad():
                for p in self.params:
                    p.add_(0)

    def state_dict(self):
        return dict(step=self.step_t.clone(), m=self.m.clone(), v=self.v.clone(), param_groups=[{k: v for k, v in self.param_groups[0].items() if k != "params"}])

    def load_state_dict(self, sd):
        self.step_t.copy_(sd["step"]); self.m.copy_(sd["m"]); self.v.copy_(sd["v"])
        self.param_groups[0].update({k: v for k, v in sd["param_groups"][0].items()})
