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
