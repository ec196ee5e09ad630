"""Multi-GPU plumbing: one process per GPU, rays sharded, one gradient all-reduce per step.

The renderer has no cross-ray operation, so a ray batch (or an image's rows) splits into contiguous
tiles with no data-path collective.  Training needs exactly one exchange per step: the sum of the MLP
gradients (precedent RG/train.py:246-253).  Both networks' gradients live in ONE flat
fp32 buffer (2 x 595 844 floats = 4.77 MB) whose slices are the parameters' ``.grad`` views, so the
all-reduce is a single NCCL call with no packing copies.  Mask-dependent loss denominators travel in the
same call.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_rays", "FlatGrads", "allreduce_masks", "gather_rows"]


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous tile [lo, hi) of n items owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays: torch.Tensor, rank: int, world_size: int, dim: int = 0) -> torch.Tensor:
    lo, hi = shard_bounds(rays.shape[dim], rank, world_size)
    return rays.narrow(dim, lo, hi - lo)


class FlatGrads:
    """Flat gradient buffer shared by a list of parameters (+ a few trailing scalar slots)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], extra_slots: int = 4):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total + extra_slots, device=dev, dtype=torch.float32)
        self.extra = self.flat[total:]
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def allreduce(self, group=None, average: bool = False, async_op: bool = False):
        """SUM over ranks in one collective.  With ``average`` the result is divided by the world size
        (use when every rank computed a mean loss over an equal share of the global batch)."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average and not async_op:
            self.flat.div_(dist.get_world_size(group))
        return work


def allreduce_masks(masks: dict, group=None) -> dict:
    """OR the per-rank partial hard masks (uint8) of consistency.build_hard_masks."""
    out = {}
    for k in sorted(masks):
        m = masks[k].to(torch.int32)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
        out[k] = m > 0
    return out


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> Optional[torch.Tensor]:
    """All-gather row tiles produced with shard_bounds back into [n_total, ...] on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0)
