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

__all__ = ["shard_bounds", "shard_rays", "FlatGrads", "global_mask_counts", "allreduce_masks", "gather_rows"]


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous tile [lo, hi) of n items owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays: torch.Tensor, rank: int, world_size: int, dim: int = 0) -> torch.Tensor:
    lo, hi = shard_bounds(rays.shape[dim], rank, world_size)
    return rays.narrow(dim, lo, hi - lo)


class FlatGrads:
    """Flat gradient buffer shared by groups of parameters (+ a few trailing scalar slots).

    ``params``: a list of parameters (one group) or a list of lists -- each group (typically one network) owns a contiguous
    segment, so a group can be reduced on its own as soon as its backward pass is done (``allreduce_group``): the fine
    network's gradients are final before the coarse network's backward starts (graph cut at NP/run_nerf.py:397), so its
    exchange runs under the coarse backward.  The parameters are marked for the fused backward's in-place accumulation
    (ops.FusedMLPFn), which adds straight into these views.

    ``optimizer.zero_grad()`` with its default ``set_to_none=True`` would silently detach the parameters from the buffer:
    use ``zero_()`` here instead; ``allreduce*`` re-attaches (and keeps the values of) any gradient that was replaced."""

    def __init__(self, params: Iterable, extra_slots: int = 4):
        params = list(params)
        groups = [list(g) for g in params] if params and isinstance(params[0], (list, tuple)) else [params]
        self.groups: List[List[torch.nn.Parameter]] = [[p for p in g if p.requires_grad] for g in groups]
        self.params: List[torch.nn.Parameter] = [p for g in self.groups for p in g]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total + extra_slots, device=dev, dtype=torch.float32)
        self.extra = self.flat[total:]
        self._slots, self._segments, self._pending = [], [], []
        off = 0
        for g in self.groups:
            start = off
            for p in g:
                view = self.flat[off:off + p.numel()].view_as(p)
                p.grad = view
                p._cnerf_accumulate_in_place = True
                self._slots.append((p, view))
                off += p.numel()
            self._segments.append((start, off))

    def zero_(self):
        self.flat.zero_()

    def _reattach(self):
        """Parameters whose .grad no longer aliases the flat buffer (zero_grad(set_to_none=True), p.grad = ...) are pointed
        back at it; a gradient that autograd accumulated into a fresh tensor meanwhile is copied over first."""
        for p, view in self._slots:
            g = p.grad
            if g is None:
                view.zero_()
                p.grad = view
            elif g.data_ptr() != view.data_ptr():
                view.copy_(g)
                p.grad = view

    def allreduce(self, group=None, average: bool = False, async_op: bool = False):
        """SUM of the whole buffer (incl. the extra slots) over ranks in one collective.  ``average`` divides by the world
        size afterwards (use when every rank computed a mean loss over an equal share of the global batch; not needed when
        the losses were normalised with ``global_mask_counts``)."""
        if average and async_op:
            raise ValueError("FlatGrads.allreduce: average=True needs the reduced values, it cannot be combined with async_op=True")
        self._reattach()
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if average:
            self.flat.div_(dist.get_world_size(group))
        return work

    def allreduce_group(self, index: int, group=None):
        """Start the SUM of one parameter group's segment (async; ordered after the work already enqueued on the current
        stream) and remember the handle; ``wait()`` joins all pending segments.  The last group also carries the extra slots."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        lo, hi = self._segments[index]
        if index == len(self._segments) - 1:
            hi = self.flat.numel()
        work = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True)
        self._pending.append(work)
        return work

    def wait(self):
        """Join the segment reductions started with allreduce_group (the current stream waits; the host does not)."""
        for w in self._pending:
            w.wait()
        self._pending = []

    def overlap_with_backward(self, nets, group=None):
        """``nets[i]`` (a NeRF whose parameters are group i) reduces its segment as soon as its fused backward has been
        enqueued -- with autograd's reverse order the fine network goes first and its exchange overlaps the coarse
        network's backward.  Call ``wait()`` before the optimizer step.  Only the fused (canonical-architecture) backward
        triggers the hook; groups whose hook did not fire are reduced by ``finish()``."""
        self._fired = set()
        for i, net in enumerate(nets):
            def ready(i=i):
                self._fired.add(i)
                self.allreduce_group(i, group=group)
            net.packed_weights().after_backward = ready
        self._overlap_group = group

    def finish(self):
        """After loss.backward(): reduce the groups whose hook did not fire, then join everything."""
        self._reattach()
        for i in range(len(self._segments)):
            if i not in getattr(self, "_fired", set()):
                self.allreduce_group(i, group=getattr(self, "_overlap_group", None))
        self.wait()
        self._fired = set()


def global_mask_counts(mask: Optional[torch.Tensor], n_rows: int, group=None, device=None) -> torch.Tensor:
    """[#mask==1, #mask==0, sum(mask), #rows] of the GLOBAL batch (one tiny all-reduce of 4 floats): the denominators of the
    masked losses (NP/run_nerf_view.py:1647-1648) so that the sum of the ranks' losses -- and of their gradients -- equals
    the single-GPU loss over the concatenated batch.  Pass the result as ``global_counts`` to masked_img_loss /
    masked_depth_loss and reduce the gradients with SUM (no averaging)."""
    if mask is None:      # plain means (img2mse): every row counts
        c = torch.full((4,), float(n_rows), dtype=torch.float32, device=device or "cuda")      # (fill kernels only: CUDA-graph capturable)
        c[1] = 0.0
    else:
        m = mask.reshape(-1).float()
        c = torch.stack([(m == 1).sum().float(), (m == 0).sum().float(), m.sum(), torch.full((), float(n_rows), dtype=torch.float32, device=m.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return c


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
