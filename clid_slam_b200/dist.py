"""Multi-GPU plumbing for the mapping step: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch on the box, gloo on CPU for the host-logic tests).

The reference is single-process, so this is new functionality whose contract
is "N ranks on one batch == 1 rank on the same batch": the sample batch is sharded across
ranks (every sample's forward / loss / backward only reads the replicated map), and the
reductions the reference does over the whole batch are completed with all-reduces:

  * decoder gradients + the three loss scalars: ONE flat fp32 buffer (833 + 3 floats at ncd128
    shapes), all-reduced right behind the backward kernel on the same stream;
  * neural-point feature gradients (the features are replicated): all-reduced with their
    `touched` flags so every rank applies the identical Adam step;
  * certainty / ts_update side effects are only consumed between frames, so they are reduced
    once per mapping() call (`reduce_side_effects`), not per iteration.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of an n-sample batch owned by `rank`; sizes differ by at
    most one and the slices tile [0, n) in rank order."""
    base, extra = divmod(n, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def decimated_count(begin: int, end: int, decimation: int) -> int:
    """How many global indices i with i % decimation == 0 fall into [begin, end)."""
    first = -(-begin // decimation) * decimation
    return 0 if first >= end else (end - 1 - first) // decimation + 1


class FlatAllReduce:
    """Sum-all-reduce several small tensors through one flat buffer (one collective launch)."""

    def __init__(self, tensors: Sequence[torch.Tensor], group=None):
        self.tensors = [t for t in tensors if t is not None]
        self.group = group
        total = sum(t.numel() for t in self.tensors)
        ref = self.tensors[0]
        self.flat = torch.empty(total, dtype=ref.dtype, device=ref.device)

    def __call__(self) -> None:
        if world()[1] == 1:
            return
        off = 0
        for t in self.tensors:
            self.flat[off:off + t.numel()].copy_(t.reshape(-1))
            off += t.numel()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        off = 0
        for t in self.tensors:
            t.copy_(self.flat[off:off + t.numel()].view_as(t))
            off += t.numel()


def all_reduce_sum(t: Optional[torch.Tensor], group=None) -> None:
    if t is not None and world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def all_reduce_max(t: Optional[torch.Tensor], group=None) -> None:
    if t is not None and world()[1] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)


def reduce_side_effects(certainty: torch.Tensor, certainty_before: torch.Tensor, ts_update: torch.Tensor,
                        group=None) -> None:
    """After a sharded mapping() call: certainty increments are summed over ranks, ts_update takes
    the maximum -- the result every rank would hold had it processed the whole batch."""
    if world()[1] == 1:
        return
    delta = certainty - certainty_before
    dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=group)
    certainty.copy_(certainty_before + delta)
    dist.all_reduce(ts_update, op=dist.ReduceOp.MAX, group=group)
