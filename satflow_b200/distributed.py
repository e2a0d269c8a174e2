"""Batch-sharded data parallelism for the ConvLSTM path: one process per GPU, each rank runs its own
rollout on its shard, and the ONLY exchange is one all-reduce of the flat fp32 gradient buffer
(what the reference gets from Lightning DDP: configs/trainer/ddp.yaml:4-5).

The reference's loss is a mean over the local shard (conv_lstm.py:63 MSELoss), so gradients are
AVERAGED over equally sized shards to equal the full-batch gradient.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class FlatGradBucket:
    """All parameter gradients as views into one contiguous fp32 buffer -> a single collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off : off + n].view_as(p)
            off += n

    def zero_(self) -> None:
        self.flat.zero_()

    def all_reduce_mean(self, group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
        """Average the gradients across ranks (no-op for a single process)."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        self.flat.div_(world)  # pre-divide: SUM of pre-scaled shards == mean, on every backend (gloo has no AVG)
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group: Optional[dist.ProcessGroup] = None) -> None:
    """Make every replica start from rank ``src``'s weights (what DDP does at construction)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous, equally sized shard of the leading (batch) dimension."""
    if x.shape[0] % world:
        raise ValueError(f"global batch {x.shape[0]} is not divisible by {world} ranks")
    per = x.shape[0] // world
    return x[rank * per : (rank + 1) * per]
