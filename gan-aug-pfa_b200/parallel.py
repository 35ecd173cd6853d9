"""Data-parallel plumbing: one process per GPU, NCCL gradient all-reduce over NVLink/NVSwitch.

The path shards by batch: every replica runs the reference iteration on its own
shard (per-replica BatchNorm statistics, like DDP without SyncBN) and the only exchange is a SUM
all-reduce of the flat gradient buffers; the 1/world factor is folded into the fused Adam kernel
(``grad_scale``).  The flat buffers are laid out in backward-completion order, so bucketed,
overlapped reduction is a matter of slicing them (see Pix2PixTrainer / DESIGN.md).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def make_allreduce(world: int, bucket_elems: int = 8 << 20) -> Optional[Callable[[torch.Tensor], None]]:
    """Returns f(flat_grad) that sum-reduces the buffer in place across ranks, in buckets."""
    if world <= 1:
        return None

    def allreduce(flat: torch.Tensor) -> None:
        n = flat.numel()
        for off in range(0, n, bucket_elems):
            dist.all_reduce(flat[off:min(n, off + bucket_elems)], op=dist.ReduceOp.SUM)

    return allreduce
