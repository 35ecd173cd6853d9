"""Data-parallel plumbing: one process per GPU, NCCL gradient all-reduce over NVLink/NVSwitch.

The path shards by batch: every replica runs the reference iteration on its own
shard (per-replica BatchNorm statistics, like DDP without SyncBN) and the only exchange is a SUM
all-reduce of the flat gradient buffers; the 1/world factor is folded into the fused Adam kernel
(``grad_scale``).  The flat buffers are laid out in backward-completion order, so a bucket is a
contiguous slice that becomes final while the rest of the backward pass is still running:
:class:`GradBucketReducer` launches each bucket's all-reduce on a side stream as soon as its last
segment is complete, overlapping NCCL with the remaining dgrad / wgrad kernels.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def make_allreduce(world: int, bucket_elems: int = 8 << 20) -> Optional[Callable[[torch.Tensor], None]]:
    """Returns f(flat_grad) that sum-reduces the buffer in place across ranks, in buckets (no overlap)."""
    if world <= 1:
        return None

    def allreduce(flat: torch.Tensor) -> None:
        n = flat.numel()
        for off in range(0, n, bucket_elems):
            dist.all_reduce(flat[off:min(n, off + bucket_elems)], op=dist.ReduceOp.SUM)

    return allreduce


def plan_buckets(segments: Sequence[Tuple[str, int, int]], total: int, bucket_elems: int) -> List[Tuple[int, int, List[str]]]:
    """Group consecutive segments (name, offset, numel — in buffer order) into buckets of at least
    ``bucket_elems`` elements.  Returns (begin, end, segment names); buckets tile [0, total)."""
    buckets: List[Tuple[int, int, List[str]]] = []
    begin, names = 0, []
    for i, (name, off, numel) in enumerate(segments):
        names.append(name)
        end = segments[i + 1][1] if i + 1 < len(segments) else total
        if end - begin >= bucket_elems or i + 1 == len(segments):
            buckets.append((begin, end, names))
            begin, names = end, []
    return buckets


class GradBucketReducer:
    """Bucketed SUM all-reduce of a flat gradient buffer, overlapped with the backward pass.

    ``segments`` lists (name, offset, numel) in the order the backward pass completes them (= buffer
    order).  The engine calls :meth:`mark_ready` when a segment's gradient is final; when every
    segment of the next bucket is ready the bucket is reduced on ``comm_stream`` (after an event on
    the compute stream), so the transfer overlaps the kernels still to come.  :meth:`finish` flushes
    what is left and makes the compute stream wait for the reductions."""

    def __init__(self, flat: torch.Tensor, segments: Sequence[Tuple[str, int, int]], bucket_elems: int = 8 << 20,
                 comm_stream: Optional["torch.cuda.Stream"] = None) -> None:
        self.flat = flat
        self.buckets = plan_buckets(segments, flat.numel(), bucket_elems)
        self.bucket_of: Dict[str, int] = {n: b for b, (_, _, names) in enumerate(self.buckets) for n in names}
        self.cuda = flat.is_cuda
        self.comm_stream = comm_stream if comm_stream is not None else (torch.cuda.Stream(flat.device) if self.cuda else None)
        self.launched_before_finish = 0          # statistics of the last step (tests / logging)
        self._pending: List[int] = []
        self._streams: List[set] = []
        self._next = 0

    def begin(self) -> None:
        self._pending = [len(names) for _, _, names in self.buckets]
        self._streams = [set() for _ in self.buckets]
        self._next = 0
        self.launched_before_finish = 0

    def _launch(self, b: int) -> None:
        begin, end, _ = self.buckets[b]
        view = self.flat[begin:end]
        if self.cuda:
            # wait for the compute stream and for every side stream that produced a segment of this bucket
            for st in {torch.cuda.current_stream(self.flat.device)} | (self._streams[b] if self._streams else set()):
                ev = torch.cuda.Event()
                ev.record(st)
                self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(view, op=dist.ReduceOp.SUM)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM)

    def mark_ready(self, name: str, stream=None) -> None:
        """``stream``: the (side) stream the segment's last kernel was enqueued on, if not the current one."""
        b = self.bucket_of.get(name)
        if b is None:
            raise KeyError(f"unknown gradient segment {name!r}")
        self._pending[b] -= 1
        if stream is not None:
            self._streams[b].add(stream)
        # buckets are launched in order on every rank (NCCL collectives must be issued in the same order)
        while self._next < len(self.buckets) and self._pending[self._next] <= 0:
            self._launch(self._next)
            self._next += 1
            self.launched_before_finish += 1

    def finish(self) -> None:
        while self._next < len(self.buckets):
            self._launch(self._next)
            self._next += 1
        if self.cuda:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)
