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


import os as _os

_SKIP = _os.environ.get("GAP_DP_SKIP_ALLREDUCE", "0") == "1"     # bring-up only: measures what the collectives cost


def make_allreduce(world: int, bucket_elems: int = 8 << 20) -> Optional[Callable[[torch.Tensor], None]]:
    """Returns f(flat_grad) that sum-reduces the buffer in place across ranks, in buckets (no overlap)."""
    if world <= 1:
        return None

    def allreduce(flat: torch.Tensor) -> None:
        n = flat.numel()
        for off in range(0, n, bucket_elems):
            dist.all_reduce(flat[off:min(n, off + bucket_elems)], op=dist.ReduceOp.SUM)

    return allreduce


def plan_buckets(segments: Sequence[Tuple[str, int, int]], total: int, bucket_elems: int,
                 tail_elems: int = 0) -> List[Tuple[int, int, List[str]]]:
    """Group consecutive segments (name, offset, numel — in buffer order) into buckets of at least
    ``bucket_elems`` elements.  Returns (begin, end, segment names); buckets tile [0, total).
    ``tail_elems`` > 0 additionally splits off the trailing segments (as many as fit in ``tail_elems`` elements) as the
    LAST bucket: that bucket can only be reduced after the backward pass has finished, so its transfer is exposed and
    should be small, while the bucket before it still overlaps the last layers."""
    n_tail = 0
    if tail_elems > 0 and len(segments) > 1:
        acc = 0
        for i in range(len(segments) - 1, 0, -1):
            end = segments[i + 1][1] if i + 1 < len(segments) else total
            if acc + (end - segments[i][1]) > tail_elems:
                break
            acc += end - segments[i][1]
            n_tail += 1
    head = segments[:len(segments) - n_tail] if n_tail else segments
    head_total = segments[len(segments) - n_tail][1] if n_tail else total
    buckets: List[Tuple[int, int, List[str]]] = []
    begin, names = 0, []
    for i, (name, off, numel) in enumerate(head):
        names.append(name)
        end = head[i + 1][1] if i + 1 < len(head) else head_total
        if end - begin >= bucket_elems or i + 1 == len(head):
            buckets.append((begin, end, names))
            begin, names = end, []
    if n_tail:
        buckets.append((head_total, total, [s[0] for s in segments[len(segments) - n_tail:]]))
    return buckets


class GradBucketReducer:
    """Bucketed SUM all-reduce of a flat gradient buffer, overlapped with the backward pass.

    ``segments`` lists (name, offset, numel) in the order the backward pass completes them (= buffer
    order).  The engine calls :meth:`mark_ready` when a segment's gradient is final; when every
    segment of the next bucket is ready the bucket is reduced on ``comm_stream`` (after an event on
    the compute stream), so the transfer overlaps the kernels still to come.  :meth:`finish` flushes
    what is left and makes the compute stream wait for the reductions."""

    def __init__(self, flat: torch.Tensor, segments: Sequence[Tuple[str, int, int]], bucket_elems: int = 8 << 20,
                 comm_stream: Optional["torch.cuda.Stream"] = None, tail_elems: int = 0) -> None:
        self.flat = flat
        self.buckets = plan_buckets(segments, flat.numel(), bucket_elems, tail_elems)
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
                if not _SKIP:
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


class TailReducer:
    """Overlapped SUM all-reduce for a flat gradient buffer laid out in FORWARD (module) order, as the Siamese
    engine's is: its backward pass finalises the buffer from the tail towards the head, so "everything from offset
    X on is final" is a watermark that only moves down.  :meth:`ready_from` reduces [X, previous watermark) on the
    communication stream once at least ``min_elems`` elements have accumulated; :meth:`finish` reduces what is left
    and makes the compute stream wait.  Every rank calls ready_from with the same sequence of offsets (they run the
    same network), so the collectives are issued in the same order everywhere."""

    def __init__(self, flat: torch.Tensor, min_elems: int = 4 << 20,
                 comm_stream: Optional["torch.cuda.Stream"] = None) -> None:
        self.flat = flat
        self.min_elems = min_elems
        self.cuda = flat.is_cuda
        self.comm_stream = comm_stream if comm_stream is not None else (torch.cuda.Stream(flat.device) if self.cuda else None)
        self.hi = flat.numel()
        self.launched: List[Tuple[int, int]] = []      # (begin, end) of the reductions of the last step
        self._streams: set = set()

    def begin(self) -> None:
        self.hi = self.flat.numel()
        self.launched = []
        self._streams = set()

    def _launch(self, lo: int, streams=()) -> None:
        view = self.flat[lo:self.hi]
        if self.cuda:
            for st in {torch.cuda.current_stream(self.flat.device), *[s for s in streams if s is not None]}:
                ev = torch.cuda.Event()
                ev.record(st)
                self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                if not _SKIP:
                    dist.all_reduce(view, op=dist.ReduceOp.SUM)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM)
        self.launched.append((lo, self.hi))
        self.hi = lo

    def ready_from(self, offset: int, streams=()) -> None:
        """``streams``: side streams (besides the current one) that produced part of [offset, watermark)."""
        if offset < 0 or offset > self.hi:
            raise ValueError(f"watermark {offset} must move down from {self.hi}")
        self._streams.update(s for s in streams if s is not None)
        if self.hi - offset >= self.min_elems:
            self._launch(offset, self._streams)
            self._streams = set()

    def finish(self) -> None:
        if self.hi > 0:
            self._launch(0, self._streams)
            self._streams = set()
        if self.cuda:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)


def broadcast_replica_state(nets, src: int = 0) -> None:
    """Make every replica start from rank ``src``'s model: parameters, Adam moments / step and BatchNorm buffers of the
    given engines are broadcast, then the bf16 operands are re-packed (what DistributedDataParallel does at construction).
    No-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for net in nets:
        st = net.store
        for buf in (st.p, st.m, st.v):
            dist.broadcast(buf, src)
        step = torch.tensor([st.step], device=st.p.device, dtype=torch.int64)
        dist.broadcast(step, src)
        st.step = int(step.item())
        if st.step_dev is not None:
            st.step_dev.fill_(st.step)
        for bn in net.bns.values():
            dist.broadcast(bn.running_mean, src)
            dist.broadcast(bn.running_var, src)
            dist.broadcast(bn.nbt, src)
        net.repack()


def replica_param_max_abs_diff(nets) -> float:
    """max over parameters of (max over ranks - min over ranks): 0.0 when the replicas hold bit-identical models, which
    the data-parallel step guarantees (identical all-reduced gradients into a deterministic optimizer)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0.0
    worst = 0.0
    for net in nets:
        hi, lo = net.store.p.clone(), net.store.p.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        worst = max(worst, float((hi - lo).abs().max().item()))
    return worst
