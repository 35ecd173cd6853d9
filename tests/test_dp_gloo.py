"""World-size-2 data-parallel host logic on CPU (gloo): the bucketed gradient all-reduce, and the DP
semantics the trainer implements (per-replica step on its own shard, SUM all-reduce, 1/world folded
into Adam) checked against the oracle's single-process emulation."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gan_aug_pfa_b200 import spec
from gan_aug_pfa_b200.parallel import make_allreduce
from oracle import pix2pix_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        sd_g, sd_d = spec.default_state_dicts(num_downs=5, ngf=8, ndf=8)
        names = O.param_names(sd_d)
        gen = torch.Generator().manual_seed(1234 + rank)
        A = torch.rand(2, 3, 32, 32, generator=gen) * 2 - 1
        B = torch.rand(2, 3, 32, 32, generator=gen) * 2 - 1
        for k in names:
            sd_d[k].requires_grad_(True)
        with torch.no_grad():
            fake = O.unet_generator_forward(sd_g, A, True, {})
        pr = O.discriminator_forward(sd_d, torch.cat((A, B), 1), True, {})
        pf = O.discriminator_forward(sd_d, torch.cat((A, fake), 1), True, {})
        loss = 0.5 * (O.bce_with_logits(pr, torch.ones_like(pr)) + O.bce_with_logits(pf, torch.zeros_like(pf)))
        grads = torch.autograd.grad(loss, [sd_d[k] for k in names])
        flat = torch.cat([g.reshape(-1) for g in grads])
        local = flat.clone()
        allreduce = make_allreduce(world, bucket_elems=1000)     # several buckets, ragged tail
        allreduce(flat)
        # numpy arrays travel by value: a torch tensor in an mp.Queue is a shared-memory handle that dies with the sender
        q.put((rank, local.numpy(), flat.numpy()))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_matches_sum_of_replica_grads():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [(r, torch.from_numpy(a), torch.from_numpy(b)) for r, a, b in res]
    total = res[0][1] + res[1][1]
    for _, _, reduced in res:
        assert torch.allclose(reduced, total, rtol=1e-6, atol=1e-9)
    # replicas saw different shards, so their local gradients differ
    assert not torch.allclose(res[0][1], res[1][1])


def test_make_allreduce_is_none_for_single_rank():
    assert make_allreduce(1) is None


def test_grad_scale_equals_averaging():
    """Adam(g_sum, grad_scale=1/R) == Adam(mean of replica grads): the fold used by the trainer."""
    g = torch.Generator().manual_seed(0)
    p = torch.randn(100, generator=g)
    g1, g2 = torch.randn(100, generator=g), torch.randn(100, generator=g)
    pa, ma, va = p.clone(), torch.zeros(100), torch.zeros(100)
    O.adam_update(pa, (g1 + g2) / 2, ma, va, 1, 1e-4, 0.5, 0.999)
    pb, mb, vb = p.clone(), torch.zeros(100), torch.zeros(100)
    O.adam_update(pb, (g1 + g2) * 0.5, mb, vb, 1, 1e-4, 0.5, 0.999)
    assert torch.equal(pa, pb)


# ---------------------------------------------------------------------------------------------------
# GradBucketReducer: bucket planning and the overlapped launch protocol (CPU tensors, gloo)
# ---------------------------------------------------------------------------------------------------
def test_plan_buckets_tiles_the_buffer_on_segment_boundaries():
    from gan_aug_pfa_b200.parallel import plan_buckets
    segs = [("a", 0, 10), ("b", 12, 100), ("c", 112, 5), ("d", 120, 300), ("e", 420, 7)]
    buckets = plan_buckets(segs, 428, 100)
    assert buckets[0][0] == 0 and buckets[-1][1] == 428
    for (b0, e0, _), (b1, _, _) in zip(buckets, buckets[1:]):
        assert e0 == b1
    assert [n for _, _, names in buckets for n in names] == ["a", "b", "c", "d", "e"]
    assert all(e - b >= 100 for b, e, _ in buckets[:-1])


def _reducer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gan_aug_pfa_b200.parallel import GradBucketReducer
        segs, off = [], 0
        for i, n in enumerate([50, 7, 300, 120, 3, 64, 900, 11]):
            segs.append((f"s{i}", off, n))
            off += (n + 3) // 4 * 4
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(off, generator=g)
        local = flat.clone()
        red = GradBucketReducer(flat, segs, bucket_elems=200)
        red.begin()
        order = ["s0", "s2", "s1", "s3", "s4", "s5", "s6"]        # s1 late: bucket 0 waits for it; s7 never marked
        launched = []
        for name in order:
            red.mark_ready(name)
            launched.append(red.launched_before_finish)
        red.finish()
        q.put((rank, local.numpy(), flat.numpy(), launched, len(red.buckets)))
    finally:
        dist.destroy_process_group()


def test_grad_bucket_reducer_world2_overlapped_protocol():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_reducer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [(r, torch.from_numpy(a), torch.from_numpy(b), l, n) for r, a, b, l, n in res]
    total = res[0][1] + res[1][1]
    for _, _, reduced, launched, n_buckets in res:
        assert torch.allclose(reduced, total, rtol=1e-6, atol=1e-9)      # every element reduced exactly once
        assert launched[0] == 0 and launched[1] == 0                      # bucket 0 = {s0,s1,s2} not complete yet
        assert launched[2] >= 1                                           # ... until s1 arrives
        assert launched[-1] >= 2 and launched[-1] < n_buckets             # buckets went out during "backward"


# ---------------------------------------------------------------------------------------------------
# TailReducer (Siamese engine: flat buffer in forward order, finalised from the tail) and the replica
# broadcast / checksum helpers
# ---------------------------------------------------------------------------------------------------
class _FakeStore:
    def __init__(self, p):
        self.p, self.m, self.v = p, torch.zeros_like(p), torch.zeros_like(p)
        self.step, self.step_dev = 0, None


class _FakeNet:
    def __init__(self, p):
        self.store, self.bns, self.repacked = _FakeStore(p), {}, 0

    def repack(self):
        self.repacked += 1


def _tail_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gan_aug_pfa_b200.parallel import TailReducer, broadcast_replica_state, replica_param_max_abs_diff
        g = torch.Generator().manual_seed(7 + rank)
        flat = torch.randn(1000, generator=g)
        local = flat.clone()
        red = TailReducer(flat, min_elems=200)
        red.begin()
        for off in (900, 760, 750, 400, 390):      # watermarks only move down; small steps are merged
            red.ready_from(off)
        before_finish = list(red.launched)
        red.finish()
        # replicas built from different seeds differ until rank 0's state is broadcast
        net = _FakeNet(torch.randn(64, generator=g))
        d0 = replica_param_max_abs_diff([net])
        broadcast_replica_state([net])
        d1 = replica_param_max_abs_diff([net])
        q.put((rank, local.numpy(), flat.numpy(), before_finish, list(red.launched), d0, d1, net.repacked))
    finally:
        dist.destroy_process_group()


def test_tail_reducer_and_replica_sync_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tail_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = torch.from_numpy(res[0][1]) + torch.from_numpy(res[1][1])
    for _, _, reduced, before, launched, d0, d1, repacked in res:
        assert torch.allclose(torch.from_numpy(reduced), total, rtol=1e-6, atol=1e-9)    # every element reduced exactly once
        assert before == [(760, 1000), (400, 760)]          # 900 and 750/390 were too small a step: merged into the next
        assert launched == before + [(0, 400)]               # finish() flushes the head
        assert d0 > 0.0 and d1 == 0.0 and repacked == 1
    import pytest as _pt
    from gan_aug_pfa_b200.parallel import TailReducer
    r = TailReducer(torch.zeros(10), min_elems=1)
    r.begin()
    r.hi = 5
    with _pt.raises(ValueError):
        r.ready_from(7)


def test_plan_buckets_splits_off_a_small_exposed_tail():
    """tail_elems: the bucket that can only be reduced after the backward pass holds just the trailing small segments."""
    from gan_aug_pfa_b200.parallel import plan_buckets
    segs = [("up", 0, 4000), ("d6", 4000, 4000), ("d5", 8000, 4000), ("d4", 12000, 4000), ("d3", 16000, 2000),
            ("d2", 18000, 500), ("d1", 18500, 128), ("d0", 18628, 4)]
    plain = plan_buckets(segs, 18632, 8000)
    assert plain[-1][2] == ["d3", "d2", "d1", "d0"]                         # 2632 elements wait for the end of backward
    split = plan_buckets(segs, 18632, 8000, tail_elems=1000)
    assert split[-1] == (18000, 18632, ["d2", "d1", "d0"])                  # only 632 do now
    assert split[-2][2] == ["d3"] and split[-2][1] == 18000
    assert split[0][0] == 0 and all(a[1] == b[0] for a, b in zip(split, split[1:]))
    assert [n for _, _, names in split for n in names] == [s[0] for s in segs]
    assert plan_buckets(segs, 18632, 8000, tail_elems=2) == plain           # nothing fits: unchanged
