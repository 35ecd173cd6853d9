"""Data-parallel correctness on real GPUs (skipped below two devices; run with `gpurun --gpus 2`): two ranks, NCCL,
the trainer's bucketed all-reduce overlapped with the backward pass (wgrad side stream -> event -> comm stream),
against the CPU oracle's 2-shard emulation (the reference step on each shard with per-replica
BatchNorm, gradients averaged, one Adam step).

Checks per rank: its own shard's losses (2e-3); the all-reduced gradient buffers divided by the world size against the
emulation's shard-averaged gradients (whole-gradient cosine >= 0.97, every tensor >= 0.9: bf16 tolerances of
test_gpu_engine.py on a smaller net); every parameter moved by at most lr (one Adam step); and bit-identical replicas
after the step (replica_param_max_abs_diff == 0) although they were constructed from DIFFERENT seeds (the trainer
broadcasts rank 0's state at construction)."""
import tempfile
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

LR = 1e-4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _shard(rank, n=2, hw=64):
    g = torch.Generator().manual_seed(1234 + rank)
    return torch.rand(n, 3, hw, hw, generator=g) * 2 - 1, torch.rand(n, 3, hw, hw, generator=g) * 2 - 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from gan_aug_pfa_b200 import parallel
        from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer
        torch.manual_seed(rank)                       # different seeds: only the broadcast makes the replicas equal
        tr = Pix2PixTrainer(dev, num_downs=5, world=world, bucket_elems=1 << 18)
        sd0 = {k: v.detach().cpu().clone() for k, v in tr.G.state_dict().items()}
        A, B = _shard(rank)
        losses = tr.train_step(A.to(dev), B.to(dev)).cpu().numpy()
        launched_early = tr.g_reducer.launched_before_finish
        diff = parallel.replica_param_max_abs_diff([tr.G, tr.D])
        from oracle import pix2pix_oracle as O
        out = {"sd_g": {k: v.detach().cpu().clone() for k, v in tr.G.state_dict().items()},
               "sd_d": {k: v.detach().cpu().clone() for k, v in tr.D.state_dict().items()},
               "g_g": {k: tr.G.grad(k).cpu().clone() / world for k in O.param_names(sd0)},
               "g_d": {k: tr.D.grad(k).cpu().clone() / world for k in O.param_names(tr.D.state_dict())},
               "sd_g0": sd0}
        path = os.path.join(tempfile.gettempdir(), f"gap_dp2_rank{rank}_{port}.pt")
        torch.save(out, path)
        q.put((rank, losses, diff, launched_early, len(tr.g_reducer.buckets), path))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_step_matches_the_oracles_two_shard_emulation():
    import torch.multiprocessing as mp
    from gan_aug_pfa_b200 import spec
    from oracle import pix2pix_oracle as O
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # ---- oracle emulation: rank 0's seeded model on both shards
    torch.manual_seed(0)
    sd_g0, sd_d0 = spec.default_state_dicts(num_downs=5)
    dumps = [torch.load(r[5]) for r in res]
    for r in res:
        os.remove(r[5])
    for k, v in dumps[1]["sd_g0"].items():            # rank 1 started from rank 0's weights (broadcast at construction)
        assert torch.equal(v, sd_g0[k]), k
    names_g, names_d = O.param_names(sd_g0), O.param_names(sd_d0)
    shard_losses, gd, gg = [], [], []
    # D step on every shard from the same weights; average; one Adam step.  Then the G step against the UPDATED D.
    reps = [(O.clone_state_dict(sd_g0), O.clone_state_dict(sd_d0)) for _ in range(world)]

    class _Capture(O.AdamState):
        """Records the gradients instead of applying them (the emulation averages across shards first)."""
        def apply(self, sd, grads):
            self.grads = {k: g.detach().clone() for k, g in grads.items()}

    # pass 1: D gradients per shard (the G part of this pass is discarded)
    for r in range(world):
        sg, sd_ = reps[r]
        cg, cd = _Capture(sg, names_g, LR, (0.5, 0.999)), _Capture(sd_, names_d, LR, (0.5, 0.999))
        A, B = _shard(r)
        O.gan_train_step(O.clone_state_dict(sg), O.clone_state_dict(sd_), cg, cd, A, B)
        gd.append(cd.grads)
    d_avg = {k: sum(g[k] for g in gd) / world for k in names_d}
    sd_d1 = O.clone_state_dict(sd_d0)
    od = O.AdamState(sd_d1, names_d, LR, (0.5, 0.999))
    with torch.no_grad():
        od.apply(sd_d1, d_avg)

    class _Fixed(O.AdamState):
        """Applies the pre-computed averaged D gradient whatever the shard's own gradient was."""
        def apply(self, sd, grads):
            with torch.no_grad():
                for k in self.names:
                    sd[k].copy_(sd_d1[k])

    # pass 2: the full iteration per shard with D jumping to the averaged update -> losses and G gradients
    for r in range(world):
        sg, sd_ = O.clone_state_dict(sd_g0), O.clone_state_dict(sd_d0)
        cg, fd = _Capture(sg, names_g, LR, (0.5, 0.999)), _Fixed(sd_, names_d, LR, (0.5, 0.999))
        A, B = _shard(r)
        ld, lg, _ = O.gan_train_step(sg, sd_, cg, fd, A, B)
        shard_losses.append((ld, lg))
        gg.append(cg.grads)
    g_avg = {k: sum(g[k] for g in gg) / world for k in names_g}
    sd_g1 = O.clone_state_dict(sd_g0)
    og = O.AdamState(sd_g1, names_g, LR, (0.5, 0.999))
    with torch.no_grad():
        og.apply(sd_g1, g_avg)
    # ---- compare
    def cos(a, b):
        a, b = a.double().flatten(), b.double().flatten()
        return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))

    for (rank, losses, diff, early, n_buckets, _), dump in zip(res, dumps):
        ld, lg = shard_losses[rank]
        assert abs(float(losses[0]) - ld) < 2e-3 * max(1.0, abs(ld)), (rank, losses, ld)
        assert abs(float(losses[1]) - lg) < 2e-3 * max(1.0, abs(lg)), (rank, losses, lg)
        assert diff == 0.0, f"replicas diverged by {diff}"
        assert n_buckets >= 3 and early >= 1          # at least one bucket went out while backward was still running
        for got, ref, names in ((dump["g_g"], g_avg, names_g), (dump["g_d"], d_avg, names_d)):
            whole = cos(torch.cat([got[k].flatten() for k in names]), torch.cat([ref[k].flatten() for k in names]))
            assert whole > 0.97, whole
            worst = min((cos(got[k], ref[k]), k) for k in names)
            assert worst[0] > 0.9, worst
        for sd_gpu, sd_init, names in ((dump["sd_g"], sd_g0, names_g), (dump["sd_d"], sd_d0, names_d)):
            for k in names:
                step = (sd_gpu[k] - sd_init[k].detach()).abs().max()
                assert 0.0 < float(step) <= 1.001 * LR, (k, float(step))
    assert torch.equal(dumps[0]["sd_g"]["model.model.0.weight"], dumps[1]["sd_g"]["model.model.0.weight"])
