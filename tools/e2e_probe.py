"""Where does the end-to-end loop lose time at world > 1?  torchrun --nproc-per-node N tools/e2e_probe.py
Variants of bench.py's e2e loop: device batches / prefetched host batches x no loss read / lag-1 / lag-2 reads."""
import os
import sys
from pathlib import Path
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200.io import PairPrefetcher  # noqa: E402
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
tr = Pix2PixTrainer(dev, world=world)
g = torch.Generator().manual_seed(1 + rank)
host = [(torch.randint(0, 256, (64, 256, 256, 3), generator=g, dtype=torch.uint8).pin_memory(),
         torch.randint(0, 256, (64, 256, 256, 3), generator=g, dtype=torch.uint8).pin_memory()) for _ in range(2)]
devb = [(a.to(dev), b.to(dev)) for a, b in host]
for i in range(5):
    tr.train_step(*devb[i % 2])
STEPS = 20


def run(h2d: bool, lag: int, zero_copy: bool = False):
    pinned = [torch.empty(2, dtype=torch.float64).pin_memory() for _ in range(4)]
    evs = [torch.cuda.Event() for _ in range(4)]
    cur = torch.cuda.current_stream()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    src = PairPrefetcher(dev, [host[i % 2] for i in range(STEPS)]) if h2d else [devb[i % 2] for i in range(STEPS)]
    for i, (a, b) in enumerate(src):
        out = tr.train_step(a, b, loss_host=pinned[i % 4] if (zero_copy and lag >= 0) else None)
        if h2d:
            src.release()
        if lag >= 0:
            if not zero_copy:
                pinned[i % 4].copy_(out, non_blocking=True)
            evs[i % 4].record(cur)
            if i >= lag:
                evs[(i - lag) % 4].synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / STEPS], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


for h2d, lag, zc in ((False, -1, False), (False, 1, False), (False, 1, True), (True, 1, False), (True, 1, True), (True, 2, True),
                     (False, -1, False), (True, 1, True)):
    t = run(h2d, lag, zc)
    if rank == 0:
        print(f"h2d={h2d} loss-read lag={lag if lag >= 0 else 'none'} zero_copy={zc}: {t:.3f} ms/step", flush=True)
if world > 1:
    dist.destroy_process_group()
