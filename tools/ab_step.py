"""A/B of engine switches inside one process (same box, same clocks): python tools/ab_step.py attr [attr...]
Each attr is a boolean class attribute of the engines (e.g. overlap_wgrad, fuse_head); every combination is timed
as eager and graph-replayed steps, interleaved twice."""
import itertools
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402

dev = torch.device("cuda:0")
N = 64
gen = torch.Generator().manual_seed(1234)
A = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)
B = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)
attrs = sys.argv[1:] or ["overlap_wgrad"]


def timeit(fn, iters=15):
    for _ in range(4):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


trainers = {}
for combo in itertools.product((True, False), repeat=len(attrs)):
    torch.manual_seed(0)
    tr = Pix2PixTrainer(dev)
    for a, v in zip(attrs, combo):
        if hasattr(tr, a):
            setattr(tr, a, v)
        else:
            setattr(tr.G, a, v)
            setattr(tr.D, a, v)
    trainers[combo] = tr
for rnd in range(2):
    for combo, tr in trainers.items():
        t_e = timeit(lambda: tr.train_step(A, B))
        t_g = timeit(lambda: tr.train_step_graphed(A, B))
        print(f"round {rnd} {dict(zip(attrs, combo))}: eager {t_e:.3f} ms  graph {t_g:.3f} ms", flush=True)
