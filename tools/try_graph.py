"""Eager vs CUDA-graph replay of the training step, interleaved blocks (thermal drift cancels) (run under gpurun)."""
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


def timeit(fn, iters=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
tr = Pix2PixTrainer(dev)
for _ in range(5):
    tr.train_step(A, B)
    tr.train_step_graphed(A, B)
te, tg = [], []
for rnd in range(6):
    te.append(timeit(lambda: tr.train_step(A, B)))
    tg.append(timeit(lambda: tr.train_step_graphed(A, B)))
print("eager:", " ".join(f"{t:.3f}" for t in te), f"  mean {sum(te) / len(te):.3f} ms")
print("graph:", " ".join(f"{t:.3f}" for t in tg), f"  mean {sum(tg) / len(tg):.3f} ms")
