"""Eager vs CUDA-graph replay, wgrad side-stream overlap on/off (run under gpurun)."""
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


for overlap in (True, False):
    torch.manual_seed(0)
    tr = Pix2PixTrainer(dev)
    tr.G.overlap_wgrad = tr.D.overlap_wgrad = overlap
    t_e = timeit(lambda: tr.train_step(A, B))
    t_g = timeit(lambda: tr.train_step_graphed(A, B))
    t_e2 = timeit(lambda: tr.train_step(A, B))
    print(f"overlap_wgrad={overlap}: eager {t_e:.3f} ms  graph {t_g:.3f} ms  eager again {t_e2:.3f} ms", flush=True)
    del tr
