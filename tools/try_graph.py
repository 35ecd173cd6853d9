"""CUDA-graph capture of one Pix2Pix training iteration (run under gpurun): time eager vs graph replay."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
tr = Pix2PixTrainer(dev)
N = 64
gen = torch.Generator().manual_seed(1234)
A = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)
B = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print(f"eager: {timeit(lambda: tr.train_step(A, B)):.3f} ms/step")
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        tr.train_step(A, B)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = tr.train_step(A, B)
torch.cuda.synchronize()
print(f"graph: {timeit(g.replay):.3f} ms/step   losses {out.tolist()}")
