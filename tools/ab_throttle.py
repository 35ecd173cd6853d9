"""Does it matter how far the host runs ahead of the GPU?  (run under gpurun)

bench.py's end-to-end loop (the host reads step i-1's losses before it enqueues step i+1) measured 3-5 % FASTER per
iteration than the device-resident loop, which enqueues all its steps without ever waiting — although it does strictly
more work.  This A/B runs the same eager training step with the host allowed to be `lag` steps ahead (an event per step,
synchronised `lag` steps later; lag = -1: never wait), interleaved, CUDA events around 20 iterations each.
usage: ab_throttle.py [rounds]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
N = 64
gen = torch.Generator().manual_seed(1234)
A = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)
B = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)
torch.manual_seed(0)
tr = Pix2PixTrainer(dev)
for _ in range(5):
    tr.train_step(A, B)
evs = [torch.cuda.Event() for _ in range(8)]


def run(lag, iters=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        tr.train_step(A, B)
        if lag >= 0:
            evs[i % 8].record()
            if i >= lag:
                evs[(i - lag) % 8].synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


lags = (-1, 0, 1, 2)
t = {k: [] for k in lags}
for rnd in range(rounds):
    for lag in (lags if rnd % 2 == 0 else lags[::-1]):
        run(lag, 3)
        t[lag].append(run(lag))
for lag in lags:
    name = "never waits" if lag < 0 else f"<= {lag + 1} step(s) ahead"
    print(f"host {name:18s}:", " ".join(f"{x:.3f}" for x in t[lag]), f"  mean {sum(t[lag]) / len(t[lag]):.3f} ms", flush=True)
