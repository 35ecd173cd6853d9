"""ABBA A/B of an engine attribute on the Siamese 512x512 batch-4 training step: ab_siamese.py attr v0 v1 [rounds]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import models as M  # noqa: E402
from gan_aug_pfa_b200.siamese import SiameseEngine  # noqa: E402

attr, v0, v1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda:0")
torch.manual_seed(0)
eng = SiameseEngine(dev)
eng.load_state_dict({k: v.detach() for k, v in M.SiameseUNet(3, 1).state_dict().items()})
g = torch.Generator().manual_seed(1)
N, S = 4, 512
b = ((torch.rand(N, 3, S, S, generator=g) * 2 - 1).to(dev), (torch.rand(N, 3, S, S, generator=g) * 2 - 1).to(dev),
     (torch.rand(N, S, S, generator=g) < 0.05).long().to(dev))


def timeit(iters=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        eng.train_step(*b)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for _ in range(3):
    eng.train_step(*b)
t = {v0: [], v1: []}
for rnd in range(rounds):
    for v in ((v0, v1) if rnd % 2 == 0 else (v1, v0)):
        setattr(eng, attr, bool(v))
        eng.train_step(*b)
        t[v].append(timeit())
for v in (v0, v1):
    print(f"{attr}={v}:", " ".join(f"{x:.3f}" for x in t[v]), f"  mean {sum(t[v]) / len(t[v]):.3f} ms")
