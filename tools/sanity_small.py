"""Smallest end-to-end exercise of every Pix2Pix kernel (compute-sanitizer target; run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
tr = Pix2PixTrainer(dev, num_downs=5)
g = torch.Generator().manual_seed(1)
A = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).to(dev)
B = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).to(dev)
for _ in range(2):
    out = tr.train_step(A, B)
a8 = torch.randint(0, 256, (2, 64, 64, 3), generator=g, dtype=torch.uint8).to(dev)
b8 = torch.randint(0, 256, (2, 64, 64, 3), generator=g, dtype=torch.uint8).to(dev)
out = tr.train_step(a8, b8)
tr.G.training = False
o8 = torch.empty(2, 64, 64, 3, device=dev, dtype=torch.uint8)
tr.G.forward(a8, out_u8=o8)
torch.cuda.synchronize()
print("ok", out.tolist(), int(o8.sum()))
