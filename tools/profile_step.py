"""One Pix2Pix training iteration bracketed by cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
WARM = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
torch.manual_seed(0)
tr = Pix2PixTrainer(dev)
gen = torch.Generator().manual_seed(1234)
A = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)
B = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)
for _ in range(WARM):
    tr.train_step(A, B)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = tr.train_step(A, B)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("losses", out.tolist())
