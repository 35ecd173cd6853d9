"""The PatchGAN-head kernels at the batch-64 shape, three launches each (ncu target; run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
N, H, C = 64, 31, 512
g = torch.Generator().manual_seed(0)
y = torch.randn(N, H, H, C, generator=g).to(torch.bfloat16).to(dev)
w = (torch.randn(16 * C, generator=g) / 90).to(torch.bfloat16).to(dev)
bias = torch.zeros(1, device=dev)
z = torch.empty(N * H * H * 16, device=dev)
logits = torch.empty(N, H - 1, H - 1, device=dev)
dlog = torch.randn(N, H - 1, H - 1, generator=g).to(dev)
scale, shift = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
gx = torch.empty_like(y)
dw = torch.zeros(16 * C, device=dev)
sums = torch.zeros(2 * C, device=dev, dtype=torch.float64)
pre = (scale, shift, 0.2)
for _ in range(3):
    ops.cout1_conv_fwd(y, w, bias, z, logits, pre=pre)
    ops.cout1_conv_dgrad(dlog, w, gx, bwd=dict(y=y, scale=scale, shift=shift, slope=0.2, sums=sums))
    ops.cout1_conv_dgrad(dlog, w, gx)
    ops.cout1_conv_wgrad(dlog, y, dw, pre=pre)
torch.cuda.synchronize()
print("ok")
