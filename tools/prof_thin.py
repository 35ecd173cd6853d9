"""The thin-layer kernels at the batch-64 shapes, two launches each (ncu target; run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf = dict(device=dev, dtype=torch.bfloat16)
N = 64
xa = torch.randn(N, 256, 256, 4, **bf)
xb = torch.randn(N, 256, 256, 4, **bf)
w8 = torch.randn(64, 128, **bf)
w4 = torch.randn(64, 64, **bf)
bias = torch.randn(64, device=dev)
o64 = torch.empty(N, 128, 128, 64, **bf)
o64b = torch.empty(N, 128, 128, 64, **bf)
o128 = torch.empty(N, 128, 128, 128, **bf)
w4b = torch.randn(128, 64, **bf)
dw8 = torch.zeros(64, 96, device=dev)
dw4 = torch.zeros(64, 48, device=dev)
db = torch.zeros(64, device=dev)
wide128 = torch.randn(N, 128, 128, 128, **bf)
wcol = torch.randn(48, 128, **bf)
b3 = torch.randn(3, device=dev)
fake_bf = torch.zeros(N, 256, 256, 4, **bf)
fake_f32 = torch.zeros(N, 256, 256, 4, device=dev)
for _ in range(2):
    ops.thin_conv_fwd(xa, xb, w8, bias, o64, ops.ACT_LRELU)                              # D.0 forward
    ops.thin_conv_fwd(xa, None, w4, None, o64, ops.ACT_LRELU, o64b, ops.ACT_RELU)         # G.0 forward
    ops.thin_conv_fwd(xa, None, w4b, None, o128)                                         # G.last input gradient
    ops.thin_conv_wgrad(o64, xa, xb, dw8, 96, db)                                        # D.0 wgrad
    ops.thin_conv_wgrad(o64, xa, None, dw4, 48, None)                                    # G.0 wgrad
    ops.thin_convT_fwd(wide128, wcol, b3, ops.ACT_TANH, fake_bf, fake_f32)               # G.last forward
torch.cuda.synchronize()
print("ok")
