"""One small-N dgrad with the backward-fused epilogue, a few launches (ncu target; run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf = dict(device=dev, dtype=torch.bfloat16)
N = 64
dy = torch.randn(N, 64, 64, 128, **bf)
w = torch.randn(4, 64, 4 * 128, **bf)
out = torch.empty(N, 128, 128, 64, **bf)
y = torch.randn(N, 128, 128, 64, **bf)
g2 = torch.randn(N, 128, 128, 64, **bf)
for _ in range(4):
    ops.conv_gemm([dy], w, ops.geom_phase_k4s2p1(), out, 64, (64, 64), bwd=dict(y=y, slope=0.2, g2=g2))
torch.cuda.synchronize()
print("ok")
