"""A dgrad with the backward-fused epilogue (BatchNorm sums on), a few launches (ncu target; run under gpurun).
usage: prof_bwd_epi.py [cin cout g]   (default 256 128 32: dy [64,g,g,cin] -> d [64,2g,2g,cout])"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf = dict(device=dev, dtype=torch.bfloat16)
N = 64
cin, cout, g = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 128, 32)
dy = torch.randn(N, g, g, cin, **bf)
w = torch.randn(4, cout, 4 * cin, **bf)
out = torch.empty(N, 2 * g, 2 * g, cout, **bf)
y = torch.randn(N, 2 * g, 2 * g, cout, **bf)
g2 = torch.randn(N, 2 * g, 2 * g, cout, **bf)
scale = torch.rand(cout, device=dev) + 0.5
shift = torch.randn(cout, device=dev)
st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(4):
    if i == 3:
        e0.record()
    ops.conv_gemm([dy], w, ops.geom_phase_k4s2p1(), out, cout, (g, g), stats=st,
                  bwd=dict(y=y, slope=0.2, g2=g2, scale=scale, shift=shift))
e1.record()
torch.cuda.synchronize()
print(f"ok {e0.elapsed_time(e1) * 1e3:.1f} us")
