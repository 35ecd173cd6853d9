"""Ablation of the fprop epilogue on the small-K / small-N layers (run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
bf = dict(device=dev, dtype=torch.bfloat16)


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
    return e0.elapsed_time(e1) / iters * 1e3


N = 64
# D layer-1 dgrad: dy [64,64,64,128] -> gH0 [64,128,128,64], four phases, act-only backward epilogue
dy = torch.randn(N, 64, 64, 128, **bf)
w = torch.randn(4, 64, 4 * 128, **bf)
out = torch.empty(N, 128, 128, 64, **bf)
y = torch.randn(N, 128, 128, 64, **bf)
g2 = torch.randn(N, 128, 128, 64, **bf)
geom = ops.geom_phase_k4s2p1()
fl = 2.0 * N * 64 * 64 * 4 * 64 * 4 * 128
# G conv 64->128 s2 fwd with stats
x = torch.randn(N, 128, 128, 64, **bf)
w2 = torch.randn(1, 128, 16 * 64, **bf)
out2 = torch.empty(N, 64, 64, 128, **bf)
st = torch.zeros(256, device=dev, dtype=torch.float64)
fl2 = 2.0 * N * 64 * 64 * 128 * 16 * 64
for skip, ld32 in ((0, 1), (0, 0), (0, 1), (0, 0), (2, 1), (7, 1)):
    _lib.debug_set("fprop_skip", skip)
    _lib.debug_set("fprop_ld32", ld32)
    t0 = timeit(lambda: ops.conv_gemm([dy], w, geom, out, 64, (64, 64)))
    t1 = timeit(lambda: ops.conv_gemm([dy], w, geom, out, 64, (64, 64), bwd=dict(y=y, slope=0.2)))
    t2 = timeit(lambda: ops.conv_gemm([dy], w, geom, out, 64, (64, 64), bwd=dict(y=y, slope=0.2, g2=g2)))
    t3 = timeit(lambda: ops.conv_gemm([x], w2, ops.geom_conv_fwd(4, 2, 1), out2, 128, (64, 64), stats=st))
    print(f"skip {skip} ld32 {ld32}: dgrad 128->64 ph4 plain {t0:6.1f} us ({fl/t0/1e6:5.0f} TF) | +bwd act {t1:6.1f} | +bwd act+g2 {t2:6.1f} |"
          f" fwd 64->128 s2 stats {t3:6.1f} us ({fl2/t3/1e6:5.0f} TF)", flush=True)
