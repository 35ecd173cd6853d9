"""Microbenchmark of the thin-layer kernels at the Pix2Pix batch-64 shapes (run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
N = 64


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


bf = dict(device=dev, dtype=torch.bfloat16)
xa = torch.randn(N, 256, 256, 4, **bf)
xb = torch.randn(N, 256, 256, 4, **bf)
w8 = torch.randn(64, 128, **bf)
w4 = torch.randn(64, 64, **bf)
w4b = torch.randn(128, 64, **bf)
bias = torch.randn(64, device=dev)
o64 = torch.empty(N, 128, 128, 64, **bf)
o64b = torch.empty(N, 128, 128, 64, **bf)
o128 = torch.empty(N, 128, 128, 128, **bf)
for cps in [int(a) for a in sys.argv[1:]] or [0]:
    _lib.debug_set("thin_skip", cps)
    t1 = timeit(lambda: ops.thin_conv_fwd(xa, xb, w8, bias, o64, ops.ACT_LRELU))
    t2 = timeit(lambda: ops.thin_conv_fwd(xa, None, w4, None, o64, ops.ACT_LRELU, o64b, ops.ACT_RELU))
    t3 = timeit(lambda: ops.thin_conv_fwd(xa, None, w4b, None, o128))
    print(f"skip {cps}: D.0 fwd (6->64) {t1:.1f} us | G.0 fwd (3->64, 2 outs) {t2:.1f} us | G.last dgrad (3->128) {t3:.1f} us",
          flush=True)
