"""Blocks-per-SM sweep for the BatchNorm elementwise kernels at the step's tensor sizes (run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
bf = dict(device=dev, dtype=torch.bfloat16)


def timeit(fn, iters=20):
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


for px, c in ((262144, 128), (65536, 256), (61504, 512), (1048576, 64), (16384, 512)):
    y = torch.randn(px, c, **bf)
    d = torch.randn(px, c, **bf)
    o = torch.empty(px, c, **bf)
    sc, sh, mu, iv = (torch.rand(c, device=dev) + 0.5 for _ in range(4))
    sums = torch.randn(2 * c, device=dev, dtype=torch.float64)
    line = [f"px{px} c{c}:"]
    for bps in (2, 4, 8):
        _lib.debug_set("bn_bwd_bps", bps)
        t = timeit(lambda: ops.bn_bwd_apply(y, d, None, 1.0, sc, sh, mu, iv, sums, px, o))
        line.append(f"bwd bps{bps} {t:6.1f}us {px * c * 6 / t / 1e3:5.0f}GB/s")
    for bps in (4, 8, 16):
        _lib.debug_set("bn_act_bps", bps)
        t = timeit(lambda: ops.bn_act(y, sc, sh, o, ops.ACT_LRELU))
        line.append(f"act bps{bps} {t:6.1f}us {px * c * 4 / t / 1e3:5.0f}GB/s")
    print(" | ".join(line), flush=True)
