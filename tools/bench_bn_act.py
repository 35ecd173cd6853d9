"""bn_act: the general kernel (bn_act_fast=0) vs the slope-activation kernel (bn_act_fast=1) at the batch-64 shapes of the
Pix2Pix step, one and two outputs, bn_act_bps (resident blocks per SM) swept (run under gpurun)."""
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


for (n, h, c) in ((64, 128, 64), (64, 64, 128), (64, 32, 256), (64, 16, 512), (64, 8, 512)):
    y = torch.randn(n, h, h, c, **bf)
    o1 = torch.empty_like(y)
    wide = torch.empty(n, h, h, 2 * c, **bf)
    sc = torch.rand(c, device=dev) + 0.5
    sh = torch.randn(c, device=dev)
    gb = y.numel() * 2 / 1e9
    row = []
    for fast, bps in ((0, 4), (1, 2), (1, 3), (1, 4), (0, 4), (1, 3)):
        _lib.debug_set("bn_act_fast", fast)
        _lib.debug_set("bn_act_bps", bps)
        t1 = timeit(lambda: ops.bn_act(y, sc, sh, o1, ops.ACT_LRELU))
        t2 = timeit(lambda: ops.bn_act(y, sc, sh, o1, ops.ACT_LRELU, wide[..., :c], ops.ACT_RELU))
        row.append(f"fast{fast}/bps{bps}: {t1:5.1f} us {2 * gb / t1 * 1e6:5.0f} GB/s, 2 outs {t2:5.1f} us {3 * gb / t2 * 1e6:5.0f} GB/s")
    print(f"n{n} {h}x{h} c{c}: " + " | ".join(row), flush=True)

