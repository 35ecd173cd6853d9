"""Microbenchmark of the HBM-bound BatchNorm kernels at Pix2Pix batch-64 shapes (run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops  # noqa: E402

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


for (n, h, c) in ((64, 128, 64), (64, 64, 128), (64, 32, 256), (64, 31, 512)):
    y = torch.randn(n, h, h, c, **bf)
    g = torch.randn(n, h, h, c, **bf)
    g2 = torch.randn(n, h, h, c, **bf)
    dy = torch.empty_like(y)
    o1 = torch.empty_like(y)
    wide = torch.empty(n, h, h, 2 * c, **bf)
    sc = torch.rand(c, device=dev) + 0.5
    sh = torch.randn(c, device=dev)
    mu = torch.randn(c, device=dev)
    iv = torch.rand(c, device=dev) + 0.5
    sums = torch.zeros(2 * c, device=dev, dtype=torch.float64)
    el = y.numel()
    t_act = timeit(lambda: ops.bn_act(y, sc, sh, o1, ops.ACT_LRELU, wide[..., :c], ops.ACT_RELU))
    t_red = timeit(lambda: ops.bn_bwd_reduce(y, g, None, 0.2, sc, sh, mu, iv, sums))
    t_app = timeit(lambda: ops.bn_bwd_apply(y, g, None, 1.0, sc, sh, mu, iv, sums, el // c, dy))
    t_app2 = timeit(lambda: ops.bn_bwd_apply(y, g, g2, 0.2, sc, sh, mu, iv, sums, el // c, dy))
    gb = el * 2 / 1e9
    print(f"n{n} {h}x{h} c{c}: bn_act(2 outs) {t_act:6.1f} us {3*gb/t_act*1e6:5.0f} GB/s | reduce {t_red:6.1f} us {2*gb/t_red*1e6:5.0f} GB/s |"
          f" apply {t_app:6.1f} us {3*gb/t_app*1e6:5.0f} GB/s | apply+g2 {t_app2:6.1f} us {4*gb/t_app2*1e6:5.0f} GB/s", flush=True)
