"""Can an HBM-bound BatchNorm pass hide under a tensor-bound GEMM of ANOTHER stream?  (run under gpurun)

Times (CUDA events, warm) a large conv_gemm launch alone, a bn_act pass alone, and both issued together on two
streams; prints how much of the shorter one the concurrent run hides.  The persistent GEMM CTA (352 threads x ~150
registers, ~200 KiB of shared memory) leaves room for about one 256-thread elementwise CTA per SM.
usage: overlap_probe.py [iters]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf = dict(device=dev, dtype=torch.bfloat16)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = 64

# D.8-like GEMM: 512 -> 256, k4 s1 dgrad geometry on a 32x32 grid (the largest launch of the iteration, ~205 us)
src = torch.randn(N, 32, 32, 512, **bf)
w = torch.randn(1, 256, 16 * 512, **bf) / 90
out = torch.empty(N, 32, 32, 256, **bf)
geom = ops.geom_conv_dgrad_s1(4, 1)


def gemm():
    ops.conv_gemm([src], w, geom, out, 256, (32, 32))


# bn_act on a 64 x 64 x 64 x 128 tensor (67 MB in, 67 MB out)
y = torch.randn(N, 64, 64, 128, **bf)
a_out = torch.empty_like(y)
scale = torch.rand(128, device=dev) + 0.5
shift = torch.randn(128, device=dev)


def bn(k=1):
    for _ in range(k):
        ops.bn_act(y, scale, shift, a_out, ops.ACT_LRELU)


def timed(fn):
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


side = torch.cuda.Stream(dev)


def both(k):
    def run():
        cur = torch.cuda.current_stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            bn(k)
        gemm()
        cur.wait_stream(side)
    return run


t_g = timed(gemm)
print(f"gemm alone            {t_g:7.1f} us", flush=True)
for k in (1, 4):
    t_b = timed(lambda: bn(k))
    t_c = timed(both(k))
    hidden = t_g + t_b - t_c
    print(f"bn_act x{k} alone       {t_b:7.1f} us | together {t_c:7.1f} us | serial sum {t_g + t_b:7.1f} us | "
          f"hidden {hidden:6.1f} us = {100 * hidden / min(t_g, t_b):.0f} % of the shorter", flush=True)
