"""Microbenchmark of the PatchGAN head kernels (Conv 512 -> 1, k4 s1 p1 on 31x31, batch 64; run under gpurun):
the general kernels (cout1_stream=0, with the given cout1_wg_mult values) and the streaming ones (cout1_stream=1).
usage: bench_cout1.py [cout1_wg_mult ...]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
N, H, C = 64, 31, 512


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
mb = y.numel() * 2 / 1e6
for stream, mult in [(0, int(a)) for a in (sys.argv[1:] or ["4"])] + [(1, 4), (0, 4), (1, 4)]:
    _lib.debug_set("cout1_wg_mult", mult)
    _lib.debug_set("cout1_stream", stream)
    t_f = timeit(lambda: ops.cout1_conv_fwd(y, w, bias, z, logits, pre=pre))
    t_f0 = timeit(lambda: ops.cout1_conv_fwd(y, w, bias, z, logits))
    t_d = timeit(lambda: ops.cout1_conv_dgrad(dlog, w, gx, bwd=dict(y=y, scale=scale, shift=shift, slope=0.2, sums=sums)))
    t_d0 = timeit(lambda: ops.cout1_conv_dgrad(dlog, w, gx))
    t_w = timeit(lambda: ops.cout1_conv_wgrad(dlog, y, dw, pre=pre))
    t_w0 = timeit(lambda: ops.cout1_conv_wgrad(dlog, y, dw))
    print(f"stream {stream} wg_mult {mult}: fwd pre {t_f:.1f} us ({mb / t_f:.0f} GB/s) plain {t_f0:.1f} | dgrad bwd-fused {t_d:.1f} "
          f"({2 * mb / t_d:.0f} GB/s) plain {t_d0:.1f} ({mb / t_d0:.0f}) | wgrad pre {t_w:.1f} ({mb / t_w:.0f} GB/s) "
          f"plain {t_w0:.1f}", flush=True)
