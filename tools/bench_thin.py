"""Microbenchmark of the thin-layer kernels at the Pix2Pix batch-64 shapes (run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
N = 64


def timeit(fn, iters=10):
    """Device time per call: the launches are replayed from a CUDA graph, so the host's launch cost (tensor-map
    encodes, ctypes) is not what is measured when a kernel is shorter than its launch."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(iters):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    graph.replay()
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
wide128 = torch.randn(N, 128, 128, 128, **bf)
wide64 = torch.randn(N, 128, 128, 64, **bf)
wcol128 = torch.randn(48, 128, **bf)
wcol64 = torch.randn(48, 64, **bf)
b3 = torch.randn(3, device=dev)
fake_bf = torch.zeros(N, 256, 256, 4, **bf)
fake_f32 = torch.zeros(N, 256, 256, 4, device=dev)
dw8 = torch.zeros(64, 96, device=dev)
dw4 = torch.zeros(64, 48, device=dev)
dw4l = torch.zeros(128, 48, device=dev)
db = torch.zeros(64, device=dev)
t4 = timeit(lambda: ops.thin_convT_fwd(wide128, wcol128, b3, ops.ACT_TANH, fake_bf, fake_f32))
t5 = timeit(lambda: ops.thin_convT_fwd(wide64, wcol64, None, ops.ACT_NONE, fake_bf, None))
t6 = timeit(lambda: ops.thin_conv_wgrad(o64, xa, xb, dw8, 96, db))
t7 = timeit(lambda: ops.thin_conv_wgrad(o64, xa, None, dw4, 48, None))
print(f"G.last fwd (convT 128->3, tanh, bf16+f32 out) {t4:.1f} us | D.0 input grad (convT 64->3, bf16 out) {t5:.1f} us | "
      f"D.0 wgrad {t6:.1f} us | G.0 wgrad {t7:.1f} us", flush=True)
for cps in [int(a) for a in sys.argv[1:]] or [0]:
  for tc in (1, 0, 1, 0):
    _lib.debug_set("thin_tc", tc)          # 1: tcgen05 row kernel (256-pixel rows), 0: the mma.sync tile kernel
    _lib.debug_set("thin_skip", cps)
    t1 = timeit(lambda: ops.thin_conv_fwd(xa, xb, w8, bias, o64, ops.ACT_LRELU))
    t2 = timeit(lambda: ops.thin_conv_fwd(xa, None, w4, None, o64, ops.ACT_LRELU, o64b, ops.ACT_RELU))
    t3 = timeit(lambda: ops.thin_conv_fwd(xa, None, w4b, None, o128))
    t6 = timeit(lambda: ops.thin_conv_wgrad(o64, xa, xb, dw8, 96, db))
    t7 = timeit(lambda: ops.thin_conv_wgrad(o64, xa, None, dw4, 48, None))
    t8 = timeit(lambda: ops.thin_conv_wgrad(wide128, xa, None, dw4l, 48, None))
    t6b = timeit(lambda: ops.thin_conv_wgrad(o64, xa, xb, dw8, 96, None))
    print(f"tc {tc}: D.0 wgrad without the bias gradient {t6b:.1f} us", flush=True)
    print(f"tc {tc}: D.0 wgrad {t6:.1f} us | G.0 wgrad {t7:.1f} us | G.last wgrad (128 wide) {t8:.1f} us", flush=True)
    print(f"tc {tc} skip {cps}: D.0 fwd (6->64) {t1:.1f} us | G.0 fwd (3->64, 2 outs) {t2:.1f} us | G.last dgrad (3->128) {t3:.1f} us",
          flush=True)
