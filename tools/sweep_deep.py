"""block_n x mt sweep on the deep (8x8 .. 2x2) layers, with BatchNorm statistics on (run under gpurun)."""
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


N = 64
cases = []
for cin, cout, g in ((512, 512, 8), (512, 1024, 8), (512, 512, 4), (512, 1024, 4), (512, 512, 2), (512, 512, 1)):
    src = torch.randn(N, 2 * g, 2 * g, cin, **bf)
    w = torch.randn(1, cout, 16 * cin, **bf)
    out = torch.empty(N, g, g, cout, **bf)
    st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
    fl = 2.0 * N * g * g * cout * 16 * cin
    cases.append((f"s2  {cin}->{cout} g{g}", fl,
                  lambda src=src, w=w, out=out, cout=cout, g=g, st=st: ops.conv_gemm([src], w, ops.geom_conv_fwd(4, 2, 1), out, cout, (g, g), stats=st)))
for cin, cout, g in ((1024, 512, 8), (512, 512, 8), (1024, 512, 4), (512, 512, 4), (1024, 512, 2), (512, 512, 2), (512, 512, 1)):
    src = torch.randn(N, g, g, cin, **bf)
    w = torch.randn(4, cout, 4 * cin, **bf)
    out = torch.empty(N, 2 * g, 2 * g, cout, **bf)
    st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
    fl = 2.0 * N * g * g * 4 * cout * 4 * cin
    cases.append((f"ph4 {cin}->{cout} g{g}", fl,
                  lambda src=src, w=w, out=out, cout=cout, g=g, st=st: ops.conv_gemm([src], w, ops.geom_phase_k4s2p1(), out, cout, (g, g), stats=st)))

for name, fl, fn in cases:
    line = [f"{name:20s}"]
    _lib.debug_set("fprop_mt", 0)
    _lib.debug_set("fprop_block_n", 0)
    line.append(f"default {timeit(fn):6.1f}us")
    for bn in (64, 128, 256):
        for mt in (1, 2):
            _lib.debug_set("fprop_mt", mt)
            _lib.debug_set("fprop_block_n", bn)
            try:
                t = timeit(fn)
                line.append(f"bn{bn} mt{mt}: {t:6.1f}")
            except RuntimeError:
                line.append(f"bn{bn} mt{mt}:   ERR ")
    print(" | ".join(line), flush=True)
