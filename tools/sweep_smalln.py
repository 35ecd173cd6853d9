"""Knob sweep (mt / halo / stages) on the small-N implicit-GEMM layers (run under gpurun)."""
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
cases = []
# dgrad / ConvT fwd, phase geometry: cin -> cout on a (g x g) input grid
for cin, cout, g in ((128, 64, 64), (256, 128, 32), (256, 64, 64), (512, 128, 32)):
    src = torch.randn(N, g, g, cin, **bf)
    w = torch.randn(4, cout, 4 * cin, **bf)
    out = torch.empty(N, 2 * g, 2 * g, cout, **bf)
    fl = 2.0 * N * g * g * 4 * cout * 4 * cin
    cases.append((f"ph4 {cin}->{cout} g{g}", fl,
                  lambda src=src, w=w, out=out, cout=cout, g=g: ops.conv_gemm([src], w, ops.geom_phase_k4s2p1(), out, cout, (g, g))))
# conv s2 forward
for cin, cout, g in ((64, 128, 64), (128, 256, 32), (64, 256, 64)):
    src = torch.randn(N, 2 * g, 2 * g, cin, **bf)
    w = torch.randn(1, cout, 16 * cin, **bf)
    out = torch.empty(N, g, g, cout, **bf)
    fl = 2.0 * N * g * g * cout * 16 * cin
    cases.append((f"s2  {cin}->{cout} g{g}", fl,
                  lambda src=src, w=w, out=out, cout=cout, g=g: ops.conv_gemm([src], w, ops.geom_conv_fwd(4, 2, 1), out, cout, (g, g))))

for name, fl, fn in cases:
    line = [f"{name:20s}"]
    for mt in (1, 2):
        for halo in (0, 1):
            for st in (0, 3):
                _lib.debug_set("fprop_mt", mt)
                _lib.debug_set("fprop_halo", halo)
                _lib.debug_set("fprop_stages", st)
                try:
                    t = timeit(fn)
                    line.append(f"mt{mt} h{halo} st{st}: {t:6.1f}us {fl / t / 1e6:5.0f}TF")
                except RuntimeError as e:
                    line.append(f"mt{mt} h{halo} st{st}: ERR")
    print(" | ".join(line), flush=True)
