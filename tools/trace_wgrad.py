"""clock64 timeline of CTA 0 of one gap_conv_wgrad launch (run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")


def to_i32(u):
    return u - (1 << 32) if u >= (1 << 31) else u


def run(n, mc, nc, gh, k, s, p, skip):
    h = (gh - 1) * s - 2 * p + k
    dy = torch.randn(n, gh, gh, mc, device=dev).to(torch.bfloat16)
    x = torch.randn(n, h, h, nc, device=dev).to(torch.bfloat16)
    out = torch.zeros(mc, k * k, nc, device=dev)
    tr = torch.zeros(1024, device=dev, dtype=torch.int64)
    _lib.debug_set("wgrad_skip", skip)
    for _ in range(2):
        ops.conv_wgrad(dy, x, out, (k, k), s, (-p, -p), k * k * nc, nc)
    ptr = tr.data_ptr()
    _lib.debug_set("trace_ptr_lo", to_i32(ptr & 0xFFFFFFFF))
    _lib.debug_set("trace_ptr_hi", to_i32(ptr >> 32))
    ops.conv_wgrad(dy, x, out, (k, k), s, (-p, -p), k * k * nc, nc)
    torch.cuda.synchronize()
    _lib.debug_set("trace_ptr_lo", 0)
    _lib.debug_set("trace_ptr_hi", 0)
    t = tr.cpu().tolist()
    t0 = min(v for v in t[:512] if v > 0)
    print(f"--- skip={skip} m{mc} n{nc} g{gh} s{s}")
    print("iter: prod(after empty wait, after issue)  mma(after full wait, after commit)")
    for i in range(0, 24):
        if t[2 * i] == 0:
            break
        print(f"{i:3d}: P {t[2*i]-t0:7d} {t[2*i+1]-t0:7d}   M {t[256+2*i]-t0:7d} {t[256+2*i+1]-t0:7d}")
    n_it = sum(1 for i in range(128) if t[256 + 2 * i] > 0)
    last = t[256 + 2 * (n_it - 1) + 1]
    print(f"iters {n_it}; mma loop total {last - t[256]} cyc -> {(last - t[256]) / max(1, n_it - 1):.0f} cyc/iter; "
          f"epilogue wait start {t[512]-t0} done-bar {t[513]-t0} epilogue end {t[514]-t0}")
    return t



def summary(n, mc, nc, gh, k, s, p, skip):
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        t = run(n, mc, nc, gh, k, s, p, skip)
    return t


import statistics
for shape in ((64, 128, 64, 64, 4, 2, 1), (64, 512, 256, 31, 4, 1, 1), (64, 512, 256, 16, 4, 2, 1)):
    for skip in (15, 12, 8, 4, 0, 2):
        t = summary(*shape, skip)
        iss = [t[2 * i + 1] - t[2 * i] for i in range(4, 40) if t[2 * i + 1] > 0]
        gap = [t[2 * i + 2] - t[2 * i + 1] for i in range(4, 40) if t[2 * i + 2] > 0]
        mma = [t[256 + 2 * i + 1] - t[256 + 2 * i] for i in range(4, 40) if t[256 + 2 * i + 1] > 0]
        per = [t[256 + 2 * i + 2] - t[256 + 2 * i] for i in range(4, 40) if t[256 + 2 * i + 2] > 0]
        print(f"shape m{shape[1]} n{shape[2]} g{shape[3]} s{shape[5]} skip={skip:2d}: producer issue {statistics.median(iss):6.0f} "
              f"other {statistics.median(gap):6.0f} | mma issue {statistics.median(mma):6.0f} iter period {statistics.median(per):6.0f}")
