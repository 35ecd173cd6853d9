"""Split-K sweep on the small-M conv layers (run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")


def bench(n, cin, cout, h, k, s, p, iters=20):
    x = torch.randn(n, h, h, cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(1, cout, k * k * cin, device=dev) / (cin * k * k) ** 0.5).to(torch.bfloat16)
    ho = (h + 2 * p - k) // s + 1
    out = torch.empty((n, ho, ho, cout), device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
    g = ops.geom_conv_fwd(k, s, p)
    for _ in range(3):
        ops.conv_gemm([x], w, g, out, cout, (ho, ho), stats=st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.conv_gemm([x], w, g, out, cout, (ho, ho), stats=st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for (n, cin, cout, h) in ((64, 512, 512, 4), (64, 512, 512, 8), (64, 512, 512, 16), (1, 256, 512, 32), (1, 128, 256, 64), (4, 512, 512, 16)):
    line = f"n{n} {cin}->{cout} h{h} k4s2:"
    _lib.debug_set("fprop_splitk", 0)
    _lib.debug_set("fprop_splits", 0)
    line += f" off {bench(n, cin, cout, h, 4, 2, 1):6.1f} us |"
    for sp in (2, 4, 8, 16):
        _lib.debug_set("fprop_splits", sp)
        for bn in (0, 128, 256):
            _lib.debug_set("fprop_block_n", bn)
            line += f" s{sp}/bn{bn} {bench(n, cin, cout, h, 4, 2, 1):6.1f}"
        line += " |"
    _lib.debug_set("fprop_block_n", 0)
    print(line, flush=True)
