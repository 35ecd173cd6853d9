"""Ablation of gap_conv_wgrad (run under gpurun): time the kernel with parts switched off."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def bench(n, mc, nc, gh, k, s, p, skip, iters=10, convT=False):
    """conv wgrad: mop = dy [n,gh,gh,mc], nop = x [n,h,h,nc]"""
    h = (gh - 1) * s - 2 * p + k if not convT else gh * 2
    dy = torch.randn(n, gh, gh, mc, device=dev).to(torch.bfloat16)
    x = torch.randn(n, h, h, nc, device=dev).to(torch.bfloat16)
    out = torch.zeros(mc, k * k, nc, device=dev)
    _lib.debug_set("wgrad_skip", skip)
    for _ in range(3):
        ops.conv_wgrad(dy, x, out, (k, k), s, (-p, -p), k * k * nc, nc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.conv_wgrad(dy, x, out, (k, k), s, (-p, -p), k * k * nc, nc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * n * gh * gh * mc * nc * k * k
    return ms * 1e3, fl / ms / 1e9


SHAPES = [
    ("m128 n64 g64 s2", 64, 128, 64, 64, 4, 2, 1),
    ("m256 n128 g32 s2", 64, 256, 128, 32, 4, 2, 1),
    ("m512 n256 g31 s1", 64, 512, 256, 31, 4, 1, 1),
    ("m512 n256 g16 s2", 64, 512, 256, 16, 4, 2, 1),
    ("m1024 n256 g16 s2 (convT)", 64, 1024, 256, 16, 4, 2, 1),
]
MASKS = [(0, "full"), (1, "no atomics"), (2, "no mma"), (3, "no mma/atomics"), (4, "no N tma"), (8, "no M tma"),
         (12, "no tma"), (13, "mma only"), (14, "nothing but epilogue"), (15, "nothing")]
if __name__ != "__main__":
    SHAPES = []
extra = [a for a in sys.argv[1:]] if __name__ == "__main__" else []
for kv in extra:
    k, v = kv.split("=")
    _lib.debug_set(k, int(v))
for name, n, mc, nc, gh, k, s, p in SHAPES:
    print(f"== {name}")
    for mask, what in MASKS:
        us, tf = bench(n, mc, nc, gh, k, s, p, mask)
        print(f"   {what:24s} {us:8.1f} us  {tf:7.1f} TF/s-equivalent", flush=True)
