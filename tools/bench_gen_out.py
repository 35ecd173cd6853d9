"""gap_gen_out_bwd / _u8 at batch 64, 256x256: the per-pixel kernel (gen_out_c3=0) vs the four-pixel kernel (run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
n, h = 64, 256
g = torch.Generator().manual_seed(0)
fake = torch.tanh(torch.randn(n, h, h, 4, generator=g)).to(dev)
dfd = (torch.randn(n, h, h, 4, generator=g) * 1e-3).to(dev)
real_f = (torch.rand(n, 3, h, h, generator=g) * 2 - 1).to(dev)
real_u8 = torch.randint(0, 256, (n, h, h, 3), generator=g, dtype=torch.uint8).to(dev)
dpre = torch.zeros(n, h, h, 4, device=dev, dtype=torch.bfloat16)
acc = torch.zeros(1, device=dev, dtype=torch.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()                       # cold L2, as inside the training step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


for knob in (0, 1, 0, 1):
    _lib.debug_set("gen_out_c3", knob)
    tf = timeit(lambda: ops.gen_out_bwd(fake, real_f, dfd, 1e-5, dpre, acc))
    tu = timeit(lambda: ops.gen_out_bwd(fake, real_u8, dfd, 1e-5, dpre, acc))
    mb_f = (fake.numel() * 4 * 2 + real_f.numel() * 4 + dpre.numel() * 2) / 1e6
    mb_u = (fake.numel() * 4 * 2 + real_u8.numel() + dpre.numel() * 2) / 1e6
    print(f"gen_out_c3={knob}: fp32 real {tf:6.1f} us ({mb_f / tf:5.2f} TB/s)   uint8 real {tu:6.1f} us ({mb_u / tu:5.2f} TB/s)", flush=True)
