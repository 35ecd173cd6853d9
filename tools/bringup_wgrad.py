"""Bring-up script for gap_conv_wgrad (run under gpurun).  Dev tool, not a test."""
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def report(name, got, ref):
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()
    ok = rel < 5e-3
    print(f"{'OK  ' if ok else 'FAIL'} {name}: max_abs={err:.4e} rel_l2={rel:.4e} ref_absmax={ref.abs().max().item():.3e}",
          flush=True)
    return ok


def test_conv_wgrad(n, cin, cout, h, k, s, p):
    x = torch.randn(n, cin, h, h, device=dev).to(torch.bfloat16)
    ho = (h + 2 * p - k) // s + 1
    dy = torch.randn(n, cout, ho, ho, device=dev).to(torch.bfloat16)
    xf = x.float().requires_grad_(False)
    w = torch.zeros(cout, cin, k, k, device=dev, requires_grad=True)
    y = F.conv2d(xf, w, None, stride=s, padding=p)
    (ref,) = torch.autograd.grad(y, w, dy.float())
    out = torch.zeros(cout, k * k, cin, device=dev)
    ops.conv_wgrad(nhwc(dy), nhwc(x), out, (k, k), s, (-p, -p), k * k * cin, cin)
    torch.cuda.synchronize()
    return report(f"conv wgrad n{n} {cin}->{cout} h{h} k{k}s{s}p{p}", out,
                  ref.permute(0, 2, 3, 1).reshape(cout, k * k, cin))


def test_convT_wgrad(n, cin, cout, h):
    x = torch.randn(n, cin, h, h, device=dev).to(torch.bfloat16)
    dy = torch.randn(n, cout, 2 * h, 2 * h, device=dev).to(torch.bfloat16)
    w = torch.zeros(cin, cout, 4, 4, device=dev, requires_grad=True)
    y = F.conv_transpose2d(x.float(), w, None, stride=2, padding=1)
    (ref,) = torch.autograd.grad(y, w, dy.float())
    out = torch.zeros(cin, 16, cout, device=dev)
    ops.conv_wgrad(nhwc(x), nhwc(dy), out, (4, 4), 2, (-1, -1), 16 * cout, cout)
    torch.cuda.synchronize()
    return report(f"convT wgrad n{n} {cin}->{cout} h{h}", out, ref.permute(0, 2, 3, 1).reshape(cin, 16, cout))


def bench(n, cin, cout, h, k, s, p, iters=10):
    x = nhwc(torch.randn(n, cin, h, h, device=dev).to(torch.bfloat16))
    ho = (h + 2 * p - k) // s + 1
    dy = nhwc(torch.randn(n, cout, ho, ho, device=dev).to(torch.bfloat16))
    out = torch.zeros(cout, k * k, cin, device=dev)
    for _ in range(3):
        ops.conv_wgrad(dy, x, out, (k, k), s, (-p, -p), k * k * cin, cin)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.conv_wgrad(dy, x, out, (k, k), s, (-p, -p), k * k * cin, cin)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * n * ho * ho * cout * cin * k * k
    print(f"bench wgrad n{n} {cin}->{cout} h{h} k{k}s{s}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    variant = sys.argv[1] if len(sys.argv) > 1 else ""
    t0 = time.time()
    ok = True
    ok &= test_conv_wgrad(2, 64, 128, 8, 1, 1, 0)
    ok &= test_conv_wgrad(2, 128, 128, 16, 1, 1, 0)
    ok &= test_conv_wgrad(2, 64, 64, 16, 3, 1, 1)
    ok &= test_conv_wgrad(2, 64, 128, 32, 4, 2, 1)
    ok &= test_conv_wgrad(4, 256, 512, 16, 4, 2, 1)
    ok &= test_conv_wgrad(2, 256, 512, 32, 4, 1, 1)
    ok &= test_conv_wgrad(8, 512, 512, 4, 4, 2, 1)
    ok &= test_convT_wgrad(2, 64, 64, 4)
    ok &= test_convT_wgrad(2, 512, 128, 16)
    print("ALL OK" if ok else "SOME FAILED", f"({time.time()-t0:.1f}s)", flush=True)
    if ok:
        bench(64, 64, 128, 128, 4, 2, 1)
        bench(64, 128, 256, 64, 4, 2, 1)
        bench(64, 256, 512, 32, 4, 2, 1)
        bench(64, 256, 512, 32, 4, 1, 1)
        bench(64, 512, 512, 16, 4, 2, 1)
        for cols in (512,):
            _lib.debug_set("wgrad_acc_cols", cols)
            print("acc_cols", cols)
            bench(64, 64, 128, 128, 4, 2, 1)
            bench(64, 128, 256, 64, 4, 2, 1)
            bench(64, 256, 512, 32, 4, 2, 1)
