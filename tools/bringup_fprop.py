"""Bring-up script for the implicit-GEMM engine (run under gpurun).  Dev tool, not a test: prints
max-abs / rel-L2 errors of gap_conv_gemm against torch fp32 convolutions on bf16-rounded inputs."""
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def report(name, got, ref):
    got = got.float()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()
    ok = rel < 2e-2
    print(f"{'OK  ' if ok else 'FAIL'} {name}: max_abs={err:.4e} rel_l2={rel:.4e} ref_absmax={ref.abs().max().item():.3e}",
          flush=True)
    return ok


def pack_conv(w):  # (Cout,Cin,kh,kw) -> [1][Cout][kh*kw*Cin]
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(1, co, kh * kw * ci).to(torch.bfloat16).contiguous()


def pack_convT_phase(w):  # (Cin,Cout,4,4) -> [4][Cout][4*Cin]
    ci, co, _, _ = w.shape
    out = torch.empty(4, co, 4 * ci, device=w.device, dtype=torch.bfloat16)
    for ph in range(2):
        for pw in range(2):
            for th in range(2):
                for tw in range(2):
                    kh, kw = 3 - ph - 2 * th, 3 - pw - 2 * tw
                    t = th * 2 + tw
                    out[ph * 2 + pw, :, t * ci:(t + 1) * ci] = w[:, :, kh, kw].t().to(torch.bfloat16)
    return out.contiguous()


def test_conv(n, cin, cout, h, k, s, p, bias=False, act=ops.ACT_NONE, stats=False, split=None, tag=""):
    x = torch.randn(n, cin, h, h, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, device=dev) if bias else None
    ho = (h + 2 * p - k) // s + 1
    xh = nhwc(x)
    if split:
        srcs = [xh[..., :split], xh[..., split:]]
    else:
        srcs = [xh]
    out = torch.full((n, ho, ho, cout), float("nan"), device=dev, dtype=torch.bfloat16)
    st = torch.zeros(2 * cout, device=dev, dtype=torch.float64) if stats else None
    ops.conv_gemm(srcs, pack_conv(w), ops.geom_conv_fwd(k, s, p), out, cout, (ho, ho), act=act, bias=b,
                  stats=st)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), w.to(torch.bfloat16).float(), b, stride=s, padding=p)
    ok = True
    if stats:
        rb = ref.to(torch.bfloat16).double()   # the statistics are those of the bf16-rounded tensor
        rs = rb.sum((0, 2, 3))
        rq = (rb ** 2).sum((0, 2, 3))
        e1 = ((st[:cout] - rs).abs().max() / rs.abs().max().clamp_min(1e-9)).item()
        e2 = ((st[cout:] - rq).abs().max() / rq.abs().max()).item()
        sok = e1 < 3e-3 and e2 < 3e-3
        print(f"{'OK  ' if sok else 'FAIL'} stats rel err sum={e1:.3e} sumsq={e2:.3e}")
        ok &= sok
    if act == ops.ACT_LRELU:
        ref = F.leaky_relu(ref, 0.2)
    elif act == ops.ACT_TANH:
        ref = torch.tanh(ref)
    ok &= report(f"conv{tag} n{n} {cin}->{cout} h{h} k{k}s{s}p{p}", out, nhwc(ref))
    return ok


def test_convT(n, cin, cout, h, split=None):
    x = torch.randn(n, cin, h, h, device=dev).to(torch.bfloat16)
    w = torch.randn(cin, cout, 4, 4, device=dev) / (cin * 4) ** 0.5
    xh = nhwc(x)
    srcs = [xh[..., :split], xh[..., split:]] if split else [xh]
    out = torch.full((n, 2 * h, 2 * h, cout), float("nan"), device=dev, dtype=torch.bfloat16)
    ops.conv_gemm(srcs, pack_convT_phase(w), ops.geom_phase_k4s2p1(), out, cout, (h, h))
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(x.float(), w.to(torch.bfloat16).float(), stride=2, padding=1)
    return report(f"convT n{n} {cin}->{cout} h{h}", out, nhwc(ref))


def bench(n, cin, cout, h, k, s, p, iters=20):
    x = nhwc(torch.randn(n, cin, h, h, device=dev).to(torch.bfloat16))
    w = pack_conv(torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5)
    ho = (h + 2 * p - k) // s + 1
    out = torch.empty((n, ho, ho, cout), device=dev, dtype=torch.bfloat16)
    g = ops.geom_conv_fwd(k, s, p)
    for _ in range(3):
        ops.conv_gemm([x], w, g, out, cout, (ho, ho))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ops.conv_gemm([x], w, g, out, cout, (ho, ho))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * n * ho * ho * cout * cin * k * k
    print(f"bench conv n{n} {cin}->{cout} h{h} k{k}s{s}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


def run_all():
    allok = True
    allok &= test_conv(2, 64, 64, 8, 1, 1, 0, tag="[1x1]")
    allok &= test_conv(2, 128, 128, 16, 1, 1, 0, tag="[1x1]")
    allok &= test_conv(4, 256, 512, 16, 1, 1, 0, tag="[1x1]")
    allok &= test_conv(2, 64, 64, 16, 3, 1, 1)
    allok &= test_conv(2, 64, 128, 32, 4, 2, 1)
    allok &= test_conv(4, 128, 256, 16, 4, 2, 1, stats=True)
    allok &= test_conv(8, 512, 512, 4, 4, 2, 1)
    allok &= test_conv(3, 64, 128, 256, 4, 2, 1, act=ops.ACT_LRELU, stats=True)
    allok &= test_conv(2, 256, 512, 32, 4, 1, 1, stats=True)
    allok &= test_conv(2, 512, 1, 31, 4, 1, 1, bias=True)
    allok &= test_conv(3, 64, 64, 31, 4, 1, 1, bias=True, stats=True)
    allok &= test_conv(2, 128, 64, 16, 4, 2, 1, split=64)
    allok &= test_convT(2, 64, 64, 4)
    allok &= test_convT(2, 128, 64, 16, split=64)
    allok &= test_convT(4, 1024, 512, 8, split=512)
    allok &= test_convT(3, 256, 64, 32)
    return allok


if __name__ == "__main__":
    from gan_aug_pfa_b200 import _lib
    t0 = time.time()
    allok = True
    for mt in (1, 2):
        print("=== forced mt", mt)
        _lib.debug_set("fprop_mt", mt)
        allok &= run_all()
    _lib.debug_set("fprop_mt", 0)
    for sp in (3, 7):
        print("=== forced split-K", sp)
        _lib.debug_set("fprop_splits", sp)
        _lib.debug_set("fprop_halo", 0)
        allok &= run_all()
    _lib.debug_set("fprop_splits", 0)
    _lib.debug_set("fprop_halo", 1)
    print("ALL OK" if allok else "SOME FAILED", f"({time.time()-t0:.1f}s)", flush=True)
    if allok:
        for mt in (1, 2):
            print("=== bench mt", mt)
            _lib.debug_set("fprop_mt", mt)
            bench(64, 64, 128, 128, 4, 2, 1)
            bench(64, 128, 256, 64, 4, 2, 1)
            bench(64, 256, 512, 32, 4, 2, 1)
            bench(64, 256, 512, 32, 4, 1, 1)
            bench(64, 512, 512, 16, 4, 2, 1)
