"""Seeded random-shape sweep of the implicit-GEMM engine against torch's fp32 CPU convolutions: ragged and
non-square grids, batch sizes that do not fill an M tile, output-channel counts that are not tile multiples, two
concatenated sources, all four geometries (Conv2d forward k1/k3/k4 at stride 1/2, stride-1 dgrad, the four-phase
transposed geometry) and the weight-gradient kernel.  Same tolerances as test_gpu_conv.py (bf16 outputs 4e-3, fp32
weight gradients 1e-4).  Twenty-eight forward / sixteen wgrad cases, fixed seeds."""
import random

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gan_aug_pfa_b200 import ops  # noqa: E402
from test_gpu_conv import FWD_TOL, WG_TOL, nhwc, pack_conv, pack_phase, rel  # noqa: E402

DEV = "cuda:0"


def _cases(seed, count, kinds):
    rng = random.Random(seed)
    out = []
    for i in range(count):
        kind = kinds[i % len(kinds)]
        n = rng.choice([1, 2, 3, 5, 7])
        cin = rng.choice([64, 128, 192, 256])
        cout = rng.choice([16, 48, 64, 80, 128, 144, 256, 272, 320])
        h, w = rng.randint(3, 37), rng.randint(3, 41)
        out.append((kind, n, cin, cout, h, w, seed * 1000 + i))
    return out


@pytest.mark.parametrize("kind,n,cin,cout,h,w,seed", _cases(11, 28, ["k4s2", "k4s1", "k3s1", "k1", "phase", "dgrad_s1", "concat"]))
def test_conv_gemm_random_shapes(kind, n, cin, cout, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "k4s2":
        h, w = 2 * max(2, h // 2), 2 * max(2, w // 2)
    if kind in ("k4s1",) and min(h, w) < 4:
        h, w = h + 3, w + 3
    x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16)
    xd = nhwc(x).to(DEV)
    if kind in ("k4s2", "k4s1", "k3s1", "k1", "concat"):
        k, s, p = {"k4s2": (4, 2, 1), "k4s1": (4, 1, 1), "k3s1": (3, 1, 1), "k1": (1, 1, 0), "concat": (3, 1, 1)}[kind]
        wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(torch.bfloat16)
        b = torch.randn(cout, generator=g)
        ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        out = torch.full((n, ho, wo, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
        srcs = [xd] if kind != "concat" or cin < 128 else [xd[..., :64], xd[..., 64:]]
        ops.conv_gemm(srcs, pack_conv(wt).to(DEV), ops.geom_conv_fwd(k, s, p), out, cout, (ho, wo), bias=b.to(DEV),
                      act=ops.ACT_LRELU)
        ref = F.leaky_relu(F.conv2d(x.float(), wt.float(), b, stride=s, padding=p), 0.2)
    elif kind == "phase":
        wt = (torch.randn(cin, cout, 4, 4, generator=g) / (4 * cin) ** 0.5).to(torch.bfloat16)
        out = torch.full((n, 2 * h, 2 * w, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.conv_gemm([xd], pack_phase(wt.permute(1, 0, 2, 3)).to(DEV), ops.geom_phase_k4s2p1(), out, cout, (h, w))
        ref = F.conv_transpose2d(x.float(), wt.float(), None, 2, 1)
    else:   # stride-1 k4 dgrad: x plays dY [n, cin, h, w] of a conv cout -> cin over an (h+1) x (w+1) input
        wt = (torch.randn(cin, cout, 4, 4, generator=g) / (16 * cin) ** 0.5).to(torch.bfloat16)      # (Cout_conv, Cin_conv)
        out = torch.full((n, h + 1, w + 1, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
        wf = wt.flip(2, 3).permute(1, 2, 3, 0).reshape(1, cout, 16 * cin).contiguous()
        ops.conv_gemm([xd], wf.to(DEV), ops.geom_conv_dgrad_s1(4, 1), out, cout, (h + 1, w + 1))
        ref = torch.nn.grad.conv2d_input((n, cout, h + 1, w + 1), wt.float(), x.float(), 1, 1)
    got = out.cpu().float()
    assert not torch.isnan(got).any()
    assert rel(got, ref.permute(0, 2, 3, 1)) < FWD_TOL, (kind, n, cin, cout, h, w)


@pytest.mark.parametrize("kind,n,cin,cout,h,w,seed", _cases(23, 16, ["k4s2", "k4s1", "k3s1", "k1"]))
def test_conv_wgrad_random_shapes(kind, n, cin, cout, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    cout = max(64, (cout + 63) // 64 * 64)          # the M operand's channels come in 64-wide TMA boxes
    if kind == "k4s2":
        h, w = 2 * max(2, h // 2), 2 * max(2, w // 2)
    if kind == "k4s1" and min(h, w) < 4:
        h, w = h + 3, w + 3
    k, s, p = {"k4s2": (4, 2, 1), "k4s1": (4, 1, 1), "k3s1": (3, 1, 1), "k1": (1, 1, 0)}[kind]
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    x = torch.randn(n, cin, h, w, generator=g).to(torch.bfloat16)
    dy = torch.randn(n, cout, ho, wo, generator=g).to(torch.bfloat16)
    ref = torch.nn.grad.conv2d_weight(x.float(), (cout, cin, k, k), dy.float(), s, p)
    out = torch.zeros(cout, k * k, cin, device=DEV)
    ops.conv_wgrad(nhwc(dy).to(DEV), nhwc(x).to(DEV), out, (k, k), s, (-p, -p), k * k * cin, cin if k > 1 else 0)
    assert rel(out.cpu(), ref.permute(0, 2, 3, 1).reshape(cout, k * k, cin)) < WG_TOL, (kind, n, cin, cout, h, w)
