"""Parity of the direct warp-MMA kernels for the thin (1-6 channel) layers against fp32 CPU convolutions on
the same bf16-rounded operands.  Tolerances: bf16 outputs rel-L2 <= 4e-3; the Cout=1 gradients round
d(logits) to bf16 before the MMA (one more 2^-9 rounding), so rel-L2 <= 6e-3 there."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gan_aug_pfa_b200 import ops  # noqa: E402

DEV = "cuda:0"


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("n,c,h", [(2, 512, 31), (3, 128, 9), (1, 64, 4)])
def test_cout1_conv_forward_and_gradients(n, c, h):
    """Conv2d(C -> 1, k4, s1, p1) + bias (models.py:243): forward, dgrad, wgrad, bias grad."""
    g = torch.Generator().manual_seed(11 + c)
    x = torch.randn(n, c, h, h, generator=g).to(torch.bfloat16)
    w = (torch.randn(1, c, 4, 4, generator=g) / (16 * c) ** 0.5).to(torch.bfloat16)
    b = torch.randn(1, generator=g)
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    ref = F.conv2d(xr, wr, b, stride=1, padding=1)
    oh = h - 1
    xd = nhwc(x).to(DEV)
    wd = w.permute(0, 2, 3, 1).reshape(-1).contiguous().to(DEV)          # [kh][kw][c]
    z = torch.empty(n * h * h, 16, device=DEV)
    logits = torch.full((n, oh, oh), float("nan"), device=DEV)
    ops.cout1_conv_fwd(xd, wd, b.to(DEV), z, logits)
    assert rel(logits.cpu(), ref.detach()[:, 0]) < 2e-3
    dl = torch.randn(n, oh, oh, generator=g) * 0.01
    ref.backward(dl.unsqueeze(1))
    gx = torch.full((n, h, h, c), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.cout1_conv_dgrad(dl.to(DEV), wd, gx)
    assert rel(gx.cpu().float(), nhwc(xr.grad)) < 6e-3
    dw = torch.zeros(16 * c, device=DEV)
    ops.cout1_conv_wgrad(dl.to(DEV), xd, dw)
    assert rel(dw.cpu().view(4, 4, c), wr.grad[0].permute(1, 2, 0)) < 6e-3
    # a second call accumulates
    ops.cout1_conv_wgrad(dl.to(DEV), xd, dw)
    assert rel(dw.cpu().view(4, 4, c), 2 * wr.grad[0].permute(1, 2, 0)) < 6e-3


def test_cout1_conv_channel_slice_input():
    """x may be a channel slice of a wider NHWC buffer (pixel stride > channels)."""
    g = torch.Generator().manual_seed(3)
    n, c, h = 2, 64, 7
    wide = torch.randn(n, h, h, 2 * c, generator=g).to(torch.bfloat16).to(DEV)
    xs = wide[..., c:]
    w = (torch.randn(1, c, 4, 4, generator=g) / 32).to(torch.bfloat16)
    wd = w.permute(0, 2, 3, 1).reshape(-1).contiguous().to(DEV)
    z = torch.empty(n * h * h, 16, device=DEV)
    logits = torch.empty(n, h - 1, h - 1, device=DEV)
    ops.cout1_conv_fwd(xs, wd, None, z, logits)
    ref = F.conv2d(xs.cpu().float().permute(0, 3, 1, 2), w.float(), None, padding=1)
    assert rel(logits.cpu(), ref[:, 0]) < 2e-3


def test_bce_logits_const_f32_loss_grad_and_bias_grad():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 30, 30, generator=g) * 3
    for t in (0.0, 1.0):
        acc = torch.zeros(1, device=DEV, dtype=torch.float64)
        dx = torch.empty(4, 30, 30, device=DEV)
        db = torch.zeros(1, device=DEV)
        ops.bce_logits_const_f32(x.to(DEV), t, 0.5 / x.numel(), dx, acc, db)
        xr = x.clone().requires_grad_(True)
        ref = F.binary_cross_entropy_with_logits(xr, torch.full_like(x, t))
        (0.5 * ref).backward()
        assert abs(float(acc) / x.numel() - float(ref)) < 1e-5
        assert rel(dx.cpu(), xr.grad) < 1e-5
        assert abs(float(db) - float(xr.grad.sum())) < 1e-6 + 1e-4 * abs(float(xr.grad.sum()))
    s = torch.zeros(1, device=DEV)
    ops.sum_f32(x.to(DEV), s)
    assert abs(float(s) - float(x.sum())) < 1e-2
