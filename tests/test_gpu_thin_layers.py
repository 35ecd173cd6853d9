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


def test_cout1_conv_dgrad_with_fused_activation_backward_and_bn_sums():
    """gap_cout1_conv_dgrad_bwd == gap_cout1_conv_dgrad followed by the LeakyReLU-after-BatchNorm mask and the
    [sum d, sum d*y] reduction (bit-exact on the stored tensor: both round g to bf16 before masking)."""
    g = torch.Generator().manual_seed(21)
    n, c, h, slope = 3, 512, 9, 0.2
    w = (torch.randn(1, c, 4, 4, generator=g) / 64).to(torch.bfloat16)
    wd = w.permute(0, 2, 3, 1).reshape(-1).contiguous().to(DEV)
    dl = (torch.randn(n, h - 1, h - 1, generator=g) * 0.05).to(DEV)
    y = torch.randn(n, h, h, c, generator=g).to(torch.bfloat16).to(DEV)
    scale = (torch.rand(c, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(c, generator=g) * 0.3).to(DEV)
    gx = torch.empty(n, h, h, c, device=DEV, dtype=torch.bfloat16)
    ops.cout1_conv_dgrad(dl, wd, gx)
    yh = y.float() * scale + shift
    d_ref = torch.where(yh > 0, gx.float(), (slope * gx.float())).to(torch.bfloat16)
    out = torch.full_like(gx, float("nan"))
    sums = torch.zeros(2 * c, device=DEV, dtype=torch.float64)
    ops.cout1_conv_dgrad(dl, wd, out, bwd=dict(y=y, scale=scale, shift=shift, slope=slope, sums=sums))
    assert torch.equal(out, d_ref)
    dd = d_ref.double()
    s1, s2 = dd.sum((0, 1, 2)), (dd * y.double()).sum((0, 1, 2))
    denom = (dd ** 2).sum((0, 1, 2)).sqrt().max()
    assert float((sums[:c] - s1).abs().max() / denom) < 1e-5
    assert float((sums[c:] - s2).abs().max() / denom) < 1e-5


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


def _slots(x):
    """NCHW (3 channels) -> NHWC bf16 with 4 channel slots (slot 3 = 0)."""
    n, c, h, w = x.shape
    out = torch.zeros(n, h, w, 4, dtype=torch.bfloat16)
    out[..., :c] = nhwc(x).to(torch.bfloat16)
    return out


def _pack_thin(w, groups):
    """(cw, 3*groups, 4, 4) -> [cw][16 taps][4*groups slots] bf16, each group's 3 channels + a zero slot."""
    cw = w.shape[0]
    out = torch.zeros(cw, 16, 4 * groups, dtype=torch.bfloat16)
    for gi in range(groups):
        out[:, :, 4 * gi:4 * gi + 3] = w[:, 3 * gi:3 * gi + 3].permute(0, 2, 3, 1).reshape(cw, 16, 3).to(torch.bfloat16)
    return out.reshape(cw, -1).contiguous()


@pytest.mark.parametrize("n,h,w,cw,groups,bias,two_out", [
    (2, 32, 64, 64, 1, False, True),      # generator first conv: LeakyReLU + ReLU outputs (models.py:177,178,208)
    (3, 48, 32, 64, 2, True, False),      # discriminator first conv on cat(A, B) (models.py:223)
    (2, 20, 36, 128, 1, False, False),    # generator last ConvT seen from its dgrad; partial tiles
    (1, 8, 8, 64, 1, False, True),        # image smaller than one tile: the TMA store box exceeds the tensor
    (2, 4, 12, 64, 2, True, False),
    (1, 256, 256, 64, 2, True, False),
    # 256-pixel rows: the tcgen05 row kernel (A operand read out of the staged image rows, no im2col)
    (2, 12, 256, 64, 2, True, False),     # one ragged unit of 6 output rows per image
    (1, 40, 256, 64, 1, False, True),     # one source, two outputs; units of 8, 8, 4 rows
    (3, 18, 256, 128, 1, False, False),   # 128 output channels (two TMEM stages), 9 rows
    (2, 2, 256, 64, 2, False, False),     # a single output row: both neighbours are padding
    (5, 64, 256, 64, 2, True, False),     # more units than one wave of CTAs handles at once on small grids
])
def test_thin_conv_fwd(n, h, w, cw, groups, bias, two_out):
    g = torch.Generator().manual_seed(h + cw)
    xs = [torch.randn(n, 3, h, w, generator=g).to(torch.bfloat16).float() for _ in range(groups)]
    wt = (torch.randn(cw, 3 * groups, 4, 4, generator=g) / (48 * groups) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(cw, generator=g) if bias else None
    ref = F.conv2d(torch.cat(xs, 1), wt, b, stride=2, padding=1)
    srcs = [_slots(x).to(DEV) for x in xs]
    guard = torch.full((n + 2, h // 2, w // 2, 2 * cw), float("nan"), device=DEV, dtype=torch.bfloat16)
    wide = guard[1:-1]                                     # guard images before and after (TMA stores must clip)
    out1 = wide[..., :cw]                                  # a channel slot of a wider buffer
    out2 = torch.full((n, h // 2, w // 2, cw), float("nan"), device=DEV, dtype=torch.bfloat16) if two_out else None
    ops.thin_conv_fwd(srcs[0], srcs[1] if groups == 2 else None, _pack_thin(wt, groups).to(DEV),
                      b.to(DEV) if bias else None, out1, ops.ACT_LRELU, out2, ops.ACT_RELU)
    assert rel(out1.cpu().float(), nhwc(F.leaky_relu(ref, 0.2))) < 4e-3
    assert torch.isnan(wide[..., cw:].float()).all()       # the neighbouring slot is untouched
    assert torch.isnan(guard[0].float()).all() and torch.isnan(guard[-1].float()).all()
    if two_out:
        assert rel(out2.cpu().float(), nhwc(F.relu(ref))) < 4e-3


@pytest.mark.parametrize("n,h,w,cw,groups,bias", [
    (2, 32, 64, 64, 1, False),       # generator first conv wgrad
    (3, 48, 32, 64, 2, True),        # discriminator first conv wgrad + bias grad
    (2, 20, 36, 128, 1, False),      # generator last ConvT wgrad (wide = its input); partial tiles
    (1, 8, 8, 64, 1, False),         # image smaller than one tile (TMA box exceeds the tensor)
    (5, 64, 64, 64, 2, True),
    # 256-pixel rows: the tcgen05 row kernel (im2col matrix read out of staged image rows, one TMEM tile per CTA)
    (2, 12, 256, 64, 2, True),       # two sources + bias gradient; every CTA owns one output row
    (1, 40, 256, 64, 1, False),      # one source (64 valid im2col rows)
    (3, 18, 256, 128, 1, False),     # 128 wide channels (two dy tiles per row), 9 rows per image
    (2, 2, 256, 64, 2, True),        # a single output row per image: both neighbours are padding
    (3, 256, 256, 64, 2, True),      # 384 rows over 148 CTAs: ranges cross image boundaries, ring and mirror wrap
    (2, 256, 256, 64, 1, False),
])
def test_thin_conv_wgrad(n, h, w, cw, groups, bias):
    g = torch.Generator().manual_seed(h * 3 + cw)
    xs = [torch.randn(n, 3, h, w, generator=g).to(torch.bfloat16).float() for _ in range(groups)]
    dy = torch.randn(n, cw, h // 2, w // 2, generator=g).to(torch.bfloat16).float()
    wt = torch.zeros(cw, 3 * groups, 4, 4, requires_grad=True)
    bt = torch.zeros(cw, requires_grad=True)
    out = F.conv2d(torch.cat(xs, 1), wt, bt, stride=2, padding=1)
    out.backward(dy)
    c = 3 * groups
    krow = 64 if groups == 1 else 128
    dw = torch.zeros(cw, krow, device=DEV)
    db = torch.zeros(cw, device=DEV) if bias else None
    srcs = [_slots(x).to(DEV) for x in xs]
    ops.thin_conv_wgrad(nhwc(dy).to(torch.bfloat16).to(DEV), srcs[0], srcs[1] if groups == 2 else None, dw.view(-1), krow, db)
    ref = wt.grad.permute(0, 2, 3, 1).reshape(cw, 16 * c)                      # [cw][(kh*4+kw)*c + ch]
    assert rel(dw.cpu()[:, :16 * c], ref) < 2e-4
    assert float(dw.cpu()[:, 16 * c:].abs().max()) == 0.0
    if bias:
        assert rel(db.cpu(), bt.grad) < 2e-4


@pytest.mark.parametrize("n,ih,iw,cw,bias,act", [
    (2, 16, 32, 128, True, "tanh"),     # generator last layer (models.py:184,186)
    (3, 24, 16, 64, False, "none"),     # discriminator first conv, input gradient
    (2, 10, 18, 64, True, "tanh"),      # partial tiles
    (1, 4, 4, 128, True, "tanh"),       # input smaller than one halo tile (TMA box exceeds the tensor)
    (2, 2, 6, 64, False, "none"),
])
def test_thin_convT_fwd(n, ih, iw, cw, bias, act):
    g = torch.Generator().manual_seed(ih + cw)
    x = torch.randn(n, cw, ih, iw, generator=g).to(torch.bfloat16).float()
    wt = (torch.randn(cw, 3, 4, 4, generator=g) / (4 * cw) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(3, generator=g) if bias else None
    ref = F.conv_transpose2d(x, wt, b, stride=2, padding=1)
    if act == "tanh":
        ref = torch.tanh(ref)
    w2 = wt.permute(2, 3, 1, 0).reshape(48, cw).to(torch.bfloat16).contiguous()       # [(kh*4+kw)*3 + co][ci]
    obf = torch.full((n, 2 * ih, 2 * iw, 4), float("nan"), device=DEV, dtype=torch.bfloat16)
    o32 = torch.full((n, 2 * ih, 2 * iw, 4), float("nan"), device=DEV)
    ops.thin_convT_fwd(nhwc(x).to(torch.bfloat16).to(DEV), w2.to(DEV), b.to(DEV) if bias else None,
                       ops.ACT_TANH if act == "tanh" else ops.ACT_NONE, obf, o32)
    assert rel(o32.cpu()[..., :3], nhwc(ref)) < 1e-4
    assert rel(obf.cpu().float()[..., :3], nhwc(ref)) < 4e-3
    assert float(o32.cpu()[..., 3].abs().max()) == 0.0


def test_uint8_image_input_and_output_paths():
    """uint8 HWC -> normalised NHWC bf16 (dataset.py:155-159) and the generator's uint8 image
    output (generate_synthetic_data.py:69-88: x*0.5+0.5 -> to_pil_image = mul(255).byte())."""
    g = torch.Generator().manual_seed(8)
    img = torch.randint(0, 256, (2, 12, 10, 3), generator=g, dtype=torch.uint8)
    out = torch.full((2, 12, 10, 4), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.u8_hwc_to_nhwc_bf16(img.to(DEV), out)
    ref = (img.float() / 255 - 0.5) / 0.5
    assert rel(out.cpu().float()[..., :3], ref) < 4e-3
    assert float(out.cpu().float()[..., 3].abs().max()) == 0.0
    # last ConvT layer with the uint8 epilogue
    n, ih, iw, cw = 2, 8, 16, 64
    x = torch.randn(n, cw, ih, iw, generator=g).to(torch.bfloat16).float()
    wt = (torch.randn(cw, 3, 4, 4, generator=g) / (4 * cw) ** 0.5).to(torch.bfloat16).float()
    b = torch.randn(3, generator=g)
    y = torch.tanh(F.conv_transpose2d(x, wt, b, stride=2, padding=1))
    w2 = wt.permute(2, 3, 1, 0).reshape(48, cw).to(torch.bfloat16).contiguous()
    o32 = torch.empty(n, 2 * ih, 2 * iw, 4, device=DEV)
    ou8 = torch.zeros(n, 2 * ih, 2 * iw, 3, device=DEV, dtype=torch.uint8)
    ops.thin_convT_fwd(nhwc(x).to(torch.bfloat16).to(DEV), w2.to(DEV), b.to(DEV), ops.ACT_TANH, None, o32, ou8)
    want = ((o32.cpu()[..., :3] * 0.5 + 0.5) * 255).clamp(0, 255).to(torch.uint8)      # same fp32 values, same truncation
    assert torch.equal(ou8.cpu(), want)
    ref_u8 = ((nhwc(y) * 0.5 + 0.5) * 255).to(torch.uint8).int()
    assert int((ou8.cpu().int() - ref_u8).abs().max()) <= 1                               # vs the fp32 reference: +-1 level


@pytest.mark.parametrize("n,c,h", [(2, 512, 31), (3, 128, 9), (64, 512, 31)])
def test_cout1_conv_with_batchnorm_leakyrelu_fused_into_the_loads(n, c, h):
    """gap_cout1_conv_fwd / _wgrad with in_scale / in_shift / in_slope: the head reads the RAW conv output y of the layer
    below and forms LeakyReLU(y*scale + shift) on the fly (models.py:239-240 fused into models.py:243) — same results as
    running on the materialised activated tensor, which then never exists in HBM."""
    g = torch.Generator().manual_seed(41 + c + n)
    y = torch.randn(n, h, h, c, generator=g).to(torch.bfloat16)
    scale = torch.rand(c, generator=g) + 0.5
    shift = torch.randn(c, generator=g) * 0.3
    w = (torch.randn(1, c, 4, 4, generator=g) / (16 * c) ** 0.5).to(torch.bfloat16)
    b = torch.randn(1, generator=g)
    act = F.leaky_relu(y.float() * scale + shift, 0.2).to(torch.bfloat16)             # what bn_act would have stored
    wd = w.permute(0, 2, 3, 1).reshape(-1).contiguous().to(DEV)
    yd, ad = y.to(DEV), act.to(DEV)
    z = torch.empty(n * h * h, 16, device=DEV)
    want = torch.empty(n, h - 1, h - 1, device=DEV)
    got = torch.full_like(want, float("nan"))
    ops.cout1_conv_fwd(ad, wd, b.to(DEV), z, want)
    ops.cout1_conv_fwd(yd, wd, b.to(DEV), z, got, pre=(scale.to(DEV), shift.to(DEV), 0.2))
    assert rel(got.cpu(), want.cpu()) < 2e-3          # one differently-rounded fma per element at most
    ref = F.conv2d(act.float().permute(0, 3, 1, 2), w.float(), b, stride=1, padding=1)[:, 0]
    assert rel(got.cpu(), ref) < 3e-3
    dl = (torch.randn(n, h - 1, h - 1, generator=g) * 0.01).to(DEV)
    dw_want = torch.zeros(16 * c, device=DEV)
    dw_got = torch.zeros(16 * c, device=DEV)
    ops.cout1_conv_wgrad(dl, ad, dw_want)
    ops.cout1_conv_wgrad(dl, yd, dw_got, pre=(scale.to(DEV), shift.to(DEV), 0.2))
    assert rel(dw_got.cpu(), dw_want.cpu()) < 2e-3
    # a ragged last slab / zero rows must stay zero after the transform (BN(0) != 0): compare against fp32 torch
    xr = act.float().permute(0, 3, 1, 2)
    ref_w = torch.nn.grad.conv2d_weight(xr, (1, c, 4, 4), dl.cpu().unsqueeze(1), 1, 1)
    assert rel(dw_got.cpu().view(4, 4, c), ref_w[0].permute(1, 2, 0)) < 6e-3


@pytest.mark.parametrize("ih,iw,oh,ow", [(300, 420, 256, 256), (512, 512, 256, 256), (100, 90, 256, 256), (256, 256, 256, 256),
                                         (777, 333, 128, 128)])
def test_on_device_resize_matches_the_reference_dataset_transforms(ih, iw, oh, ow):
    """gap_resize_u8_to_nhwc_bf16 == dataset.py's ToTensor -> JointResize -> JointNormalize (dataset.py:28-29,136-159):
    torchvision's tensor resize with BILINEAR is F.interpolate(mode="bilinear", align_corners=False, antialias=True)
    (pinned on the CPU against torchvision itself in tests/test_oracle_golden.py).  bf16 output: rel-L2 <= 4e-3 and every
    pixel within one bf16 ulp of the fp32 value.  Labels: NEAREST, bit-exact."""
    g = torch.Generator().manual_seed(ih + iw)
    img = torch.randint(0, 256, (2, ih, iw, 3), generator=g, dtype=torch.uint8)
    x = img.permute(0, 3, 1, 2).float() / 255.0                                        # ToTensor
    ref = F.interpolate(x, size=(oh, ow), mode="bilinear", align_corners=False, antialias=True) * 2.0 - 1.0
    out = torch.full((2, oh, ow, 4), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.resize_u8_to_nhwc_bf16(img.to(DEV), out)
    got = out.cpu().float()
    assert rel(got[..., :3], nhwc(ref)) < 4e-3
    assert float((got[..., :3] - nhwc(ref)).abs().max()) <= 2.0 ** -7             # bf16 spacing below 2 is 2^-7 at most... one ulp
    assert float(got[..., 3].abs().max()) == 0.0
    f32 = torch.full((2, 3, oh, ow), float("nan"), device=DEV)
    ops.resize_u8_to_nhwc_bf16(img.to(DEV), None, f32)                            # the DataLoader's fp32 NCHW form
    assert float((f32.cpu() - ref).abs().max()) < 2e-6
    lab = (torch.rand(2, ih, iw, generator=g) < 0.3).long()
    want = F.interpolate(lab.unsqueeze(1).float(), size=(oh, ow), mode="nearest").squeeze(1).long()
    res = torch.full((2, oh, ow), -1, device=DEV, dtype=torch.int64)
    ops.resize_nearest_i64(lab.to(DEV), res)
    assert torch.equal(res.cpu(), want)
