"""Parity of the HBM-bound kernels (BatchNorm, losses, Adam, packing, im2col / col2im) against fp32 CPU
restatements of the same formulas (oracle functions where they exist)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gan_aug_pfa_b200 import ops  # noqa: E402
from oracle import pix2pix_oracle as O  # noqa: E402

DEV = "cuda:0"


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def test_bn_finalize_act_and_buffers():
    g = torch.Generator().manual_seed(0)
    n, h, c = 3, 10, 64
    y = (torch.randn(n, h, h, c, generator=g) * 2 + 0.5).to(torch.bfloat16)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    yf = y.float()
    stats = torch.cat([yf.double().sum((0, 1, 2)), (yf.double() ** 2).sum((0, 1, 2))]).to(DEV)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    nbt = torch.zeros((), device=DEV, dtype=torch.int64)
    scale, shift, mean, invstd = (torch.empty(c, device=DEV) for _ in range(4))
    ops.bn_finalize(stats, n * h * h, gamma.to(DEV), beta.to(DEV), 1e-5, 0.1, 2, rm, rv, nbt, scale, shift, mean, invstd)
    sd = {"bn.weight": gamma, "bn.bias": beta, "bn.running_mean": torch.zeros(c), "bn.running_var": torch.ones(c),
          "bn.num_batches_tracked": torch.tensor(0)}
    nb = {}
    ref = O.batchnorm2d(yf.permute(0, 3, 1, 2), sd, "bn", True, nb)
    ref = O.batchnorm2d(yf.permute(0, 3, 1, 2), sd, "bn", True, nb)      # two identical updates (repeat=2)
    assert torch.allclose(rm.cpu(), nb["bn.running_mean"], atol=1e-5)
    assert torch.allclose(rv.cpu(), nb["bn.running_var"], atol=1e-5)
    assert int(nbt) == 2 and float(stats.abs().max()) == 0.0
    o1 = torch.empty(n, h, h, c, device=DEV, dtype=torch.bfloat16)
    cat = torch.zeros(n, h, h, 2 * c, device=DEV, dtype=torch.bfloat16)
    ops.bn_act(y.to(DEV), scale, shift, o1, ops.ACT_LRELU, cat[..., c:], ops.ACT_RELU)
    refh = ref.permute(0, 2, 3, 1)
    assert rel(o1.cpu().float(), F.leaky_relu(refh, 0.2)) < 4e-3
    assert rel(cat[..., c:].cpu().float(), F.relu(refh)) < 4e-3
    assert float(cat[..., :c].abs().max()) == 0.0
    # eval mode scale/shift
    ops.bn_eval_scale_shift(gamma.to(DEV), beta.to(DEV), rm, rv, 1e-5, scale, shift)
    ref_s = gamma / torch.sqrt(rv.cpu() + 1e-5)
    assert torch.allclose(scale.cpu(), ref_s, rtol=1e-5)


@pytest.mark.parametrize("n,h,c", [(3, 10, 64), (64, 2, 512), (5, 33, 128)])
def test_bn_backward_from_dgrad_epilogue_sums(n, h, c):
    """gap_bn_bwd_finalize turns the raw [sum d, sum d*y] a backward-fused dgrad epilogue leaves behind into the
    [sum d, sum d*xhat] gap_bn_bwd_apply expects: same dy, dgamma, dbeta as the stand-alone reduce pass (slope 1 = the
    activation backward already applied), the raw sums are re-zeroed, parameter gradients are accumulated."""
    g = torch.Generator().manual_seed(n + c)
    y = (torch.randn(n, h, h, c, generator=g) * 2 + 0.5).to(torch.bfloat16).to(DEV)
    d = torch.randn(n, h, h, c, generator=g).to(torch.bfloat16).to(DEV)
    gamma, beta = (torch.rand(c, generator=g) + 0.5).to(DEV), torch.randn(c, generator=g).to(DEV)
    yd, dd = y.double(), d.double()
    count = n * h * h
    mean64 = yd.mean((0, 1, 2))
    var64 = (yd ** 2).mean((0, 1, 2)) - mean64 ** 2
    invstd = torch.rsqrt(var64 + 1e-5).float()
    mean = mean64.float()
    scale = gamma * invstd
    shift = beta - mean * scale
    raw = torch.cat([dd.sum((0, 1, 2)), (dd * yd).sum((0, 1, 2))])
    dg_a, db_a = torch.ones(c, device=DEV), torch.ones(c, device=DEV)
    sums = torch.zeros(2 * c, device=DEV, dtype=torch.float64)
    dy_a = torch.full_like(d, float("nan"))
    ops.bn_bwd_finalize(raw, mean, invstd, dg_a, db_a, sums)
    assert float(raw.abs().max()) == 0.0
    ops.bn_bwd_apply(y, d, None, 1.0, scale, shift, mean, invstd, sums, count, dy_a)
    # the stand-alone path: reduce (slope 1: d passes through whatever the sign of yhat) + apply + parameter gradients
    sums_b = torch.zeros(2 * c, device=DEV, dtype=torch.float64)
    dy_b = torch.full_like(d, float("nan"))
    ops.bn_bwd_reduce(y, d, None, 1.0, scale, shift, mean, invstd, sums_b)
    ops.bn_bwd_apply(y, d, None, 1.0, scale, shift, mean, invstd, sums_b, count, dy_b)
    dg_b, db_b = torch.ones(c, device=DEV), torch.ones(c, device=DEV)
    ops.bn_param_grads(sums_b, dg_b, db_b)
    assert rel(dy_a.float(), dy_b.float()) < 4e-3
    assert rel(dg_a, dg_b) < 1e-4 and rel(db_a, db_b) < 1e-5
    # and fp64 math
    xhat = (yd - mean64) * invstd.double()
    db_ref, dg_ref = dd.sum((0, 1, 2)), (dd * xhat).sum((0, 1, 2))
    dy_ref = scale.double() * (dd - db_ref / count - xhat * dg_ref / count)
    assert rel(dy_a.double(), dy_ref) < 4e-3
    assert rel(dg_a.double() - 1.0, dg_ref) < 1e-4 and rel(db_a.double() - 1.0, db_ref) < 1e-4


@pytest.mark.parametrize("c,slope,two", [(64, 0.2, True), (512, 0.0, False), (128, 0.2, False)])
def test_bn_backward(c, slope, two):
    g = torch.Generator().manual_seed(c)
    n, h = 2, 6
    y = (torch.randn(n, h, h, c, generator=g) + 0.2).to(torch.bfloat16)
    g1 = torch.randn(n, h, h, c, generator=g).to(torch.bfloat16)
    g2 = torch.randn(n, h, h, c, generator=g).to(torch.bfloat16) if two else None
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    yv = y.float().requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    mean = yv.mean((0, 1, 2))
    var = ((yv - mean) ** 2).mean((0, 1, 2))
    yhat = (yv - mean) * torch.rsqrt(var + 1e-5) * gm + bt
    a = F.leaky_relu(yhat, slope) if slope > 0 else F.relu(yhat)
    loss = (a * g1.float()).sum()
    if two:
        loss = loss + (F.relu(a) * g2.float()).sum()
    dy_ref, dg_ref, db_ref = torch.autograd.grad(loss, [yv, gm, bt])
    cnt = n * h * h
    md, vd = mean.detach(), var.detach()
    invstd = torch.rsqrt(vd + 1e-5)
    scale = (gamma * invstd).to(DEV)
    shift = (beta - md * gamma * invstd).to(DEV)
    sums = torch.zeros(2 * c, device=DEV, dtype=torch.float64)
    dy = torch.empty(n, h, h, c, device=DEV, dtype=torch.bfloat16)
    yd, g1d = y.to(DEV), g1.to(DEV)
    g2d = g2.to(DEV) if two else None
    ops.bn_bwd_reduce(yd, g1d, g2d, slope, scale, shift, md.to(DEV), invstd.to(DEV), sums)
    ops.bn_bwd_apply(yd, g1d, g2d, slope, scale, shift, md.to(DEV), invstd.to(DEV), sums, cnt, dy)
    dgam, dbet = torch.zeros(c, device=DEV), torch.zeros(c, device=DEV)
    ops.bn_param_grads(sums, dgam, dbet)
    assert rel(dy.cpu().float(), dy_ref) < 6e-3
    assert rel(dgam.cpu(), dg_ref) < 1e-3 and rel(dbet.cpu(), db_ref) < 1e-3
    assert float(sums.abs().max()) == 0.0
    # activation-only mode
    ops.bn_bwd_apply(yd, g1d, g2d, slope, None, None, None, None, None, 0, dy)
    yv2 = y.float().requires_grad_(True)
    a2 = F.leaky_relu(yv2, slope) if slope > 0 else F.relu(yv2)
    l2 = (a2 * g1.float()).sum() + ((F.relu(a2) * g2.float()).sum() if two else 0)
    (ref2,) = torch.autograd.grad(l2, yv2)
    assert rel(dy.cpu().float(), ref2) < 4e-3


def test_bce_and_l1_kernels():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 30, 30, 1, generator=g) * 3
    for t in (0.0, 1.0):
        acc = torch.zeros(1, device=DEV, dtype=torch.float64)
        dx = torch.zeros(4, 30, 30, 64, device=DEV, dtype=torch.bfloat16)
        ops.bce_logits_const(x.to(DEV), t, 0.5 / x.numel(), dx, acc)
        xv = x.clone().requires_grad_(True)
        ref = O.bce_with_logits(xv, torch.full_like(xv, t))
        (gref,) = torch.autograd.grad(0.5 * ref, xv)
        assert abs(float(acc) / x.numel() - float(ref)) < 1e-5
        assert rel(dx[..., 0].cpu().float(), gref[..., 0]) < 4e-3
        assert float(dx[..., 1:].abs().max()) == 0.0
    # generator output backward: L1 + tanh'
    n, h = 2, 16
    fake = torch.tanh(torch.randn(n, h, h, 4, generator=g))
    real = torch.rand(n, 3, h, h, generator=g) * 2 - 1
    dfd = torch.randn(n, h, h, 4, generator=g) * 1e-3
    acc = torch.zeros(1, device=DEV, dtype=torch.float64)
    dpre = torch.zeros(n, h, h, 4, device=DEV, dtype=torch.bfloat16)
    lam = 100.0 / (n * 3 * h * h)
    ops.gen_out_bwd(fake.to(DEV), real.to(DEV), dfd.to(DEV), lam, dpre, acc)
    f3 = fake[..., :3]
    r3 = real.permute(0, 2, 3, 1)
    assert abs(float(acc) - float((f3 - r3).abs().sum())) / float(acc) < 1e-5
    ref = (dfd[..., :3] + lam * torch.sign(f3 - r3)) * (1 - f3 * f3)
    assert rel(dpre[..., :3].cpu().float(), ref) < 4e-3


@pytest.mark.parametrize("n,h,w,c", [(3, 7, 5, 64), (2, 9, 11, 128), (1, 5, 3, 512), (2, 4, 4, 1024)])
def test_bn_act_slope_kernel_equals_the_general_one(n, h, w, c):
    """gap_bn_act: the kernel for identity / LeakyReLU / ReLU (channel group per thread, `bn_act_fast` knob) against the
    general kernel: bit-identical outputs, one and two outputs, outputs that are channel slots of a wider tensor, ragged
    pixel counts."""
    from gan_aug_pfa_b200 import _lib
    g = torch.Generator().manual_seed(17 + c)
    y = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).to(DEV)
    sc = (torch.rand(c, generator=g) + 0.5).to(DEV)
    sh = torch.randn(c, generator=g).to(DEV)
    try:
        for act1, act2 in ((ops.ACT_LRELU, None), (ops.ACT_RELU, ops.ACT_LRELU), (ops.ACT_NONE, ops.ACT_RELU)):
            got = []
            for knob in (0, 1):
                _lib.debug_set("bn_act_fast", knob)
                o1 = torch.full((n, h, w, c), float("nan"), device=DEV, dtype=torch.bfloat16)
                wide = torch.full((n, h, w, 2 * c), float("nan"), device=DEV, dtype=torch.bfloat16)
                if act2 is None:
                    ops.bn_act(y, sc, sh, wide[..., c:], act1)
                else:
                    ops.bn_act(y, sc, sh, o1, act1, wide[..., :c], act2)
                got.append((o1.cpu(), wide.cpu()))
            for a, b in zip(got[0], got[1]):
                assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    finally:
        _lib.debug_set("bn_act_fast", 1)


def test_gen_out_bwd_four_pixel_kernel_equals_the_scalar_one():
    """gap_gen_out_bwd / _u8: the four-pixels-per-thread kernel (3 channels in 4-slot tensors) against the per-pixel
    kernel it replaces on that layout (`gen_out_c3` knob): identical bf16 gradients, the same L1 sum."""
    from gan_aug_pfa_b200 import _lib
    g = torch.Generator().manual_seed(9)
    n, h, w = 3, 20, 36
    fake = torch.tanh(torch.randn(n, h, w, 4, generator=g)).to(DEV)
    dfd = (torch.randn(n, h, w, 4, generator=g) * 1e-3).to(DEV)
    real_f = (torch.rand(n, 3, h, w, generator=g) * 2 - 1).to(DEV)
    real_u8 = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8).to(DEV)
    lam = 100.0 / (n * 3 * h * w)
    try:
        for real in (real_f, real_u8):
            for d in (dfd, None):
                out = []
                for knob in (0, 1):
                    _lib.debug_set("gen_out_c3", knob)
                    acc = torch.zeros(1, device=DEV, dtype=torch.float64)
                    dpre = torch.zeros(n, h, w, 4, device=DEV, dtype=torch.bfloat16)
                    ops.gen_out_bwd(fake, real, d, lam, dpre, acc)
                    out.append((dpre.cpu(), float(acc)))
                assert torch.equal(out[0][0], out[1][0])
                assert float(out[1][0][..., 3].abs().max()) == 0.0
                assert abs(out[0][1] - out[1][1]) <= 1e-6 * abs(out[0][1])
    finally:
        _lib.debug_set("gen_out_c3", 1)


def test_adam_kernel_matches_oracle():
    g = torch.Generator().manual_seed(1)
    n = 1027                                            # exercises the vector body and the scalar tail
    p0, m0, v0 = torch.randn(n + 1, generator=g), torch.zeros(n + 1), torch.zeros(n + 1)
    for decoupled, wd in ((False, 0.0), (True, 0.01)):
        p, m, v = p0.clone(), m0.clone(), v0.clone()
        pd, md, vd = (t.clone().to(DEV)[:n] for t in (p0, m0, v0))
        for step in range(1, 4):
            grad = torch.randn(n + 1, generator=g)
            O.adam_update(p, grad / 2, m, v, step, 1e-3, 0.5, 0.999, 1e-8, wd, decoupled)
            ops.adam_flat(pd, grad.to(DEV)[:n], md, vd, 1e-3, 0.5, 0.999, 1e-8, wd, decoupled, step, grad_scale=0.5)
        assert torch.allclose(pd.cpu(), p[:n], atol=2e-6)
        assert torch.allclose(vd.cpu(), v[:n], rtol=1e-4, atol=1e-9)


def test_layout_conversion_roundtrip_at_the_module_boundary():
    g = torch.Generator().manual_seed(2)
    n, h = 2, 12
    a = torch.randn(n, 3, h, h, generator=g)
    b = torch.randn(n, 3, h, h, generator=g)
    an = torch.zeros(n, h, h, 4, device=DEV, dtype=torch.bfloat16)
    bn = torch.zeros_like(an)
    ops.nchw_to_nhwc_bf16(a.to(DEV), an)
    ops.nchw_to_nhwc_bf16(b.to(DEV), bn)
    assert torch.equal(an[..., :3].cpu().float(), a.to(torch.bfloat16).float().permute(0, 2, 3, 1))
    out = torch.empty(n, 3, h, h, device=DEV)
    ops.nhwc_to_nchw_f32(an, out, 3)
    assert torch.equal(out.cpu(), a.to(torch.bfloat16).float())


def test_pack_weights_modes():
    g = torch.Generator().manual_seed(4)
    co, ci = 96, 64
    w = torch.randn(co, ci, 4, 4, generator=g)
    native = w.permute(0, 2, 3, 1).contiguous().view(-1).to(DEV)        # [co][kh][kw][ci]
    out = torch.empty(1, co, 16 * ci, device=DEV, dtype=torch.bfloat16)
    ops.pack_weights(native, 0, out, 0, 1, co, co, (4, 4), ci, ci, 16 * ci, (16 * ci, 1, 4 * ci, ci))
    assert torch.equal(out.cpu().float().view(co, 4, 4, ci), w.permute(0, 2, 3, 1).to(torch.bfloat16).float())
    ph = torch.empty(4, ci, 4 * co, device=DEV, dtype=torch.bfloat16)
    ops.pack_weights(native, 0, ph, 2, 4, ci, ci, (2, 2), co, co, 4 * co, (1, 16 * ci, 4 * ci, ci))
    ref = torch.empty(4, ci, 4 * co)
    for p in range(4):
        for t in range(4):
            kh, kw = 3 - (p >> 1) - 2 * (t >> 1), 3 - (p & 1) - 2 * (t & 1)
            ref[p, :, t * co:(t + 1) * co] = w[:, :, kh, kw].t()
    assert torch.equal(ph.cpu().float(), ref.to(torch.bfloat16).float())
    fl = torch.empty(1, ci, 16 * 128, device=DEV, dtype=torch.bfloat16)          # flipped, channel-padded 96 -> 128
    ops.pack_weights(native, 0, fl, 1, 1, ci, ci, (4, 4), co, 128, 16 * 128, (1, 16 * ci, 4 * ci, ci))
    got = fl.cpu().float().view(ci, 4, 4, 128)
    assert torch.equal(got[..., :co], w.flip(2, 3).permute(1, 2, 3, 0).to(torch.bfloat16).float())
    assert float(got[..., co:].abs().max()) == 0.0


def test_pack_plan_matches_single_tensor_pack():
    """gap_pack_weights_multi (one launch, tiled transposes) == gap_pack_weights on every mode the engines use."""
    g = torch.Generator().manual_seed(5)
    co, ci = 96, 72
    flat = torch.randn(2 * co * ci * 16 + 40, generator=g).to(DEV)
    off2 = co * ci * 16 + 24
    cases = [  # (w_off, mode, n_phase, rows, rows_pad, taps, c, c_pad, krow, strides)
        (0, 0, 1, co, co, (4, 4), ci, ci, 16 * ci, (16 * ci, 1, 4 * ci, ci)),
        (0, 2, 4, ci, ci, (2, 2), co, co, 4 * co, (1, 16 * ci, 4 * ci, ci)),
        (off2, 1, 1, ci, ci, (4, 4), co, 128, 16 * 128, (1, 16 * ci, 4 * ci, ci)),
        (off2, 0, 1, 64, 64, (1, 1), 128, 128, 128, (1, 64, 0, 0)),
        (off2, 1, 1, ci, 80, (4, 4), 1, 64, 16 * 64, (1, 16 * ci, 4 * ci, ci)),
    ]
    plan = ops.PackPlan()
    outs, refs = [], []
    for (o, mode, nph, rows, rpad, taps, c, cpad, krow, st) in cases:
        a = torch.zeros(nph, rpad, krow, device=DEV, dtype=torch.bfloat16)
        b = torch.full((nph, rpad, krow), 7.0, device=DEV, dtype=torch.bfloat16)
        plan.add(flat, o, a, mode, nph, rows, rpad, taps, c, cpad, krow, st)
        ops.pack_weights(flat, o, b, mode, nph, rows, rpad, taps, c, cpad, krow, st)
        outs.append(a)
        refs.append(b)
    plan.run()
    plan.run()
    torch.cuda.synchronize()
    for a, b in zip(outs, refs):
        assert torch.equal(a.cpu().float(), b.cpu().float())


@pytest.mark.parametrize("n,h,w,slots", [(2, 8, 12, 4), (1, 5, 7, 4), (3, 16, 16, 8)])
def test_nchw_f32_to_nhwc_bf16_image_layouts(n, h, w, slots):
    """fp32 NCHW images -> NHWC bf16 channel slots: the vectorised 3-into-4 path (h*w % 4 == 0), its scalar fallback
    (odd sizes) and a wider pixel stride; unused slots stay zero."""
    g = torch.Generator().manual_seed(n * h + w)
    x = torch.rand(n, 3, h, w, generator=g) * 2 - 1
    out = torch.zeros(n, h, w, slots, device=DEV, dtype=torch.bfloat16)
    ops.nchw_to_nhwc_bf16(x.to(DEV), out)
    ref = torch.zeros(n, h, w, slots, dtype=torch.bfloat16)
    ref[..., :3] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(out.cpu(), ref)
