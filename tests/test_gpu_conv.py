"""Parity of the tcgen05 implicit-GEMM kernels (through the C ABI) against fp32 CPU convolutions on the
same bf16-rounded inputs.  Tolerances: outputs are stored in bf16 (2^-9 relative rounding) after fp32
accumulation, so rel-L2 <= 4e-3 for forward/dgrad; weight gradients stay fp32, rel-L2 <= 1e-4."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gan_aug_pfa_b200 import ops  # noqa: E402

DEV = "cuda:0"
FWD_TOL = 4e-3
WG_TOL = 1e-4


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def pack_conv(w):
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(1, co, kh * kw * ci).to(torch.bfloat16).contiguous()


def pack_phase(w_rows_first):
    """w_rows_first: (rows, c, 4, 4) -> [4][rows][4*c] with kh = 3-ph-2th, kw = 3-pw-2tw."""
    r, c = w_rows_first.shape[:2]
    out = torch.empty(4, r, 4 * c, dtype=torch.bfloat16)
    for ph in range(2):
        for pw in range(2):
            for th in range(2):
                for tw in range(2):
                    t = th * 2 + tw
                    out[ph * 2 + pw, :, t * c:(t + 1) * c] = w_rows_first[:, :, 3 - ph - 2 * th, 3 - pw - 2 * tw]
    return out.contiguous()


CONV_CASES = [
    # n, cin, cout, h, k, s, p
    (2, 64, 64, 8, 1, 1, 0),
    (1, 64, 128, 32, 4, 2, 1),
    (3, 128, 256, 16, 4, 2, 1),
    (5, 512, 512, 4, 4, 2, 1),       # 2x2 outputs, several images per M tile
    (1, 512, 512, 2, 4, 2, 1),       # 1x1 output, batch 1 (tile almost empty)
    (2, 256, 512, 32, 4, 1, 1),      # PatchGAN 32 -> 31 (ragged tiles)
    (2, 512, 1, 31, 4, 1, 1),        # Cout = 1
    (2, 64, 64, 20, 3, 1, 1),        # Siamese 3x3, non power-of-two width
    (1, 64, 192, 16, 3, 1, 1),       # Cout not a power of two
]


@pytest.mark.parametrize("n,cin,cout,h,k,s,p", CONV_CASES)
def test_conv2d_forward(n, cin, cout, h, k, s, p):
    g = torch.Generator().manual_seed(n * 1000 + cin + cout + h)
    x = torch.randn(n, cin, h, h, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ho = (h + 2 * p - k) // s + 1
    out = torch.full((n, ho, ho, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * cout, device=DEV, dtype=torch.float64)
    ops.conv_gemm([nhwc(x).to(DEV)], pack_conv(w).to(DEV), ops.geom_conv_fwd(k, s, p), out, cout, (ho, ho),
                  bias=b.to(DEV), stats=stats)
    ref = F.conv2d(x.float(), w.to(torch.bfloat16).float(), b, stride=s, padding=p)
    assert rel(out.cpu().float(), nhwc(ref)) < FWD_TOL
    assert not torch.isnan(out.float()).any()
    s1 = ref.double().sum((0, 2, 3))
    s2 = (ref.double() ** 2).sum((0, 2, 3))
    # the epilogue's statistics are those of the stored (bf16-rounded) tensor: exact against the output itself,
    # bf16-rounding close to the fp32 reference
    o = out.cpu().double()
    o1, o2 = o.sum((0, 1, 2)), (o ** 2).sum((0, 1, 2))
    assert float((stats[:cout].cpu() - o1).abs().max() / o2.sqrt().max()) < 1e-5
    assert rel(stats[cout:].cpu(), o2) < 1e-5
    assert float((stats[:cout].cpu() - s1).abs().max() / s2.sqrt().max()) < 5e-3
    assert rel(stats[cout:].cpu(), s2) < 5e-3


def test_activations_dual_output_and_concat():
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 128, 16, 16, generator=g).to(torch.bfloat16)
    w = torch.randn(64, 128, 4, 4, generator=g) / 45.0
    xh = nhwc(x).to(DEV)
    buf = torch.zeros(2, 8, 8, 192, device=DEV, dtype=torch.bfloat16)      # write into a channel slot
    out1 = torch.empty(2, 8, 8, 64, device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([xh[..., :64], xh[..., 64:]], pack_conv(w).to(DEV), ops.geom_conv_fwd(4, 2, 1), out1, 64, (8, 8),
                  act=ops.ACT_LRELU, out2=buf[..., 64:128], act2=ops.ACT_RELU)
    ref = F.conv2d(x.float(), w.to(torch.bfloat16).float(), None, stride=2, padding=1)
    assert rel(out1.cpu().float(), nhwc(F.leaky_relu(ref, 0.2))) < FWD_TOL
    assert rel(buf[..., 64:128].cpu().float(), nhwc(F.relu(ref))) < FWD_TOL
    assert float(buf[..., :64].abs().max()) == 0 and float(buf[..., 128:].abs().max()) == 0
    # fp32 output + tanh (cold epilogue path)
    o32 = torch.empty(2, 8, 8, 64, device=DEV, dtype=torch.float32)
    ops.conv_gemm([xh], pack_conv(w).to(DEV), ops.geom_conv_fwd(4, 2, 1), o32, 64, (8, 8), act=ops.ACT_TANH)
    assert rel(o32.cpu(), nhwc(torch.tanh(ref))) < 1e-4


@pytest.mark.parametrize("n,cin,cout,h", [(2, 64, 64, 4), (1, 128, 64, 16), (3, 1024, 512, 2), (2, 512, 128, 8)])
def test_conv_transpose2d_forward(n, cin, cout, h):
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(n, cin, h, h, generator=g).to(torch.bfloat16)
    w = torch.randn(cin, cout, 4, 4, generator=g) / (cin * 4) ** 0.5
    out = torch.full((n, 2 * h, 2 * h, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    wp = pack_phase(w.to(torch.bfloat16).permute(1, 0, 2, 3)).to(DEV)
    ops.conv_gemm([nhwc(x).to(DEV)], wp, ops.geom_phase_k4s2p1(), out, cout, (h, h))
    ref = F.conv_transpose2d(x.float(), w.to(torch.bfloat16).float(), stride=2, padding=1)
    assert rel(out.cpu().float(), nhwc(ref)) < FWD_TOL


def test_dgrad_geometries():
    g = torch.Generator().manual_seed(11)
    # stride-2 conv dgrad through the four phases
    n, cin, cout, h = 2, 128, 256, 16
    dy = torch.randn(n, cout, h // 2, h // 2, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 4, 4, generator=g) / 64.0
    x = torch.zeros(n, cin, h, h, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(x, w.to(torch.bfloat16).float(), None, 2, 1), x, dy.float())
    out = torch.empty(n, h, h, cin, device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([nhwc(dy).to(DEV)], pack_phase(w.to(torch.bfloat16).permute(1, 0, 2, 3)).to(DEV),
                  ops.geom_phase_k4s2p1(), out, cin, (h // 2, h // 2))
    assert rel(out.cpu().float(), nhwc(ref)) < FWD_TOL
    # stride-1 conv dgrad with flipped taps (PatchGAN 32 -> 31)
    n, cin, cout, h = 2, 64, 128, 12
    dy = torch.randn(n, cout, h - 1, h - 1, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 4, 4, generator=g) / 32.0
    x = torch.zeros(n, cin, h, h, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(x, w.to(torch.bfloat16).float(), None, 1, 1), x, dy.float())
    wf = w.to(torch.bfloat16).flip(2, 3).permute(1, 2, 3, 0).reshape(1, cin, 16 * cout).contiguous()
    out = torch.empty(n, h, h, cin, device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([nhwc(dy).to(DEV)], wf.to(DEV), ops.geom_conv_dgrad_s1(4, 1), out, cin, (h, h))
    assert rel(out.cpu().float(), nhwc(ref)) < FWD_TOL
    # transposed-conv dgrad = strided gather over dY
    n, cin, cout, h = 2, 128, 64, 8
    dy = torch.randn(n, cout, 2 * h, 2 * h, generator=g).to(torch.bfloat16)
    w = torch.randn(cin, cout, 4, 4, generator=g) / 32.0
    x = torch.zeros(n, cin, h, h, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv_transpose2d(x, w.to(torch.bfloat16).float(), None, 2, 1), x, dy.float())
    wd = w.to(torch.bfloat16).permute(0, 2, 3, 1).reshape(1, cin, 16 * cout).contiguous()
    out = torch.empty(n, h, h, cin, device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([nhwc(dy).to(DEV)], wd.to(DEV), ops.geom_conv_fwd(4, 2, 1), out, cin, (h, h))
    assert rel(out.cpu().float(), nhwc(ref)) < FWD_TOL


@pytest.mark.parametrize("with_bn,with_g2,c0", [(True, True, 0), (True, False, 0), (False, True, 0), (True, True, 64)])
def test_dgrad_backward_fused_epilogue(with_bn, with_g2, c0):
    """Stride-2 conv dgrad whose epilogue applies the activation backward of the layer below and accumulates the
    BatchNorm-backward sums: d = (y*scale+shift > 0) ? g + g2 : slope*g on channels >= c0 (plain g below c0),
    stats = [sum d | sum d*y] of the stored bf16 d."""
    g = torch.Generator().manual_seed(5 + c0)
    n, cin, cout, h, slope = 3, 128, 256, 16, 0.2
    dy = torch.randn(n, cout, h // 2, h // 2, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 4, 4, generator=g) / 64.0
    x = torch.zeros(n, cin, h, h, requires_grad=True)
    (gref,) = torch.autograd.grad(F.conv2d(x, w.to(torch.bfloat16).float(), None, 2, 1), x, dy.float())
    gref = nhwc(gref)                                                    # [n, h, h, cin]
    nb = cin - c0
    y = torch.randn(n, h, h, nb, generator=g).to(torch.bfloat16)
    g2 = torch.randn(n, h, h, nb, generator=g).to(torch.bfloat16) if with_g2 else None
    scale = (torch.rand(nb, generator=g) + 0.5) if with_bn else None
    shift = torch.randn(nb, generator=g) * 0.3 if with_bn else None
    yh = y.float() * scale + shift if with_bn else y.float()
    dref = gref.clone()
    gb = gref[..., c0:]
    dref[..., c0:] = torch.where(yh > 0, gb + (g2.float() if with_g2 else 0.0), slope * gb)
    out = torch.full((n, h, h, cin), float("nan"), device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * nb, device=DEV, dtype=torch.float64)
    bwd = dict(y=y.to(DEV), slope=slope, c0=c0)
    if with_bn:
        bwd.update(scale=scale.to(DEV), shift=shift.to(DEV))
    if with_g2:
        bwd["g2"] = g2.to(DEV)
    ops.conv_gemm([nhwc(dy).to(DEV)], pack_phase(w.to(torch.bfloat16).permute(1, 0, 2, 3)).to(DEV),
                  ops.geom_phase_k4s2p1(), out, cin, (h // 2, h // 2), stats=stats, bwd=bwd)
    o = out.cpu().float()
    # a mask decision can flip only where |y*scale+shift| is at rounding level: compare away from the boundary
    safe = torch.ones_like(o, dtype=torch.bool)
    safe[..., c0:] = yh.abs() > 1e-3
    assert float(((o - dref) * safe).norm() / dref.norm()) < FWD_TOL
    assert float((~safe).float().mean()) < 0.01
    od = o[..., c0:].double()
    s1, s2 = od.sum((0, 1, 2)), (od * y.double()).sum((0, 1, 2))
    denom = (od ** 2).sum((0, 1, 2)).sqrt().max()
    assert float((stats[:nb].cpu() - s1).abs().max() / denom) < 1e-5
    assert float((stats[nb:].cpu() - s2).abs().max() / denom) < 1e-5


WG_CASES = [(2, 64, 128, 8, 1, 1, 0), (2, 64, 64, 16, 3, 1, 1), (2, 64, 128, 32, 4, 2, 1), (3, 256, 512, 16, 4, 2, 1),
            (2, 256, 512, 12, 4, 1, 1), (5, 512, 512, 4, 4, 2, 1)]


@pytest.mark.parametrize("n,cin,cout,h,k,s,p", WG_CASES)
def test_conv2d_wgrad(n, cin, cout, h, k, s, p):
    g = torch.Generator().manual_seed(cin * 7 + cout + h)
    x = torch.randn(n, cin, h, h, generator=g).to(torch.bfloat16)
    ho = (h + 2 * p - k) // s + 1
    dy = torch.randn(n, cout, ho, ho, generator=g).to(torch.bfloat16)
    w = torch.zeros(cout, cin, k, k, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(x.float(), w, None, s, p), w, dy.float())
    out = torch.zeros(cout, k * k, cin, device=DEV)
    ops.conv_wgrad(nhwc(dy).to(DEV), nhwc(x).to(DEV), out, (k, k), s, (-p, -p), k * k * cin, cin)
    assert rel(out.cpu(), ref.permute(0, 2, 3, 1).reshape(cout, k * k, cin)) < WG_TOL
    # accumulation semantics: a second call doubles the gradient (the caller zeroes, like zero_grad())
    ops.conv_wgrad(nhwc(dy).to(DEV), nhwc(x).to(DEV), out, (k, k), s, (-p, -p), k * k * cin, cin)
    assert rel(out.cpu(), 2 * ref.permute(0, 2, 3, 1).reshape(cout, k * k, cin)) < WG_TOL


def test_conv_transpose2d_wgrad_and_row_mask():
    g = torch.Generator().manual_seed(5)
    n, cin, cout, h = 2, 512, 128, 8
    x = torch.randn(n, cin, h, h, generator=g).to(torch.bfloat16)
    dy = torch.randn(n, cout, 2 * h, 2 * h, generator=g).to(torch.bfloat16)
    w = torch.zeros(cin, cout, 4, 4, requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv_transpose2d(x.float(), w, None, 2, 1), w, dy.float())
    out = torch.zeros(cin, 16, cout, device=DEV)
    ops.conv_wgrad(nhwc(x).to(DEV), nhwc(dy).to(DEV), out, (4, 4), 2, (-1, -1), 16 * cout, cout)
    assert rel(out.cpu(), ref.permute(0, 2, 3, 1).reshape(cin, 16, cout)) < WG_TOL
    # m_rows: only the first row exists (the discriminator's 1-channel head, dY padded to 64 channels)
    dyp = torch.zeros(2, 6, 6, 64, dtype=torch.bfloat16)
    dyp[..., 0] = torch.randn(2, 6, 6, generator=g).to(torch.bfloat16)
    xx = torch.randn(2, 64, 7, 7, generator=g).to(torch.bfloat16)
    w1 = torch.zeros(1, 64, 4, 4, requires_grad=True)
    (ref1,) = torch.autograd.grad(F.conv2d(xx.float(), w1, None, 1, 1), w1, dyp[..., :1].permute(0, 3, 1, 2).float())
    o1 = torch.zeros(2, 16, 64, device=DEV)
    ops.conv_wgrad(dyp.to(DEV), nhwc(xx).to(DEV), o1, (4, 4), 1, (-1, -1), 16 * 64, 64, m_rows=1)
    assert rel(o1[0].cpu(), ref1.permute(0, 2, 3, 1).reshape(16, 64)) < WG_TOL
    assert float(o1[1].abs().max()) == 0.0


@pytest.mark.parametrize("n,cin,cout,h,k,p,slot", [(2, 64, 64, 20, 3, 1, False),     # Siamese 3x3 dgrad shape
                                                   (1, 128, 192, 16, 3, 1, True),    # into a channel slot of a wider buffer
                                                   (3, 256, 64, 12, 1, 0, False),    # gate 1x1
                                                   (2, 64, 24, 8, 1, 0, False)])     # ragged last chunk: scalar path
def test_accumulating_epilogue(n, cin, cout, h, k, p, slot):
    """accumulate=True adds the result to what `out` holds (gradient buffers with several contributors): one rounding
    of prev + acc from fp32, vector and scalar store paths, neighbouring channel slots untouched."""
    g = torch.Generator().manual_seed(cin + cout + h)
    x = torch.randn(n, cin, h, h, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    prev = torch.randn(n, h, h, cout, generator=g).to(torch.bfloat16)
    wide = torch.full((n, h, h, cout + 32), 7.0, device=DEV, dtype=torch.bfloat16)
    out = wide[..., 16:16 + cout] if slot else torch.empty(n, h, h, cout, device=DEV, dtype=torch.bfloat16)
    out.copy_(prev.to(DEV))
    ops.conv_gemm([nhwc(x).to(DEV)], pack_conv(w).to(DEV), ops.geom_conv_fwd(k, 1, p), out, cout, (h, h),
                  accumulate=True)
    ref = nhwc(F.conv2d(x.float(), w.to(torch.bfloat16).float(), None, stride=1, padding=p)) + prev.float()
    assert rel(out.cpu().float(), ref) < FWD_TOL
    if slot:
        assert float((wide[..., :16] - 7.0).abs().max()) == 0 and float((wide[..., 16 + cout:] - 7.0).abs().max()) == 0
    with pytest.raises(RuntimeError, match="accumulate excludes"):
        ops.conv_gemm([nhwc(x).to(DEV)], pack_conv(w).to(DEV), ops.geom_conv_fwd(k, 1, p), out, cout, (h, h),
                      accumulate=True, stats=torch.zeros(2 * cout, device=DEV, dtype=torch.float64))


def test_invalid_configurations_are_errors():
    x = torch.zeros(1, 4, 4, 48, device=DEV, dtype=torch.bfloat16)     # 48 channels: not a multiple of 64
    w = torch.zeros(1, 64, 48, device=DEV, dtype=torch.bfloat16)
    out = torch.zeros(1, 4, 4, 64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="multiples of 64"):
        ops.conv_gemm([x], w, ops.geom_conv_fwd(1, 1, 0), out, 64, (4, 4))


def test_full_size_adjointness_property():
    """<conv(x), y> == <x, conv_dgrad(y)> at BASELINE size (batch 64, 128 -> 256 @ 64x64): a
    size-independent check of forward and dgrad against each other."""
    g = torch.Generator(device=DEV).manual_seed(1)
    n, cin, cout, h = 64, 128, 256, 64
    x = torch.randn(n, h, h, cin, device=DEV, generator=g).to(torch.bfloat16)
    y = torch.randn(n, h // 2, h // 2, cout, device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 4, 4, device=DEV, generator=g) / 45.0).to(torch.bfloat16)
    fx = torch.empty(n, h // 2, h // 2, cout, device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([x], w.permute(0, 2, 3, 1).reshape(1, cout, 16 * cin).contiguous(), ops.geom_conv_fwd(4, 2, 1), fx,
                  cout, (h // 2, h // 2))
    gy = torch.empty(n, h, h, cin, device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([y], pack_phase(w.cpu().permute(1, 0, 2, 3)).to(DEV), ops.geom_phase_k4s2p1(), gy, cin,
                  (h // 2, h // 2))
    lhs = float((fx.double() * y.double()).sum())
    rhs = float((x.double() * gy.double()).sum())
    scale = float(fx.double().norm() * y.double().norm())
    assert abs(lhs - rhs) / scale < 1e-4


@pytest.fixture
def force_cta_pairs():
    """fprop_pair = 2: tcgen05 cta_group::2 (M = 256 across two CTAs of a cluster) on every launch where it is legal; by
    default only the 256-wide single-phase layers with long K loops use it (profiles/r2_pair_mode_per_layer.txt)."""
    from gan_aug_pfa_b200 import _lib
    _lib.debug_set("fprop_pair", 2)
    yield
    _lib.debug_set("fprop_pair", 1)


def test_cta_pair_mode_on_every_geometry(force_cta_pairs):
    """The CTA-pair path (each CTA stages half of the B rows, the leader issues the MMAs, completion multicast to both
    CTAs) against fp32 CPU convolutions: stride-2 forward with statistics, the four-phase transposed geometry, halo
    mode (k3 s1), a ragged stride-1 grid, two N tiles, and the backward-fused epilogue."""
    g = torch.Generator().manual_seed(77)
    # stride-2 forward + BatchNorm statistics (N = 128)
    n, cin, cout, h = 16, 64, 128, 64
    x = torch.randn(n, cin, h, h, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 4, 4, generator=g) / 32).to(torch.bfloat16)
    out = torch.full((n, h // 2, h // 2, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * cout, device=DEV, dtype=torch.float64)
    ops.conv_gemm([nhwc(x).to(DEV)], pack_conv(w).to(DEV), ops.geom_conv_fwd(4, 2, 1), out, cout, (h // 2, h // 2), stats=stats)
    assert rel(out.cpu().float(), nhwc(F.conv2d(x.float(), w.float(), None, 2, 1))) < FWD_TOL
    o = out.double()
    assert rel(stats[cout:].cpu(), (o ** 2).sum((0, 1, 2)).cpu()) < 1e-5
    # ConvTranspose2d forward: four phases, halo tiles (N = 64)
    n, cin, cout, h = 8, 128, 64, 32
    x = torch.randn(n, cin, h, h, generator=g).to(torch.bfloat16)
    w = (torch.randn(cin, cout, 4, 4, generator=g) / 24).to(torch.bfloat16)
    out = torch.full((n, 2 * h, 2 * h, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([nhwc(x).to(DEV)], pack_phase(w.permute(1, 0, 2, 3)).to(DEV), ops.geom_phase_k4s2p1(), out, cout, (h, h))
    assert rel(out.cpu().float(), nhwc(F.conv_transpose2d(x.float(), w.float(), None, 2, 1))) < FWD_TOL
    # stride-1 k4 on a ragged 31x31 grid with two 256-wide N tiles (the PatchGAN layer, models.py:238)
    n, cin, cout, h = 8, 128, 512, 32
    x = torch.randn(n, cin, h, h, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 4, 4, generator=g) / 45).to(torch.bfloat16)
    out = torch.full((n, h - 1, h - 1, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([nhwc(x).to(DEV)], pack_conv(w).to(DEV), ops.geom_conv_fwd(4, 1, 1), out, cout, (h - 1, h - 1))
    assert rel(out.cpu().float(), nhwc(F.conv2d(x.float(), w.float(), None, 1, 1))) < FWD_TOL
    # k3 s1 p1 (Siamese double_conv) with an odd number of M tiles per phase: the last pair has a missing partner tile
    n, cin, cout, h = 5, 64, 64, 48
    x = torch.randn(n, cin, h, h, generator=g).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / 24).to(torch.bfloat16)
    out = torch.full((n, h, h, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([nhwc(x).to(DEV)], pack_conv(w).to(DEV), ops.geom_conv_fwd(3, 1, 1), out, cout, (h, h))
    assert rel(out.cpu().float(), nhwc(F.conv2d(x.float(), w.float(), None, 1, 1))) < FWD_TOL
    assert not torch.isnan(out.float()).any()
    # backward-fused epilogue under pairs
    _pair_bwd_epilogue(g)


def _pair_bwd_epilogue(g):
    n, cin, cout, h, slope = 16, 128, 256, 32, 0.2
    dy = torch.randn(n, cout, h // 2, h // 2, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 4, 4, generator=g) / 64.0
    x = torch.zeros(n, cin, h, h, requires_grad=True)
    (gref,) = torch.autograd.grad(F.conv2d(x, w.to(torch.bfloat16).float(), None, 2, 1), x, dy.float())
    gref = nhwc(gref)
    y = torch.randn(n, h, h, cin, generator=g).to(torch.bfloat16)
    g2 = torch.randn(n, h, h, cin, generator=g).to(torch.bfloat16)
    scale, shift = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.3
    yh = y.float() * scale + shift
    dref = torch.where(yh > 0, gref + g2.float(), slope * gref)
    out = torch.full((n, h, h, cin), float("nan"), device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * cin, device=DEV, dtype=torch.float64)
    ops.conv_gemm([nhwc(dy).to(DEV)], pack_phase(w.to(torch.bfloat16).permute(1, 0, 2, 3)).to(DEV), ops.geom_phase_k4s2p1(),
                  out, cin, (h // 2, h // 2), stats=stats,
                  bwd=dict(y=y.to(DEV), slope=slope, g2=g2.to(DEV), scale=scale.to(DEV), shift=shift.to(DEV)))
    o = out.cpu().float()
    safe = yh.abs() > 1e-3
    assert float(((o - dref) * safe).norm() / dref.norm()) < FWD_TOL
    od = o.double()
    denom = (od ** 2).sum((0, 1, 2)).sqrt().max()
    assert float((stats[:cin].cpu() - od.sum((0, 1, 2))).abs().max() / denom) < 1e-5
    assert float((stats[cin:].cpu() - (od * y.double()).sum((0, 1, 2))).abs().max() / denom) < 1e-5
