"""Parity at the configuration the benchmark is quoted on (BASELINE.json config 2: batch 64 per GPU, 256x256).

One known-answer test per distinct layer of SURVEY.md App. A at batch 64 — forward, dgrad and wgrad through the
C ABI against torch's fp32 CPU convolutions (F.conv2d / F.conv_transpose2d and their torch.nn.grad counterparts)
on the same bf16-rounded operands — plus one full batch-64 Pix2PixTrainer.train_step against the CPU oracle's
gan_train_step.  These sizes reach the code paths the small KATs of test_gpu_conv.py never see: two M tiles per
work item, the 148-CTA persistent wrap-around, wgrad split-K, both TMA producers.

Tolerances: bf16-stored outputs rel-L2 <= 4e-3 (one 2^-9 rounding after fp32 accumulation); fp32 weight
gradients rel-L2 <= 2e-4 (the fp32 CPU reference itself carries ~1e-5 of summation noise over 10^6 pixels); the
Cout = 1 gradients round d(logits) to bf16 first, <= 6e-3; the full step uses the tolerances of
test_gpu_engine.py (losses 2e-3, per-tensor gradient cosine >= 0.97, BatchNorm buffers 1e-3 / 1e-2)."""
import pytest
import torch
import torch.nn.functional as F
from torch.nn import grad as G

pytestmark = pytest.mark.gpu

from gan_aug_pfa_b200 import ops  # noqa: E402
from test_gpu_conv import nhwc, pack_conv, pack_phase, rel  # noqa: E402
from test_gpu_thin_layers import _pack_thin, _slots  # noqa: E402

DEV = "cuda:0"
N = 64
FWD_TOL = 4e-3
WG_TOL = 2e-4


def _bf(t):
    return t.to(torch.bfloat16)


def _rand(gen, *shape):
    return _bf(torch.randn(*shape, generator=gen))


def _cmp_nhwc(out, ref_nchw, tol, what):
    """rel-L2 of a device NHWC tensor against an NCHW fp32 CPU reference, without materialising a permuted copy."""
    r = rel(out.cpu().float(), ref_nchw.permute(0, 2, 3, 1))
    assert r < tol, f"{what}: rel-L2 {r:.3e} > {tol}"


# ---- stride-2 Conv2d layers (generator down path models.py:177; discriminator models.py:230): cin, cout, input side
DOWN = [(64, 128, 128), (128, 256, 64), (256, 512, 32), (512, 512, 16), (512, 512, 8), (512, 512, 4)]


@pytest.mark.parametrize("cin,cout,h", DOWN)
def test_conv_k4s2_fwd_dgrad_wgrad_batch64(cin, cout, h):
    g = torch.Generator().manual_seed(cin + cout + h)
    x = _rand(g, N, cin, h, h)
    w = _bf(torch.randn(cout, cin, 4, 4, generator=g) / (16 * cin) ** 0.5)
    dy = _rand(g, N, cout, h // 2, h // 2)
    xd, dyd = nhwc(x).to(DEV), nhwc(dy).to(DEV)
    ho = h // 2
    # forward (+ BatchNorm statistics in the epilogue)
    out = torch.full((N, ho, ho, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * cout, device=DEV, dtype=torch.float64)
    ops.conv_gemm([xd], pack_conv(w).to(DEV), ops.geom_conv_fwd(4, 2, 1), out, cout, (ho, ho), stats=stats)
    ref = F.conv2d(x.float(), w.float(), None, 2, 1)
    _cmp_nhwc(out, ref, FWD_TOL, "forward")
    o = out.double()
    assert float((stats[:cout] - o.sum((0, 1, 2))).abs().max() / (o ** 2).sum((0, 1, 2)).sqrt().max()) < 1e-5
    assert rel(stats[cout:].cpu(), (o ** 2).sum((0, 1, 2)).cpu()) < 1e-5
    del out, o, ref
    # dgrad: four output-parity phases
    gx = torch.full((N, h, h, cin), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([dyd], pack_phase(w.permute(1, 0, 2, 3)).to(DEV), ops.geom_phase_k4s2p1(), gx, cin, (ho, ho))
    ref = G.conv2d_input((N, cin, h, h), w.float(), dy.float(), 2, 1)
    _cmp_nhwc(gx, ref, FWD_TOL, "dgrad")
    del gx, ref
    # wgrad into the master layout [cout][kh][kw][cin]
    dw = torch.zeros(cout, 16, cin, device=DEV)
    ops.conv_wgrad(dyd, xd, dw, (4, 4), 2, (-1, -1), 16 * cin, cin)
    ref = G.conv2d_weight(x.float(), (cout, cin, 4, 4), dy.float(), 2, 1)
    r = rel(dw.cpu(), ref.permute(0, 2, 3, 1).reshape(cout, 16, cin))
    assert r < WG_TOL, f"wgrad rel-L2 {r:.3e}"


def test_patchgan_conv_k4s1_fwd_dgrad_wgrad_batch64():
    """Conv2d(256 -> 512, k4, s1, p1) at 32 -> 31 (models.py:238): the largest GEMM of the step (258 GFLOP)."""
    cin, cout, h = 256, 512, 32
    g = torch.Generator().manual_seed(8)
    x = _rand(g, N, cin, h, h)
    w = _bf(torch.randn(cout, cin, 4, 4, generator=g) / (16 * cin) ** 0.5)
    dy = _rand(g, N, cout, h - 1, h - 1)
    xd, dyd = nhwc(x).to(DEV), nhwc(dy).to(DEV)
    out = torch.full((N, h - 1, h - 1, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([xd], pack_conv(w).to(DEV), ops.geom_conv_fwd(4, 1, 1), out, cout, (h - 1, h - 1))
    _cmp_nhwc(out, F.conv2d(x.float(), w.float(), None, 1, 1), FWD_TOL, "forward")
    del out
    wf = w.flip(2, 3).permute(1, 2, 3, 0).reshape(1, cin, 16 * cout).contiguous()
    gx = torch.full((N, h, h, cin), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.conv_gemm([dyd], wf.to(DEV), ops.geom_conv_dgrad_s1(4, 1), gx, cin, (h, h))
    _cmp_nhwc(gx, G.conv2d_input((N, cin, h, h), w.float(), dy.float(), 1, 1), FWD_TOL, "dgrad")
    del gx
    dw = torch.zeros(cout, 16, cin, device=DEV)
    ops.conv_wgrad(dyd, xd, dw, (4, 4), 1, (-1, -1), 16 * cin, cin)
    ref = G.conv2d_weight(x.float(), (cout, cin, 4, 4), dy.float(), 1, 1)
    r = rel(dw.cpu(), ref.permute(0, 2, 3, 1).reshape(cout, 16, cin))
    assert r < WG_TOL, f"wgrad rel-L2 {r:.3e}"


# ---- ConvTranspose2d layers of the generator up path (models.py:189,194): cin (after the skip concat), cout, input side
UP = [(512, 512, 2), (1024, 512, 4), (1024, 512, 8), (1024, 256, 16), (512, 128, 32), (256, 64, 64)]


@pytest.mark.parametrize("cin,cout,h", UP)
def test_conv_transpose_k4s2_fwd_dgrad_wgrad_batch64(cin, cout, h):
    g = torch.Generator().manual_seed(cin * 3 + cout + h)
    x = _rand(g, N, cin, h, h)
    w = _bf(torch.randn(cin, cout, 4, 4, generator=g) / (4 * cin) ** 0.5)
    dy = _rand(g, N, cout, 2 * h, 2 * h)
    xd, dyd = nhwc(x).to(DEV), nhwc(dy).to(DEV)
    # forward: four phases, BatchNorm statistics
    out = torch.full((N, 2 * h, 2 * h, cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * cout, device=DEV, dtype=torch.float64)
    ops.conv_gemm([xd], pack_phase(w.permute(1, 0, 2, 3)).to(DEV), ops.geom_phase_k4s2p1(), out, cout, (h, h),
                  stats=stats)
    _cmp_nhwc(out, F.conv_transpose2d(x.float(), w.float(), None, 2, 1), FWD_TOL, "forward")
    o = out.double()
    assert rel(stats[cout:].cpu(), (o ** 2).sum((0, 1, 2)).cpu()) < 1e-5
    del out, o
    # dgrad: a stride-2 gather over dY (= Conv2d forward with the weight read as (cin, cout) rows)
    gx = torch.full((N, h, h, cin), float("nan"), device=DEV, dtype=torch.bfloat16)
    wd = w.permute(0, 2, 3, 1).reshape(1, cin, 16 * cout).contiguous()
    ops.conv_gemm([dyd], wd.to(DEV), ops.geom_conv_fwd(4, 2, 1), gx, cin, (h, h))
    _cmp_nhwc(gx, F.conv2d(dy.float(), w.float(), None, 2, 1), FWD_TOL, "dgrad")
    del gx
    # wgrad into [cin][kh][kw][cout]; conv_transpose2d's weight gradient = conv2d_weight with x and dy swapped
    dw = torch.zeros(cin, 16, cout, device=DEV)
    ops.conv_wgrad(xd, dyd, dw, (4, 4), 2, (-1, -1), 16 * cout, cout)
    ref = G.conv2d_weight(dy.float(), (cin, cout, 4, 4), x.float(), 2, 1)
    r = rel(dw.cpu(), ref.permute(0, 2, 3, 1).reshape(cin, 16, cout))
    assert r < WG_TOL, f"wgrad rel-L2 {r:.3e}"


# ---- thin layers -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cw,groups,bias,two_out", [
    (64, 1, False, True),       # generator first conv 3 -> 64 (models.py:177): LeakyReLU + ReLU outputs
    (64, 2, True, False),       # discriminator first conv 6 -> 64 + bias (models.py:223)
    (128, 1, False, False),     # generator last ConvTranspose2d seen from its dgrad (d(pre-tanh) -> gR[0])
])
def test_thin_conv_fwd_and_wgrad_batch64(cw, groups, bias, two_out):
    h = 256
    g = torch.Generator().manual_seed(cw + groups)
    xs = [_rand(g, N, 3, h, h).float() for _ in range(groups)]
    wt = _bf(torch.randn(cw, 3 * groups, 4, 4, generator=g) / (48 * groups) ** 0.5).float()
    b = torch.randn(cw, generator=g) if bias else None
    srcs = [_slots(x).to(DEV) for x in xs]
    out1 = torch.full((N, h // 2, h // 2, cw), float("nan"), device=DEV, dtype=torch.bfloat16)
    out2 = torch.full_like(out1, float("nan")) if two_out else None
    ops.thin_conv_fwd(srcs[0], srcs[1] if groups == 2 else None, _pack_thin(wt, groups).to(DEV),
                      b.to(DEV) if bias else None, out1, ops.ACT_LRELU, out2, ops.ACT_RELU)
    xc = torch.cat(xs, 1)
    ref = F.conv2d(xc, wt, b, stride=2, padding=1)
    _cmp_nhwc(out1, F.leaky_relu(ref, 0.2), FWD_TOL, "forward LeakyReLU")
    if two_out:
        _cmp_nhwc(out2, F.relu(ref), FWD_TOL, "forward ReLU")
    del out1, out2, ref
    dy = _rand(g, N, cw, h // 2, h // 2)
    krow = 64 if groups == 1 else 128
    dw = torch.zeros(cw, krow, device=DEV)
    db = torch.zeros(cw, device=DEV) if bias else None
    ops.thin_conv_wgrad(nhwc(dy).to(DEV), srcs[0], srcs[1] if groups == 2 else None, dw.view(-1), krow, db)
    c = 3 * groups
    ref = G.conv2d_weight(xc, (cw, c, 4, 4), dy.float(), 2, 1).permute(0, 2, 3, 1).reshape(cw, 16 * c)
    assert rel(dw.cpu()[:, :16 * c], ref) < 4e-4
    assert float(dw[:, 16 * c:].abs().max()) == 0.0
    if bias:
        assert rel(db.cpu(), dy.float().sum((0, 2, 3))) < 4e-4


@pytest.mark.parametrize("cw,bias,act", [(128, True, "tanh"), (64, False, "none")])
def test_thin_convT_fwd_batch64(cw, bias, act):
    """Generator last layer ConvTranspose2d(128 -> 3) + bias + Tanh (models.py:184,186), and the discriminator first
    conv's input gradient (64 -> 3 slots, no activation)."""
    ih = 128
    g = torch.Generator().manual_seed(cw)
    x = _rand(g, N, cw, ih, ih)
    wt = _bf(torch.randn(cw, 3, 4, 4, generator=g) / (4 * cw) ** 0.5).float()
    b = torch.randn(3, generator=g) if bias else None
    ref = F.conv_transpose2d(x.float(), wt, b, stride=2, padding=1)
    if act == "tanh":
        ref = torch.tanh(ref)
    w2 = _bf(wt.permute(2, 3, 1, 0).reshape(48, cw)).contiguous()
    obf = torch.full((N, 2 * ih, 2 * ih, 4), float("nan"), device=DEV, dtype=torch.bfloat16)
    o32 = torch.full((N, 2 * ih, 2 * ih, 4), float("nan"), device=DEV)
    ops.thin_convT_fwd(nhwc(x).to(DEV), w2.to(DEV), b.to(DEV) if bias else None,
                       ops.ACT_TANH if act == "tanh" else ops.ACT_NONE, obf, o32)
    assert rel(o32.cpu()[..., :3], ref.permute(0, 2, 3, 1)) < 1e-4
    assert rel(obf.cpu().float()[..., :3], ref.permute(0, 2, 3, 1)) < FWD_TOL
    assert float(o32[..., 3].abs().max()) == 0.0


def test_cout1_conv_fwd_dgrad_wgrad_batch64():
    """PatchGAN head Conv2d(512 -> 1, k4, s1, p1) + bias at 31 -> 30 (models.py:243)."""
    c, h = 512, 31
    g = torch.Generator().manual_seed(30)
    x = _rand(g, N, c, h, h)
    w = _bf(torch.randn(1, c, 4, 4, generator=g) / (16 * c) ** 0.5)
    b = torch.randn(1, generator=g)
    xd = nhwc(x).to(DEV)
    wd = w.permute(0, 2, 3, 1).reshape(-1).contiguous().to(DEV)
    z = torch.empty(N * h * h, 16, device=DEV)
    logits = torch.full((N, h - 1, h - 1), float("nan"), device=DEV)
    ops.cout1_conv_fwd(xd, wd, b.to(DEV), z, logits)
    assert rel(logits.cpu(), F.conv2d(x.float(), w.float(), b, 1, 1)[:, 0]) < 2e-3
    dl = torch.randn(N, h - 1, h - 1, generator=g) * 0.01
    gx = torch.full((N, h, h, c), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.cout1_conv_dgrad(dl.to(DEV), wd, gx)
    _cmp_nhwc(gx, G.conv2d_input((N, c, h, h), w.float(), dl.unsqueeze(1), 1, 1), 6e-3, "dgrad")
    dw = torch.zeros(16 * c, device=DEV)
    ops.cout1_conv_wgrad(dl.to(DEV), xd, dw)
    ref = G.conv2d_weight(x.float(), (1, c, 4, 4), dl.unsqueeze(1), 1, 1)
    assert rel(dw.cpu().view(4, 4, c), ref[0].permute(1, 2, 0)) < 6e-3


# ---- the whole iteration -----------------------------------------------------------------------------------------
def test_full_batch64_train_step_against_oracle():
    """One Pix2PixTrainer.train_step at batch 64, 256x256 against the CPU oracle's train_gan_one_epoch iteration
    (oracle/pix2pix_oracle.py::gan_train_step, pinned to the reference in tests/test_oracle_vs_reference.py):
    both losses, every parameter gradient, BatchNorm running statistics, and the updated parameters."""
    from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer
    from oracle import pix2pix_oracle as O
    from test_gpu_engine import _cpu_sd, cos
    dev = torch.device(DEV)
    torch.manual_seed(0)
    tr = Pix2PixTrainer(dev)
    sd_g, sd_d = _cpu_sd(tr.G), _cpu_sd(tr.D)
    og = O.AdamState(sd_g, O.param_names(sd_g), 1e-4, (0.5, 0.999))
    od = O.AdamState(sd_d, O.param_names(sd_d), 1e-4, (0.5, 0.999))
    gen = torch.Generator().manual_seed(1234)
    A = torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1
    B = torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1
    ld, lg, aux = O.gan_train_step(sd_g, sd_d, og, od, A, B, return_grads=True)
    losses = tr.train_step(A.to(dev), B.to(dev)).cpu()
    assert abs(float(losses[0]) - ld) < 2e-3 * max(1, abs(ld)), (float(losses[0]), ld)
    assert abs(float(losses[1]) - lg) < 2e-3 * max(1, abs(lg)), (float(losses[1]), lg)
    fake = tr.G.output_nchw().cpu()
    assert rel(fake, aux["fake_B"]) < 2e-2
    # per-tensor gradient cosine >= 0.97, or — where torch's OWN bf16 arithmetic does worse than that on this very
    # fixture (tests/golden/gan_yardstick_b64.json, made by make_gan_yardstick.py: the oracle under CPU bf16 autocast
    # reaches only 0.9645 on the depth-4 down-norm bias and 0.926 on a discriminator norm bias) — within 1.5 x the
    # yardstick's error:  1 - cos <= max(0.03, 2.25 x (1 - yardstick cos))
    import json
    from pathlib import Path
    yard = json.loads((Path(__file__).resolve().parent / "golden" / "gan_yardstick_b64.json").read_text())
    assert abs(yard["loss_d"] - ld) < 1e-4 * abs(ld) and abs(yard["loss_g"] - lg) < 1e-4 * abs(lg)   # same fixture
    bad = []
    for net, grads, yc in ((tr.D, aux["grads_d"], yard["cos_d"]), (tr.G, aux["grads_g"], yard["cos_g"])):
        for k, gref in grads.items():
            c = cos(net.grad(k).cpu(), gref)
            if 1 - c > max(0.03, 2.25 * (1 - yc[k])):
                bad.append((k, round(c, 4), round(yc[k], 4)))
    assert not bad, f"gradient cosines outside the bf16 band (tensor, ours, torch-bf16 yardstick): {bad}"
    assert rel(tr.G.grad("model.model.3.weight").cpu(), aux["grads_g"]["model.model.3.weight"]) < 2e-2
    assert rel(tr.D.grad("model.11.weight").cpu(), aux["grads_d"]["model.11.weight"]) < 2e-2
    for net, sd, inc in ((tr.G, sd_g, 2), (tr.D, sd_d, 3)):
        for k, v in net.state_dict().items():
            if k.endswith("num_batches_tracked"):
                assert int(v) == int(sd[k]) == inc
            elif "running" in k:
                assert rel(v.cpu(), sd[k]) < (1e-3 if "var" in k else 1e-2), k
    # Adam(1e-4): after one step every parameter moved by ~lr*sign(grad); the update direction must agree
    for net, sd0, sd1 in ((tr.G, _cpu_sd(tr.G), sd_g), (tr.D, _cpu_sd(tr.D), sd_d)):
        for k in O.param_names(sd1):
            assert float((sd0[k] - sd1[k]).abs().max()) < 2.5e-4, k
