"""The models.py sub-module entry points called DIRECTLY (surface that must not change):
`UnetSkipConnectionBlock.forward(x)` (models.py:204-208), `AttentionGate.forward(g, x)` (models.py:39-44) and
`SiameseUNet.forward_encoder(x)` (models.py:92-102) compute through the native kernels and plug into torch autograd.
Checker: the CPU oracle's restatement of the same functions (fp32).  Tolerances are those of the network-level tests:
bf16 activations, fp32 accumulation -> outputs rel-L2 <= 2e-2, input gradients cosine >= 0.99, parameter gradients
cosine >= 0.97 whole-gradient.  The encoder fixture (64x64, batch 2: ten conv + train-mode BatchNorm layers, the deepest
normalising over 32 values per channel) is bounded by the torch-bf16 yardstick measured on it instead (ENCODER_YARDSTICK
below): 1.5 x its per-level feature error, and 1 - cos <= 2.25 x (1 - cos_yardstick) for the gradients."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from gan_aug_pfa_b200 import models as M  # noqa: E402
from oracle import pix2pix_oracle as O  # noqa: E402

DEV = torch.device("cuda:0")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))


def _sd_with_grad(module):
    sd = {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}
    names = O.param_names(sd)
    for k in names:
        sd[k].requires_grad_(True)
    return sd, names


def _whole(grads, names):
    return torch.cat([grads[k].flatten() for k in names])


def test_inner_unet_block_called_directly():
    torch.manual_seed(3)
    blk = M.UnetSkipConnectionBlock(128, 256, submodule=M.UnetSkipConnectionBlock(
        256, 512, submodule=M.UnetSkipConnectionBlock(512, 512, innermost=True)))
    sd, names = _sd_with_grad(blk)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 128, 32, 32, generator=g)
    cot = torch.randn(4, 256, 32, 32, generator=g)
    xr = x.clone().requires_grad_(True)
    nb = {}
    ref = O._unet_block(sd, "model", xr, True, nb, False)
    gref = torch.autograd.grad((ref * cot).sum(), [xr] + [sd[k] for k in names])
    blk = blk.to(DEV).train()
    # (a non-leaf input that requires grad: the reference cannot take a leaf there either — its first op is in place)
    leaf = x.clone().to(DEV).requires_grad_(True)
    xd = leaf * 1.0
    out = blk(xd)
    assert out.shape == (4, 256, 32, 32) and out.dtype == torch.float32
    assert rel(out.detach().cpu(), ref.detach()) < 2e-2
    # the skip half is LeakyReLU(x), not x (models.py:178 + 208)
    assert rel(out.detach().cpu()[:, :128], F.leaky_relu(x, 0.2)) < 4e-3
    (out * cot.to(DEV)).sum().backward()
    assert cos(leaf.grad.cpu(), gref[0]) > 0.99 and rel(leaf.grad.cpu(), gref[0]) < 8e-2
    got = {k: p.grad.detach().cpu() for k, p in blk.named_parameters()}
    assert cos(_whole(got, names), _whole(dict(zip(names, gref[1:])), names)) > 0.97
    for k, v in nb.items():                                   # BatchNorm buffers were updated like the reference's
        mine = blk.state_dict()[k].cpu()
        assert torch.equal(mine, v) if k.endswith("num_batches_tracked") else rel(mine, v) < 1e-2, k
    # eval mode + no_grad: running statistics, and the caller's tensor is modified in place like nn.LeakyReLU(0.2, True)
    blk.eval()
    sd_e = {k: v.detach().cpu() for k, v in blk.state_dict().items()}
    x2 = x.clone().to(DEV)
    with torch.no_grad():
        out_e = blk(x2)
        ref_e = O._unet_block(sd_e, "model", x.clone(), False, None, False)
    assert rel(out_e.cpu(), ref_e) < 2e-2
    assert rel(x2.cpu(), F.leaky_relu(x, 0.2)) < 4e-3


def test_outermost_unet_block_called_directly_is_a_small_generator():
    torch.manual_seed(4)
    blk = M.UnetSkipConnectionBlock(3, 64, input_nc=3, outermost=True, submodule=M.UnetSkipConnectionBlock(
        64, 128, submodule=M.UnetSkipConnectionBlock(128, 128, innermost=True)))
    sd, names = _sd_with_grad(blk)
    g = torch.Generator().manual_seed(6)
    x = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    cot = torch.randn(2, 3, 64, 64, generator=g)
    ref = O._unet_block(sd, "model", x, True, {}, True)
    gref = dict(zip(names, torch.autograd.grad((ref * cot).sum(), [sd[k] for k in names])))
    blk = blk.to(DEV).train()
    out = blk(x.to(DEV))
    assert rel(out.detach().cpu(), ref.detach()) < 2e-2
    (out * cot.to(DEV)).sum().backward()
    got = {k: p.grad.detach().cpu() for k, p in blk.named_parameters()}
    assert cos(_whole(got, names), _whole(gref, names)) > 0.97
    assert cos(got["model.3.weight"], gref["model.3.weight"]) > 0.995


def test_attention_gate_called_directly():
    torch.manual_seed(5)
    gate = M.AttentionGate(F_g=128, F_l=64, F_int=64)
    sd, names = _sd_with_grad(gate)
    psd = {"att." + k: v for k, v in sd.items()}
    gen = torch.Generator().manual_seed(7)
    g = torch.randn(2, 128, 24, 16, generator=gen)
    x = torch.randn(2, 64, 24, 16, generator=gen)
    cot = torch.randn(2, 64, 24, 16, generator=gen)
    gr, xr = g.clone().requires_grad_(True), x.clone().requires_grad_(True)
    nb = {}
    ref = O._attention_gate(psd, "att", gr, xr, True, nb)
    gref = torch.autograd.grad((ref * cot).sum(), [gr, xr] + [sd[k] for k in names])
    gate = gate.to(DEV).train()
    gd, xd = g.clone().to(DEV).requires_grad_(True), x.clone().to(DEV).requires_grad_(True)
    out = gate(gd, xd)
    assert out.shape == (2, 64, 24, 16)
    assert rel(out.detach().cpu(), ref.detach()) < 1e-2
    (out * cot.to(DEV)).sum().backward()
    assert cos(gd.grad.cpu(), gref[0]) > 0.99 and cos(xd.grad.cpu(), gref[1]) > 0.995
    assert rel(xd.grad.cpu(), gref[1]) < 3e-2
    got = {k: p.grad.detach().cpu() for k, p in gate.named_parameters()}
    assert cos(_whole(got, names), _whole(dict(zip(names, gref[2:])), names)) > 0.97
    assert int(gate.state_dict()["psi.1.num_batches_tracked"]) == 1
    for k, v in nb.items():
        if "running" in k:
            assert rel(gate.state_dict()[k[len("att."):]].cpu(), v) < 1e-2, k


# The oracle's _siamese_encoder on this test's fixture under torch.autocast("cpu", bfloat16) vs fp32 (same seeds, same
# cotangents): rel-L2 of (conv1, conv2, conv3, conv4, bottleneck) and cosine of the whole encoder gradient.
ENCODER_YARDSTICK = {"feature_rel_l2": (0.0059, 0.0128, 0.0224, 0.0358, 0.0536), "grad_cos_whole": 0.9556,
                     "grad_cos_first_conv": 0.9474}


def test_siamese_forward_encoder_called_directly():
    torch.manual_seed(0)
    net = M.SiameseUNet(3, 1)
    sd, names = _sd_with_grad(net)
    gen = torch.Generator().manual_seed(8)
    x = torch.rand(2, 3, 64, 64, generator=gen) * 2 - 1
    nb = {}
    ref = O._siamese_encoder(sd, x, True, nb)
    cots = [torch.randn(f.shape, generator=gen) for f in ref]
    enc = [k for k in names if k.startswith(("dconv_down", "bottleneck"))]
    gref = dict(zip(enc, torch.autograd.grad(sum((f * c).sum() for f, c in zip(ref, cots)), [sd[k] for k in enc])))
    net = net.to(DEV).train()
    feats = net.forward_encoder(x.to(DEV))
    assert len(feats) == 5 and [tuple(f.shape) for f in feats] == [tuple(f.shape) for f in ref]
    for lvl, (a, b) in enumerate(zip(feats, ref)):
        assert rel(a.detach().cpu(), b.detach()) < 1.5 * ENCODER_YARDSTICK["feature_rel_l2"][lvl], lvl
    sum((f * c.to(DEV)).sum() for f, c in zip(feats, cots)).backward()
    got = {k: p.grad.detach().cpu() for k, p in net.named_parameters()}
    assert 1 - cos(_whole(got, enc), _whole(gref, enc)) < 2.25 * (1 - ENCODER_YARDSTICK["grad_cos_whole"])
    assert 1 - cos(got["dconv_down1.0.weight"], gref["dconv_down1.0.weight"]) < 2.25 * (1 - ENCODER_YARDSTICK["grad_cos_first_conv"])
    assert cos(got["bottleneck.3.weight"], gref["bottleneck.3.weight"]) > 0.97      # next to the cotangents: tight
    assert all(float(got[k].abs().max()) == 0.0 for k in names if k not in enc)      # the decoder was not involved
    assert int(net.state_dict()["dconv_down1.1.num_batches_tracked"]) == 1           # ONE encoder pass (forward() does two)
