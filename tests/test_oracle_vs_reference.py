"""Pins the oracle against the UNMODIFIED reference imported from /root/reference (build container
only; skipped where the reference tree is absent, e.g. on the GPU box)."""
import importlib.util
import os
import sys
import types
from pathlib import Path

import pytest
import torch

from oracle import pix2pix_oracle as O

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "models.py").exists(), reason="reference tree not present")


def _load(name, stubs=()):
    for s in stubs:
        sys.modules.setdefault(s, types.ModuleType(s))
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    spec = importlib.util.spec_from_file_location("ref_" + name, REF / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref_models():
    return _load("models")


def test_generator_discriminator_train_and_eval(ref_models):
    torch.manual_seed(1)
    G = ref_models.UNetGenerator(3, 3, num_downs=6, ngf=4)
    D = ref_models.NLayerDiscriminator(6, ndf=4)
    x = torch.randn(3, 3, 64, 64)
    y = torch.randn(3, 3, 64, 64)
    for train in (True, False):
        G.train(train)
        D.train(train)
        sd_g = O.clone_state_dict(G.state_dict())
        sd_d = O.clone_state_dict(D.state_dict())
        nb_g, nb_d = {}, {}
        with torch.no_grad():
            ref_f = G(x)
            ref_p = D(torch.cat((x, y), 1))
            got_f = O.unet_generator_forward(sd_g, x, train, nb_g)
            got_p = O.discriminator_forward(sd_d, torch.cat((x, y), 1), train, nb_d)
        assert torch.allclose(got_f, ref_f, atol=1e-6)
        assert torch.allclose(got_p, ref_p, atol=1e-6)
        if train:
            for k, v in nb_g.items():
                assert torch.allclose(v.float(), G.state_dict()[k].float(), atol=1e-6), k
            for k, v in nb_d.items():
                assert torch.allclose(v.float(), D.state_dict()[k].float(), atol=1e-6), k


def test_skip_branch_carries_leaky_relu_of_input(ref_models):
    """The in-place LeakyReLU quirk (models.py:178 + 208): the skip half of the block output equals
    LeakyReLU(x), both in the reference and in the oracle."""
    torch.manual_seed(2)
    blk = ref_models.UnetSkipConnectionBlock(8, 8, innermost=True)
    x = torch.randn(2, 8, 4, 4)
    out = blk(x.clone())
    assert torch.allclose(out[:, :8], torch.nn.functional.leaky_relu(x, 0.2))


def test_train_gan_one_epoch_three_steps():
    real_makedirs = os.makedirs
    os.makedirs = lambda *a, **k: None
    try:
        _load("dataset")
        sys.modules["dataset"] = sys.modules.get("dataset") or _load("dataset")
        sys.modules["models"] = _load("models")
        tg = _load("train_gan")
    finally:
        os.makedirs = real_makedirs

    class NoBar:
        def __init__(self, it, **k):
            self.it = it

        def __iter__(self):
            return iter(self.it)

        def set_postfix(self, **k):
            pass

    tg.tqdm = NoBar
    torch.manual_seed(0)
    G = sys.modules["models"].UNetGenerator(3, 3, num_downs=5, ngf=8)
    D = sys.modules["models"].NLayerDiscriminator(6, ndf=8)
    sd_g, sd_d = O.clone_state_dict(G.state_dict()), O.clone_state_dict(D.state_dict())
    opt_g = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.5, 0.999))
    opt_d = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.5, 0.999))
    og = O.AdamState(sd_g, O.param_names(sd_g), 1e-4, (0.5, 0.999))
    od = O.AdamState(sd_d, O.param_names(sd_d), 1e-4, (0.5, 0.999))
    gen = torch.Generator().manual_seed(9)
    for _ in range(3):
        A = torch.rand(2, 3, 32, 32, generator=gen) * 2 - 1
        B = torch.rand(2, 3, 32, 32, generator=gen) * 2 - 1
        ld_ref, lg_ref = tg.train_gan_one_epoch(G, D, [{"image1": A, "image2": B}], opt_g, opt_d)
        ld, lg, _ = O.gan_train_step(sd_g, sd_d, og, od, A, B)
        assert abs(ld - ld_ref) < 1e-5 and abs(lg - lg_ref) < 1e-4
    for k, v in G.state_dict().items():
        assert torch.allclose(sd_g[k].detach().float(), v.float(), atol=1e-5), k
    for k, v in D.state_dict().items():
        assert torch.allclose(sd_d[k].detach().float(), v.float(), atol=1e-5), k


def test_siamese_forward_and_losses(ref_models):
    tr = _load("train", stubs=("optuna",))
    torch.manual_seed(3)
    S = ref_models.SiameseUNet(3, 1)
    x1, x2 = torch.randn(1, 3, 32, 32), torch.randn(1, 3, 32, 32)
    lab = (torch.rand(1, 32, 32) < 0.1).long()
    S.train()
    sd = O.clone_state_dict(S.state_dict())
    nb = {}
    with torch.no_grad():
        ref = S(x1, x2)
        got = O.siamese_forward(sd, x1, x2, True, nb)
    assert torch.allclose(got, ref, atol=2e-5)
    for k, v in nb.items():
        assert torch.allclose(v.float(), S.state_dict()[k].float(), atol=1e-5), k
    assert abs(float(tr.CombinedLoss()(ref, lab)) - float(O.combined_loss(ref, lab))) < 1e-6
    fd = tr.FocalDiceLoss(beta=0.67, focal_gamma=1.79, focal_alpha=0.6, dice_smooth=1.96e-6)
    assert abs(float(fd(ref, lab)) - float(O.focal_dice_loss(ref, lab, 0.67, 1.79, 0.6, 1.96e-6))) < 1e-6


def test_calculate_metrics_restatement_matches_the_reference_function():
    """O.calculate_metrics == evaluate.py's own calculate_metrics.  Only that function is executed (extracted from the
    file's AST): importing evaluate.py would create directories under /Users/... and needs matplotlib."""
    import ast
    tree = ast.parse((REF / "evaluate.py").read_text())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "calculate_metrics")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), str(REF / "evaluate.py"), "exec"), ns)
    g = torch.Generator().manual_seed(0)
    for p_pos, p_pred in ((0.05, 0.07), (0.5, 0.5), (0.0, 0.2), (0.3, 0.0), (1.0, 1.0)):
        targets = (torch.rand(1, 48, 48, generator=g) < p_pos).float()
        probs = torch.sigmoid(torch.randn(1, 1, 48, 48, generator=g) + (2.0 if p_pred >= 1.0 else -2.0 + 4 * p_pred))
        ref = ns["calculate_metrics"](probs, targets)
        got, _ = O.calculate_metrics(probs, targets)
        assert set(got) == set(ref)
        for k in ref:
            assert got[k] == pytest.approx(ref[k], rel=1e-7, abs=1e-12), k


def test_uint8_normalisation_matches_the_reference_transforms():
    """The device-side input pipeline (gap_u8_hwc_to_nhwc_bf16, gap_gen_out_bwd_u8) assumes
    JointNormalize(JointToTensor(img)) == (uint8 / 255) * 2 - 1 in fp32, HWC -> CHW: check that against the reference's
    own transform classes on a PIL image (dataset.py:21-36,155-159)."""
    import numpy as np
    from PIL import Image
    ds = _load("dataset")
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (24, 40, 3), generator=g, dtype=torch.uint8)
    img = Image.fromarray(u8.numpy(), mode="RGB")
    sample = {"image1": img, "image2": img, "label": None}
    sample = ds.JointNormalize()(ds.JointToTensor()(sample))
    mine = ((u8.float() / 255.0) * 2.0 - 1.0).permute(2, 0, 1)
    assert torch.equal(sample["image1"], mine) and torch.equal(sample["image2"], mine)


def test_attention_gate_conv_biases_have_zero_gradient_in_the_reference(ref_models):
    """AttentionGate's W_g / W_x / psi convs carry a bias and feed a training-mode BatchNorm (models.py:21-34), which
    subtracts the batch mean: d(loss)/d(bias) is identically zero and autograd returns rounding noise.  The native
    engine therefore leaves the W_g / W_x bias gradients at zero instead of summing dy (siamese.py _conv1x1_bn)."""
    torch.manual_seed(3)
    gate = ref_models.AttentionGate(F_g=16, F_l=16, F_int=8).train()
    g, x = torch.randn(2, 16, 12, 12), torch.randn(2, 16, 12, 12)
    (gate(g, x) * torch.randn(2, 16, 12, 12)).sum().backward()
    for conv in (gate.W_g[0], gate.W_x[0], gate.psi[0]):
        assert float(conv.weight.grad.abs().max()) > 1e-3
        assert float(conv.bias.grad.abs().max()) < 1e-5 * float(conv.weight.grad.abs().max())
