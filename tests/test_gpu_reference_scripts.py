"""The reference's OWN step loops, unmodified, running on the drop-in modules.

oracle/_ref holds the reference byte-compiled by oracle/stage_ref.py (no reference source in the repo; the bytecode
travels to the GPU box).  oracle/ref_loader.py imports `train_gan` / `train` twice: once against the reference's
`models` (the checker, on the CPU) and once against `gan_aug_pfa_b200.models` under the name `models` (the product
path, on the GPU).  The real `train_gan.train_gan_one_epoch` (train_gan.py:46-75), `train.train_one_epoch` and
`train.validate` (train.py:131-164) then run on identical seeded weights and batches, and their returned losses are
compared.  Tolerances: GAN losses 3e-3 relative per iteration (bf16 activations, fp32 accumulation; measured ~5e-4);
Siamese CombinedLoss / FocalDiceLoss 3e-2 (the 31-conv network with train-mode BatchNorm on a small fixture)."""
import pytest
import torch
import torch.optim as optim

pytestmark = pytest.mark.gpu

from oracle import ref_loader  # noqa: E402

DEV = torch.device("cuda:0")
needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not staged (python oracle/stage_ref.py)")


@pytest.fixture(scope="module")
def ref():
    ns = ref_loader.load(tag="cpu")
    ns.train_gan.DEVICE = torch.device("cpu")          # configuration constant (train_gan.py:25), not code
    ns.train.DEVICE = torch.device("cpu")
    return ns


@pytest.fixture(scope="module")
def dropin():
    from gan_aug_pfa_b200 import models
    ns = ref_loader.load(models=models, tag="gpu")
    ns.train_gan.DEVICE = DEV
    ns.train.DEVICE = DEV
    return ns


@needs_ref
def test_real_train_gan_one_epoch_runs_on_the_dropin_modules(ref, dropin):
    assert dropin.train_gan.UNetGenerator.__module__.startswith("gan")      # the script really bound the drop-in classes
    assert ref.train_gan.UNetGenerator.__module__.startswith("_gapref")
    nets = {}
    for tag, ns, dev in (("ref", ref, torch.device("cpu")), ("gpu", dropin, DEV)):
        torch.manual_seed(0)
        gen = ns.train_gan.UNetGenerator(input_nc=3, output_nc=3).to(dev)          # train_gan.py:138-141
        disc = ns.train_gan.NLayerDiscriminator(input_nc=6).to(dev)
        opt_g = optim.Adam(gen.parameters(), lr=1e-4, betas=(0.5, 0.999))
        opt_d = optim.Adam(disc.parameters(), lr=1e-4, betas=(0.5, 0.999))
        nets[tag] = (ns, gen, disc, opt_g, opt_d)
    for (k1, v1), (k2, v2) in zip(nets["ref"][1].state_dict().items(), nets["gpu"][1].state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2.cpu()), k1
    g = torch.Generator().manual_seed(1234)
    seq = {"ref": [], "gpu": []}
    for _ in range(3):
        batch = {"image1": torch.rand(2, 3, 256, 256, generator=g) * 2 - 1,
                 "image2": torch.rand(2, 3, 256, 256, generator=g) * 2 - 1}
        for tag in ("ref", "gpu"):
            ns, gen, disc, opt_g, opt_d = nets[tag]
            seq[tag].append(ns.train_gan.train_gan_one_epoch(gen, disc, [batch], opt_g, opt_d))
    for (ld_r, lg_r), (ld_g, lg_g) in zip(seq["ref"], seq["gpu"]):
        assert abs(ld_g - ld_r) < 3e-3 * max(1.0, abs(ld_r)), seq
        assert abs(lg_g - lg_r) < 3e-3 * max(1.0, abs(lg_r)), seq
    # parameters after three Adam steps: every tensor moved by at most 3*lr, and in (almost) the same direction
    sd_r, sd_g = nets["ref"][1].state_dict(), nets["gpu"][1].state_dict()
    for k in sd_r:
        if k.endswith("num_batches_tracked"):
            assert int(sd_r[k]) == int(sd_g[k]) == 6
        elif "running" not in k:
            assert float((sd_r[k] - sd_g[k].cpu()).abs().max()) < 6.5e-4, k
    w0 = "model.model.3.weight"
    torch.manual_seed(0)
    init = ref.models.UNetGenerator(3, 3).state_dict()[w0]
    du_r, du_g = (sd_r[w0] - init).flatten().double(), (sd_g[w0].cpu() - init).flatten().double()
    assert float(du_r @ du_g / (du_r.norm() * du_g.norm())) > 0.9
    # save_samples' call pattern (train_gan.py:78-92): eval() + no_grad forward
    gen = nets["gpu"][1]
    gen.eval()
    with torch.no_grad():
        out = gen(batch["image1"].to(DEV))
    nets["ref"][1].eval()
    with torch.no_grad():
        want = nets["ref"][1](batch["image1"])
    assert float((out.cpu() - want).norm() / want.norm()) < 5e-2


@needs_ref
@pytest.mark.parametrize("crit", ["combined", "focal_dice"])
def test_real_train_one_epoch_and_validate_run_on_the_dropin_siamese(ref, dropin, crit):
    nets = {}
    for tag, ns, dev in (("ref", ref, torch.device("cpu")), ("gpu", dropin, DEV)):
        torch.manual_seed(0)
        model = ns.train.SiameseUNet(n_channels=3, n_classes=1).to(dev)            # train.py:293-295
        if crit == "combined":
            criterion = ns.train.CombinedLoss()
        else:
            criterion = ns.train.FocalDiceLoss(beta=0.6701, focal_gamma=1.7929, focal_alpha=0.6032, dice_smooth=1.96e-6)
        opt = optim.AdamW(model.parameters(), lr=1.0152e-4, weight_decay=1.118e-5)
        nets[tag] = (ns, model, criterion, opt, dev)
    g = torch.Generator().manual_seed(99)
    seq = {"ref": [], "gpu": []}
    val = {}
    for it in range(2):
        batch = {"image1": torch.rand(4, 3, 64, 64, generator=g) * 2 - 1,
                 "image2": torch.rand(4, 3, 64, 64, generator=g) * 2 - 1,
                 "label": (torch.rand(4, 64, 64, generator=g) < 0.05).long()}
        for tag in ("ref", "gpu"):
            ns, model, criterion, opt, dev = nets[tag]
            seq[tag].append(ns.train.train_one_epoch(model, [batch, None], opt, criterion, dev) * 2)   # None batches are skipped
    for tag in ("ref", "gpu"):
        ns, model, criterion, opt, dev = nets[tag]
        val[tag] = ns.train.validate(model, [batch], criterion, dev)
    for a, b in zip(seq["ref"], seq["gpu"]):
        assert abs(a - b) < 3e-2 * abs(a), seq
    assert abs(val["ref"] - val["gpu"]) < 5e-2 * abs(val["ref"]), val
