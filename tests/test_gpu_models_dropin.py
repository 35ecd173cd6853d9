"""The drop-in ``models.py`` surface: reference constructors / state_dict layout, and the reference's own
training-loop call pattern (train_gan.py:52-74: module calls, torch losses, loss.backward(), optim.Adam)
running against the native engine through torch autograd."""
import json

import pytest
import torch
import torch.nn as nn
import torch.optim as optim

pytestmark = pytest.mark.gpu

from gan_aug_pfa_b200.models import NLayerDiscriminator, UNetGenerator  # noqa: E402
from oracle import pix2pix_oracle as O  # noqa: E402

DEV = torch.device("cuda:0")


def test_constructors_and_state_dict_match_reference_layout(golden_dir):
    gold = json.loads((golden_dir / "gan_full.json").read_text())
    torch.manual_seed(0)
    G = UNetGenerator(input_nc=3, output_nc=3)
    D = NLayerDiscriminator(input_nc=6)
    assert list(G.state_dict().keys()) == gold["keys_g"]
    assert list(D.state_dict().keys()) == gold["keys_d"]
    assert sum(p.numel() for p in G.parameters()) == gold["n_params_g"]
    with pytest.raises(RuntimeError, match="no CPU"):
        G(torch.zeros(1, 3, 256, 256))


def test_reference_loop_runs_unchanged_on_the_native_modules():
    torch.manual_seed(0)
    gen = UNetGenerator(input_nc=3, output_nc=3).to(DEV)
    disc = NLayerDiscriminator(input_nc=6).to(DEV)
    sd_g = {k: v.detach().cpu().clone() for k, v in gen.state_dict().items()}
    sd_d = {k: v.detach().cpu().clone() for k, v in disc.state_dict().items()}
    opt_g = optim.Adam(gen.parameters(), lr=1e-4, betas=(0.5, 0.999))
    opt_d = optim.Adam(disc.parameters(), lr=1e-4, betas=(0.5, 0.999))
    og = O.AdamState(sd_g, O.param_names(sd_g), 1e-4, (0.5, 0.999))
    od = O.AdamState(sd_d, O.param_names(sd_d), 1e-4, (0.5, 0.999))
    loss_GAN, loss_L1 = nn.BCEWithLogitsLoss(), nn.L1Loss()
    g = torch.Generator().manual_seed(1234)
    gen.train()
    disc.train()
    for _ in range(2):
        A = torch.rand(2, 3, 256, 256, generator=g) * 2 - 1
        B = torch.rand(2, 3, 256, 256, generator=g) * 2 - 1
        ld_ref, lg_ref, _ = O.gan_train_step(sd_g, sd_d, og, od, A, B)
        # --- the reference's loop body, verbatim in structure (train_gan.py:53-71)
        real_A, real_B = A.to(DEV), B.to(DEV)
        opt_d.zero_grad()
        fake_B = gen(real_A).detach()
        pred_real = disc(torch.cat((real_A, real_B), 1))
        loss_d_real = loss_GAN(pred_real, torch.ones_like(pred_real))
        pred_fake = disc(torch.cat((real_A, fake_B), 1))
        loss_d_fake = loss_GAN(pred_fake, torch.zeros_like(pred_fake))
        loss_d = (loss_d_real + loss_d_fake) * 0.5
        loss_d.backward()
        opt_d.step()
        opt_g.zero_grad()
        fake_B_for_g = gen(real_A)
        pred_fake_for_g = disc(torch.cat((real_A, fake_B_for_g), 1))
        loss_g = loss_GAN(pred_fake_for_g, torch.ones_like(pred_fake_for_g)) + loss_L1(fake_B_for_g, real_B) * 100.0
        loss_g.backward()
        opt_g.step()
        assert abs(loss_d.item() - ld_ref) < 3e-3 * max(1, abs(ld_ref)), (loss_d.item(), ld_ref)
        assert abs(loss_g.item() - lg_ref) < 3e-3 * max(1, abs(lg_ref)), (loss_g.item(), lg_ref)
    # BatchNorm buffers follow the reference's bookkeeping: G forward x2, D forward x3 per iteration
    assert int(gen.state_dict()["model.model.1.model.2.num_batches_tracked"]) == 4
    assert int(disc.state_dict()["model.3.num_batches_tracked"]) == 6
    rv = gen.state_dict()["model.model.1.model.2.running_var"].cpu()
    assert float((rv - sd_g["model.model.1.model.2.running_var"]).abs().max() / rv.abs().max()) < 5e-3
    # eval-mode inference (generate_synthetic_data.py:55-68)
    gen.eval()
    with torch.no_grad():
        out = gen(A.to(DEV))
        ref = O.unet_generator_forward(sd_g, A, False, None)
    assert out.shape == (2, 3, 256, 256) and out.dtype == torch.float32
    assert float((out.cpu() - ref).norm() / ref.norm()) < 5e-2
    # checkpoints round-trip in the reference's format
    gen2 = UNetGenerator(3, 3).to(DEV)
    gen2.load_state_dict(gen.state_dict())
    gen2.eval()
    with torch.no_grad():
        assert torch.equal(gen2(A.to(DEV)), out)
