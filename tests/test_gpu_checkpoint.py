"""Checkpoint / resume: a run that is saved after two iterations and resumed in a NEW trainer
continues like the uninterrupted run (same losses to fp32-atomic noise: Adam moments, step counters and BatchNorm
buffers all came back), and the model part of the checkpoint is a reference-format state_dict."""
import io

import pytest
import torch

pytestmark = pytest.mark.gpu

from gan_aug_pfa_b200 import checkpoint  # noqa: E402

DEV = torch.device("cuda:0")


def _batches(n, seed, hw=64):
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(2, 3, hw, hw, generator=g) * 2 - 1, torch.rand(2, 3, hw, hw, generator=g) * 2 - 1) for _ in range(n)]


def test_pix2pix_trainer_resumes_bit_faithfully(golden_dir):
    from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer
    data = _batches(5, 3)
    torch.manual_seed(0)
    a = Pix2PixTrainer(DEV, num_downs=5)
    for x, y in data[:2]:
        a.train_step(x.to(DEV), y.to(DEV))
    buf = io.BytesIO()
    checkpoint.save(buf, checkpoint.trainer_state(a, extra={"epoch": 7}))
    want = [a.train_step(x.to(DEV), y.to(DEV)).cpu() for x, y in data[2:]]
    torch.manual_seed(123)                               # a different seed: everything must come from the checkpoint
    b = Pix2PixTrainer(DEV, num_downs=5)
    buf.seek(0)
    state = checkpoint.load(buf)
    assert checkpoint.load_trainer_state(b, state) == {"epoch": 7}
    assert int(b.G.store.step) == 2 and int(b.D.store.step) == 2
    got = [b.train_step(x.to(DEV), y.to(DEV)).cpu() for x, y in data[2:]]
    for w, g in zip(want, got):
        assert torch.allclose(w, g, rtol=2e-3, atol=2e-4), (want, got)
    # without the optimizer state the continuation differs (Adam's bias correction restarts): the check above is not vacuous
    torch.manual_seed(123)
    c = Pix2PixTrainer(DEV, num_downs=5)
    c.G.load_state_dict(state["G"]["state_dict"])
    c.D.load_state_dict(state["D"]["state_dict"])
    cold = [c.train_step(x.to(DEV), y.to(DEV)).cpu() for x, y in data[2:]]
    assert not torch.allclose(want[-1], cold[-1], rtol=2e-3, atol=2e-4)
    # the model part is what the reference's loaders expect
    ref_sd = checkpoint.export_reference_state_dicts(state)
    assert set(ref_sd) == {"generator", "discriminator"}
    assert list(ref_sd["discriminator"].keys())[:2] == ["model.0.weight", "model.0.bias"]
    assert ref_sd["generator"]["model.model.0.weight"].shape == (64, 3, 4, 4)
    with pytest.raises(ValueError):
        checkpoint.load_trainer_state(Pix2PixTrainer(DEV, num_downs=6), state)


def test_siamese_engine_resumes():
    from gan_aug_pfa_b200 import models as M
    from gan_aug_pfa_b200.siamese import SiameseEngine
    g = torch.Generator().manual_seed(4)
    data = [((torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).to(DEV), (torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).to(DEV),
             (torch.rand(2, 32, 32, generator=g) < 0.1).long().to(DEV)) for _ in range(4)]
    torch.manual_seed(0)
    a = SiameseEngine(DEV)
    a.load_state_dict({k: v.detach() for k, v in M.SiameseUNet(3, 1).state_dict().items()})
    for b in data[:2]:
        a.train_step(*b)
    state = checkpoint.siamese_state(a)
    want = [float(a.train_step(*b).cpu()) for b in data[2:]]
    r = SiameseEngine(DEV)
    checkpoint.load_siamese_state(r, state)
    got = [float(r.train_step(*b).cpu()) for b in data[2:]]
    assert got == pytest.approx(want, rel=5e-3)
