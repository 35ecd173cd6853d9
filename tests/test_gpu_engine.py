"""End-to-end parity of the native Pix2Pix engine (every layer through the C ABI) with the CPU oracle
and with golden values recorded from the reference itself.

Tolerances (bf16 storage of activations / activation gradients, fp32 accumulation):
  forward outputs          rel-L2 <= 2e-2  (measured ~8e-3; torch's own bf16 autocast: 0.4-1.2 % per layer)
  parameter gradients      cosine >= 0.97 per tensor (torch-bf16 yardstick reaches 0.972 at the innermost block)
  losses                   |d| <= 2e-3 * max(1, |loss|)
  BatchNorm running stats  rel <= 1e-3, num_batches_tracked exact (G: +2, D: +3 per iteration)
"""
import json

import pytest
import torch

pytestmark = pytest.mark.gpu

from gan_aug_pfa_b200 import ops  # noqa: E402
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402
from oracle import pix2pix_oracle as O  # noqa: E402

DEV = torch.device("cuda:0")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))


@pytest.fixture(scope="module")
def trainer():
    torch.manual_seed(0)
    return Pix2PixTrainer(DEV)


def _cpu_sd(net):
    return {k: v.detach().cpu().clone().contiguous() for k, v in net.state_dict().items()}


def test_state_dict_layout_and_seeded_weights(trainer, golden_dir):
    gold = json.loads((golden_dir / "gan_full.json").read_text())
    sd_g, sd_d = trainer.G.state_dict(), trainer.D.state_dict()
    assert list(sd_g.keys()) == gold["keys_g"] and list(sd_d.keys()) == gold["keys_d"]
    import hashlib

    def sd_hash(sd):
        h = hashlib.sha256()
        for k, v in sd.items():
            h.update(k.encode())
            h.update(str(tuple(v.shape)).encode())
            h.update(str(v.dtype).encode())
            h.update(v.detach().cpu().contiguous().numpy().tobytes())
        return h.hexdigest()

    assert sd_hash(sd_g) == gold["sd_g_sha256"] and sd_hash(sd_d) == gold["sd_d_sha256"]
    # round trip through load_state_dict
    trainer.G.load_state_dict(_cpu_sd(trainer.G))
    assert sd_hash(trainer.G.state_dict()) == gold["sd_g_sha256"]
    with pytest.raises(KeyError):
        trainer.D.load_state_dict({"model.0.weight": torch.zeros(64, 6, 4, 4)})


def test_forward_train_and_eval_against_oracle(trainer):
    sd_g, sd_d = _cpu_sd(trainer.G), _cpu_sd(trainer.D)
    gen = torch.Generator().manual_seed(42)
    A = torch.rand(2, 3, 256, 256, generator=gen) * 2 - 1
    B = torch.rand(2, 3, 256, 256, generator=gen) * 2 - 1
    bn = torch.zeros(2, 256, 256, 4, device=DEV, dtype=torch.bfloat16)
    ops.nchw_to_nhwc_bf16(B.to(DEV), bn)
    for train in (True, False):
        if not train:
            # eval mode normalises with the running statistics: make them meaningful first (20 momentum
            # updates on the oracle side), then load the same buffers into the engine
            ev_g, ev_d = O.clone_state_dict(sd_g), O.clone_state_dict(sd_d)
            with torch.no_grad():
                for _ in range(20):
                    nb_g, nb_d = {}, {}
                    f = O.unet_generator_forward(ev_g, A, True, nb_g)
                    O.discriminator_forward(ev_d, torch.cat((A, f), 1), True, nb_d)
                    ev_g.update(nb_g)
                    ev_d.update(nb_d)
            trainer.G.load_state_dict(ev_g)
            trainer.D.load_state_dict(ev_d)
        else:
            ev_g, ev_d = sd_g, sd_d
        trainer.G.training = trainer.D.training = train
        with torch.no_grad():
            ref_f = O.unet_generator_forward(ev_g, A, train, None)
            ref_p = O.discriminator_forward(ev_d, torch.cat((A, B), 1), train, None)
        trainer.G.forward(A.to(DEV))
        fake = trainer.G.output_nchw().cpu()
        assert rel(fake, ref_f) < 2e-2, f"train={train}"
        logits = trainer.D.forward(trainer.G.x_nhwc, bn).permute(0, 3, 1, 2).cpu()
        assert logits.shape == ref_p.shape == (2, 1, 30, 30)
        assert rel(logits, ref_p) < 2e-2, f"train={train}"
    trainer.G.load_state_dict(sd_g)
    trainer.D.load_state_dict(sd_d)
    trainer.G.training = trainer.D.training = True


def test_train_step_grads_losses_buffers_against_oracle(trainer):
    sd_g, sd_d = _cpu_sd(trainer.G), _cpu_sd(trainer.D)
    og = O.AdamState(sd_g, O.param_names(sd_g), 1e-4, (0.5, 0.999))
    od = O.AdamState(sd_d, O.param_names(sd_d), 1e-4, (0.5, 0.999))
    gen = torch.Generator().manual_seed(1234)
    A = torch.rand(2, 3, 256, 256, generator=gen) * 2 - 1
    B = torch.rand(2, 3, 256, 256, generator=gen) * 2 - 1
    ld, lg, aux = O.gan_train_step(sd_g, sd_d, og, od, A, B, return_grads=True)
    losses = trainer.train_step(A.to(DEV), B.to(DEV)).cpu()
    assert abs(float(losses[0]) - ld) < 2e-3 * max(1, abs(ld))
    assert abs(float(losses[1]) - lg) < 2e-3 * max(1, abs(lg))
    worst = 1.0
    for k, gref in aux["grads_d"].items():
        worst = min(worst, cos(trainer.D.grad(k).cpu(), gref))
    for k, gref in aux["grads_g"].items():
        worst = min(worst, cos(trainer.G.grad(k).cpu(), gref))
    assert worst >= 0.97, f"worst per-tensor gradient cosine {worst}"
    # the outermost layers see the least accumulated bf16 noise: tight check there
    assert rel(trainer.G.grad("model.model.3.weight").cpu(), aux["grads_g"]["model.model.3.weight"]) < 2e-2
    assert rel(trainer.D.grad("model.11.weight").cpu(), aux["grads_d"]["model.11.weight"]) < 2e-2
    for net, sd, inc in ((trainer.G, sd_g, 2), (trainer.D, sd_d, 3)):
        for k, v in net.state_dict().items():
            if k.endswith("num_batches_tracked"):
                assert int(v) == int(sd[k]) == inc
            elif "running" in k:
                assert rel(v.cpu(), sd[k]) < (1e-3 if "var" in k else 1e-2), k


def test_three_step_loss_sequence_against_reference_golden(golden_dir):
    """Batch 1, 256x256, seed 0: the loss sequence recorded from the reference's own
    train_gan_one_epoch (tests/golden/gan_full.json)."""
    gold = json.loads((golden_dir / "gan_full.json").read_text())
    torch.manual_seed(0)
    tr = Pix2PixTrainer(DEV)
    gen = torch.Generator().manual_seed(1234)
    for ld_ref, lg_ref in gold["loss_sequence"]:
        A = torch.rand(1, 3, 256, 256, generator=gen) * 2 - 1
        B = torch.rand(1, 3, 256, 256, generator=gen) * 2 - 1
        losses = tr.train_step(A.to(DEV), B.to(DEV)).cpu()
        assert abs(float(losses[0]) - ld_ref) < 5e-3, (float(losses[0]), ld_ref)
        assert abs(float(losses[1]) - lg_ref) < 5e-3 * lg_ref, (float(losses[1]), lg_ref)
    nbt = [int(v) for k, v in tr.G.state_dict().items() if k.endswith("num_batches_tracked")]
    assert set(nbt) == {6}


def test_full_batch_step_is_finite_and_reproducible_in_loss():
    """BASELINE size (batch 64): two trainers from the same seed give the same losses to fp32-atomic
    noise, and every parameter stays finite."""
    vals = []
    for _ in range(2):
        torch.manual_seed(0)
        tr = Pix2PixTrainer(DEV)
        gen = torch.Generator().manual_seed(7)
        A = (torch.rand(64, 3, 256, 256, generator=gen) * 2 - 1).to(DEV)
        B = (torch.rand(64, 3, 256, 256, generator=gen) * 2 - 1).to(DEV)
        out = None
        for _ in range(2):
            out = tr.train_step(A, B)
        vals.append(out.cpu())
        assert torch.isfinite(tr.G.store.p).all() and torch.isfinite(tr.D.store.p).all()
        del tr
        torch.cuda.empty_cache()
    assert torch.allclose(vals[0], vals[1], rtol=1e-3)


def test_graphed_train_step_matches_eager():
    """Pix2PixTrainer.train_step_graphed (CUDA-graph replay, device-side Adam step counter) == train_step."""
    from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(9)
    batches = [(torch.rand(2, 3, 64, 64, generator=gen) * 2 - 1, torch.rand(2, 3, 64, 64, generator=gen) * 2 - 1)
               for _ in range(4)]
    res = []
    for graphed in (False, True):
        torch.manual_seed(0)
        tr = Pix2PixTrainer(dev, num_downs=5)
        seq = []
        for a, b in batches:
            fn = tr.train_step_graphed if graphed else tr.train_step
            seq.append(fn(a.to(dev), b.to(dev)).cpu().clone())
        res.append(torch.stack(seq))
        if graphed:
            assert int(tr.G.store.step_dev) == len(batches)
    # fp32 atomics make the two runs differ in the last bits only
    assert torch.allclose(res[0], res[1], rtol=2e-3, atol=2e-4), (res[0], res[1])


def test_train_step_accepts_raw_uint8_images():
    """Device-side input pipeline: uint8 [n,h,w,3] pairs give the same iteration as their
    ToTensor + JointNormalize fp32 NCHW form (dataset.py:28-29,155-159), eagerly and through the CUDA graph."""
    from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(12)
    a8 = torch.randint(0, 256, (2, 64, 64, 3), generator=gen, dtype=torch.uint8)
    b8 = torch.randint(0, 256, (2, 64, 64, 3), generator=gen, dtype=torch.uint8)
    to_f32 = lambda u: ((u.float() / 255.0) * 2.0 - 1.0).permute(0, 3, 1, 2).contiguous()
    res = []
    for mode in ("f32", "u8", "u8_graph"):
        torch.manual_seed(0)
        tr = Pix2PixTrainer(dev, num_downs=5)
        fn = tr.train_step_graphed if mode == "u8_graph" else tr.train_step
        args = (to_f32(a8).to(dev), to_f32(b8).to(dev)) if mode == "f32" else (a8.to(dev), b8.to(dev))
        res.append(torch.stack([fn(*args).cpu().clone() for _ in range(3)]))
    assert torch.allclose(res[0], res[1], rtol=2e-3, atol=2e-4), (res[0], res[1])
    assert torch.allclose(res[1], res[2], rtol=2e-3, atol=2e-4), (res[1], res[2])
    with pytest.raises(ValueError):
        tr.train_step(a8.to(dev), to_f32(b8).to(dev))


def test_dropout_kernel_mask_is_regenerable_and_fair():
    """gap_dropout_bf16 (nn.Dropout(0.5), models.py:197-198): x -> {0, 2x}, same (seed, offset) -> same mask (that is how
    the backward pass re-applies it to the gradient), keep probability 0.5."""
    x = (torch.rand(4, 8, 8, 512, device=DEV) + 0.5).to(torch.bfloat16)
    a, b, c = x.clone(), x.clone(), x.clone()
    ops.dropout_(a, 0.5, 7, 1 << 40)
    ops.dropout_(b, 0.5, 7, 1 << 40)
    ops.dropout_(c, 0.5, 7, 2 << 40)
    assert torch.equal(a, b) and not torch.equal(a, c)
    kept = a != 0
    assert abs(float(kept.float().mean()) - 0.5) < 0.01
    assert torch.equal(a[kept].float(), (2 * x[kept].float()).to(torch.bfloat16).float())
    half = x[..., 256:].clone()
    wide = x.clone()
    ops.dropout_(wide[..., 256:], 0.5, 3, 0)          # channel slice of a wider buffer
    assert torch.equal(wide[..., :256], x[..., :256]) and not torch.equal(wide[..., 256:], half)


def test_generator_with_dropout_trains_and_evaluates():
    """UNetGenerator(use_dropout=True): the two ngf*8 blocks end in Dropout(0.5) (models.py:156-157).  Mask parity with
    torch's RNG is not possible, so: eval mode equals the no-dropout network, training forwards differ between calls
    (fresh masks), and the full GAN iteration (second generator forward NOT elided) runs with finite losses."""
    gen = torch.Generator().manual_seed(3)
    A = (torch.rand(2, 3, 256, 256, generator=gen) * 2 - 1).to(DEV)
    B = (torch.rand(2, 3, 256, 256, generator=gen) * 2 - 1).to(DEV)
    torch.manual_seed(0)
    t_plain = Pix2PixTrainer(DEV)
    torch.manual_seed(0)
    t_drop = Pix2PixTrainer(DEV, use_dropout=True)
    for t in (t_plain, t_drop):
        t.G.training = False
        t.G.forward(A)
    assert torch.equal(t_plain.G.fake_f32, t_drop.G.fake_f32)
    t_drop.G.training = True
    t_drop.G.forward(A)
    f1 = t_drop.G.fake_f32.clone()
    t_drop.G.forward(A)
    assert not torch.equal(f1, t_drop.G.fake_f32)
    nbt0 = int(t_drop.G.state_dict()["model.model.1.model.2.num_batches_tracked"])
    losses = t_drop.train_step(A, B).cpu()
    assert torch.isfinite(losses).all()
    # two real generator forwards per iteration -> BatchNorm buffers advance by 2, as in the reference
    assert int(t_drop.G.state_dict()["model.model.1.model.2.num_batches_tracked"]) == nbt0 + 2
    g = t_drop.G.store.g
    assert torch.isfinite(g).all() and float(g.abs().sum()) > 0


def test_graphed_step_recaptures_when_buffers_or_hyperparameters_change():
    """a captured iteration bakes in the engines' activation buffers (raw pointers, TMA maps) and lr / betas.
    An eager step at another batch size re-allocates those buffers; the next graphed call at the original shape must
    re-capture instead of replaying into freed memory, results of successive replays must not alias, and the host step
    counters must follow the device-side ones."""
    from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(21)
    mk = lambda n: ((torch.rand(n, 3, 64, 64, generator=gen) * 2 - 1).to(dev), (torch.rand(n, 3, 64, 64, generator=gen) * 2 - 1).to(dev))
    b2 = [mk(2) for _ in range(4)]
    b3 = mk(3)
    seqs = []
    for graphed in (False, True):
        torch.manual_seed(0)
        tr = Pix2PixTrainer(dev, num_downs=5)
        fn = tr.train_step_graphed if graphed else tr.train_step
        out = [fn(*b2[0]), fn(*b2[1])]
        if graphed:
            assert out[0].data_ptr() != out[1].data_ptr()          # replays hand out fresh tensors
            sig = tr._graph_sig
        tr.train_step(*b3)                                          # eager, another shape: buffers are re-allocated
        out.append(fn(*b2[2]))                                      # must re-capture (graphed) — not replay a stale graph
        if graphed:
            assert tr._graph_sig != sig
            sig = tr._graph_sig
        tr.lr_g = 5e-5                                              # a kernel argument of the captured Adam launch
        out.append(fn(*b2[3]))
        if graphed:
            assert tr._graph_sig != sig
            assert int(tr.G.store.step_dev) == tr.G.store.step == 5 and tr.D.store.step == 5
        seqs.append(torch.stack([o.cpu() for o in out]))
    assert torch.allclose(seqs[0], seqs[1], rtol=3e-3, atol=3e-4), (seqs[0], seqs[1])
