"""Loss-curve parity over 1000 GAN iterations against the reference's own train_gan_one_epoch runs.

Fixtures (tests/golden/, produced by make_curve.py from the UNMODIFIED reference on the deterministic learnable pairs of
tests/curve_data.py; batch 1, 256x256, seed 0, 1000 iterations each):
  gan_curve.json              the reference
  gan_curve_perturbed.json    the reference with its initial weights rounded once to bf16
  gan_curve_perturbed{1,2}.json   ... multiplied by (1 + 2^-9 u), u ~ U(-1, 1), seeds 1 and 2
The three perturbed runs measure how far the REFERENCE drifts from itself under a perturbation the size of one bf16
rounding.  GAN training is chaotic: up to ~250 iterations the four runs stay close (EMA(0.98) loss_g within 4 %,
loss_d within 15 %); from ~300 on the discriminator's short-term wins and collapses happen at different times in every
run (100-iteration window means differ by up to 2.6x for loss_g and 30x for loss_d between reference runs), while the
long-run level is stable (mean over iterations 300-999: loss_g 5.74 ... 6.41, loss_d 0.20 ... 0.30).  The band is
therefore stated in three parts, all on the native run vs the reference family:
  1. iterations 50-249, pointwise on EMA(0.98) curves:  |native - reference| <= max(3 x the family's own deviation
     from the reference at that iteration, 10 % for loss_g / 40 % for loss_d)          (measured: 0.29-0.55 of the band)
  2. iterations 300-999, long-run level: the native mean lies in [0.85 x family min, 1.15 x family max] for loss_g and
     [0.6 x min, 1.6 x max] for loss_d                   (measured loss_g 5.67-6.43 in [4.88, 7.37]; loss_d 0.18-0.29 in [0.12, 0.49])
  3. iterations 300-999, pointwise EMA envelope (no blow-up, no collapse): loss_g within [family min / 1.5, family max x 1.5],
     loss_d within [family min / 10, family max x 3]     (measured ratios: loss_g 0.88 ... 1.21, loss_d 0.17 ... 1.82)
plus the first iteration (not chaotic yet) to 5e-3 and real training progress."""
import json
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

from curve_data import ema, pairs  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"
FAMILY = ("gan_curve.json", "gan_curve_perturbed.json", "gan_curve_perturbed1.json", "gan_curve_perturbed2.json")


def test_gan_loss_curves_stay_in_the_reference_band_over_1000_iterations():
    from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer
    fam = [json.loads((GOLD / f).read_text()) for f in FAMILY]
    steps = fam[0]["steps"]
    assert steps == 1000 and all(f["steps"] == steps for f in fam)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    tr = Pix2PixTrainer(dev)
    data = [(a.to(dev), b.to(dev)) for a, b in pairs()]
    out = []
    for s in range(steps):
        a, b = data[s % len(data)]
        out.append(tr.train_step(a, b))
    got = torch.stack(out).cpu().tolist()
    assert all(torch.isfinite(torch.tensor(got)).flatten())
    for col, name, floor, lvl_lo, lvl_hi, env_lo, env_hi in ((1, "loss_g", 0.10, 0.85, 1.15, 1.5, 1.5),
                                                             (0, "loss_d", 0.40, 0.60, 1.60, 10.0, 3.0)):
        fe = [ema([x[col] for x in f["loss_d_g"]]) for f in fam]
        g = ema([x[col] for x in got])
        ref = fe[0]
        # 1. early, pointwise
        worst = 0.0
        for i in range(50, 250):
            env = max(abs(e[i] - ref[i]) for e in fe[1:]) / abs(ref[i])
            worst = max(worst, abs(g[i] - ref[i]) / abs(ref[i]) / max(3.0 * env, floor))
        assert worst <= 1.0, f"{name}: EMA curve leaves the early band (worst deviation / band = {worst:.2f})"
        # 2. long-run level
        fm = [sum(x[col] for x in f["loss_d_g"][300:]) / (steps - 300) for f in fam]
        gm = sum(x[col] for x in got[300:]) / (steps - 300)
        assert lvl_lo * min(fm) <= gm <= lvl_hi * max(fm), f"{name}: long-run mean {gm:.4f} vs reference family {fm}"
        # 3. pointwise envelope
        for i in range(300, steps):
            lo, hi = min(e[i] for e in fe), max(e[i] for e in fe)
            assert lo / env_lo <= g[i] <= hi * env_hi, f"{name}: EMA {g[i]:.4f} at iteration {i} outside [{lo:.4f}/{env_lo}, {hi:.4f}x{env_hi}]"
    # the first iteration is not chaotic yet: it must match the reference closely
    first = fam[0]["loss_d_g"][0]
    assert abs(got[0][0] - first[0]) < 5e-3
    assert abs(got[0][1] - first[1]) < 5e-3 * first[1]
    # and training must actually make progress on this learnable data
    assert ema([x[1] for x in got])[-1] < 0.3 * got[0][1]
