"""Loss-curve parity over 300 GAN iterations against the reference's own train_gan_one_epoch run
(tests/golden/gan_curve.json, produced by tests/golden/make_curve.py from the unmodified reference on the
deterministic learnable pairs of tests/curve_data.py; batch 1, 256x256, seed 0).

GAN training is chaotic: the reference re-run with its initial weights merely rounded to bf16
(tests/golden/gan_curve_perturbed.json) already drifts from itself (EMA loss_g by up to 4.1 %, EMA loss_d by up to
44 %), and two native runs differ from each other through fp32-atomic ordering (EMA loss_d by ~25 % mid-run).  The
band is therefore stated on EMA(0.98)-smoothed curves after a 50-step burn-in:
  loss_g  pointwise within max(3 x the reference's own perturbation deviation, 10 %)   (measured: 2.8 %)
  loss_d  mean over steps 100..299 within 40 % of the reference's, pointwise EMA within a factor of 3
          (measured over several runs: mean within 20 %; pointwise the EMA swings between -52 % and +73 % of the
          reference's because the discriminator's short-term wins and losses are not reproducible — the reference
          perturbed by one bf16 rounding does the same)."""
import json
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

from curve_data import ema, pairs  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"


def test_gan_loss_curves_stay_in_the_reference_band():
    from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer
    ref = json.loads((GOLD / "gan_curve.json").read_text())
    per = json.loads((GOLD / "gan_curve_perturbed.json").read_text())
    steps = ref["steps"]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    tr = Pix2PixTrainer(dev)
    data = [(a.to(dev), b.to(dev)) for a, b in pairs()]
    out = []
    for s in range(steps):
        a, b = data[s % len(data)]
        out.append(tr.train_step(a, b))
    got = torch.stack(out).cpu().tolist()
    r = ema([x[1] for x in ref["loss_d_g"]])
    p = ema([x[1] for x in per["loss_d_g"]])
    g = ema([x[1] for x in got])
    worst = max(abs(g[i] - r[i]) / abs(r[i]) / max(3.0 * abs(p[i] - r[i]) / abs(r[i]), 0.10) for i in range(50, steps))
    assert worst <= 1.0, f"loss_g: EMA curve leaves the band (worst deviation / band = {worst:.2f})"
    rd = ema([x[0] for x in ref["loss_d_g"]])
    gd = ema([x[0] for x in got])
    mean_r = sum(x[0] for x in ref["loss_d_g"][100:]) / (steps - 100)
    mean_g = sum(x[0] for x in got[100:]) / (steps - 100)
    assert abs(mean_g - mean_r) < 0.40 * mean_r, (mean_g, mean_r)
    assert all(rd[i] / 3.0 < gd[i] < 3.0 * rd[i] for i in range(50, steps))
    # the first iteration is not chaotic yet: it must match the reference closely
    assert abs(got[0][0] - ref["loss_d_g"][0][0]) < 5e-3
    assert abs(got[0][1] - ref["loss_d_g"][0][1]) < 5e-3 * ref["loss_d_g"][0][1]
    # and training must actually make progress on this learnable data
    assert ema([x[1] for x in got])[-1] < 0.3 * got[0][1]
