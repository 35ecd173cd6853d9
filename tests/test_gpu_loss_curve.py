"""Loss-curve parity over 1000 GAN iterations against the reference's own train_gan_one_epoch runs.

Fixtures (tests/golden/; batch 1, 256x256, seed 0, the deterministic learnable pairs of tests/curve_data.py, 1000
iterations each, all produced by the UNMODIFIED reference):
  gan_curve.json                     the reference on the CPU (make_curve.py)
  gan_curve_perturbed.json           ... with its initial weights rounded once to bf16
  gan_curve_perturbed{1..6}.json     ... multiplied by (1 + 2^-9 u), u ~ U(-1, 1), seeds 1-6
  gan_curve_cuda.json                the reference on a B200 through torch eager / cuDNN (make_curve_cuda.py): 6 runs in
                                     fp32 (TF32 off) and 6 under torch.autocast(bfloat16), same perturbations
Twenty reference runs, because GAN training is chaotic: for ~250 iterations they stay close (EMA(0.98) loss_g within
4 %, loss_d within 15 % on the CPU family), from ~300 on the discriminator's short-term wins and collapses happen at
different times in every run (100-iteration window means differ by up to 2.6x for loss_g and 30x for loss_d between
reference runs) while the long-run level (mean over iterations 300-999) spreads over loss_g 5.04 ... 6.84 and loss_d
0.156 ... 0.409 (bf16 runs sit lower in loss_g: torch's own bf16 arithmetic shifts the level by about as much).  Ten
native runs measured loss_g 4.59 ... 6.58 and loss_d 0.164 ... 0.504.  The band has three parts:
  1. iterations 50-199, pointwise on EMA(0.98) curves against the CPU reference:  |native - reference| <= max(3 x the CPU
     family's own deviation at that iteration, 10 % for loss_g / 40 % for loss_d)
     (ten native runs: 0.22-0.42 / 0.31-0.58 of the band; the reference's own cuda-bf16 runs: up to 0.60 / 0.67)
  2. iterations 300-999, long-run level: the native mean lies in [0.85 x min, 1.15 x max] of the twenty reference means
     for loss_g and in [0.6 x min, 1.6 x max] for loss_d           (i.e. [4.28, 7.87] and [0.094, 0.654])
  3. iterations 300-999, EMA envelope (no blow-up, no collapse): within [family min / 1.5, family max x 1.5] for loss_g
     and [family min / 3, family max x 3] for loss_d, the family min / max taken over the twenty runs AND over +-50
     iterations (one EMA time constant) around the iteration: the discriminator's dips come at different times in every
     run, so a bound that is pointwise in time is marginal for the reference ITSELF -- leave-one-out over the twenty
     reference runs, the worst one sits at 0.37 x the others' pointwise minimum (the bound is 0.333) but at 0.79 x the
     windowed one; a native run measured 0.0132 at iteration 987, where the pointwise family minimum is 0.0437 and was
     0.0182 twenty iterations earlier.              (native runs vs the windowed envelope: >= 0.73 x min, <= 1.3 x max)
plus the first iteration (not chaotic yet) to 5e-3 and real training progress."""
import json
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu

from curve_data import ema, pairs  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"
FAMILY = ("gan_curve.json", "gan_curve_perturbed.json") + tuple(f"gan_curve_perturbed{k}.json" for k in range(1, 7))


def test_gan_loss_curves_stay_in_the_reference_band_over_1000_iterations():
    from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer
    fam = [json.loads((GOLD / f).read_text()) for f in FAMILY]
    steps = fam[0]["steps"]
    assert steps == 1000 and all(f["steps"] == steps for f in fam)
    cuda = json.loads((GOLD / "gan_curve_cuda.json").read_text())
    wide = [f["loss_d_g"] for f in fam] + cuda["fp32"] + cuda["bf16"]        # all twenty reference runs
    assert len(wide) == 20 and all(len(r) == steps for r in wide)
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
                                                             (0, "loss_d", 0.40, 0.60, 1.60, 3.0, 3.0)):
        fe = [ema([x[col] for x in f["loss_d_g"]]) for f in fam]
        g = ema([x[col] for x in got])
        ref = fe[0]
        # 1. early, pointwise
        worst = 0.0
        for i in range(50, 200):
            env = max(abs(e[i] - ref[i]) for e in fe[1:]) / abs(ref[i])
            worst = max(worst, abs(g[i] - ref[i]) / abs(ref[i]) / max(3.0 * env, floor))
        assert worst <= 1.0, f"{name}: EMA curve leaves the early band (worst deviation / band = {worst:.2f})"
        # 2. long-run level
        fm = [sum(x[col] for x in r[300:]) / (steps - 300) for r in wide]
        gm = sum(x[col] for x in got[300:]) / (steps - 300)
        assert lvl_lo * min(fm) <= gm <= lvl_hi * max(fm), \
            f"{name}: long-run mean {gm:.4f} outside [{lvl_lo} x {min(fm):.4f}, {lvl_hi} x {max(fm):.4f}]"
        # 3. envelope over the twenty runs and +-50 iterations
        we = [ema([x[col] for x in r]) for r in wide]
        for i in range(300, steps):
            w0, w1 = max(300, i - 50), min(steps, i + 51)
            lo, hi = min(min(e[w0:w1]) for e in we), max(max(e[w0:w1]) for e in we)
            assert lo / env_lo <= g[i] <= hi * env_hi, f"{name}: EMA {g[i]:.4f} at iteration {i} outside [{lo:.4f}/{env_lo}, {hi:.4f}x{env_hi}]"
    # the first iteration is not chaotic yet: it must match the reference closely
    first = fam[0]["loss_d_g"][0]
    assert abs(got[0][0] - first[0]) < 5e-3
    assert abs(got[0][1] - first[1]) < 5e-3 * first[1]
    # and training must actually make progress on this learnable data
    assert ema([x[1] for x in got])[-1] < 0.3 * got[0][1]
