"""Print the EMA loss-curve deviation of the native run from the reference golden (run under gpurun)."""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from curve_data import ema, pairs  # noqa: E402
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402

ref = json.loads((ROOT / "tests/golden/gan_curve.json").read_text())
per = json.loads((ROOT / "tests/golden/gan_curve_perturbed.json").read_text())
dev = torch.device("cuda:0")
for run in range(2):
    torch.manual_seed(0)
    tr = Pix2PixTrainer(dev)
    data = [(a.to(dev), b.to(dev)) for a, b in pairs()]
    got = torch.stack([tr.train_step(*data[s % len(data)]) for s in range(ref["steps"])]).cpu().tolist()
    for col, name in ((0, "loss_d"), (1, "loss_g")):
        r = ema([x[col] for x in ref["loss_d_g"]]); p = ema([x[col] for x in per["loss_d_g"]]); g = ema([x[col] for x in got])
        dv = [abs(g[i] - r[i]) / abs(r[i]) for i in range(50, len(r))]
        pv = [abs(p[i] - r[i]) / abs(r[i]) for i in range(50, len(r))]
        print(f"run {run} {name}: max dev {max(dv):.3f} at {50 + dv.index(max(dv))}, end dev {dv[-1]:.3f} | reference-perturbed max {max(pv):.3f} end {pv[-1]:.3f}"
              f" | ema at 100/200/299: native {g[100]:.3f} {g[200]:.3f} {g[299]:.3f} ref {r[100]:.3f} {r[200]:.3f} {r[299]:.3f}")
    print("first step", got[0], ref["loss_d_g"][0])
