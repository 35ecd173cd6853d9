"""bf16 yardstick + reference scalars for the Siamese U-Net at the sizes the numbers are quoted on.

For each fixture — (batch 4, 128x128): the reference defaults train.py:330,333; (batch 1, 512x512): BASELINE.json
config 5's image size — run the UNMODIFIED reference `models.SiameseUNet(3, 1)` (seed 0) with `train.CombinedLoss()`
on seeded synthetic inputs twice: in fp32, and under `torch.autocast("cpu", dtype=torch.bfloat16)`.  The difference
between the two is what bf16 arithmetic alone does to this network on this input ("torch-bf16 yardstick", SURVEY.md
§8c); tests/test_gpu_siamese_sizes.py bounds the native kernels by 1.5 x that yardstick.  Also stored: the fp32
reference's loss and logits moments, so the GPU box (which has no /root/reference) can check the oracle it compares
against.  Writes tests/golden/siamese_yardstick.json.   Run in the build container:
    python tests/golden/make_siamese_yardstick.py"""
import json
import os
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import make_golden as MG  # noqa: E402

FIXTURES = [(4, 128), (1, 512)]


def inputs(n, s):
    g = torch.Generator().manual_seed(77 + s)
    x1 = torch.rand(n, 3, s, s, generator=g) * 2 - 1
    x2 = torch.rand(n, 3, s, s, generator=g) * 2 - 1
    lab = (torch.rand(n, s, s, generator=g) < 0.05).long()
    return x1, x2, lab


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))


def main():
    sys.path.insert(0, str(MG.REF))
    real = os.makedirs
    os.makedirs = lambda *a, **k: None
    models = MG._load("models")
    MG._load("dataset")
    train = MG._load("train", stubs=("optuna",))
    os.makedirs = real
    torch.set_num_threads(int(os.environ.get("YARD_THREADS", "8")))
    out = {}
    for n, s in FIXTURES:
        x1, x2, lab = inputs(n, s)
        res = {}
        for mode in ("fp32", "bf16"):
            torch.manual_seed(0)
            m = models.SiameseUNet(3, 1)
            m.train()
            crit = train.CombinedLoss()
            if mode == "bf16":
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    logits = m(x1, x2)
                    loss = crit(logits.float(), lab)
            else:
                logits = m(x1, x2)
                loss = crit(logits, lab)
            loss.backward()
            res[mode] = (logits.detach().float(), float(loss), {k: p.grad.detach().clone() for k, p in m.named_parameters()})
            print(n, s, mode, float(loss), flush=True)
        (l32, c32, g32), (l16, c16, g16) = res["fp32"], res["bf16"]
        names = list(g32)
        w3 = [k for k in names if g32[k].dim() == 4 and g32[k].shape[2] == 3]
        out[f"n{n}_s{s}"] = {
            "n": n, "s": s,
            "ref_loss_combined": c32,
            "ref_logits_mean": float(l32.double().mean()), "ref_logits_std": float(l32.double().std()),
            "yard_logits_rel_l2": float((l16.double() - l32.double()).norm() / l32.double().norm()),
            "yard_loss_rel": abs(c16 - c32) / abs(c32),
            "yard_grad_cos_whole": cos(torch.cat([g16[k].flatten() for k in names]), torch.cat([g32[k].flatten() for k in names])),
            "yard_grad_cos_min_3x3": min(cos(g16[k], g32[k]) for k in w3),
            "yard_grad_cos_per_tensor": {k: cos(g16[k], g32[k]) for k in w3},
            "yard_grad_norm_ratio_3x3": [min(float(g16[k].norm() / g32[k].norm()) for k in w3),
                                         max(float(g16[k].norm() / g32[k].norm()) for k in w3)],
        }
        print(json.dumps({k: v for k, v in out[f"n{n}_s{s}"].items() if k != "yard_grad_cos_per_tensor"}), flush=True)
    (HERE / "siamese_yardstick.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
