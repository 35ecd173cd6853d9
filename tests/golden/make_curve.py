"""Loss curve of the UNMODIFIED reference (train_gan.train_gan_one_epoch, models.py) on tests/curve_data.pairs():
STEPS iterations at batch 1, 256x256, seed 0, Adam(1e-4, (0.5, 0.999)).  Writes tests/golden/gan_curve.json.
Run in the build container:  python tests/golden/make_curve.py [steps]"""
import json
import os
import sys
import time
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent))
import make_golden as MG  # noqa: E402
from curve_data import pairs  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 300
# argv[2] = "perturb":  initial weights rounded to bf16 — the reference's own sensitivity to one bf16 rounding;
# argv[2] = "perturbK" (K = 1, 2, ...): initial weights multiplied by (1 + 2^-9 u), u ~ U(-1, 1) from seed K (the same
# relative magnitude as a bf16 rounding, a different direction).  Together they calibrate the band of the GPU test.
PERTURB = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2].startswith("perturb") else ""
THREADS = int(os.environ.get("CURVE_THREADS", "8"))


def main():
    sys.path.insert(0, str(MG.REF))
    real = os.makedirs
    os.makedirs = lambda *a, **k: None
    models = MG._load("models")
    MG._load("dataset")
    train_gan = MG._load("train_gan")
    os.makedirs = real
    torch.set_num_threads(THREADS)
    torch.manual_seed(0)
    G = models.UNetGenerator(3, 3)
    D = models.NLayerDiscriminator(6)
    if PERTURB == "perturb":
        with torch.no_grad():
            for q in list(G.parameters()) + list(D.parameters()):
                q.copy_(q.to(torch.bfloat16).float())
    elif PERTURB:
        gp = torch.Generator().manual_seed(int(PERTURB[len("perturb"):]))
        with torch.no_grad():
            for q in list(G.parameters()) + list(D.parameters()):
                q.mul_(1 + 2.0 ** -9 * (torch.rand(q.shape, generator=gp) * 2 - 1))
    opt_g = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.5, 0.999))
    opt_d = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.5, 0.999))
    train_gan.tqdm = lambda it, **k: MG._NoBar(it)
    data = pairs()
    seq = []
    t0 = time.time()
    for s in range(STEPS):
        a, b = data[s % len(data)]
        ld, lg = train_gan.train_gan_one_epoch(G, D, [{"image1": a, "image2": b}], opt_g, opt_d)
        seq.append([float(ld), float(lg)])
        if s % 25 == 0:
            print(s, seq[-1], f"{time.time() - t0:.0f}s", flush=True)
    name = "gan_curve.json" if not PERTURB else "gan_curve_perturbed" + PERTURB[len("perturb"):] + ".json"
    (HERE / name).write_text(json.dumps({"steps": STEPS, "loss_d_g": seq,
                                                     "note": "reference train_gan_one_epoch, batch 1, 256x256, "
                                                             "tests/curve_data.pairs() cycled, seed 0"}))


if __name__ == "__main__":
    main()
