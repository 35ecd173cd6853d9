"""Loss curves of the UNMODIFIED reference (oracle/_ref: models.py + train_gan.train_gan_one_epoch) run on cuda through
torch eager / cuDNN, in fp32 (TF32 off) and under bf16 autocast: what does bf16 arithmetic alone do to the long-run level
of this GAN?  Calibration data for tests/test_gpu_loss_curve.py (tests/golden/gan_curve_cuda.json = this script's output
with the values rounded to 5 significant digits).  Needs a GPU and the staged reference (oracle/stage_ref.py):
    gpurun -- python tests/golden/make_curve_cuda.py 1000 6      -> gpurun_out/ref_cuda_curves.json"""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from curve_data import pairs  # noqa: E402
from oracle import ref_loader  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ns = ref_loader.load(tag="curve")
ns.train_gan.DEVICE = dev
out = {}
for mode in ("fp32", "bf16"):
    out[mode] = []
    for run in range(runs):
        torch.manual_seed(0)
        G = ns.models.UNetGenerator(3, 3).to(dev)
        D = ns.models.NLayerDiscriminator(6).to(dev)
        if run > 0:          # perturbation of the size of one bf16 rounding, as in tests/golden/make_curve.py
            gp = torch.Generator().manual_seed(run)
            with torch.no_grad():
                for q in list(G.parameters()) + list(D.parameters()):
                    q.mul_(1 + 2.0 ** -9 * (torch.rand(q.shape, generator=gp) * 2 - 1).to(dev))
        og = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.5, 0.999))
        od = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.5, 0.999))
        data = [(a.to(dev), b.to(dev)) for a, b in pairs()]
        seq = []
        for s in range(steps):
            a, b = data[s % len(data)]
            if mode == "bf16":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    ld, lg = ns.train_gan.train_gan_one_epoch(G, D, [{"image1": a, "image2": b}], og, od)
            else:
                ld, lg = ns.train_gan.train_gan_one_epoch(G, D, [{"image1": a, "image2": b}], og, od)
            seq.append([float(ld), float(lg)])
        out[mode].append(seq)
        m = sum(x[1] for x in seq[300:]) / max(1, steps - 300)
        d = sum(x[0] for x in seq[300:]) / max(1, steps - 300)
        print(mode, run, f"mean[300:] loss_g {m:.3f} loss_d {d:.3f}", flush=True)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "ref_cuda_curves.json").write_text(json.dumps(out))
