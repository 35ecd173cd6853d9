"""Dump RUNS native loss curves (STEPS iterations each, tests/curve_data.pairs(), batch 1) to gpurun_out/ for calibrating
tests/test_gpu_loss_curve.py.  usage: curve_dump.py [steps] [runs]"""
import json
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from curve_data import pairs  # noqa: E402
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
out = []
for run in range(runs):
    torch.manual_seed(0)
    tr = Pix2PixTrainer(dev)
    data = [(a.to(dev), b.to(dev)) for a, b in pairs()]
    out.append(torch.stack([tr.train_step(*data[s % len(data)]) for s in range(steps)]).cpu().tolist())
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "native_curves.json").write_text(json.dumps(out))
print("dumped", runs, "runs of", steps, "steps")
