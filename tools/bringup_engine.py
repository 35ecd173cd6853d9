"""Bring-up of the native Pix2Pix engine against the CPU oracle (run under gpurun).  Dev tool."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402
from oracle import pix2pix_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 256
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 3


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


torch.manual_seed(0)
tr = Pix2PixTrainer(dev)
sd_g = {k: v.detach().cpu().clone().contiguous() for k, v in tr.G.state_dict().items()}
sd_d = {k: v.detach().cpu().clone().contiguous() for k, v in tr.D.state_dict().items()}
print("G params", sum(v.numel() for k, v in sd_g.items() if v.dtype == torch.float32 and "running" not in k),
      "D params", sum(v.numel() for k, v in sd_d.items() if v.dtype == torch.float32 and "running" not in k))
gen = torch.Generator().manual_seed(1234)
batches = [(torch.rand(N, 3, HW, HW, generator=gen) * 2 - 1, torch.rand(N, 3, HW, HW, generator=gen) * 2 - 1)
           for _ in range(STEPS)]

# ---- forward-only checks (train-mode BN), no state change on the oracle side
A, B = batches[0]
with torch.no_grad():
    ref_fake = O.unet_generator_forward(sd_g, A, True, None)
    ref_pred = O.discriminator_forward(sd_d, torch.cat((A, B), 1), True, None)
tr.G.training = True
nbt_before = {k: v.clone() for k, v in tr.G.state_dict().items() if "running" in k or "num_batches" in k}
tr.G.forward(A.to(dev))
fake = tr.G.output_nchw().cpu()
print(f"G fwd: rel_l2={rel(fake, ref_fake):.3e} max_abs={(fake-ref_fake).abs().max():.3e}")
from gan_aug_pfa_b200 import ops  # noqa: E402
b_nhwc = torch.zeros(N, HW, HW, 4, device=dev, dtype=torch.bfloat16)
ops.nchw_to_nhwc_bf16(B.to(dev), b_nhwc)
logits = tr.D.forward(tr.G.x_nhwc, b_nhwc)
lg = logits.permute(0, 3, 1, 2).cpu()
print(f"D fwd: rel_l2={rel(lg, ref_pred):.3e} max_abs={(lg-ref_pred).abs().max():.3e}")
# restore BN buffers touched by the forward-only checks
tr.G.load_state_dict({k: v.to(dev) for k, v in sd_g.items()})
tr.D.load_state_dict({k: v.to(dev) for k, v in sd_d.items()})

# ---- training steps
names_g, names_d = O.param_names(sd_g), O.param_names(sd_d)
og = O.AdamState(sd_g, names_g, 1e-4, (0.5, 0.999))
od = O.AdamState(sd_d, names_d, 1e-4, (0.5, 0.999))
for step, (A, B) in enumerate(batches):
    t0 = time.time()
    ld, lg_, aux = O.gan_train_step(sd_g, sd_d, og, od, A, B, return_grads=True)
    t1 = time.time()
    losses = tr.train_step(A.to(dev), B.to(dev)).cpu()
    print(f"step {step}: oracle loss_d={ld:.6f} loss_g={lg_:.6f} | gpu loss_d={losses[0]:.6f} loss_g={losses[1]:.6f}"
          f"  (oracle {t1-t0:.1f}s)")
    if step == 0:
        worst = (1.0, "")
        for k in names_d:
            g = tr.D.grad(k).cpu()
            c = cos(g, aux["grads_d"][k])
            print(f"   D grad {k:28s} cos={c:.5f} rel={rel(g, aux['grads_d'][k]):.3e}")
        for k in names_g:
            g = tr.G.grad(k).cpu()
            c = cos(g, aux["grads_g"][k])
            print(f"   G grad {k:58s} cos={c:.5f} rel={rel(g, aux['grads_g'][k]):.3e}")
        for k, v in tr.G.state_dict().items():
            if "running_var" in k or "num_batches" in k:
                r = sd_g[k]
                print(f"   G buf {k:58s} rel={rel(v.cpu().double(), r.double()):.3e}")
# final parameter drift
print("param drift after steps:")
mx = 0.0
for k in names_g:
    mx = max(mx, rel(tr.G.param(k).cpu(), sd_g[k].detach()))
print(f"   G max rel param diff {mx:.3e}")
mx = 0.0
for k in names_d:
    mx = max(mx, rel(tr.D.param(k).cpu(), sd_d[k].detach()))
print(f"   D max rel param diff {mx:.3e}")
