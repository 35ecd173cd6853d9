"""Per-parameter gradient comparison of the native Siamese engine against the CPU oracle (run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import models  # noqa: E402
from gan_aug_pfa_b200.siamese import SiameseEngine  # noqa: E402
from oracle import pix2pix_oracle as O  # noqa: E402

HW = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = models.SiameseUNet(3, 1)
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
gen = torch.Generator().manual_seed(77)
x1 = torch.rand(N, 3, HW, HW, generator=gen) * 2 - 1
x2 = torch.rand(N, 3, HW, HW, generator=gen) * 2 - 1
lab = (torch.rand(N, HW, HW, generator=gen) < 0.05).long()
names = O.param_names(sd)
for k in names:
    sd[k].requires_grad_(True)
out_ref = O.siamese_forward(sd, x1, x2, True, {})
loss_ref = O.combined_loss(out_ref, lab)
ref_g = dict(zip(names, torch.autograd.grad(loss_ref, [sd[k] for k in names])))
eng = SiameseEngine(dev)
eng.load_state_dict({k: v.detach() for k, v in sd.items()})
eng.training = True
eng.zero_grad()
logits = eng.forward(x1.to(dev), x2.to(dev))
loss = eng.loss_and_grad(lab.to(dev), "combined")
eng.backward()
torch.cuda.synchronize()
o = logits.cpu().unsqueeze(1)
print(f"logits rel {float((o - out_ref.detach()).norm() / out_ref.norm()):.4f}  loss {float(loss):.5f} ref {float(loss_ref):.5f}")
a_all, b_all = [], []
for k in names:
    g = eng.grad(k).detach().cpu().double().reshape(-1)
    r = ref_g[k].double().reshape(-1)
    cos = float(g @ r / (g.norm() * r.norm()).clamp_min(1e-30))
    ratio = float(g.norm() / r.norm().clamp_min(1e-30))
    flag = "" if cos > 0.9 else "   <<<<"
    print(f"{k:32s} cos {cos:7.4f}  |g|/|ref| {ratio:8.4f}  |ref| {float(r.norm()):.3e}{flag}")
    a_all.append(g)
    b_all.append(r)
a, b = torch.cat(a_all), torch.cat(b_all)
print("TOTAL cos", float(a @ b / (a.norm() * b.norm())))
