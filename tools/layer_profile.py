"""Per-launch device timing of one Pix2Pix training iteration (run under gpurun).  Dev tool.

Every gap_* call is bracketed by CUDA events on the launching stream (warm caches, real clocks — unlike
the serialised cold-cache ncu launch list) and labelled with its shape; GEMM launches also get their
algorithmic TFLOP/s.  usage: layer_profile.py [batch] [steps] [min_us]
"""
import collections
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import _lib  # noqa: E402
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
MIN_US = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0

real = _lib.lib()
RECORDS = []
ENABLED = [False]


def label_of(name, args):
    a = getattr(args[0], "_obj", None) if args else None
    if name == "gap_conv_gemm":
        ctot = a.src_c[0] + a.src_c[1]
        fl = 2.0 * a.n * a.gh * a.gw * a.n_phase * a.n_out * a.taps_h * a.taps_w * ctot
        return (f"gemm n{a.n} {ctot}->{a.n_out} grid{a.gh}x{a.gw} ph{a.n_phase} taps{a.taps_h}x{a.taps_w} "
                f"s{a.in_stride}{' stats' if a.stats else ''}{' out2' if a.out2 else ''}", fl)
    if name == "gap_conv_wgrad":
        mr = a.m_rows if a.m_rows > 0 else a.m_c
        fl = 2.0 * a.n * a.gh * a.gw * mr * a.n_c * a.taps_h * a.taps_w
        return f"wgrad n{a.n} m{a.m_c} n{a.n_c} grid{a.gh}x{a.gw} taps{a.taps_h}x{a.taps_w} s{a.stride}", fl
    def _i(v):
        return int(getattr(v, "value", v) or 0)
    if name == "gap_bn_bwd_apply":     # y ld g1 ld g2 ld slope scale shift mean invstd pixels c ...
        px, c = _i(args[11]), _i(args[12])
        nb = 3 if args[4] is not None and getattr(args[4], "value", args[4]) else 2
        return f"bn_bwd_apply px{px} c{c} tensors{nb + 1}", -float(px * c * 2 * (nb + 1))
    if name == "gap_bn_act":           # y ld scale shift pixels c o1 ld1 a1 o2 ...
        px, c = _i(args[4]), _i(args[5])
        no = 2 if args[9] is not None and getattr(args[9], "value", args[9]) else 1
        return f"bn_act px{px} c{c} outs{no}", -float(px * c * 2 * (1 + no))
    return name[4:], 0.0


class Proxy:
    def __getattr__(self, name):
        fn = getattr(real, name)
        if not name.startswith("gap_") or name in ("gap_last_error_string", "gap_debug_set", "gap_version",
                                                   "gap_sm_count"):
            return fn

        def wrapped(*args):
            if not ENABLED[0]:
                return fn(*args)
            lab, fl = label_of(name, args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            RECORDS.append((lab, fl, e0, e1))
            return rc

        return wrapped


_lib._lib = Proxy()

dev = torch.device("cuda:0")
torch.manual_seed(0)
WORK = sys.argv[4] if len(sys.argv) > 4 else "pix2pix"
gen = torch.Generator().manual_seed(1234)
if WORK == "siamese":
    from gan_aug_pfa_b200.siamese import SiameseEngine
    from gan_aug_pfa_b200 import models as M

    class _T:
        def __init__(self):
            self.e = SiameseEngine(dev)
            self.e.load_state_dict({k: v.detach() for k, v in M.SiameseUNet(3, 1).state_dict().items()})

        def train_step(self, a, b):
            return self.e.train_step(a, b, LAB, kind="combined")
    tr = _T()
    import os
    if os.environ.get("GAP_NO_OVERLAP"):
        tr.e.overlap_wgrad = False
    A = (torch.rand(N, 3, 512, 512, generator=gen) * 2 - 1).to(dev)
    B = (torch.rand(N, 3, 512, 512, generator=gen) * 2 - 1).to(dev)
    LAB = (torch.rand(N, 512, 512, generator=gen) < 0.05).long().to(dev)
else:
    tr = Pix2PixTrainer(dev)
    import os
    if os.environ.get("GAP_NO_OVERLAP"):      # serialise the wgrad side stream: clean per-kernel times
        tr.G.overlap_wgrad = tr.D.overlap_wgrad = False
        tr.overlap_g_fwd = False
    A = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)
    B = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)
for _ in range(3):
    tr.train_step(A, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(STEPS):
    tr.train_step(A, B)
e1.record()
torch.cuda.synchronize()
print(f"untraced: {e0.elapsed_time(e1) / STEPS * 1e3:.0f} us/step")
ENABLED[0] = True
for _ in range(STEPS):
    tr.train_step(A, B)
torch.cuda.synchronize()
per = len(RECORDS) // STEPS
agg = collections.OrderedDict()
seq = []
for i, (lab, fl, a, b) in enumerate(RECORDS):
    us = a.elapsed_time(b) * 1e3
    d = agg.setdefault(lab, [0, 0.0, 0.0])
    d[0] += 1
    d[1] += us
    d[2] += fl
    if i >= len(RECORDS) - per:
        seq.append((lab, us, fl))
tot = sum(v[1] for v in agg.values()) / STEPS
print(f"traced sum: {tot:.0f} us/step over {per} launches")
kinds = collections.OrderedDict()
for lab, v in agg.items():
    k = lab.split(" ")[0]
    d = kinds.setdefault(k, [0, 0.0, 0.0])
    d[0] += v[0] / STEPS
    d[1] += v[1] / STEPS
    d[2] += v[2] / STEPS
print("--- by kind")
for k, v in sorted(kinds.items(), key=lambda kv: -kv[1][1]):
    tf = (f"{v[2] / v[1] / 1e6:7.1f} TF/s" if v[2] > 0 else f"{-v[2] / v[1] / 1e3:7.0f} GB/s") if v[2] else ""
    print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}% x{v[0]:4.0f}  {k} {tf}")
print("--- by shape")
for lab, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    us = v[1] / STEPS
    if us < MIN_US:
        continue
    tf = (f"{v[2] / v[1] / 1e6:7.1f} TF/s" if v[2] > 0 else f"{-v[2] / v[1] / 1e3:7.0f} GB/s") if v[2] else ""
    print(f"{us:9.1f} us {100 * us / tot:5.1f}% x{v[0] / STEPS:4.0f}  {lab} {tf}")
