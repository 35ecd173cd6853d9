"""Interleaved (ABBA) A/B on the eager training step of a gap_debug_set knob, or of a trainer / engine attribute when the
name starts with "tr." (e.g. tr.overlap_g_fwd, tr.G.overlap_wgrad): python tools/ab_knob.py knob v0 v1 [rounds]
Several knobs in one process (each tested on its own, the others at their defaults):
python tools/ab_knob.py knobA v0 v1 knobB v0 v1 ... [rounds]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import _lib  # noqa: E402
from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer  # noqa: E402

argv = sys.argv[1:]
rounds = int(argv.pop()) if len(argv) % 3 == 1 else 6
specs = [(argv[i], int(argv[i + 1]), int(argv[i + 2])) for i in range(0, len(argv), 3)]
dev = torch.device("cuda:0")
N = 64
gen = torch.Generator().manual_seed(1234)
A = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)
B = (torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1).to(dev)


def timeit(fn, iters=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.manual_seed(0)
tr = Pix2PixTrainer(dev)
for _ in range(5):
    tr.train_step(A, B)


def set_knob(knob, v):
    if knob.startswith("tr."):
        obj = tr
        parts = knob.split(".")[1:]
        for a in parts[:-1]:
            obj = getattr(obj, a)
        setattr(obj, parts[-1], bool(v))
    else:
        _lib.debug_set(knob, v)


for knob, v0, v1 in specs:
    t = {v0: [], v1: []}
    for rnd in range(rounds):
        for v in ((v0, v1) if rnd % 2 == 0 else (v1, v0)):      # ABBA order: drift cancels
            set_knob(knob, v)
            tr.train_step(A, B)
            t[v].append(timeit(lambda: tr.train_step(A, B)))
    for v in (v0, v1):
        print(f"{knob}={v}:", " ".join(f"{x:.3f}" for x in t[v]), f"  mean {sum(t[v]) / len(t[v]):.3f} ms", flush=True)
    set_knob(knob, v0)       # the first value is the one the following knobs are measured with
