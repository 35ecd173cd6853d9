"""wgrad shape sweep with knob variants (run under gpurun)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402
from ablate_wgrad import bench  # noqa: E402

# (name, n, m_c, n_c, gh, k, s, p)  as they occur in the Pix2Pix step at batch 64
SHAPES = [
    ("G/D conv 64->128   m128 n64  g64 s2", 64, 128, 64, 64, 4, 2, 1),
    ("G/D conv 128->256  m256 n128 g32 s2", 64, 256, 128, 32, 4, 2, 1),
    ("D conv 256->512 s1 m512 n256 g31 s1", 64, 512, 256, 31, 4, 1, 1),
    ("G conv 256->512    m512 n256 g16 s2", 64, 512, 256, 16, 4, 2, 1),
    ("G conv 512->512    m512 n512 g8  s2", 64, 512, 512, 8, 4, 2, 1),
    ("G conv 512->512    m512 n512 g4  s2", 64, 512, 512, 4, 4, 2, 1),
    ("G conv 512->512    m512 n512 g2  s2", 64, 512, 512, 2, 4, 2, 1),
    ("G convT 256->64    m256 n64  g64 s2", 64, 256, 64, 64, 4, 2, 1),
    ("G convT 512->128   m512 n128 g32 s2", 64, 512, 128, 32, 4, 2, 1),
    ("G convT 1024->256  m1024 n256 g16 s2", 64, 1024, 256, 16, 4, 2, 1),
    ("G convT 1024->512  m1024 n512 g8 s2", 64, 1024, 512, 8, 4, 2, 1),
    ("G convT 1024->512  m1024 n512 g4 s2", 64, 1024, 512, 4, 4, 2, 1),
]
variants = [dict()]
for a in sys.argv[1:]:
    variants.append(dict(kv.split("=") for kv in a.split(",")))
for name, n, mc, nc, gh, k, s, p in SHAPES:
    line = f"{name:40s}"
    for v in variants:
        for kk in ("wgrad_mt", "wgrad_tpc", "wgrad_splits", "wgrad_acc_cols", "wgrad_block_n"):
            _lib.debug_set(kk, int(v.get(kk, 0)) if kk != "wgrad_acc_cols" else int(v.get(kk, 512)))
        _lib.debug_set("wgrad_dual", int(v.get("wgrad_dual", -1)))
        try:
            us, tf = bench(n, mc, nc, gh, k, s, p, 0, iters=10, convT=False)
            line += f" | {us:7.1f} us {tf:6.0f} TF"
        except Exception as e:  # noqa: BLE001
            line += f" | ERR {str(e)[:30]}"
    print(line, flush=True)
