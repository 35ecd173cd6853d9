import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from gan_aug_pfa_b200 import ops, _lib  # noqa: E402
from ablate_wgrad import bench  # noqa: E402
for skip in (15, 0):
    print("skip", skip)
    for splits in (37, 74, 148, 296, 592):
        _lib.debug_set("wgrad_splits", splits)
        us, tf = bench(64, 128, 64, 64, 4, 2, 1, skip, iters=20)
        print(f"  m128 n64 g64 splits {splits:4d} (grid {4*splits}): {us:8.1f} us", flush=True)
    for splits in (2, 5, 10, 20, 40):
        _lib.debug_set("wgrad_splits", splits)
        us, tf = bench(64, 512, 256, 31, 4, 1, 1, skip, iters=20)
        print(f"  m512 n256 g31 s1 splits {splits:4d} (grid {64*splits}): {us:8.1f} us", flush=True)
# launch overhead probe: tiny problem
_lib.debug_set("wgrad_splits", 0)
us, tf = bench(1, 64, 64, 8, 1, 1, 0, 0, iters=50)
print(f"tiny problem: {us:.1f} us per call (host launch overhead bound)")
