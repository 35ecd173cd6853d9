"""torch-bf16 yardstick for the batch-64 Pix2Pix iteration ("bf16 error, not bugs, dominates
end-to-end gradient differences").

Runs the CPU oracle's gan_train_step (pinned to the reference, tests/test_oracle_vs_reference.py) on the batch-64
fixture of tests/test_gpu_batch64.py twice — in fp32 and under torch.autocast("cpu", dtype=torch.bfloat16), i.e. with
torch's own bf16 convolutions — and records, per parameter tensor, the cosine between the two gradients, plus the
relative differences of the losses and of the generator output.  tests/test_gpu_batch64.py bounds the native kernels
by   1 - cos <= max(0.03, 2.25 x (1 - yardstick cos))   per tensor (1.5 x the yardstick's error; error^2 ~ 1 - cos).
Writes tests/golden/gan_yardstick_b64.json.     python tests/golden/make_gan_yardstick.py [batch]"""
import json
import os
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from gan_aug_pfa_b200 import spec  # noqa: E402
from oracle import pix2pix_oracle as O  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))


def run(bf16: bool):
    torch.manual_seed(0)
    sd_g, sd_d = spec.default_state_dicts()
    og = O.AdamState(sd_g, O.param_names(sd_g), 1e-4, (0.5, 0.999))
    od = O.AdamState(sd_d, O.param_names(sd_d), 1e-4, (0.5, 0.999))
    gen = torch.Generator().manual_seed(1234)
    A = torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1
    B = torch.rand(N, 3, 256, 256, generator=gen) * 2 - 1
    if bf16:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            return O.gan_train_step(sd_g, sd_d, og, od, A, B, return_grads=True)
    return O.gan_train_step(sd_g, sd_d, og, od, A, B, return_grads=True)


def main():
    torch.set_num_threads(int(os.environ.get("YARD_THREADS", "8")))
    ld, lg, aux = run(False)
    print("fp32", ld, lg, flush=True)
    ld16, lg16, aux16 = run(True)
    print("bf16", ld16, lg16, flush=True)
    out = {"batch": N, "loss_d": ld, "loss_g": lg, "yard_loss_d_rel": abs(ld16 - ld) / abs(ld),
           "yard_loss_g_rel": abs(lg16 - lg) / abs(lg),
           "yard_fake_rel_l2": float((aux16["fake_B"].float() - aux["fake_B"]).norm() / aux["fake_B"].norm()),
           "cos_g": {k: cos(aux16["grads_g"][k].float(), v) for k, v in aux["grads_g"].items()},
           "cos_d": {k: cos(aux16["grads_d"][k].float(), v) for k, v in aux["grads_d"].items()}}
    print("worst G", min(out["cos_g"].items(), key=lambda kv: kv[1]), "worst D", min(out["cos_d"].items(), key=lambda kv: kv[1]))
    (HERE / f"gan_yardstick_b{N}.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
