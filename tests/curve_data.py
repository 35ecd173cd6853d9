"""Deterministic learnable synthetic pairs for the loss-curve parity test (no files needed on the GPU box).

A_i: a smooth random field (8x8 uniform noise, bilinearly upsampled to 256x256, 3 channels, in [-1, 1]);
B_i: a fixed smooth function of A_i (channel mix + tanh), i.e. something a Pix2Pix generator can learn, unlike white
noise.  Eight pairs, cycled, batch 1 — the reference's own default batch size (train_gan.py:26)."""
import torch
import torch.nn.functional as F

N_PAIRS = 8
HW = 256


def pairs():
    g = torch.Generator().manual_seed(2024)
    out = []
    mix = torch.tensor([[0.2, 0.7, -0.4], [-0.6, 0.3, 0.5], [0.5, -0.5, 0.6]])
    for _ in range(N_PAIRS):
        low = torch.rand(1, 3, 8, 8, generator=g) * 2 - 1
        a = F.interpolate(low, size=(HW, HW), mode="bilinear", align_corners=True).clamp(-1, 1)
        b = torch.tanh(1.5 * torch.einsum("oc,nchw->nohw", mix, a))
        out.append((a.contiguous(), b.contiguous()))
    return out


def ema(xs, beta=0.98):
    out, m = [], None
    for x in xs:
        m = x if m is None else beta * m + (1 - beta) * x
        out.append(m)
    return out
