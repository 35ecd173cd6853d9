"""Host-side launchers: torch tensors in, C-ABI calls out.

PyTorch is used for device memory and streams only; every computation below is a ``gap_*`` call
into libgap_b200.so.  Activations are NHWC bf16 tensors (possibly channel slices of a wider
buffer: ``x[..., a:b]`` keeps the parent's pixel stride).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH  # noqa: F401


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _nhwc_view(t: torch.Tensor) -> tuple[int, int, int, int, int]:
    """Return (n, h, w, c, pixel_stride) of an NHWC bf16 tensor that may be a channel slice."""
    if t.dtype != torch.bfloat16 or t.dim() != 4 or not t.is_cuda:
        raise ValueError(f"expected a CUDA NHWC bf16 tensor, got {t.dtype} {tuple(t.shape)} {t.device}")
    n, h, w, c = t.shape
    sn, sh, sw, sc = t.stride()
    if c > 1 and sc != 1:
        raise ValueError("NHWC tensor must be channel-contiguous")
    ld = sw
    if (w > 1 and sh != w * ld and h > 1) or (n > 1 and sn != h * w * ld):
        raise ValueError(f"NHWC tensor must be dense in n/h/w (strides {t.stride()})")
    return n, h, w, c, ld


@dataclass(frozen=True)
class Geometry:
    """Iteration geometry of one gap_conv_gemm call (see gap_b200.h)."""
    n_phase: int
    taps_h: int
    taps_w: int
    in_stride: int
    in_off_h: tuple[int, int]
    in_off_w: tuple[int, int]
    out_stride: int


def geom_conv_fwd(k: int, stride: int, pad: int) -> Geometry:
    """Conv2d forward (models.py:177,223,230,238,243,9,12) and ConvTranspose2d dgrad."""
    return Geometry(1, k, k, stride, (-pad, -pad), (-pad, -pad), 1)


def geom_conv_dgrad_s1(k: int, pad: int) -> Geometry:
    """Stride-1 Conv2d dgrad: a stride-1 conv over dY with flipped taps and pad k-1-pad."""
    return Geometry(1, k, k, 1, (-(k - 1 - pad),) * 2, (-(k - 1 - pad),) * 2, 1)


def geom_phase_k4s2p1() -> Geometry:
    """ConvTranspose2d(k4,s2,p1) forward (models.py:184,189,194) and Conv2d(k4,s2,p1) dgrad:
    four output-parity phases of 2x2 taps; phase (ph,pw), tap (th,tw) uses kernel element
    (3-ph-2*th, 3-pw-2*tw)."""
    return Geometry(4, 2, 2, 1, (-1, 0), (-1, 0), 2)


def conv_gemm(srcs: Sequence[torch.Tensor], wpk: torch.Tensor, geom: Geometry, out: torch.Tensor,
              n_out: int, grid_hw: tuple[int, int], *, act: int = ACT_NONE,
              out2: Optional[torch.Tensor] = None, act2: int = ACT_NONE,
              bias: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None) -> None:
    """Launch the implicit-GEMM engine.  ``wpk`` is [n_phase, rows, taps*ctot] bf16."""
    a = _lib.ConvGemmArgs()
    n = ih = iw = None
    for i in range(2):
        if i < len(srcs):
            sn, sh, sw, sc, ld = _nhwc_view(srcs[i])
            if n is None:
                n, ih, iw = sn, sh, sw
            elif (n, ih, iw) != (sn, sh, sw):
                raise ValueError("concatenated sources must share n/h/w")
            a.src[i] = srcs[i].data_ptr()
            a.src_c[i] = sc
            a.src_ld[i] = ld
        else:
            a.src[i] = None
            a.src_c[i] = 0
            a.src_ld[i] = 0
    a.n, a.ih, a.iw = n, ih, iw
    a.gh, a.gw = grid_hw
    a.n_phase = geom.n_phase
    a.taps_h, a.taps_w = geom.taps_h, geom.taps_w
    a.in_stride = geom.in_stride
    a.in_off_h[0], a.in_off_h[1] = geom.in_off_h
    a.in_off_w[0], a.in_off_w[1] = geom.in_off_w
    a.out_stride = geom.out_stride
    if wpk.dtype != torch.bfloat16 or wpk.dim() != 3 or not wpk.is_contiguous():
        raise ValueError("wpk must be a contiguous [n_phase, rows, K] bf16 tensor")
    ctot = sum(int(s.shape[3]) for s in srcs)
    if wpk.shape[0] != geom.n_phase or wpk.shape[2] != geom.taps_h * geom.taps_w * ctot:
        raise ValueError(f"wpk shape {tuple(wpk.shape)} does not match geometry/channels {ctot}")
    a.wpk = wpk.data_ptr()
    a.w_rows = wpk.shape[1]
    a.n_out = n_out
    on, oh, ow, oc, old = _nhwc_view(out)
    if on != n or oc < n_out:
        raise ValueError("output tensor does not match")
    a.oh, a.ow = oh, ow
    a.out = out.data_ptr()
    a.out_ld = old
    a.act = act
    if out2 is not None:
        o2n, o2h, o2w, o2c, o2ld = _nhwc_view(out2)
        if (o2n, o2h, o2w) != (on, oh, ow) or o2c < n_out:
            raise ValueError("out2 does not match out")
        a.out2 = out2.data_ptr()
        a.out2_ld = o2ld
    else:
        a.out2 = None
        a.out2_ld = 0
    a.act2 = act2
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() < n_out:
            raise ValueError("bias must be fp32 with >= n_out elements")
        a.bias = bias.data_ptr()
    else:
        a.bias = None
    if stats is not None:
        if stats.dtype != torch.float64 or stats.numel() != 2 * n_out:
            raise ValueError("stats must be fp64 [2*n_out]")
        a.stats = stats.data_ptr()
    else:
        a.stats = None
    _lib.check(_lib.lib().gap_conv_gemm(C.byref(a), _stream()), "gap_conv_gemm")


def conv_wgrad(mop: torch.Tensor, nop: torch.Tensor, out: torch.Tensor, taps: tuple[int, int], stride: int,
               off: tuple[int, int], ld_m: int, ld_tap: int) -> None:
    """out[m*ld_m + tap*ld_tap + c] += sum_pix mop[pix, m] * nop[gather(pix, tap), c]   (fp32 out)."""
    a = _lib.WgradArgs()
    n, gh, gw, mc, mld = _nhwc_view(mop)
    n2, nh, nw, nc, nld = _nhwc_view(nop)
    if n != n2:
        raise ValueError("mop / nop batch mismatch")
    if out.dtype != torch.float32 or not out.is_cuda:
        raise ValueError("wgrad output must be a CUDA fp32 tensor")
    a.mop, a.m_c, a.m_ld = mop.data_ptr(), mc, mld
    a.nop, a.n_c, a.n_ld = nop.data_ptr(), nc, nld
    a.n, a.gh, a.gw, a.nh, a.nw = n, gh, gw, nh, nw
    a.taps_h, a.taps_w = taps
    a.stride = stride
    a.off_h, a.off_w = off
    a.out = out.data_ptr()
    a.ld_m, a.ld_tap = ld_m, ld_tap
    _lib.check(_lib.lib().gap_conv_wgrad(C.byref(a), _stream()), "gap_conv_wgrad")
