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


# bench.py sets PROFILE to a list to time every GEMM launch with CUDA events:
# entries are (kernel name, algorithmic FLOPs, start event, end event).
PROFILE = None


def _timed(name: str, flops: float, fn) -> None:
    if PROFILE is None:
        fn()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    PROFILE.append((name, flops, e0, e1))


def _stream() -> C.c_void_p:
    """The current stream of the CURRENT device: the engines pin the current device to their own for the duration of
    every entry point (pix2pix._on_device), so this is the stream of the tensors being launched on."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _nhwc_view(t: torch.Tensor, dtype=torch.bfloat16) -> tuple[int, int, int, int, int]:
    """Return (n, h, w, c, pixel_stride) of an NHWC bf16 tensor that may be a channel slice."""
    if t.dtype != dtype or t.dim() != 4 or not t.is_cuda:
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
              bias: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
              flops: Optional[float] = None, bwd: Optional[dict] = None, scale: Optional[torch.Tensor] = None,
              accumulate: bool = False) -> None:
    """Launch the implicit-GEMM engine.  ``wpk`` is [n_phase, rows, taps*ctot] bf16.  ``flops`` overrides
    the algorithmic FLOP count reported to the profiler (layers that pad channels pass the true one).
    ``accumulate``: add the result to what ``out`` already holds instead of overwriting it."""
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
    on, oh, ow, oc, old = _nhwc_view(out, out.dtype if out.dtype == torch.float32 else torch.bfloat16)
    a.out_f32 = 1 if out.dtype == torch.float32 else 0
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
    c0 = int(bwd.get("c0", 0)) if bwd is not None else 0
    if scale is not None:
        if scale.dtype != torch.float32 or scale.numel() < n_out:
            raise ValueError("scale must be fp32 with >= n_out elements")
        a.scale = scale.data_ptr()
    else:
        a.scale = None
    a.accumulate = 1 if accumulate else 0
    if stats is not None:
        if stats.dtype != torch.float64 or stats.numel() != 2 * (n_out - c0):
            raise ValueError("stats must be fp64 [2*(n_out - c0)]")
        a.stats = stats.data_ptr()
    else:
        a.stats = None
    if bwd is not None:
        # backward-fused epilogue: see gap_conv_gemm_args.bwd_* in gap_b200.h
        y = bwd["y"]
        yn, yh, yw, yc, yld = _nhwc_view(y)
        if (yn, yh, yw) != (on, oh, ow) or yc != n_out - c0:
            raise ValueError("bwd y must be [n, oh, ow, n_out - c0]")
        a.bwd_y, a.bwd_y_ld = y.data_ptr(), yld
        sc, sh = bwd.get("scale"), bwd.get("shift")
        a.bwd_scale = None if sc is None else sc.data_ptr()
        a.bwd_shift = None if sh is None else sh.data_ptr()
        g2 = bwd.get("g2")
        if g2 is not None:
            gn, gh_, gw_, gc, gld = _nhwc_view(g2)
            if (gn, gh_, gw_, gc) != (on, oh, ow, n_out - c0):
                raise ValueError("bwd g2 must match y")
            a.bwd_g2, a.bwd_g2_ld = g2.data_ptr(), gld
        else:
            a.bwd_g2, a.bwd_g2_ld = None, 0
        a.bwd_slope = float(bwd.get("slope", 0.0))
        a.bwd_c0 = c0
    else:
        a.bwd_y = None
    if flops is None:
        flops = 2.0 * n * grid_hw[0] * grid_hw[1] * geom.n_phase * n_out * geom.taps_h * geom.taps_w * ctot
    _timed("conv_fprop_kernel", flops,
           lambda: _lib.check(_lib.lib().gap_conv_gemm(C.byref(a), _stream()), "gap_conv_gemm"))


def conv_wgrad(mop: torch.Tensor, nop: torch.Tensor, out: torch.Tensor, taps: tuple[int, int], stride: int,
               off: tuple[int, int], ld_m: int, ld_tap: int, m_rows: int = 0,
               flops: Optional[float] = None) -> None:
    """out[m*ld_m + tap*ld_tap + c] += sum_pix mop[pix, m] * nop[gather(pix, tap), c]   (fp32 out)."""
    a = _lib.WgradArgs()
    n, gh, gw, mc, mld = _nhwc_view(mop)
    n2, nh, nw, nc, nld = _nhwc_view(nop)
    if n != n2:
        raise ValueError("mop / nop batch mismatch")
    if out.dtype != torch.float32 or not out.is_cuda:
        raise ValueError("wgrad output must be a CUDA fp32 tensor")
    a.mop, a.m_c, a.m_ld = mop.data_ptr(), mc, mld
    a.m_rows = m_rows
    a.nop, a.n_c, a.n_ld = nop.data_ptr(), nc, nld
    a.n, a.gh, a.gw, a.nh, a.nw = n, gh, gw, nh, nw
    a.taps_h, a.taps_w = taps
    a.stride = stride
    a.off_h, a.off_w = off
    a.out = out.data_ptr()
    a.ld_m, a.ld_tap = ld_m, ld_tap
    if flops is None:
        flops = 2.0 * n * gh * gw * (m_rows if m_rows > 0 else mc) * nc * taps[0] * taps[1]
    _timed("conv_wgrad_kernel", flops,
           lambda: _lib.check(_lib.lib().gap_conv_wgrad(C.byref(a), _stream()), "gap_conv_wgrad"))


# ------------------------------------------------------------------------------------------------
# thin wrappers over the elementwise / reduction entry points
# ------------------------------------------------------------------------------------------------
def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _rows_ld(t: torch.Tensor) -> tuple[int, int, int]:
    """(pixels, channels, pixel stride) of a channel-contiguous [..., C] tensor dense in the rest."""
    c = t.shape[-1]
    ld = t.stride(-2) if t.dim() >= 2 else c
    return t.numel() // c, c, ld


def nchw_to_nhwc_bf16(x: torch.Tensor, out: torch.Tensor) -> None:
    n, c, h, w = x.shape
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("expected a contiguous fp32 NCHW tensor")
    _lib.check(_lib.lib().gap_nchw_f32_to_nhwc_bf16(_ptr(x), _ptr(out), n, c, h, w, out.stride(2), _stream()),
               "gap_nchw_f32_to_nhwc_bf16")


def nhwc_to_nchw_f32(x: torch.Tensor, out: torch.Tensor, c: int, c_total: Optional[int] = None, c_off: int = 0) -> None:
    n, h, w, _ = x.shape
    _lib.check(_lib.lib().gap_nhwc_to_nchw_f32(_ptr(x), 1 if x.dtype == torch.float32 else 0, _ptr(out), n, c, h, w,
                                               x.stride(2), c if c_total is None else c_total, c_off, _stream()),
               "gap_nhwc_to_nchw_f32")


def tanh_bwd(gout_nchw: torch.Tensor, y_nhwc_f32: torch.Tensor, dpre: torch.Tensor) -> None:
    n, c, h, w = gout_nchw.shape
    _lib.check(_lib.lib().gap_tanh_bwd(_ptr(gout_nchw), _ptr(y_nhwc_f32), y_nhwc_f32.stride(2), _ptr(dpre),
                                       dpre.stride(2), n, c, h, w, _stream()), "gap_tanh_bwd")


def gen_out_bwd(fake_f32: torch.Tensor, real: torch.Tensor, dfake_d: Optional[torch.Tensor], l1_scale: float,
                dpre: torch.Tensor, loss_acc: torch.Tensor) -> None:
    """L1 * lambda + Tanh backward.  ``real`` is real_B either as fp32 NCHW (what the reference's DataLoader yields) or
    as the raw uint8 [n, h, w, 3] image (normalised in the kernel like dataset.py does)."""
    n, h, w, _ = fake_f32.shape
    if real.dtype == torch.uint8:
        if tuple(real.shape) != (n, h, w, 3) or not real.is_contiguous():
            raise ValueError("uint8 real_B must be contiguous [n, h, w, 3]")
        fn, name, c = _lib.lib().gap_gen_out_bwd_u8, "gap_gen_out_bwd_u8", 3
    else:
        fn, name, c = _lib.lib().gap_gen_out_bwd, "gap_gen_out_bwd", real.shape[1]
    _lib.check(fn(_ptr(fake_f32), fake_f32.stride(2), _ptr(real), h * w, _ptr(dfake_d),
                  0 if dfake_d is None else dfake_d.stride(2), l1_scale, _ptr(dpre), dpre.stride(2), n * h * w, c,
                  _ptr(loss_acc), _stream()), name)


def gan_losses(acc4: torch.Tensor, count: int, l1_weight: float, numel: int, out2: torch.Tensor) -> None:
    """[loss_D, loss_G] of train_gan.py:61,68-69 from the iteration's four fp64 loss sums (re-zeroed)."""
    if acc4.dtype != torch.float64 or out2.dtype != torch.float64 or acc4.numel() < 4 or out2.numel() < 2:
        raise ValueError("acc4 / out2 must be fp64 with 4 / 2 elements")
    _lib.check(_lib.lib().gap_gan_losses(_ptr(acc4), float(count), float(l1_weight), float(numel), _ptr(out2), _stream()),
               "gap_gan_losses")


def bce_logits_const(logits: torch.Tensor, target: float, grad_scale: float, dlogits: Optional[torch.Tensor],
                     loss_acc: torch.Tensor) -> None:
    _lib.check(_lib.lib().gap_bce_logits_const(_ptr(logits), logits.numel(), target, grad_scale, _ptr(dlogits),
                                               0 if dlogits is None else dlogits.stride(2), _ptr(loss_acc), _stream()),
               "gap_bce_logits_const")


def bce_logits_const_f32(logits: torch.Tensor, target: float, grad_scale: float, dlogits: Optional[torch.Tensor],
                         loss_acc: torch.Tensor, dbias: Optional[torch.Tensor] = None) -> None:
    """BCEWithLogitsLoss vs a constant target (train_gan.py:58,60,67): fp32 gradient (+ the producing conv's
    bias gradient)."""
    if dlogits is not None and (dlogits.dtype != torch.float32 or dlogits.numel() != logits.numel()):
        raise ValueError("dlogits must be fp32 with the shape of logits")
    _lib.check(_lib.lib().gap_bce_logits_const_f32(_ptr(logits), logits.numel(), target, grad_scale, _ptr(dlogits),
                                                   _ptr(loss_acc), _ptr(dbias), _stream()), "gap_bce_logits_const_f32")


def sum_f32(x: torch.Tensor, out: torch.Tensor) -> None:
    _lib.check(_lib.lib().gap_sum_f32(_ptr(x), x.numel(), _ptr(out), _stream()), "gap_sum_f32")


def _pre(pre: Optional[tuple], c: int):
    """(scale, shift, slope) of a fused input transform LeakyReLU_slope(x*scale + shift) -> ctypes arguments."""
    if pre is None:
        return None, None, 0.0
    sc, sh, slope = pre
    if sc.dtype != torch.float32 or sh.dtype != torch.float32 or sc.numel() < c or sh.numel() < c:
        raise ValueError("pre = (scale, shift, slope) needs fp32 vectors with >= c elements")
    return _ptr(sc), _ptr(sh), float(slope)


def cout1_conv_fwd(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], z_ws: torch.Tensor,
                   logits: torch.Tensor, ksize: int = 4, pad: int = 1, pre: Optional[tuple] = None) -> None:
    """Conv2d(C -> 1, k4, s1, p1) + bias (models.py:243): x NHWC bf16, w bf16 [16*C] ([kh][kw][c]), logits fp32
    [n, oh, ow(, 1)].  pre = (scale, shift, slope): x is the raw conv output below and BatchNorm + LeakyReLU
    (models.py:239-240) are applied while it is read."""
    n, ih, iw, c, ld = _nhwc_view(x)
    if z_ws.dtype != torch.float32 or z_ws.numel() < n * ih * iw * 16 or logits.dtype != torch.float32:
        raise ValueError("z_ws / logits must be fp32 (z_ws >= n*ih*iw*16 elements)")
    if w.dtype != torch.bfloat16 or w.numel() != ksize * ksize * c:
        raise ValueError("w must be bf16 [k*k*C]")
    if logits.numel() != n * (ih + 2 * pad - ksize + 1) * (iw + 2 * pad - ksize + 1):
        raise ValueError("logits shape mismatch")
    sc, sh, slope = _pre(pre, c)
    _lib.check(_lib.lib().gap_cout1_conv_fwd(_ptr(x), ld, n, ih, iw, c, _ptr(w), _ptr(bias), ksize, pad, _ptr(z_ws),
                                             _ptr(logits), sc, sh, slope, _stream()), "gap_cout1_conv_fwd")


def cout1_conv_dgrad(dlogits: torch.Tensor, w: torch.Tensor, gx: torch.Tensor, ksize: int = 4, pad: int = 1,
                     bwd: Optional[dict] = None) -> None:
    """Input gradient of Conv2d(C -> 1, k4, s1).  ``bwd`` = dict(y, scale, shift, slope, sums) fuses the activation
    backward of the layer below and its BatchNorm-backward sums into the kernel (see gap_cout1_conv_dgrad_bwd)."""
    n, ih, iw, c, ld = _nhwc_view(gx)
    oh, ow = ih + 2 * pad - ksize + 1, iw + 2 * pad - ksize + 1
    if dlogits.dtype != torch.float32 or dlogits.numel() != n * oh * ow:
        raise ValueError("dlogits must be fp32 [n, oh, ow]")
    if bwd is None:
        _lib.check(_lib.lib().gap_cout1_conv_dgrad(_ptr(dlogits), n, oh, ow, _ptr(w), ksize, pad, c, _ptr(gx), ld, ih, iw,
                                                   _stream()), "gap_cout1_conv_dgrad")
        return
    y = bwd["y"]
    yn, yh, yw, yc, yld = _nhwc_view(y)
    if (yn, yh, yw, yc) != (n, ih, iw, c) or bwd["sums"].dtype != torch.float64 or bwd["sums"].numel() != 2 * c:
        raise ValueError("bwd y must match gx and sums must be fp64 [2c]")
    _lib.check(_lib.lib().gap_cout1_conv_dgrad_bwd(_ptr(dlogits), n, oh, ow, _ptr(w), ksize, pad, c, _ptr(gx), ld, ih, iw,
                                                   _ptr(y), yld, _ptr(bwd["scale"]), _ptr(bwd["shift"]),
                                                   float(bwd.get("slope", 0.0)), _ptr(bwd["sums"]), _stream()),
               "gap_cout1_conv_dgrad_bwd")


def cout1_conv_wgrad(dlogits: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, ksize: int = 4, pad: int = 1,
                     pre: Optional[tuple] = None) -> None:
    n, ih, iw, c, ld = _nhwc_view(x)
    oh, ow = ih + 2 * pad - ksize + 1, iw + 2 * pad - ksize + 1
    if dlogits.dtype != torch.float32 or dlogits.numel() != n * oh * ow:
        raise ValueError("dlogits must be fp32 [n, oh, ow]")
    if dw.dtype != torch.float32 or dw.numel() != ksize * ksize * c:
        raise ValueError("dw must be fp32 [k*k*C]")
    sc, sh, slope = _pre(pre, c)
    _lib.check(_lib.lib().gap_cout1_conv_wgrad(_ptr(dlogits), n, oh, ow, _ptr(x), ld, ih, iw, c, ksize, pad, _ptr(dw),
                                               sc, sh, slope, _stream()), "gap_cout1_conv_wgrad")


def thin_conv_fwd(s0: torch.Tensor, s1: Optional[torch.Tensor], wpk: torch.Tensor, bias: Optional[torch.Tensor],
                  out1: torch.Tensor, act1: int = ACT_NONE, out2: Optional[torch.Tensor] = None,
                  act2: int = ACT_NONE) -> None:
    """Conv2d(k4,s2,p1) over one or two 4-slot NHWC sources (3 channels + a zero slot each) -> cw = 64/128
    channels, fused bias + activation(s); wpk bf16 [cw, 16*CT] with CT = 4 (one source) or 8 (two)."""
    n, h, w, _, ld0 = _nhwc_view(s0)
    ld1 = 0
    if s1 is not None:
        n1, h1, w1, _, ld1 = _nhwc_view(s1)
        if (n1, h1, w1) != (n, h, w):
            raise ValueError("sources must share n/h/w")
    on, oh, ow, cw, ldo1 = _nhwc_view(out1)
    if (on, oh, ow) != (n, h // 2, w // 2):
        raise ValueError("out1 shape mismatch")
    ct = 8 if s1 is not None else 4
    if wpk.dtype != torch.bfloat16 or wpk.numel() != cw * 16 * ct or not wpk.is_contiguous():
        raise ValueError(f"wpk must be contiguous bf16 [{cw}, {16 * ct}]")
    ldo2 = 0
    if out2 is not None:
        o2 = _nhwc_view(out2)
        if o2[:4] != (on, oh, ow, cw):
            raise ValueError("out2 shape mismatch")
        ldo2 = o2[4]
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() < cw):
        raise ValueError("bias must be fp32 [cw]")
    _lib.check(_lib.lib().gap_thin_conv_fwd(_ptr(s0), ld0, _ptr(s1), ld1, n, h, w, _ptr(wpk), _ptr(bias), cw, _ptr(out1),
                                            ldo1, act1, _ptr(out2), ldo2, act2, _stream()), "gap_thin_conv_fwd")


def thin_conv_wgrad(wide: torch.Tensor, s0: torch.Tensor, s1: Optional[torch.Tensor], dw: torch.Tensor, ld_m: int,
                    dbias: Optional[torch.Tensor] = None) -> None:
    """dw[cw][(kh*4+kw)*c + ch] += sum_pix wide[pix][cw] * thin[2*pix+tap-1][ch] (fp32 master layout, row stride
    ld_m); dbias[cw] += sum_pix wide[pix][cw]."""
    n, oh, ow, cw, ldw = _nhwc_view(wide)
    n0, h, w, _, ld0 = _nhwc_view(s0)
    if (n0, h, w) != (n, 2 * oh, 2 * ow):
        raise ValueError("thin source must be at twice the resolution of the wide tensor")
    ld1 = 0
    if s1 is not None:
        n1, h1, w1, _, ld1 = _nhwc_view(s1)
        if (n1, h1, w1) != (n, h, w):
            raise ValueError("sources must share n/h/w")
    c = 6 if s1 is not None else 3
    if dw.dtype != torch.float32 or dw.numel() < (cw - 1) * ld_m + 16 * c:
        raise ValueError("dw too small")
    _lib.check(_lib.lib().gap_thin_conv_wgrad(_ptr(wide), ldw, _ptr(s0), ld0, _ptr(s1), ld1, n, h, w, cw, _ptr(dw), ld_m,
                                              _ptr(dbias), _stream()), "gap_thin_conv_wgrad")


def u8_hwc_to_nhwc_bf16(x: torch.Tensor, out: torch.Tensor) -> None:
    """uint8 [n,h,w,3] -> normalised NHWC bf16 with 4 channel slots (dataset.py:155-159 on the device)."""
    if x.dtype != torch.uint8 or x.shape[-1] != 3 or not x.is_contiguous() or not x.is_cuda:
        raise ValueError("expected a contiguous CUDA uint8 [n,h,w,3] tensor")
    _lib.check(_lib.lib().gap_u8_hwc_to_nhwc_bf16(_ptr(x), _ptr(out), out.stride(2), x.numel() // 3, _stream()),
               "gap_u8_hwc_to_nhwc_bf16")


def resize_u8_to_nhwc_bf16(x: torch.Tensor, out: Optional[torch.Tensor], out_nchw_f32: Optional[torch.Tensor] = None) -> None:
    """dataset.py's ToTensor + JointResize(BILINEAR, antialias) + JointNormalize on the device: uint8 [n,ih,iw,3] ->
    NHWC bf16 [n,oh,ow,4 slots] and / or the fp32 NCHW [n,3,oh,ow] tensor the reference's DataLoader yields (the output
    tensors' h / w are the target size)."""
    if x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3 or not x.is_contiguous() or not x.is_cuda:
        raise ValueError("expected a contiguous CUDA uint8 [n,h,w,3] tensor")
    n = x.shape[0]
    ld = 0
    if out is not None:
        on, oh, ow, _, ld = _nhwc_view(out)
        if on != n or out.shape[-1] < 4:
            raise ValueError("output must be NHWC bf16 [n, oh, ow, >=4]")
    if out_nchw_f32 is not None:
        f = out_nchw_f32
        if f.dtype != torch.float32 or f.dim() != 4 or f.shape[0] != n or f.shape[1] != 3 or not f.is_contiguous() or not f.is_cuda:
            raise ValueError("out_nchw_f32 must be a contiguous CUDA fp32 [n,3,oh,ow] tensor")
        if out is not None and (f.shape[2], f.shape[3]) != (oh, ow):
            raise ValueError("both outputs must have the same size")
        oh, ow = f.shape[2], f.shape[3]
    if out is None and out_nchw_f32 is None:
        raise ValueError("no output given")
    _lib.check(_lib.lib().gap_resize_u8_to_nhwc_bf16(_ptr(x), n, x.shape[1], x.shape[2], oh, ow, _ptr(out), ld,
                                                     _ptr(out_nchw_f32), _stream()), "gap_resize_u8_to_nhwc_bf16")


def resize_nearest_i64(x: torch.Tensor, out: torch.Tensor) -> None:
    """JointResize's label half (dataset.py:143-146): nearest-neighbour resize of an int64 [n,ih,iw] map into [n,oh,ow]."""
    if x.dtype != torch.int64 or out.dtype != torch.int64 or x.dim() != 3 or out.dim() != 3 or x.shape[0] != out.shape[0] \
            or not (x.is_contiguous() and out.is_contiguous() and x.is_cuda and out.is_cuda):
        raise ValueError("expected contiguous CUDA int64 [n,h,w] tensors")
    _lib.check(_lib.lib().gap_resize_nearest_i64(_ptr(x), x.shape[0], x.shape[1], x.shape[2], out.shape[1], out.shape[2],
                                                 _ptr(out), _stream()), "gap_resize_nearest_i64")


def thin_convT_fwd(wide: torch.Tensor, wcol: torch.Tensor, bias: Optional[torch.Tensor], act: int,
                   out_bf16: Optional[torch.Tensor], out_f32: Optional[torch.Tensor],
                   out_u8: Optional[torch.Tensor] = None) -> None:
    """ConvTranspose2d(cw -> 3, k4, s2, p1) (+bias, Tanh) into 4-slot NHWC outputs; wcol bf16 [48 = tap*3+co, cw]."""
    n, ih, iw, cw, ldw = _nhwc_view(wide)
    if wcol.dtype != torch.bfloat16 or tuple(wcol.shape) != (48, cw) or not wcol.is_contiguous():
        raise ValueError(f"wcol must be contiguous bf16 [48, {cw}]")
    for o, dt in ((out_bf16, torch.bfloat16), (out_f32, torch.float32)):
        if o is not None and (o.dtype != dt or tuple(o.shape[:3]) != (n, 2 * ih, 2 * iw) or o.shape[3] < 4):
            raise ValueError("output must be [n, 2ih, 2iw, >=4]")
    _lib.check(_lib.lib().gap_thin_convT_fwd(_ptr(wide), ldw, n, ih, iw, cw, _ptr(wcol), _ptr(bias), act, _ptr(out_bf16),
                                             0 if out_bf16 is None else out_bf16.stride(2), _ptr(out_f32),
                                             0 if out_f32 is None else out_f32.stride(2), _ptr(out_u8), _stream()),
               "gap_thin_convT_fwd")


def bn_finalize(stats, count, gamma, beta, eps, momentum, repeat, running_mean, running_var, nbt, scale, shift,
                save_mean, save_invstd) -> None:
    c = scale.numel()
    _lib.check(_lib.lib().gap_bn_finalize(_ptr(stats), c, float(count), _ptr(gamma), _ptr(beta), eps, momentum, repeat,
                                          _ptr(running_mean), _ptr(running_var), _ptr(nbt), _ptr(scale), _ptr(shift),
                                          _ptr(save_mean), _ptr(save_invstd), _stream()), "gap_bn_finalize")


def bn_eval_scale_shift(gamma, beta, running_mean, running_var, eps, scale, shift) -> None:
    _lib.check(_lib.lib().gap_bn_eval_scale_shift(scale.numel(), _ptr(gamma), _ptr(beta), _ptr(running_mean),
                                                  _ptr(running_var), eps, _ptr(scale), _ptr(shift), _stream()),
               "gap_bn_eval_scale_shift")


def bn_act(y: torch.Tensor, scale, shift, out1: torch.Tensor, act1: int, out2: Optional[torch.Tensor] = None,
           act2: int = ACT_NONE) -> None:
    pixels, c, ld = _rows_ld(y)
    _lib.check(_lib.lib().gap_bn_act(_ptr(y), ld, _ptr(scale), _ptr(shift), pixels, c, _ptr(out1), out1.stride(-2),
                                     act1, _ptr(out2), 0 if out2 is None else out2.stride(-2), act2, _stream()),
               "gap_bn_act")


def bn_bwd_reduce(y, g1, g2, slope, scale, shift, mean, invstd, sums) -> None:
    pixels, c, ld = _rows_ld(y)
    _lib.check(_lib.lib().gap_bn_bwd_reduce(_ptr(y), ld, _ptr(g1), g1.stride(-2), _ptr(g2),
                                            0 if g2 is None else g2.stride(-2), slope, _ptr(scale), _ptr(shift),
                                            _ptr(mean), _ptr(invstd), pixels, c, _ptr(sums), _stream()),
               "gap_bn_bwd_reduce")


def bn_bwd_apply(y, g1, g2, slope, scale, shift, mean, invstd, sums, count, dy) -> None:
    pixels, c, ld = _rows_ld(y)
    _lib.check(_lib.lib().gap_bn_bwd_apply(_ptr(y), ld, _ptr(g1), g1.stride(-2), _ptr(g2),
                                           0 if g2 is None else g2.stride(-2), slope, _ptr(scale), _ptr(shift),
                                           _ptr(mean), _ptr(invstd), pixels, c, _ptr(sums), float(count), _ptr(dy),
                                           dy.stride(-2), _stream()), "gap_bn_bwd_apply")


def bn_param_grads(sums, dgamma, dbeta) -> None:
    _lib.check(_lib.lib().gap_bn_param_grads(_ptr(sums), sums.numel() // 2, _ptr(dgamma), _ptr(dbeta), _stream()),
               "gap_bn_param_grads")


def bn_bwd_finalize(raw, mean, invstd, dgamma, dbeta, sums) -> None:
    """raw [sum d, sum d*y] (re-zeroed) -> sums [sum d, sum d*xhat]; dgamma / dbeta accumulated (may be None)."""
    _lib.check(_lib.lib().gap_bn_bwd_finalize(_ptr(raw), _ptr(mean), _ptr(invstd), mean.numel(), _ptr(dgamma),
                                              _ptr(dbeta), _ptr(sums), _stream()), "gap_bn_bwd_finalize")


def colsum_bf16(x: torch.Tensor, c: int, out: torch.Tensor) -> None:
    pixels = x.numel() // x.shape[-1]
    _lib.check(_lib.lib().gap_colsum_bf16(_ptr(x), x.stride(-2), pixels, c, _ptr(out), _stream()), "gap_colsum_bf16")


def adam_flat(p, g, m, v, lr, beta1, beta2, eps, weight_decay, decoupled, step, grad_scale=1.0) -> None:
    _lib.check(_lib.lib().gap_adam_flat(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), lr, beta1, beta2, eps,
                                        weight_decay, 1 if decoupled else 0, step, grad_scale, _stream()),
               "gap_adam_flat")


def adam_flat_devstep(p, g, m, v, lr, beta1, beta2, eps, weight_decay, decoupled, step_dev, grad_scale=1.0) -> None:
    """Adam / AdamW with the step counter in device memory (int32 tensor), for CUDA-graph replay."""
    _lib.check(_lib.lib().gap_adam_flat_devstep(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), lr, beta1, beta2, eps,
                                                weight_decay, 1 if decoupled else 0, _ptr(step_dev), grad_scale, _stream()),
               "gap_adam_flat_devstep")


def pack_weights(w: torch.Tensor, w_off: int, out: torch.Tensor, mode: int, n_phase: int, rows: int, rows_pad: int,
                 taps: tuple[int, int], c: int, c_pad: int, krow: int, strides: tuple[int, int, int, int],
                 kdim: int = 0) -> None:
    """``w`` is a flat fp32 buffer, ``w_off`` the element offset of this layer's weights in it."""
    src = C.c_void_p(w.data_ptr() + 4 * w_off)
    _lib.check(_lib.lib().gap_pack_weights(src, _ptr(out), mode, n_phase, rows, rows_pad, taps[0], taps[1], c, c_pad,
                                           krow, strides[0], strides[1], strides[2], strides[3], kdim, _stream()),
               "gap_pack_weights")


class PackPlan:
    """All fp32-master -> bf16-operand repacks of one network as ONE launch (gap_pack_weights_multi).

    ``add`` takes the arguments of :func:`pack_weights`; ``run`` launches.  Padding elements of the
    operands are never written, so the operand buffers must be zero-initialised by the owner."""

    def __init__(self) -> None:
        self.entries: list[_lib.PackEntry] = []
        self.total_tiles = 0
        self._table = None
        self._keep = []

    def add(self, w: torch.Tensor, w_off: int, out: torch.Tensor, mode: int, n_phase: int, rows: int, rows_pad: int,
            taps: tuple[int, int], c: int, c_pad: int, krow: int, strides: tuple[int, int, int, int],
            out_off: int = 0) -> None:
        """``out_off`` (elements) shifts the destination inside each operand row (channel-slot packing)."""
        if mode not in (0, 1, 2):
            raise ValueError("PackPlan supports modes 0-2")
        last = out_off + ((n_phase - 1) * rows_pad + rows - 1) * krow + (taps[0] * taps[1] - 1) * c_pad + c - 1
        if out.dtype != torch.bfloat16 or not out.is_contiguous() or last >= out.numel():
            raise ValueError("operand buffer is too small for this entry")
        e = _lib.PackEntry()
        e.w = w.data_ptr() + 4 * w_off
        e.out = out.data_ptr() + 2 * out_off
        e.mode, e.n_phase, e.rows, e.rows_pad = mode, n_phase, rows, rows_pad
        e.taps_h, e.taps_w, e.c, e.c_pad, e.krow = taps[0], taps[1], c, c_pad, krow
        e.tiles_r, e.tiles_c = (rows + 63) // 64, (c + 63) // 64
        e.tile_begin = self.total_tiles
        e.s_r, e.s_c, e.s_kh, e.s_kw = strides
        self.total_tiles += n_phase * taps[0] * taps[1] * e.tiles_r * e.tiles_c
        self.entries.append(e)
        self._keep.append((w, out))
        self._table = None

    def run(self) -> None:
        if not self.entries:
            return
        if self._table is None:
            arr = (_lib.PackEntry * len(self.entries))(*self.entries)
            raw = bytes(arr)
            dev = self._keep[0][1].device
            self._table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        _lib.check(_lib.lib().gap_pack_weights_multi(_ptr(self._table), len(self.entries), self.total_tiles, _stream()),
                   "gap_pack_weights_multi")


# ------------------------------------------------------------------------------------------------
# Siamese U-Net extras (gap_b200.h "Siamese U-Net extras")
# ------------------------------------------------------------------------------------------------
def _px(t: torch.Tensor) -> int:
    return t.numel() // t.shape[-1]


def im2col_k3s1p1_c3(x: torch.Tensor, col: torch.Tensor) -> None:
    n, h, w, _, ld = _nhwc_view(x)
    _lib.check(_lib.lib().gap_im2col_k3s1p1_c3(_ptr(x), ld, _ptr(col), n, h, w, _stream()), "gap_im2col_k3s1p1_c3")


def maxpool2x2_fwd(x: torch.Tensor, out: torch.Tensor) -> None:
    n, h, w, c, ld = _nhwc_view(x)
    _lib.check(_lib.lib().gap_maxpool2x2_fwd(_ptr(x), ld, _ptr(out), out.stride(2), n, h, w, c, _stream()),
               "gap_maxpool2x2_fwd")


def maxpool2x2_bwd(x: torch.Tensor, gout: torch.Tensor, gin: torch.Tensor, accumulate: bool) -> None:
    n, h, w, c, ld = _nhwc_view(x)
    _lib.check(_lib.lib().gap_maxpool2x2_bwd(_ptr(x), ld, _ptr(gout), gout.stride(2), _ptr(gin), gin.stride(2), n, h, w, c,
                                             1 if accumulate else 0, _stream()), "gap_maxpool2x2_bwd")


def upsample2x_fwd(x: torch.Tensor, out: torch.Tensor) -> None:
    n, h, w, c, ld = _nhwc_view(x)
    _lib.check(_lib.lib().gap_upsample_bilinear2x_fwd(_ptr(x), ld, _ptr(out), out.stride(2), n, h, w, c, _stream()),
               "gap_upsample_bilinear2x_fwd")


def upsample2x_bwd(gout: torch.Tensor, gin: torch.Tensor, accumulate: bool) -> None:
    n, h, w, c, ld = _nhwc_view(gin)
    _lib.check(_lib.lib().gap_upsample_bilinear2x_bwd(_ptr(gout), gout.stride(2), _ptr(gin), ld, n, h, w, c,
                                                      1 if accumulate else 0, _stream()), "gap_upsample_bilinear2x_bwd")


def dropout_(x: torch.Tensor, p_drop: float, seed: int, offset: int) -> None:
    """In-place nn.Dropout(p) with a regenerable mask (same (seed, offset) -> same mask; used on the gradient too)."""
    _lib.check(_lib.lib().gap_dropout_bf16(_ptr(x), x.stride(-2), _px(x), x.shape[-1], p_drop, seed & (2**64 - 1),
                                           offset & (2**64 - 1), _stream()), "gap_dropout_bf16")


def att_add_relu_fwd(yg, scg, shg, yx, scx, shx, s) -> None:
    _lib.check(_lib.lib().gap_att_add_relu_fwd(_ptr(yg), _ptr(scg), _ptr(shg), _ptr(yx), _ptr(scx), _ptr(shx), _ptr(s),
                                               _px(s), s.shape[-1], _stream()), "gap_att_add_relu_fwd")


def relu_bwd(s: torch.Tensor, gs: torch.Tensor, d: torch.Tensor) -> None:
    _lib.check(_lib.lib().gap_relu_bwd(_ptr(s), _ptr(gs), _ptr(d), s.numel(), _stream()), "gap_relu_bwd")


def lrelu_bwd(y: torch.Tensor, g: torch.Tensor, slope: float, d: torch.Tensor, accumulate: bool) -> None:
    """d (+)= (y > 0) ? g : slope*g on NHWC bf16 tensors / channel slices of the same shape."""
    _lib.check(_lib.lib().gap_lrelu_bwd_bf16(_ptr(y), y.stride(-2), _ptr(g), g.stride(-2), slope, _ptr(d), d.stride(-2),
                                             _px(y), y.shape[-1], 1 if accumulate else 0, _stream()), "gap_lrelu_bwd_bf16")


def att_gate_fwd(ypsi, scale, shift, psi, x, out) -> None:
    _lib.check(_lib.lib().gap_att_gate_fwd(_ptr(ypsi), _ptr(scale), _ptr(shift), _ptr(psi), _ptr(x), x.stride(-2), _ptr(out),
                                           out.stride(-2), _px(x), x.shape[-1], _stream()), "gap_att_gate_fwd")


def att_gate_bwd(gout, x, psi, gx, accumulate: bool, dz) -> None:
    _lib.check(_lib.lib().gap_att_gate_bwd(_ptr(gout), gout.stride(-2), _ptr(x), x.stride(-2), _ptr(psi), _ptr(gx),
                                           gx.stride(-2), 1 if accumulate else 0, _ptr(dz), _px(x), x.shape[-1], _stream()),
               "gap_att_gate_bwd")


def vec_stats(y: torch.Tensor, stats: torch.Tensor) -> None:
    _lib.check(_lib.lib().gap_vec_stats(_ptr(y), y.numel(), _ptr(stats), _stream()), "gap_vec_stats")


def vec_bn_bwd(y, dz, scale, mean, invstd, sums, dy) -> None:
    _lib.check(_lib.lib().gap_vec_bn_bwd(_ptr(y), _ptr(dz), y.numel(), _ptr(scale), _ptr(mean), _ptr(invstd), _ptr(sums),
                                         _ptr(dy), _stream()), "gap_vec_bn_bwd")


def conv1x1_cout1_fwd(x, w, bias, out) -> None:
    _lib.check(_lib.lib().gap_conv1x1_cout1_fwd(_ptr(x), x.stride(-2), _ptr(w), _ptr(bias), _ptr(out), _px(x), x.shape[-1],
                                                _stream()), "gap_conv1x1_cout1_fwd")


def conv1x1_cout1_dgrad(dl, w, gx) -> None:
    _lib.check(_lib.lib().gap_conv1x1_cout1_dgrad(_ptr(dl), _ptr(w), _ptr(gx), gx.stride(-2), _px(gx), gx.shape[-1],
                                                  _stream()), "gap_conv1x1_cout1_dgrad")


def conv1x1_cout1_wgrad(dl, x, dw, db) -> None:
    _lib.check(_lib.lib().gap_conv1x1_cout1_wgrad(_ptr(dl), _ptr(x), x.stride(-2), _px(x), x.shape[-1], _ptr(dw), _ptr(db),
                                                  _stream()), "gap_conv1x1_cout1_wgrad")


def seg_loss(logits: torch.Tensor, labels: torch.Tensor, mode: int, w_point: float, w_dice: float, pos_weight: float,
             smooth: float, gamma: float, focal_alpha: float, sums4: torch.Tensor, grad: Optional[torch.Tensor],
             grad_scale: float, loss: torch.Tensor) -> None:
    """CombinedLoss (mode 0) / FocalDiceLoss (mode 1) of train.py:82-128 with the gradient w.r.t. the logits."""
    if logits.dtype != torch.float32 or labels.dtype != torch.int64 or logits.numel() != labels.numel():
        raise ValueError("logits must be fp32 and labels int64 with the same number of elements")
    _lib.check(_lib.lib().gap_seg_loss(_ptr(logits), _ptr(labels), logits.numel(), mode, w_point, w_dice, pos_weight, smooth,
                                       gamma, focal_alpha, _ptr(sums4), _ptr(grad), grad_scale, _ptr(loss), _stream()),
               "gap_seg_loss")


def seg_confusion(logits: torch.Tensor, labels: torch.Tensor, counts: torch.Tensor) -> None:
    """counts[n, 4] (int64, accumulated) += per-sample [TP, FP, FN, TN] of sigmoid(logits) > 0.5 vs {0,1} labels
    (evaluate.py:34-46).  logits fp32 [n, ...]; labels int64 or fp32 with the same number of elements."""
    n = logits.shape[0]
    if logits.dtype != torch.float32 or not logits.is_contiguous() or not labels.is_contiguous():
        raise ValueError("logits must be contiguous fp32, labels contiguous")
    if labels.dtype not in (torch.int64, torch.float32) or labels.numel() != logits.numel():
        raise ValueError("labels must be int64 or fp32 with as many elements as logits")
    if counts.dtype != torch.int64 or tuple(counts.shape) != (n, 4) or not counts.is_contiguous():
        raise ValueError("counts must be a contiguous int64 [n, 4] tensor")
    if not (logits.is_cuda and labels.is_cuda and counts.is_cuda):
        raise ValueError("seg_confusion needs CUDA tensors (there is no CPU path)")
    _lib.check(_lib.lib().gap_seg_confusion(_ptr(logits), _ptr(labels), 1 if labels.dtype == torch.int64 else 0, n,
                                            logits.numel() // n, _ptr(counts), _stream()), "gap_seg_confusion")
