/*
 * gap_b200.h — C ABI of the B200-native conv hot path for GAN-AUG-PFA (Pix2Pix U-Net generator,
 * PatchGAN discriminator, Siamese U-Net reuse).
 *
 * The reference (Affi-Amine/GAN-AUG-PFA) is pure Python and defines no FFI of its own: its hot path
 * is whatever torch dispatches for the nn.* modules built in models.py.  Every entry point below
 * therefore cites the reference call site whose torch library kernel it replaces (file:line into
 * the reference tree).  INTEGRATION.md shows the ctypes stub a maintainer adds to models.py.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says host; the caller owns every buffer
 *     (including workspaces); the library never allocates or frees device memory;
 *   - activations are NHWC bf16 (pixel stride `ld` in elements, so a tensor may live in a channel
 *     slot of a wider concat buffer); weights are packed bf16 K-major matrices (gap_pack_weights);
 *     gradients of weights, BatchNorm statistics, losses and optimizer state are fp32/fp64;
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising;
 *   - return value: 0 = ok, negative = gap_status (bad argument / unsupported configuration),
 *     positive = cudaError_t.  gap_last_error_string() describes the last failure of the calling
 *     thread.  There is no CPU fallback and no cuDNN/cuBLAS fallback.
 */
#ifndef GAP_B200_H
#define GAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum gap_status {
  GAP_OK = 0,
  GAP_ERR_BAD_ARG = -1,
  GAP_ERR_UNSUPPORTED = -2,
  GAP_ERR_ALIGNMENT = -3,
  GAP_ERR_DRIVER = -4
};

enum gap_act {
  GAP_ACT_NONE = 0,
  GAP_ACT_LRELU = 1, /* LeakyReLU(0.2)   models.py:178,223,232,240 */
  GAP_ACT_RELU = 2,  /* ReLU             models.py:180,11,14,37   */
  GAP_ACT_TANH = 3,  /* Tanh             models.py:186             */
  GAP_ACT_SIGMOID = 4 /* Sigmoid          models.py:34              */
};

const char* gap_last_error_string(void);
int gap_version(void);
/* Number of SMs the persistent kernels size their grids for (148 on B200). */
int gap_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution engine (tcgen05 / TMEM / TMA).
 *
 * One call computes, for every phase p = (ph, pw), every image n and every grid point (gy, gx):
 *
 *   acc[co] = sum over taps (th, tw), sources s, channels c of
 *             src_s[n, gy*in_stride + in_off_h[ph] + th, gx*in_stride + in_off_w[pw] + tw, c]
 *             * wpk[p][co][(th*taps_w + tw)*ctot + chan_off(s) + c]
 *   v = acc[co] + bias[co];  stats[co] += v, stats[n_out+co] += v*v   (fp64, optional)
 *   out [n, gy*out_stride + ph, gx*out_stride + pw, co] = act (v)   (bf16)
 *   out2[same pixel, co]                                 = act2(v)   (bf16, optional)
 *
 * Out-of-range source coordinates read as zero (that is the convolution padding).  With the right
 * taps / offsets / packed weights this one contraction is Conv2d forward (models.py:177,223,230,
 * 238,243,9,12,22,27,32,90), Conv2d dgrad, ConvTranspose2d forward (models.py:184,189,194) and
 * ConvTranspose2d dgrad; the named wrappers further down fill this struct.
 * ---------------------------------------------------------------------------------------------- */
typedef struct gap_conv_gemm_args {
  /* A operand: one or two NHWC bf16 sources concatenated along channels (torch.cat, models.py:208) */
  const void* src[2];
  int src_c[2];      /* channels taken from each source; multiple of 64; src_c[1] = 0 if unused */
  int64_t src_ld[2]; /* pixel stride in elements (>= src_c, multiple of 8) */
  int n, ih, iw;     /* source images / rows / cols */
  /* M iteration space */
  int gh, gw;    /* grid rows / cols per image per phase */
  int n_phase;   /* 1, or 4 for the stride-2 transposed geometry */
  int taps_h, taps_w;
  int in_stride;
  int in_off_h[2], in_off_w[2];
  int out_stride;
  /* B operand: packed weights, bf16, [n_phase][w_rows][taps_h*taps_w*ctot] */
  const void* wpk;
  int w_rows; /* rows present in wpk per phase (>= n_out) */
  /* epilogue */
  int n_out; /* valid output channels */
  int oh, ow; /* full output rows / cols */
  void* out;
  int64_t out_ld;
  int act;
  void* out2; /* may be NULL */
  int64_t out2_ld;
  int act2;
  const float* bias; /* may be NULL */
  double* stats;     /* may be NULL; [2*n_out] */
} gap_conv_gemm_args;

int gap_conv_gemm(const gap_conv_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight-gradient contraction (tcgen05, MN-major operands, split-K with fp32 red.add):
 *
 *   out[m*ld_m + tap*ld_tap + c] += sum over images and grid points (gy, gx) of
 *        mop[img, gy, gx, m] * nop[img, gy*stride + off_h + th, gx*stride + off_w + tw, c]
 *
 * Conv2d wgrad (autograd of models.py:177,223,230,238,243,9,12,...): mop = dY, nop = X,
 * out = dW viewed as [Cout][kh*kw][Cin].  ConvTranspose2d wgrad (models.py:184,189,194):
 * mop = X, nop = dY, out = dW viewed as [Cin][kh*kw][Cout].  The caller zeroes `out` (that is
 * optimizer.zero_grad(), train_gan.py:55,64); concatenated inputs are handled by one call per
 * source with `out` offset to the source's channel range.
 * ---------------------------------------------------------------------------------------------- */
typedef struct gap_wgrad_args {
  const void* mop; /* NHWC bf16 [n, gh, gw, m_c] */
  int m_c;
  int64_t m_ld;
  const void* nop; /* NHWC bf16 [n, nh, nw, n_c] */
  int n_c;
  int64_t n_ld;
  int n, gh, gw, nh, nw;
  int taps_h, taps_w, stride, off_h, off_w;
  float* out;
  int64_t ld_m, ld_tap;
} gap_wgrad_args;

int gap_conv_wgrad(const gap_wgrad_args* args, void* stream);

/* Debug knobs for bring-up (descriptor conventions); not part of the stable surface. */
int gap_debug_set(const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* GAP_B200_H */
