// Direct kernels for the HBM-bound "thin" layers of the PatchGAN discriminator and the U-Net
// generator: layers whose input or output has 1-6 channels move hundreds of MB for a few GFLOP, so
// they are fused, coalesced, warp-MMA (mma.sync) kernels that touch each activation once instead of
// im2col / channel-padded tcgen05 GEMMs.
//
//   gap_cout1_conv_{fwd,dgrad,wgrad}: Conv2d(C -> 1, k4, s1, p1) + bias, the discriminator's last layer
//   (models.py:243) and its two gradients.
#include "common.h"
#include "mma_sync.cuh"
#include "ptx.cuh"

namespace gap {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// ------------------------------------------------------------------------------------------------
// Cout = 1 forward, stage 1: z[pix][tap] = sum_c x[pix][c] * w[tap][c] for the 16 taps of a 4x4
// filter — every input element is read once (the 16 shifted sums are formed in stage 2 from z).
// CTA = 128 pixels, warp = 16 pixels x 16 taps, K = channels in chunks of 64.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cout1_z_kernel(const bf16* __restrict__ x, long long ld_x, long long npix,
                                                      int c, const bf16* __restrict__ w, float* __restrict__ z) {
  __shared__ __align__(16) bf16 xs[128][72];
  __shared__ __align__(16) bf16 ws[16][72];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t xs_a = smem_u32(&xs[0][0]), ws_a = smem_u32(&ws[0][0]);
  for (long long tile = blockIdx.x; tile * 128 < npix; tile += gridDim.x) {
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    for (int c0 = 0; c0 < c; c0 += 64) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int v = tid + i * 256, px = v >> 3, seg = v & 7;
        const long long pix = tile * 128 + px;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (pix < npix) val = ldg128(x + pix * ld_x + c0 + seg * 8);
        *reinterpret_cast<uint4*>(&xs[px][seg * 8]) = val;
      }
      if (tid < 128) {
        const int t = tid >> 3, seg = tid & 7;
        *reinterpret_cast<uint4*>(&ws[t][seg * 8]) = ldg128(w + static_cast<long long>(t) * c + c0 + seg * 8);
      }
      __syncthreads();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t a[4], b[4];
        lda_16x16(a, xs_a + warp * 16 * 144 + ks * 32, 144, lane);
        ldb_16x16(b, ws_a + ks * 32, 144, lane);
        mma_bf16_16816(acc[0], a, b[0], b[1]);
        mma_bf16_16816(acc[1], a, b[2], b[3]);
      }
      __syncthreads();
    }
    const int g = lane >> 2, t = lane & 3;
    const long long row0 = tile * 128 + warp * 16 + g, row1 = row0 + 8;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      if (row0 < npix) *reinterpret_cast<float2*>(z + row0 * 16 + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
      if (row1 < npix) *reinterpret_cast<float2*>(z + row1 * 16 + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
    }
  }
}

// stage 2: logits[n][oy][ox] = bias + sum_{kh,kw} z[n][oy+kh-pad][ox+kw-pad][kh*4+kw]
__global__ void cout1_gather_kernel(const float* __restrict__ z, const float* __restrict__ bias, int n, int ih, int iw,
                                    int oh, int ow, int pad, float* __restrict__ logits) {
  const long long total = static_cast<long long>(n) * oh * ow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = static_cast<int>(i % ow);
    const int oy = static_cast<int>((i / ow) % oh);
    const long long img = i / (static_cast<long long>(ow) * oh);
    float s = bias ? __ldg(bias) : 0.f;
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const int iy = oy + kh - pad;
      if (iy < 0 || iy >= ih) continue;
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int ix = ox + kw - pad;
        if (ix < 0 || ix >= iw) continue;
        s += __ldg(z + ((img * ih + iy) * iw + ix) * 16 + kh * 4 + kw);
      }
    }
    logits[i] = s;
  }
}

// u[pix][tap] = dlogits[n][y - kh + pad][x - kw + pad] (zero outside): the column of output gradients
// that input pixel (n, y, x) sees through tap (kh, kw).
__device__ __forceinline__ float gather_dlogit(const float* __restrict__ dlog, long long pix, long long npix, int tap,
                                               int ih, int iw, int oh, int ow, int pad) {
  if (pix >= npix) return 0.f;
  const int x = static_cast<int>(pix % iw);
  const int y = static_cast<int>((pix / iw) % ih);
  const long long img = pix / (static_cast<long long>(iw) * ih);
  const int oy = y - (tap >> 2) + pad, ox = x - (tap & 3) + pad;
  if (oy < 0 || oy >= oh || ox < 0 || ox >= ow) return 0.f;
  return __ldg(dlog + (img * oh + oy) * ow + ox);
}

// ------------------------------------------------------------------------------------------------
// Cout = 1 dgrad: gx[pix][c] = sum_tap u[pix][tap] * w[tap][c]   (one k16 MMA step per output tile)
// CTA = 128 input pixels; warp = one 64-channel range, its 8 B fragments held in registers.
// dynamic smem: wT[c][24] | us[128][24] | out_s[8][16][72]   (bf16)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cout1_dgrad_kernel(const float* __restrict__ dlog, int ih, int iw, int oh, int ow,
                                                          int pad, long long npix, const bf16* __restrict__ w, int c,
                                                          bf16* __restrict__ gx, long long ld_gx) {
  extern __shared__ __align__(16) uint8_t dsm[];
  bf16* wT = reinterpret_cast<bf16*>(dsm);
  bf16* us = wT + static_cast<size_t>(c) * 24;
  bf16* out_s = us + 128 * 24;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int idx = tid; idx < c * 16; idx += 256) {
    const int tap = idx / c, cc = idx - tap * c;
    wT[cc * 24 + tap] = w[idx];
  }
  const uint32_t us_a = smem_u32(us);
  bf16* my_out = out_s + warp * 16 * 72;
  for (long long tile = blockIdx.x; tile * 128 < npix; tile += gridDim.x) {
    __syncthreads();  // wT ready (first pass) / previous tile's us consumed
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = tid + i * 256, px = idx >> 4, tap = idx & 15;
      us[px * 24 + tap] = __float2bfloat16(gather_dlogit(dlog, tile * 128 + px, npix, tap, ih, iw, oh, ow, pad));
    }
    __syncthreads();
    for (int nr = warp; nr < (c >> 6); nr += 8) {
      const int n0 = nr * 64;
      uint32_t bfr[8][2];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const bf16* row = wT + (n0 + nt * 8 + g) * 24;
        bfr[nt][0] = *reinterpret_cast<const uint32_t*>(row + 2 * t);
        bfr[nt][1] = *reinterpret_cast<const uint32_t*>(row + 2 * t + 8);
      }
      for (int mt = 0; mt < 8; ++mt) {
        uint32_t a[4];
        lda_16x16(a, us_a + mt * 16 * 48, 48, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          mma_bf16_16816(acc, a, bfr[nt][0], bfr[nt][1]);
          *reinterpret_cast<uint32_t*>(my_out + g * 72 + nt * 8 + 2 * t) = pack_bf16x2(acc[0], acc[1]);
          *reinterpret_cast<uint32_t*>(my_out + (g + 8) * 72 + nt * 8 + 2 * t) = pack_bf16x2(acc[2], acc[3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int idx = lane + i * 32, r = idx >> 3, seg = idx & 7;
          const long long pix = tile * 128 + mt * 16 + r;
          if (pix < npix)
            *reinterpret_cast<uint4*>(gx + pix * ld_gx + n0 + seg * 8) = *reinterpret_cast<const uint4*>(my_out + r * 72 + seg * 8);
        }
        __syncwarp();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cout = 1 wgrad: dw[tap][c] += sum_pix u[pix][tap] * x[pix][c]   (K = pixels)
// CTA = a contiguous range of input pixels; warp = one 64-channel range (c <= 512).
// dynamic smem: xs[16][c + 8] | ut[16][24]   (bf16)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cout1_wgrad_kernel(const float* __restrict__ dlog, int ih, int iw, int oh, int ow,
                                                          int pad, long long npix, long long chunk,
                                                          const bf16* __restrict__ x, long long ld_x, int c,
                                                          float* __restrict__ dw) {
  extern __shared__ __align__(16) uint8_t dsm[];
  bf16* xs = reinterpret_cast<bf16*>(dsm);
  const int xstride = c + 8;
  bf16* ut = xs + 16 * xstride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const long long p0 = blockIdx.x * chunk, p1 = min(npix, p0 + chunk);
  const uint32_t xs_a = smem_u32(xs), ut_a = smem_u32(ut);
  const int n0 = warp * 64;
  const bool active = n0 < c;
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[nt][j] = 0.f;
  const int vec_per_px = c >> 3;
  for (long long k0 = p0; k0 < p1; k0 += 16) {
    for (int idx = tid; idx < 16 * vec_per_px; idx += 256) {
      const int px = idx / vec_per_px, seg = idx - px * vec_per_px;
      const long long pix = k0 + px;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (pix < p1) val = ldg128(x + pix * ld_x + seg * 8);
      *reinterpret_cast<uint4*>(xs + px * xstride + seg * 8) = val;
    }
    {
      const int tap = tid >> 4, px = tid & 15;
      const long long pix = k0 + px;
      ut[tap * 24 + px] = __float2bfloat16(pix < p1 ? gather_dlogit(dlog, pix, npix, tap, ih, iw, oh, ow, pad) : 0.f);
    }
    __syncthreads();
    if (active) {
      uint32_t a[4];
      lda_16x16(a, ut_a, 48, lane);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, xs_a + ((lane & 15) * xstride + n0 + nt * 8) * 2);
        mma_bf16_16816(acc[nt], a, b0, b1);
      }
    }
    __syncthreads();
  }
  if (active && p1 > p0) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = n0 + nt * 8 + 2 * t;
      atomicAdd(dw + static_cast<long long>(g) * c + col, acc[nt][0]);
      atomicAdd(dw + static_cast<long long>(g) * c + col + 1, acc[nt][1]);
      atomicAdd(dw + static_cast<long long>(g + 8) * c + col, acc[nt][2]);
      atomicAdd(dw + static_cast<long long>(g + 8) * c + col + 1, acc[nt][3]);
    }
  }
}

// BCE-with-logits against a constant target with an fp32 gradient and the bias gradient of the
// producing Cout = 1 conv:  dlogits = grad_scale*(sigmoid(x) - t);  dbias += sum dlogits
__global__ void bce_logits_const_f32_kernel(const float* __restrict__ x, long long count, float t, float grad_scale,
                                            float* __restrict__ dx, double* __restrict__ loss_acc,
                                            float* __restrict__ dbias) {
  float part = 0.f, gsum = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    part += fmaxf(v, 0.f) - v * t + log1pf(__expf(-fabsf(v)));
    const float d = grad_scale * (1.f / (1.f + __expf(-v)) - t);
    if (dx) dx[i] = d;
    gsum += d;
  }
  part = warp_sum(part);
  gsum = warp_sum(gsum);
  __shared__ float red[2][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    red[0][wid] = part;
    red[1][wid] = gsum;
  }
  __syncthreads();
  if (wid == 0) {
    float v = lane < (blockDim.x >> 5) ? red[0][lane] : 0.f;
    float gg = lane < (blockDim.x >> 5) ? red[1][lane] : 0.f;
    v = warp_sum(v);
    gg = warp_sum(gg);
    if (lane == 0) {
      atomicAdd(loss_acc, static_cast<double>(v));
      if (dbias) atomicAdd(dbias, gg);
    }
  }
}

__global__ void sum_f32_kernel(const float* __restrict__ x, long long count, float* __restrict__ out) {
  float part = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    part += x[i];
  part = warp_sum(part);
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = part;
  __syncthreads();
  if (wid == 0) {
    float v = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(out, v);
  }
}

}  // namespace gap

using namespace gap;

extern "C" {

int gap_cout1_conv_fwd(const void* x, int64_t ld_x, int n, int ih, int iw, int c, const void* w, const float* bias,
                       int ksize, int pad, float* z_ws, float* logits, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GAP_CHECK_ARG(x && w && z_ws && logits && n > 0 && ih > 0 && iw > 0, "gap_cout1_conv_fwd: bad arguments");
  if (ksize != 4 || c % 64 != 0 || c <= 0 || ld_x % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(w) & 15)) {
    set_error("gap_cout1_conv_fwd: needs ksize 4, channels %% 64 == 0 and 16-byte aligned rows (c=%d ld=%lld)", c,
              (long long)ld_x);
    return GAP_ERR_UNSUPPORTED;
  }
  const int oh = ih + 2 * pad - 3, ow = iw + 2 * pad - 3;
  GAP_CHECK_ARG(oh > 0 && ow > 0, "gap_cout1_conv_fwd: empty output");
  const long long npix = static_cast<long long>(n) * ih * iw;
  const int tiles = static_cast<int>((npix + 127) / 128);
  cout1_z_kernel<<<std::min(tiles, 8 * sm_count()), 256, 0, st>>>(static_cast<const bf16*>(x), ld_x, npix, c,
                                                                   static_cast<const bf16*>(w), z_ws);
  GAP_CUDA(cudaGetLastError());
  const long long total = static_cast<long long>(n) * oh * ow;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 8LL * sm_count()));
  cout1_gather_kernel<<<blocks, 256, 0, st>>>(z_ws, bias, n, ih, iw, oh, ow, pad, logits);
  GAP_CUDA(cudaGetLastError());
  return 0;
}

int gap_cout1_conv_dgrad(const float* dlogits, int n, int oh, int ow, const void* w, int ksize, int pad, int c, void* gx,
                         int64_t ld_gx, int ih, int iw, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GAP_CHECK_ARG(dlogits && w && gx && n > 0 && oh > 0 && ow > 0 && ih > 0 && iw > 0, "gap_cout1_conv_dgrad: bad arguments");
  if (ksize != 4 || c % 64 != 0 || c <= 0 || ld_gx % 8 != 0 || (reinterpret_cast<uintptr_t>(gx) & 15)) {
    set_error("gap_cout1_conv_dgrad: needs ksize 4, channels %% 64 == 0 and 16-byte aligned rows");
    return GAP_ERR_UNSUPPORTED;
  }
  const long long npix = static_cast<long long>(n) * ih * iw;
  const size_t smem = (static_cast<size_t>(c) * 24 + 128 * 24 + 8 * 16 * 72) * sizeof(bf16);
  if (smem > 200 * 1024) {
    set_error("gap_cout1_conv_dgrad: %d channels do not fit shared memory", c);
    return GAP_ERR_UNSUPPORTED;
  }
  static size_t attr = 0;
  if (smem > attr) {
    GAP_CUDA(cudaFuncSetAttribute(cout1_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  const int tiles = static_cast<int>((npix + 127) / 128);
  cout1_dgrad_kernel<<<std::min(tiles, 2 * sm_count()), 256, smem, st>>>(dlogits, ih, iw, oh, ow, pad, npix,
                                                                         static_cast<const bf16*>(w), c,
                                                                         static_cast<bf16*>(gx), ld_gx);
  GAP_CUDA(cudaGetLastError());
  return 0;
}

int gap_cout1_conv_wgrad(const float* dlogits, int n, int oh, int ow, const void* x, int64_t ld_x, int ih, int iw, int c,
                         int ksize, int pad, float* dw, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GAP_CHECK_ARG(dlogits && x && dw && n > 0 && oh > 0 && ow > 0 && ih > 0 && iw > 0, "gap_cout1_conv_wgrad: bad arguments");
  if (ksize != 4 || c % 64 != 0 || c <= 0 || c > 512 || ld_x % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15)) {
    set_error("gap_cout1_conv_wgrad: needs ksize 4, channels %% 64 == 0, <= 512, 16-byte aligned rows");
    return GAP_ERR_UNSUPPORTED;
  }
  const long long npix = static_cast<long long>(n) * ih * iw;
  const size_t smem = (static_cast<size_t>(16) * (c + 8) + 16 * 24) * sizeof(bf16);
  const int want = 4 * sm_count();
  long long chunk = (npix + want - 1) / want;
  chunk = (chunk + 15) / 16 * 16;
  const int grid = static_cast<int>((npix + chunk - 1) / chunk);
  cout1_wgrad_kernel<<<grid, 256, smem, st>>>(dlogits, ih, iw, oh, ow, pad, npix, chunk, static_cast<const bf16*>(x),
                                              ld_x, c, dw);
  GAP_CUDA(cudaGetLastError());
  return 0;
}

int gap_bce_logits_const_f32(const float* logits, int64_t count, float target, float grad_scale, float* dlogits,
                             double* loss_acc, float* dbias, void* stream) {
  GAP_CHECK_ARG(logits && loss_acc && count > 0, "gap_bce_logits_const_f32: bad arguments");
  const int blocks = static_cast<int>(std::min<int64_t>((count + 255) / 256, 148));
  bce_logits_const_f32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, count, target, grad_scale,
                                                                                     dlogits, loss_acc, dbias);
  GAP_CUDA(cudaGetLastError());
  return 0;
}

int gap_sum_f32(const float* x, int64_t count, float* out, void* stream) {
  GAP_CHECK_ARG(x && out && count > 0, "gap_sum_f32: bad arguments");
  const int blocks = static_cast<int>(std::min<int64_t>((count + 255) / 256, 148));
  sum_f32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, count, out);
  GAP_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
