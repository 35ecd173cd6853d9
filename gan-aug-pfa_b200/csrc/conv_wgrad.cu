// Weight-gradient contraction for sm_100a (tcgen05 / TMEM / TMA), shared by Conv2d and
// ConvTranspose2d:
//
//   dW[m][tap][n] += sum over pixels (img, gy, gx) of
//        Mop[img, gy, gx, m] * Nop[img, gy*stride + off_h + th, gx*stride + off_w + tw, n]
//
// Conv2d wgrad:          Mop = dY (output grid, Cout),  Nop = X gathered per tap (Cin)
// ConvTranspose2d wgrad: Mop = X  (input grid, Cin),    Nop = dY gathered per tap (Cout)
//
// The reduction dimension is the pixel index, so both operands sit in shared memory "MN-major":
// each 64-pixel K tile is fetched by 4-D TMA boxes of 64 channels x 64 pixels (128-byte swizzle)
// and handed to tcgen05.mma through MN-major matrix descriptors — no transpose pass.  One CTA owns
// a 128-channel M tile, a BN-channel N tile, a group of taps (each tap has its own TMEM
// accumulator columns, so the M-operand tile is fetched once per tap group) and a slice of the
// pixel range (split-K); partial sums are added to the fp32 gradient with red.global.add.f32.
#include "common.h"
#include "ptx.cuh"

namespace gap {

constexpr int kWgThreads = 224;   // warps: 0 TMA, 1 MMA, 2..5 epilogue, 6 second TMA lane
constexpr int kWgProducer2Warp = 6;
constexpr int kWgBlockK = 64;             // pixels per K tile
constexpr int kWgBoxBytes = 64 * 64 * 2;  // one 64ch x 64px TMA box
constexpr int kWgMaxStages = 6;
constexpr int kWgSmemBudget = 227 * 1024;

struct alignas(64) WgradParams {
  CUtensorMap tmM;
  CUtensorMap tmN;
  int m_c, n_c, m_rows;
  int n_img, gh, gw;
  int taps_w, n_taps, stride, off_h, off_w;
  int log_bw, log_bh;
  int tiles_w, tiles_h, tiles_n, pix_tiles;
  int m_tiles, n_tiles, tap_groups, splits;
  int block_n, tpc;
  int mt;  // 128-row M tiles per CTA (1 or 2): they share every N-operand box
  int num_stages, tmem_cols;
  uint32_t idesc;
  int dual;   // two TMA producer lanes share each stage (see the producer loop)
  float* out;
  long long ld_m, ld_tap;
  long long* trace;  // bring-up: clock64 samples of CTA 0 ([0..255] producer, [256..511] mma, [512..] epilogue)
  int vec_red;  // 1: 16-byte aligned rows and n_c % 16 == 0 -> red.global.add.v4.f32
  int skip;  // bring-up ablation mask: 1 no atomics, 2 no MMA, 4 no N-operand TMA, 8 no M-operand TMA
};

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int a_bytes = p.mt * 2 * kWgBoxBytes;
  const int b_tap_bytes = (p.block_n / 64) * kWgBoxBytes;
  const int stage_bytes = a_bytes + p.tpc * b_tap_bytes;
  const uint32_t bar_base = smem_base + p.num_stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWgMaxStages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kWgMaxStages);
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem_gen + p.num_stages * stage_bytes + 8 * (2 * kWgMaxStages + 1));

  // Warp index broadcast from lane 0 so the compiler knows the role dispatch is warp-uniform: the
  // producer / MMA loops below are executed by all 32 lanes and only the asynchronous instructions
  // sit under elect_one(), which keeps descriptors, coordinates and barrier addresses in uniform
  // registers (a per-lane `if (lane == 0)` loop makes ptxas wrap every UTMALDG / UTCHMMA in a
  // lane-serialising loop with R2UR moves: ~150 / ~75 cycles per issue, measured).
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const uint32_t lane = threadIdx.x & 31;

  // work item
  int w = blockIdx.x;
  const int m_tile = w % p.m_tiles;
  w /= p.m_tiles;
  const int n_tile = w % p.n_tiles;
  w /= p.n_tiles;
  const int tap_group = w % p.tap_groups;
  const int split = w / p.tap_groups;
  const int tap0 = tap_group * p.tpc;
  const int ntap = min(p.tpc, p.n_taps - tap0);
  const int pt0 = static_cast<int>(static_cast<long long>(p.pix_tiles) * split / p.splits);
  const int pt1 = static_cast<int>(static_cast<long long>(p.pix_tiles) * (split + 1) / p.splits);
  const bool tracing = p.trace != nullptr && blockIdx.x == 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmM);
    tma_prefetch_desc(&p.tmN);
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(full_bar(s), p.dual ? 2 : 1);   // with two producer lanes both post their bytes
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();      // the prologue above overlapped the previous kernel in the stream

  const int BW = 1 << p.log_bw, BH = 1 << p.log_bh;
  const int BNI = kWgBlockK >> (p.log_bw + p.log_bh);
  const int m_base = m_tile * 128 * p.mt;  // first M channel of this CTA
  const int m_boxes = (p.skip & 8) ? 0 : min(2 * p.mt, (p.m_c - m_base + 63) / 64);
  const int n_boxes = (p.skip & 4) ? 0 : min(p.block_n / 64, (p.n_c - n_tile * p.block_n + 63) / 64);

  if (warp == 0 || warp == kWgProducer2Warp) {
    // ------------------------------------------------------------------ TMA producers
    // Two lanes (one per warp) share every stage: issuing one stage costs a single lane ~280 cycles + ~50 per TMA (up
    // to 8 boxes here), more than the 512 MMA cycles a 128 x 256 stage lasts.  Lane 0 issues the M boxes and the first
    // taps, lane 1 the remaining taps; each posts its own byte count on the stage's barrier (2 arrivals).
    // (Alternating whole stages between the lanes was slower on the 3-stage configurations: two stages' boxes then
    // interleave in the TMA queue and both complete late.)
    // Measured (tools/sweep_wgrad.py, wgrad_dual=0/1 in one process): +15 % on the 128 x (4 taps x 64) stage of the
    // 64->128 layers, neutral elsewhere.
    const bool second = warp == kWgProducer2Warp;
    if ((!second || p.dual) && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int t_split = !p.dual ? ntap : (ntap > 1 ? ntap / 2 : 0);   // lane 0: taps [0, t_split), lane 1: the rest
      const int t_lo = second ? t_split : 0, t_hi = second ? ntap : t_split;
      const int my_m = second ? 0 : m_boxes;
      const uint32_t tx = static_cast<uint32_t>((my_m + (t_hi - t_lo) * n_boxes) * kWgBoxBytes);
      int tw = pt0 % p.tiles_w;
      int th = (pt0 / p.tiles_w) % p.tiles_h;
      int tn = pt0 / (p.tiles_w * p.tiles_h);
      const int tap_lo = tap0 + t_lo;
      const int t_h0 = tap_lo / p.taps_w, t_w0 = tap_lo - t_h0 * p.taps_w;
      for (int pt = pt0; pt < pt1; ++pt) {
        const int gx0 = tw * BW, gy0 = th * BH, n0 = tn * BNI;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (tracing && !second && pt - pt0 < 128) p.trace[2 * (pt - pt0)] = clock64();
        mbar_expect_tx(full_bar(stage), tx);
        const uint32_t a_dst = smem_base + stage * stage_bytes;
        for (int b = 0; b < my_m; ++b)
          tma_load_4d(a_dst + b * kWgBoxBytes, &p.tmM, full_bar(stage), m_base + b * 64, gx0, gy0, n0);
        int t_h = t_h0, t_w = t_w0;
        for (int t_i = t_lo; t_i < t_hi; ++t_i) {
          const uint32_t b_dst = a_dst + a_bytes + t_i * b_tap_bytes;
          for (int b = 0; b < n_boxes; ++b)
            tma_load_4d(b_dst + b * kWgBoxBytes, &p.tmN, full_bar(stage), n_tile * p.block_n + b * 64,
                        gx0 * p.stride + p.off_w + t_w, gy0 * p.stride + p.off_h + t_h, n0);
          if (++t_w == p.taps_w) {
            t_w = 0;
            ++t_h;
          }
        }
        if (tracing && !second && pt - pt0 < 128) p.trace[2 * (pt - pt0) + 1] = clock64();
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) {
            th = 0;
            ++tn;
          }
        }
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int ntap_mma = (p.skip & 2) ? 0 : ntap;
      const int mt_eff = min(p.mt, (p.m_c - m_base + 127) / 128);
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (tracing && pt - pt0 < 128) p.trace[256 + 2 * (pt - pt0)] = clock64();
        const uint32_t a_addr = smem_base + stage * stage_bytes;
        for (int j = 0; j < mt_eff; ++j) {
          for (int t_i = 0; t_i < ntap_mma; ++t_i) {
            const uint32_t b_addr = a_addr + a_bytes + t_i * b_tap_bytes;
#pragma unroll
            for (int k = 0; k < kWgBlockK / 16; ++k) {
              // MN-major SW128: 8 pixel rows (1024 B) per K atom -> a K=16 step advances 2048 B;
              // 64-channel groups are one box (8192 B) apart.
              const uint64_t adesc = make_sw128_desc(a_addr + j * 2 * kWgBoxBytes + k * 2048, kWgBoxBytes, 1024);
              const uint64_t bdesc = make_sw128_desc(b_addr + k * 2048, kWgBoxBytes, 1024);
              umma_bf16(tmem_base + (j * p.tpc + t_i) * p.block_n, adesc, bdesc, p.idesc,
                        (pt > pt0 || k > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(stage));
        if (tracing && pt - pt0 < 128) p.trace[256 + 2 * (pt - pt0) + 1] = clock64();
        if (++stage == p.num_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(done_bar);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (pt1 > pt0) {
      if (tracing && threadIdx.x == 64) p.trace[512] = clock64();
      mbar_wait(done_bar, 0);
      tc_fence_after();
      if (tracing && threadIdx.x == 64) p.trace[513] = clock64();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      for (int j = 0; j < p.mt; ++j) {
        const int m = m_base + j * 128 + row;
        if (m_base + j * 128 >= p.m_c) break;
        const bool row_ok = m < p.m_rows && !(p.skip & 1);
        for (int t_i = 0; t_i < ntap; ++t_i) {
          float* dst_row = p.out + static_cast<long long>(m) * p.ld_m + (tap0 + t_i) * p.ld_tap;
          for (int c = 0; c < p.block_n / 16; ++c) {
            uint32_t raw[16];
            tmem_ld16(t_row + (j * p.tpc + t_i) * p.block_n + c * 16, raw);
            tmem_ld_wait();
            const int col0 = n_tile * p.block_n + c * 16;
            if (row_ok && col0 < p.n_c) {
              if (p.vec_red) {
#pragma unroll
                for (int jj = 0; jj < 16; jj += 4)
                  red_add_v4_f32(dst_row + col0 + jj, __uint_as_float(raw[jj]), __uint_as_float(raw[jj + 1]),
                                 __uint_as_float(raw[jj + 2]), __uint_as_float(raw[jj + 3]));
              } else {
#pragma unroll
                for (int jj = 0; jj < 16; ++jj)
                  if (col0 + jj < p.n_c) atomicAdd(dst_row + col0 + jj, __uint_as_float(raw[jj]));
              }
            }
          }
        }
      }
    }
  }

  if (p.trace && blockIdx.x == 0 && threadIdx.x == 64) p.trace[514] = clock64();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static int wg_ilog2_ceil(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

}  // namespace gap

using namespace gap;

extern "C" int gap_conv_wgrad(const gap_wgrad_args* a, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  GAP_CHECK_ARG(a && a->mop && a->nop && a->out, "gap_conv_wgrad: null pointer");
  GAP_CHECK_ARG(a->m_c > 0 && a->m_c % 64 == 0 && a->n_c > 0 && a->n_c % 64 == 0,
                "gap_conv_wgrad: channel counts (%d, %d) must be multiples of 64", a->m_c, a->n_c);
  GAP_CHECK_ARG(a->n > 0 && a->gh > 0 && a->gw > 0 && a->nh > 0 && a->nw > 0,
                "gap_conv_wgrad: empty shape");
  GAP_CHECK_ARG(a->taps_h >= 1 && a->taps_w >= 1 && a->stride >= 1 && a->stride <= 2,
                "gap_conv_wgrad: taps/stride out of range");
  if (a->m_ld % 8 != 0 || a->n_ld % 8 != 0 || a->m_ld < a->m_c || a->n_ld < a->n_c) {
    set_error("gap_conv_wgrad: pixel strides must be multiples of 8 and >= channels");
    return GAP_ERR_ALIGNMENT;
  }

  WgradParams p;
  memset(&p, 0, sizeof(p));
  const int log_bw = std::min(6, wg_ilog2_ceil(a->gw));
  const int log_bh = std::min(6 - log_bw, wg_ilog2_ceil(a->gh));
  const int BW = 1 << log_bw, BH = 1 << log_bh, BNI = kWgBlockK / (BW * BH);
  p.log_bw = log_bw;
  p.log_bh = log_bh;
  p.tiles_w = (a->gw + BW - 1) / BW;
  p.tiles_h = (a->gh + BH - 1) / BH;
  p.tiles_n = (a->n + BNI - 1) / BNI;
  p.pix_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
  p.m_c = a->m_c;
  p.m_rows = a->m_rows > 0 ? std::min(a->m_rows, a->m_c) : a->m_c;
  p.n_c = a->n_c;
  p.n_img = a->n;
  p.gh = a->gh;
  p.gw = a->gw;
  p.taps_w = a->taps_w;
  p.n_taps = a->taps_h * a->taps_w;
  p.stride = a->stride;
  p.off_h = a->off_h;
  p.off_w = a->off_w;
  int block_n = std::min(256, a->n_c);
  const int force_bn = debug_get("wgrad_block_n", 0);
  if (force_bn > 0) block_n = force_bn;
  p.block_n = block_n;
  p.n_tiles = (a->n_c + block_n - 1) / block_n;
  // Two M tiles per CTA share every N-operand box (the TMA feed, ~60-100 B/clk per SM, is the bound:
  // a 256 x 256 output tile needs 64 KiB per 1024 MMA cycles, a 128 x 256 one 48 KiB per 512).
  int mt = a->m_c >= 256 ? 2 : 1;
  const int force_mt = debug_get("wgrad_mt", 0);
  if (force_mt > 0) mt = std::min(force_mt, 2);
  p.mt = mt;
  p.m_tiles = (a->m_c + 128 * mt - 1) / (128 * mt);
  int acc_cols = debug_get("wgrad_acc_cols", 512);
  int tpc = std::max(1, acc_cols / (mt * block_n));
  tpc = std::min(tpc, p.n_taps);
  // keep at least three pipeline stages
  while (tpc > 1 && ((mt * 2 + tpc * (block_n / 64)) * kWgBoxBytes > 72 * 1024 || p.n_taps % tpc != 0)) --tpc;
  const int force_tpc = debug_get("wgrad_tpc", 0);
  if (force_tpc > 0) tpc = std::min(force_tpc, p.n_taps);
  if (mt * tpc * block_n > 512 || block_n % 64 != 0) {
    set_error("gap_conv_wgrad: tile mt=%d tpc=%d block_n=%d does not fit 512 TMEM columns", mt, tpc, block_n);
    return GAP_ERR_UNSUPPORTED;
  }
  p.tpc = tpc;
  p.tap_groups = (p.n_taps + tpc - 1) / tpc;
  int cols = 32;
  while (cols < mt * tpc * block_n) cols *= 2;
  p.tmem_cols = cols;
  const int sms = sm_count();
  const int base = p.m_tiles * p.n_tiles * p.tap_groups;
  // split-K: minimise waves(base*s) * (K tiles per CTA + epilogue), the epilogue (fp32 red.add of the
  // whole TMEM tile) costing about as much as 15 K tiles (measured).
  int splits = 1;
  {
    const long long kEpi = 15;
    long long best = -1;
    const int s_max = std::max(1, std::min(p.pix_tiles, (4 * sms + base - 1) / base));
    for (int s = 1; s <= s_max; ++s) {
      const long long waves = (static_cast<long long>(base) * s + sms - 1) / sms;
      const long long cost = waves * ((p.pix_tiles + s - 1) / s + kEpi);
      if (best < 0 || cost < best) {
        best = cost;
        splits = s;
      }
    }
  }
  const int force_sp = debug_get("wgrad_splits", 0);
  if (force_sp > 0) splits = std::min(force_sp, p.pix_tiles);
  p.splits = splits;
  p.idesc = make_idesc_bf16(128, block_n, 1, 1);
  p.dual = 1;
  {
    const int force_dual = debug_get("wgrad_dual", -1);
    if (force_dual >= 0) p.dual = force_dual ? 1 : 0;
  }
  p.out = a->out;
  p.ld_m = a->ld_m;
  p.ld_tap = a->ld_tap;
  p.skip = debug_get("wgrad_skip", 0);
  p.vec_red = (a->ld_m % 4 == 0 && a->ld_tap % 4 == 0 && a->n_c % 16 == 0 &&
               (reinterpret_cast<uintptr_t>(a->out) & 15) == 0 && debug_get("wgrad_scalar_red", 0) == 0)
                  ? 1
                  : 0;
  {
    const long long lo = static_cast<unsigned int>(debug_get("trace_ptr_lo", 0));
    const long long hi = static_cast<unsigned int>(debug_get("trace_ptr_hi", 0));
    p.trace = reinterpret_cast<long long*>((hi << 32) | lo);
  }

  const int stage_bytes = mt * 2 * kWgBoxBytes + tpc * (block_n / 64) * kWgBoxBytes;
  int stages = (kWgSmemBudget - 1024 - 256) / stage_bytes;
  stages = std::max(1, std::min(stages, kWgMaxStages));
  p.num_stages = stages;
  const size_t smem_bytes = 1024 + static_cast<size_t>(stages) * stage_bytes + 256;
  if (smem_bytes > static_cast<size_t>(kWgSmemBudget)) {
    set_error("gap_conv_wgrad: stage of %d bytes does not fit shared memory", stage_bytes);
    return GAP_ERR_UNSUPPORTED;
  }

  {
    const uint64_t ld_b = static_cast<uint64_t>(a->m_ld) * 2;
    uint64_t dims[4] = {static_cast<uint64_t>(a->m_c), static_cast<uint64_t>(a->gw),
                        static_cast<uint64_t>(a->gh), static_cast<uint64_t>(a->n)};
    uint64_t strides[3] = {ld_b, ld_b * a->gw, ld_b * a->gw * a->gh};
    uint32_t box[4] = {64, static_cast<uint32_t>(BW), static_cast<uint32_t>(BH), static_cast<uint32_t>(BNI)};
    int rc = encode_tmap_bf16(&p.tmM, a->mop, 4, dims, strides, box, nullptr, true);
    if (rc) return rc;
  }
  {
    const uint64_t ld_b = static_cast<uint64_t>(a->n_ld) * 2;
    uint64_t dims[4] = {static_cast<uint64_t>(a->n_c), static_cast<uint64_t>(a->nw),
                        static_cast<uint64_t>(a->nh), static_cast<uint64_t>(a->n)};
    uint64_t strides[3] = {ld_b, ld_b * a->nw, ld_b * a->nw * a->nh};
    uint32_t box[4] = {64, static_cast<uint32_t>(BW * a->stride), static_cast<uint32_t>(BH * a->stride),
                       static_cast<uint32_t>(BNI)};
    uint32_t es[4] = {1, static_cast<uint32_t>(a->stride), static_cast<uint32_t>(a->stride), 1};
    int rc = encode_tmap_bf16(&p.tmN, a->nop, 4, dims, strides, box, es, true);
    if (rc) return rc;
  }

  static bool attr_set = false;
  if (!attr_set) {
    GAP_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kWgSmemBudget));
    attr_set = true;
  }
  const int grid = base * splits;
  GAP_CUDA(launch_pdl(conv_wgrad_kernel, dim3(grid), dim3(kWgThreads), smem_bytes, stream, p));
  GAP_CUDA(cudaGetLastError());
  return 0;
}
