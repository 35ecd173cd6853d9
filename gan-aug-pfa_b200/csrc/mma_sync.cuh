// Warp-level tensor-core helpers (mma.sync m16n8k16 bf16 -> fp32, ldmatrix) for the HBM-bound thin
// layers (3/6-channel inputs, 3- and 1-channel outputs).  Those layers move 100-300 MB for a few
// GFLOP, so they are written as direct, fused, coalesced kernels on the legacy warp MMA path instead
// of being reshaped into tcgen05 tiles (which needs an im2col round trip through HBM).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace gap {

// D(16x8, fp32) += A(16x16, bf16, row-major) * B(16x8, bf16, "col": pairs along k)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Four 8x8 b16 matrices; lane l supplies the address of row (l & 7) of matrix (l >> 3).
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t& r0, uint32_t& r1, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}

// A-fragment (16 rows x 16 k) of a row-major bf16 smem tile with `stride_b` bytes per row:
// lane l addresses row (l & 15), k-column ((l >> 4) * 8).
__device__ __forceinline__ void lda_16x16(uint32_t (&a)[4], uint32_t tile_saddr, int stride_b, int lane) {
  ldmatrix_x4(a, tile_saddr + (lane & 15) * stride_b + (lane >> 4) * 16);
}
// Two B-fragments (n = 16 rows of an [n][k] row-major smem tile, k = 16): returns {b0,b1} of n-tile 0 in
// r[0],r[1] and of n-tile 1 in r[2],r[3].
__device__ __forceinline__ void ldb_16x16(uint32_t (&r)[4], uint32_t tile_saddr, int stride_b, int lane) {
  ldmatrix_x4(r, tile_saddr + ((lane & 7) + ((lane >> 4) << 3)) * stride_b + ((lane >> 3) & 1) * 16);
}

}  // namespace gap
