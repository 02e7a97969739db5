// Warp-level tensor-core helpers (mma.sync m16n8k16, ldmatrix) shared by the attention kernels, plus the cooperative
// channel-LayerNorm tile loader that feeds them.  Fragment conventions (PTX ISA, m16n8k16, fp16 in / fp32 accumulate):
//   lane = 4*g + q ;  A (row-major 16x16): a0=(g, 2q..2q+1) a1=(g+8, 2q..) a2=(g, 2q+8..) a3=(g+8, 2q+8..)
//   B (col-major 16x8): b0=(k=2q..2q+1, n=g) b1=(k=2q+8.., n=g) ;  C/D (16x8): c0,c1=(g, 2q..2q+1) c2,c3=(g+8, 2q..2q+1)
// Weights are pre-packed on the host in fragment order (wdno_b200/attn_fused.py) so that one coalesced 16-byte load per
// lane yields the fragments of two consecutive k-steps (B) or of one k-step (A).
#pragma once
#include <cuda_fp16.h>
#include "cvt_sat.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

namespace wdno {

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float x, float y) {
  const __half2 h = wdno::h2_sat(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Channel LayerNorm (no bias) of ROWS consecutive voxels (C fp16 channels each, contiguous) into a shared-memory tile
// xn[ROWS][C + 8] (fp16; the 16-byte row pad keeps ldmatrix conflict-free).  All THREADS threads of the block take
// part; rows >= rows_valid are written as zeros.  reference: conv3d.py:165-174 (biased variance, eps inside sqrt).
template <int C, int ROWS, int THREADS>
__device__ __forceinline__ void ln_tile(const __half* __restrict__ x, int rows_valid, const float* __restrict__ gamma,
                                        __half* __restrict__ xn, float eps) {
  constexpr int LP = C / 8;             // lanes per voxel (16-byte chunk each)
  constexpr int RPP = THREADS / LP;     // rows per pass
  constexpr int PASSES = ROWS / RPP;
  static_assert(LP <= 32 && (THREADS % LP) == 0 && (ROWS % RPP) == 0, "ln_tile shape");
  const int l = threadIdx.x % LP;
  const int r0 = threadIdx.x / LP;
  uint4 raw[PASSES];
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    const int r = r0 + ps * RPP;
    raw[ps] = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid) raw[ps] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(r) * C) + l);
  }
  float gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gm[j] = __ldg(gamma + l * 8 + j);
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) {
    const int r = r0 + ps * RPP;
    const __half2* h = reinterpret_cast<const __half2*>(&raw[ps]);
    float f[8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __half22float2(h[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
      sum += t.x + t.y;
    }
#pragma unroll
    for (int sh = LP / 2; sh > 0; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
    const float mean = sum * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] -= mean;
      sq = fmaf(f[j], f[j], sq);
    }
#pragma unroll
    for (int sh = LP / 2; sh > 0; sh >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, sh);
    const float rstd = rsqrtf(sq * (1.0f / C) + eps);
    uint4 ov;
    uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = pack_h2(f[2 * j] * rstd * gm[2 * j], f[2 * j + 1] * rstd * gm[2 * j + 1]);
    if (r >= rows_valid) ov = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(xn + static_cast<size_t>(r) * (C + 8) + l * 8) = ov;
  }
}

}  // namespace wdno
