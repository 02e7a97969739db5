// Micro-benchmark: tcgen05.mma (kind::f16, M=128, cta_group::1) issue rate from shared memory in the no-swizzle K-major
// layout used by tapgemm, as a function of N, of the A start-address alignment, and of the accumulator count.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../wdno_b200/csrc/ptx.cuh"
using namespace wdno;

__global__ void __launch_bounds__(128, 1) k(int N, int a_shift16, int nacc, int iters, int ksteps, unsigned long long* cyc, int swz) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  uint32_t* tm = reinterpret_cast<uint32_t*>(sm + 8);
  uint8_t* a = sm + 128;                 // A slab: [4 chunks][S_pad=400][16B]
  uint8_t* b = sm + 128 + 4 * 400 * 16;  // B tile: [4 chunks][N][16B]
  for (int i = threadIdx.x; i < (4 * 400 * 16 + 4 * 256 * 16) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(tm, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tbase = *tm;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_f16(N, 0);
    // swz: 0 none (SBO 128 B, LBO = chunk stride), 2 = SWIZZLE_128B (SBO 1024 B), 4 = SWIZZLE_64B (SBO 512 B)
    uint64_t hi = static_cast<uint64_t>(8u | (1u << 14)) << 32;
    uint32_t a_lo = (ptx::smem_u32(a) >> 4) + (400u << 16) + a_shift16;
    uint32_t b_lo = (ptx::smem_u32(b) >> 4) + (static_cast<uint32_t>(N) << 16);
    uint32_t akstep = 800u, bkstep = 2u * N;
    if (swz) {
      const uint32_t sbo = (swz == 2) ? 64u : 32u;
      const uint32_t rowb = (swz == 2) ? 8u : 4u;   // row bytes / 16
      const uint32_t base_off = (a_shift16 * rowb) & 7u ? 0u : 0u;
      hi = (static_cast<uint64_t>(sbo | (1u << 14)) << 32) | (static_cast<uint64_t>(swz) << 61);
      a_lo = ((ptx::smem_u32(a) + 1023u) & ~1023u) >> 4;
      a_lo += a_shift16 * rowb + (1u << 16);
      b_lo = (((ptx::smem_u32(b) + 1023u) & ~1023u) >> 4) + (1u << 16);
      akstep = 2u; bkstep = 2u;   // +32 B per K=16 step inside the swizzled row
      (void)base_off;
    }
    const int npad = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    unsigned long long t0 = clock64();
    if (a_shift16 == 1) {
      // constant descriptors, unrolled: pure hardware MMA rate (no per-MMA descriptor arithmetic / R2UR chains)
      const uint64_t ad = hi | a_lo, bd = hi | b_lo;
      for (int it = 0; it < iters * nacc * ksteps / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) ptx::tc_mma_f16(tbase + (u & 1) * npad, ad, bd, idesc, 1u);
      }
    } else
    for (int it = 0; it < iters; ++it) {
      for (int acc = 0; acc < nacc; ++acc)
        for (int ks = 0; ks < ksteps; ++ks)
          ptx::tc_mma_f16(tbase + acc * npad, hi | (a_lo + ks * akstep + acc * 16 * (swz ? 8u : 1u)), hi | (b_lo + ks * bkstep), idesc, 1u);
    }
    ptx::tc_commit(bar);
    ptx::mbar_wait(bar, 0);
    cyc[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tbase, 512);
}

int main() {
  unsigned long long* dc; cudaMalloc(&dc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int swz : {0, 2, 4})
  for (int N : {64, 128, 192, 256})
    for (int shift : {0, 1, 3})
      for (int nacc : {2, 4}) {
        if ((N <= 64 ? 64 : N) * nacc > 512) continue;
        const int ksteps = 2;
        for (int rep = 0; rep < 2; ++rep) {
          k<<<148, 128, 100 * 1024>>>(N, shift, nacc, iters, ksteps, dc, swz);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        }
        unsigned long long hc[148]; cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
        double per = mx / (double(iters) * nacc * ksteps);
        printf("swz=%d N=%3d a_shift=%d(x16B) nacc=%d : %.1f cycles per MMA (math floor %d) -> %.0f%% of peak\n", swz, N, shift, nacc, per, N / 2, 100.0 * (N / 2) / per);
      }
  return 0;
}
