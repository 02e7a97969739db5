// Which code shape lets one warp issue tcgen05.mma at the hardware rate when descriptors change every MMA?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../wdno_b200/csrc/ptx.cuh"
using namespace wdno;

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, unsigned long long* cyc, uint32_t slot_u, uint32_t nslot) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  uint32_t* tm = reinterpret_cast<uint32_t*>(sm + 8);
  uint8_t* a = sm + 128;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(tm, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tbase = *tm;
  if (threadIdx.x < 32) {
    const uint32_t idesc = ptx::make_idesc_f16(N, 0);
    constexpr uint64_t hi = static_cast<uint64_t>(8u | (1u << 14)) << 32;
    const uint32_t a_lo0 = (ptx::smem_u32(a) >> 4) + (218u << 16);
    const uint32_t b_lo0 = (ptx::smem_u32(a + 120 * 1024) >> 4) + (static_cast<uint32_t>(N) << 16);
    unsigned long long t0 = clock64();
    if (MODE == 0) {          // everything inside the elected branch (current tapgemm shape)
      if (ptx::elect_one()) {
        for (int it = 0; it < iters; ++it) {
          const uint32_t shift = (it * 7) & 63;
#pragma unroll
          for (int za = 0; za < 4; ++za) {
            uint32_t sa = (it & 3) + za; if (sa >= nslot) sa -= nslot;
            const uint32_t al = a_lo0 + shift + sa * slot_u;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              ptx::tc_mma_f16(tbase + za * 64, hi | (al + ks * 436u), hi | (b_lo0 + ks * 2u * N), idesc, 1u);
          }
        }
      }
    } else {                  // uniform control flow; only the instruction itself is predicated on the elected lane
      const bool leader = ptx::elect_one();
      for (int it = 0; it < iters; ++it) {
        const uint32_t shift = (it * 7) & 63;
#pragma unroll
        for (int za = 0; za < 4; ++za) {
          uint32_t sa = (it & 3) + za; if (sa >= nslot) sa -= nslot;
          const uint32_t al = a_lo0 + shift + sa * slot_u;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t ad = hi | (al + ks * 436u), bd = hi | (b_lo0 + ks * 2u * N);
            if (leader) ptx::tc_mma_f16(tbase + za * 64, ad, bd, idesc, 1u);
          }
        }
      }
    }
    if (ptx::elect_one()) { ptx::tc_commit(bar); }
    ptx::mbar_wait(bar, 0);
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tbase, 512);
}

int main() {
  unsigned long long* dc; cudaMalloc(&dc, 148 * 8);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {64, 128}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, 128, 190 * 1024>>>(N, iters, dc, 872, 8); else k<1><<<148, 128, 190 * 1024>>>(N, iters, dc, 872, 8);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      }
      unsigned long long hc[148]; cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
      double mx = 0; for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
      printf("mode=%d N=%3d : %.1f cycles per MMA (hardware floor %d)\n", mode, N, mx / (iters * 8.0), N == 64 ? 48 : 64);
    }
  return 0;
}
