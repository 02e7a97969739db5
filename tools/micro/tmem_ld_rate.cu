// Micro-benchmark: tcgen05.ld 32x32b.x32 (4 KB per warp instruction) throughput and latency, 1..4 warps (one per TMEM lane
// quarter) and 8 warps (two per quarter) reading concurrently.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../wdno_b200/csrc/ptx.cuh"
using namespace wdno;

template <int DEP>
__global__ void __launch_bounds__(256, 1) k(unsigned long long* out, int nwarps, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (warp < nwarps) {
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
    for (int it = 0; it < iters; ++it) {
      uint32_t a[32], b[32];
      ptx::tmem_ld32(base + ((it * 64) & 255), a);
      if (DEP) { ptx::tmem_ld_wait(); }
      ptx::tmem_ld32(base + ((it * 64 + 32) & 255), b);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= a[i] + b[i];
    }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) :: "memory");
  }
  if (acc == 0x12345u) out[2] = acc;
  if ((threadIdx.x & 31) == 0 && warp < nwarps && blockIdx.x == 0) atomicMax(out, static_cast<unsigned long long>(t1 - t0));
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(slot, 512); }
}

int main() {
  unsigned long long* d; cudaMalloc(&d, 32);
  const int iters = 512;
  for (int dep = 0; dep < 2; ++dep)
    for (int nw : {1, 2, 4, 8}) {
      unsigned long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(d, 0, 32);
        if (dep) k<1><<<148, 256>>>(d, nw, iters); else k<0><<<148, 256>>>(d, nw, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double cyc = double(h) / (2.0 * iters);
      printf("%s  warps=%d : %.1f cycles per ld32 per warp -> %.1f B/clk per warp, %.1f B/clk per SM\n", dep ? "wait after every ld" : "two lds per wait  ",
             nw, cyc, 4096.0 / cyc, nw * 4096.0 / cyc);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
