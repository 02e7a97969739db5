// Micro-benchmark: how much does shared-memory traffic of OTHER warps slow tcgen05.mma (M = 128, operands in shared memory)?
// One thread issues a fixed number of MMAs (constant descriptors, as in mma_rate.cu); `bg` warps stream conflict-free LDS.128 +
// STS.128 over a separate 32 KB region until the MMAs are done.  Reports cycles per MMA and the background traffic in B/clk.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../wdno_b200/csrc/ptx.cuh"
using namespace wdno;

__global__ void __launch_bounds__(512, 1) k(int N, int bg, int n_mma, unsigned long long* out, int store) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  uint32_t* tm = reinterpret_cast<uint32_t*>(sm + 8);
  volatile int* done = reinterpret_cast<volatile int*>(sm + 16);
  uint8_t* a = sm + 128;                     // A: [2 chunks][128 rows][16 B]
  uint8_t* b = a + 2 * 128 * 16;             // B: [2 chunks][256 rows][16 B]
  uint8_t* scratch = b + 2 * 256 * 16;       // 64 KB for the background warps
  for (int i = threadIdx.x; i < (2 * 128 * 16 + 2 * 256 * 16 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); *done = 0; }
  if (threadIdx.x < 32) { ptx::tmem_alloc(tm, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tbase = *tm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_f16(N, 0);
      const uint64_t hi = static_cast<uint64_t>(8u | (1u << 14)) << 32;
      const uint64_t ad = hi | (((ptx::smem_u32(a) & 0x3FFFFu) >> 4) + (128u << 16));
      const uint64_t bd = hi | (((ptx::smem_u32(b) & 0x3FFFFu) >> 4) + (256u << 16));
      const int npad = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
      const unsigned long long t0 = clock64();
      for (int it = 0; it < n_mma / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) ptx::tc_mma_f16(tbase + (u & 1) * npad, ad, bd, idesc, 1u);
      }
      ptx::tc_commit(bar);
      ptx::mbar_wait(bar, 0);
      out[blockIdx.x * 2] = clock64() - t0;
      *done = 1;
    }
  } else if (warp <= bg) {
    // background: each warp owns 4 KB of scratch; 512 B per instruction, conflict-free
    uint4* p = reinterpret_cast<uint4*>(scratch + (warp - 1) * 4096) + lane;
    uint4 acc = make_uint4(0, 0, 0, 0);
    unsigned long long n = 0;
    while (!*done) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint4 v = p[u * 32];
        acc.x ^= v.x; acc.y += v.y;
        if (store) p[u * 32] = acc;
      }
      n += 8;
    }
    if (lane == 0) atomicAdd(&out[blockIdx.x * 2 + 1], n * 512ull * (store ? 2 : 1));
    if (acc.x == 0x12345678u) out[0] = 0;
  }
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tbase, 512);
}

int main() {
  unsigned long long* d; cudaMalloc(&d, 148 * 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int n_mma = 16000;
  for (int N : {64, 128, 256})
    for (int store : {0, 1})
      for (int bg : {0, 1, 2, 4, 8, 12}) {
        for (int rep = 0; rep < 2; ++rep) {
          cudaMemset(d, 0, 148 * 16);
          k<<<148, 512, 100 * 1024>>>(N, bg, n_mma, d, store);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        }
        unsigned long long h[2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const double per = double(h[0]) / n_mma;
        const double op = (4096.0 + N * 32.0) / per;
        printf("N=%3d background %2d warps (%s): %.1f cycles per MMA (math floor %d) -> %.0f%% of peak; MMA operands %.0f B/clk, background %.0f B/clk, sum %.0f B/clk\n",
               N, bg, store ? "LDS+STS" : "LDS    ", per, N / 2, 100.0 * (N / 2) / per, op, double(h[1]) / double(h[0]), op + double(h[1]) / double(h[0]));
      }
  return 0;
}
