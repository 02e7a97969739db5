// Micro-benchmark: does tcgen05.commit (mbarrier arrive on completion of prior MMAs) cost tensor-pipe time when it is
// interleaved with the MMA stream?  One thread issues n MMAs (M = 128, constant descriptors, two accumulators) with a commit to a
// scratch mbarrier (mode 0), a fence.proxy.async (mode 1) or a tcgen05.fence::after_thread_sync (mode 2) after every `every` MMAs
// (0 = only the final commit).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../wdno_b200/csrc/ptx.cuh"
using namespace wdno;

__global__ void __launch_bounds__(128, 1) k(int N, int every, int n_mma, unsigned long long* out, int mode) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);          // [0] final, [1..8] scratch
  uint32_t* tm = reinterpret_cast<uint32_t*>(sm + 96);
  uint8_t* a = sm + 128;
  uint8_t* b = a + 2 * 128 * 16;
  for (int i = threadIdx.x; i < (2 * 128 * 16 + 2 * 256 * 16) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 9; ++i) ptx::mbar_init(&bar[i], 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(tm, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tbase = *tm;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::make_idesc_f16(N, 0);
    const uint64_t hi = static_cast<uint64_t>(8u | (1u << 14)) << 32;
    const uint64_t ad = hi | (((ptx::smem_u32(a) & 0x3FFFFu) >> 4) + (128u << 16));
    const uint64_t bd = hi | (((ptx::smem_u32(b) & 0x3FFFFu) >> 4) + (256u << 16));
    const int npad = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    const unsigned long long t0 = clock64();
    int since = 0, slot = 0;
    for (int it = 0; it < n_mma; it += 2) {
      ptx::tc_mma_f16(tbase, ad, bd, idesc, 1u);
      ptx::tc_mma_f16(tbase + npad, ad, bd, idesc, 1u);
      since += 2;
      if (every && since >= every) {
        since = 0;
        if (mode == 0) {
          ptx::tc_commit(&bar[1 + slot]);
          slot = (slot + 1) & 7;
        } else if (mode == 1) {
          ptx::fence_proxy_async_smem();
        } else {
          ptx::tc_fence_after();
        }
      }
    }
    ptx::tc_commit(&bar[0]);
    ptx::mbar_wait(&bar[0], 0);
    out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tbase, 512);
}

int main() {
  unsigned long long* d; cudaMalloc(&d, 148 * 8);
  const int n_mma = 16000;
  for (int mode : {0, 1, 2})
  for (int N : {64, 128, 256})
    for (int every : {0, 32, 16, 8, 4, 2}) {
      if (mode && !every) continue;
      for (int rep = 0; rep < 2; ++rep) {
        k<<<148, 128, 20 * 1024>>>(N, every, n_mma, d, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      }
      unsigned long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double per = double(h) / n_mma;
      const char* what[] = {"tcgen05.commit", "fence.proxy.async", "tcgen05.fence::after_thread_sync"};
      printf("N=%3d %s every %2d MMAs: %.1f cycles per MMA (floor %d)%s\n", N, what[mode], every, per, N <= 64 ? 48 : N / 2,
             every ? "" : "  [single commit at the end]");
    }
  return 0;
}
