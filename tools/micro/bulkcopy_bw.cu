// Micro-benchmark: throughput of 1-D cp.async.bulk (UBLKCP) global->shared from an L2-resident buffer.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulkcopy_bw bulkcopy_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../wdno_b200/csrc/ptx.cuh"
using namespace wdno;

__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, size_t per_cta_stride, uint32_t buf_bytes, uint32_t copy_bytes,
                                            int nstage, int iters, unsigned long long* cycles) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint8_t* data = sm + 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstage; ++i) ptx::mbar_init(&full[i], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = src + blockIdx.x * per_cta_stride;
    unsigned long long t0 = clock64();
    uint32_t off = 0;
    // prime
    for (int s = 0; s < nstage; ++s) {
      ptx::mbar_arrive_expect_tx(&full[s], copy_bytes);
      ptx::bulk_g2s(data + (size_t)s * copy_bytes, base + off, copy_bytes, &full[s]);
      off += copy_bytes; if (off + copy_bytes > buf_bytes) off = 0;
    }
    for (int it = 0; it < iters; ++it) {
      int s = it % nstage; uint32_t ph = (it / nstage) & 1;
      ptx::mbar_wait(&full[s], ph);
      if (it + nstage < iters + nstage) {  // always refill
        ptx::mbar_arrive_expect_tx(&full[s], copy_bytes);
        ptx::bulk_g2s(data + (size_t)s * copy_bytes, base + off, copy_bytes, &full[s]);
        off += copy_bytes; if (off + copy_bytes > buf_bytes) off = 0;
      }
    }
    for (int s = 0; s < nstage; ++s) { int it = iters + s; ptx::mbar_wait(&full[it % nstage], (it / nstage) & 1); }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const uint32_t buf = 216 * 1024;
  uint8_t* d; cudaMalloc(&d, (size_t)buf * 148); cudaMemset(d, 1, (size_t)buf * 148);
  unsigned long long* dc; cudaMalloc(&dc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int iters = 2000;
  for (int shared_src = 1; shared_src >= 0; --shared_src)
    for (uint32_t cb : {2048u, 4096u, 16384u, 32768u})
      for (int ns : {2, 4, 6}) {
        if ((size_t)cb * ns > 190 * 1024) continue;
        for (int rep = 0; rep < 2; ++rep) {
          k<<<148, 128, 128 + cb * ns>>>(d, shared_src ? 0 : buf, buf, cb, ns, iters, dc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        }
        unsigned long long hc[148]; cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
        double bpc = (double)cb * (iters + ns) / mx;
        printf("shared_src=%d copy=%6u B stages=%d : %.1f B/cycle/SM  (%.0f cycles per copy)\n", shared_src, cb, ns, bpc, mx / (iters + ns));
      }
  return 0;
}
