// Micro-benchmark: mma.sync.m16n8k16 (fp16 in, fp32 accumulate) issue rate per SM on sm_100a (legacy tensor path).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k(unsigned long long* out, int iters, float* sink) {
  float d[8][4];
  uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
  uint32_t b0 = 0x38003800u + threadIdx.x, b1 = 0x38003800u;
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  __syncthreads();
  long long t0; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  long long t1; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) :: "memory");
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 1234.5f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  unsigned long long* d; cudaMalloc(&d, 16);
  float* sink; cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int ctas : {1, 2, 4}) {
    k<<<148 * ctas, 128>>>(d, iters, sink); cudaDeviceSynchronize();
    k<<<148 * ctas, 128>>>(d, iters, sink); cudaDeviceSynchronize();
    unsigned long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double mmas_per_sm = double(iters) * 8 * 4 * ctas;
    printf("%d CTA/SM x 4 warps: %.2f cycles per MMA per SM -> %.0f MAC/clk/SM -> %.0f TFLOP/s at 1.9 GHz\n", ctas, h / mmas_per_sm,
           2048.0 * mmas_per_sm / h, 2 * 2048.0 * mmas_per_sm / h * 148 * 1.9e9 / 1e12);
  }
  return 0;
}
