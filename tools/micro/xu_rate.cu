// Micro-benchmark: issue cost (cycles per warp-instruction per SM sub-partition) of the conversion / transcendental
// instructions the fused GroupNorm+SiLU operand prologue uses.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512, 1) k(unsigned long long* out, float seed) {
  float x[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = seed + 0.01f * (threadIdx.x + i); h[i] = 0x38003800u + threadIdx.x + i; }
  __syncthreads();
  long long t0; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 3) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(h[i]) : "f"(x[i])); x[i] = __uint_as_float(h[i]); }
      if (OP == 4) { asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(x[i]) : "r"(h[i])); h[i] = __float_as_uint(x[i]); }
      if (OP == 5) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(h[i]));
      if (OP == 6) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
      if (OP == 7) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
  }
  long long t1; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) :: "memory");
  float s = 0; uint32_t u = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += x[i]; u ^= h[i]; }
  if (s == 12345.f && u == 77) out[1] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  unsigned long long* d; cudaMalloc(&d, 16);
  const char* names[] = {"tanh.approx.f32", "tanh.approx.f16x2", "ex2.approx.f32", "cvt.rn.f16x2.f32 (pack)", "cvt.f32.f16 (unpack)", "fma.f16x2", "fma.f32", "rcp.approx.f32"};
  unsigned long long h;
#define RUN(OP) k<OP><<<148, 512>>>(d, 0.1f); cudaDeviceSynchronize(); k<OP><<<148, 512>>>(d, 0.1f); cudaDeviceSynchronize(); \
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); \
  printf("%-28s: %.2f cycles per warp-instruction per SMSP (4 warps/SMSP, 8 independent chains)\n", names[OP], double(h) / (256.0 * 8 * 4));
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
