// Probe: numeric shared-memory addresses (cvta.to.shared) of the two CTAs of a cluster, and what mapa returns
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../wdno_b200/csrc/ptx.cuh"
using namespace wdno;
__global__ void __launch_bounds__(128, 1) k(uint32_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t slot;
  const uint32_t a = ptx::smem_u32(smem);
  uint32_t m0, m1;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(m0) : "r"(a));
  asm volatile("mapa.shared::cluster.u32 %0, %1, 1;" : "=r"(m1) : "r"(a));
  if (threadIdx.x < 32) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    out[blockIdx.x * 8 + 0] = ptx::cluster_ctarank();
    out[blockIdx.x * 8 + 1] = a;
    out[blockIdx.x * 8 + 2] = m0;
    out[blockIdx.x * 8 + 3] = m1;
    out[blockIdx.x * 8 + 4] = slot;
    out[blockIdx.x * 8 + 5] = ptx::smem_u32(&slot);
  }
  __syncthreads();
  ptx::cluster_sync();
  if (threadIdx.x < 32) ptx::tmem_dealloc(slot, 512);
}
int main() {
  uint32_t* out; cudaMalloc(&out, 4 * 8 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, out);
  printf("launch %s; sync %s\n", cudaGetErrorString(e), cudaGetErrorString(cudaDeviceSynchronize()));
  uint32_t h[32]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  for (int b = 0; b < 4; ++b) printf("cta %d: rank %u smem 0x%08x mapa0 0x%08x mapa1 0x%08x tmem 0x%08x &slot 0x%08x\n", b, h[b*8], h[b*8+1], h[b*8+2], h[b*8+3], h[b*8+4], h[b*8+5]);
  return 0;
}
