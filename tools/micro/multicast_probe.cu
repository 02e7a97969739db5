// Probe: CTA-pair weight multicast protocol of tapgemm.cu in isolation (cluster of 2, leader multicasts, peer arms + remote arrive)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../wdno_b200/csrc/ptx.cuh"
using namespace wdno;

constexpr int kStages = 2, kBytes = 4096, kIters = 6;

__global__ void __launch_bounds__(128, 1) k(const uint32_t* __restrict__ src, uint32_t* __restrict__ out, int use_commit) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem);          // [2]
  uint64_t* b_empty = b_full + kStages;                          // [2]
  uint64_t* b_peer = b_empty + kStages;                          // [2]
  uint8_t* stage = smem + 128;
  const uint32_t rank = ptx::cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); ptx::mbar_init(&b_peer[i], 1); }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  ptx::cluster_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    // producer
    uint32_t bst = 0, bph = 0;
    for (int it = 0; it < kIters; ++it) {
      ptx::mbar_wait(&b_empty[bst], bph ^ 1u);
      if (rank != 0) {
        if (lane == 0) { ptx::mbar_arrive_expect_tx(&b_full[bst], kBytes); ptx::mbar_arrive_remote(&b_peer[bst], 0u); }
      } else {
        ptx::mbar_wait_cluster(&b_peer[bst], bph);
        if (lane == 0) {
          ptx::mbar_arrive_expect_tx(&b_full[bst], kBytes);
          ptx::bulk_g2s_multicast(stage + bst * kBytes, src + (blockIdx.x / 2 * kIters + it) * (kBytes / 4), kBytes, &b_full[bst], static_cast<uint16_t>(3));
        }
      }
      __syncwarp();
      if (++bst == kStages) { bst = 0; bph ^= 1u; }
    }
  } else if (warp == 1) {
    // consumer: checksum of every stage, then free it
    uint32_t bst = 0, bph = 0;
    for (int it = 0; it < kIters; ++it) {
      ptx::mbar_wait(&b_full[bst], bph);
      uint32_t s = 0;
      for (int i = lane; i < kBytes / 4; i += 32) s += reinterpret_cast<const uint32_t*>(stage + bst * kBytes)[i];
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) out[blockIdx.x * kIters + it] = s;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&b_empty[bst]);
      if (++bst == kStages) { bst = 0; bph ^= 1u; }
    }
  }
  __syncthreads();
  ptx::cluster_sync();
}

int main() {
  const int grid = 8;
  uint32_t *src, *out;
  const int n = grid / 2 * kIters * (kBytes / 4);
  cudaMalloc(&src, n * 4);
  cudaMalloc(&out, grid * kIters * 4);
  uint32_t* h = new uint32_t[n];
  for (int i = 0; i < n; ++i) h[i] = i * 2654435761u;
  cudaMemcpy(src, h, n * 4, cudaMemcpyHostToDevice);
  cudaMemset(out, 0, grid * kIters * 4);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 128 + kStages * kBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, (const uint32_t*)src, out, 0);
  printf("launch: %s\n", cudaGetErrorString(e));
  e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  uint32_t ho[grid * kIters];
  cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int b = 0; b < grid; ++b)
    for (int it = 0; it < kIters; ++it) {
      uint32_t s = 0;
      for (int i = 0; i < kBytes / 4; ++i) s += h[(b / 2 * kIters + it) * (kBytes / 4) + i];
      if (s != ho[b * kIters + it]) { ++bad; if (bad < 10) printf("mismatch cta %d it %d: %08x vs %08x\n", b, it, ho[b * kIters + it], s); }
    }
  printf("bad = %d of %d\n", bad, grid * kIters);
  return 0;
}
