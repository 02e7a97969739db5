// Row-slice of a full-batch normal draw, bit-identical to torch.randn(full_shape, device='cuda')[lo:hi].
//
// Why: SURVEY.md section 8e -- a batch-sharded run reproduces the single-GPU trajectory only if every rank consumes the
// rows it owns of the FULL-batch noise tensor the single process would have drawn (the reference draws
// torch.randn(shape) / torch.randn_like(img): smoke/ddpm/diffusion_2d.py:866,907; burgers/ddpm_burgers/diffusion_1d.py:389,431).
// Drawing the full batch on every rank and slicing costs N x the noise traffic.  torch's CUDA normal_ is counter based
// (Philox4x32-10): element e of a tensor of `numel` values is component ii of the iter-th curand_normal4() of the Philox
// stream (seed, subsequence = idx, offset), with  e = idx + stride * (4 * iter + ii),  stride = 256 * grid,
// grid = min(SMs * (max threads per SM / 256), ceil(numel / 256))  (ATen distribution_elementwise_grid_stride_kernel /
// calc_execution_policy).  This kernel evaluates exactly those draws for e in [e0, e0 + n_local) only.  The host wrapper
// (wdno_b200/ops.py::randn_rows) verifies the mapping against torch once per process and falls back to draw-and-slice
// if a torch build ever changes it.
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

#include "../../include/wdno_b200.h"
#include "common.cuh"

namespace wdno {

__global__ void __launch_bounds__(256) randn_slice_kernel(float* __restrict__ out, long long n_local, long long e0,
                                                          long long numel_full, long long stride, unsigned long long seed,
                                                          unsigned long long offset, long long iter_lo, long long items) {
  const long long e1 = e0 + n_local;
  for (long long it = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; it < items;
       it += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long iter = iter_lo + it / stride;
    const long long idx = it - (it / stride) * stride;
    const long long eb = idx + stride * 4 * iter;
    bool any = false;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long e = eb + stride * ii;
      any |= (e >= e0 && e < e1 && e < numel_full);
    }
    if (!any) continue;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, static_cast<unsigned long long>(idx), offset + 4ull * static_cast<unsigned long long>(iter), &st);
    const float4 r = curand_normal4(&st);
    const float v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const long long e = eb + stride * ii;
      if (e >= e0 && e < e1 && e < numel_full) out[e - e0] = v[ii] * 1.0f + 0.0f;  // torch: rand * std + mean
    }
  }
}

}  // namespace wdno

extern "C" int wdno_randn_slice(float* out, int64_t n_local, int64_t elem_offset, int64_t numel_full, int grid_full,
                                uint64_t seed, uint64_t philox_offset, void* stream) {
  using namespace wdno;
  if (!out || n_local < 0 || elem_offset < 0 || numel_full < 1 || elem_offset + n_local > numel_full || grid_full < 1)
    return set_error(WDNO_E_INVALID, "randn_slice: bad arguments");
  if (n_local == 0) return WDNO_OK;
  const long long stride = 256ll * grid_full;
  const long long iter_lo = (elem_offset / stride) / 4, iter_hi = ((elem_offset + n_local - 1) / stride) / 4;
  const long long items = (iter_hi - iter_lo + 1) * stride;
  long long g = (items + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (g > cap) g = cap;
  randn_slice_kernel<<<static_cast<int>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out, n_local, elem_offset, numel_full, stride, seed, philox_offset, iter_lo, items);
  return check_launch("randn_slice");
}
