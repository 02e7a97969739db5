// fp32 -> fp16 conversions of everything the engine stores or feeds to the tensor cores: round-to-nearest with
// SATURATION to the largest finite fp16 (+-65504) instead of +-inf.  The reference keeps fp32 activations
// (video_diffusion_pytorch_conv3d.py, unet.py); the engine's fp16 residual stream must not turn a large-but-finite
// activation of a trained checkpoint into inf/NaN that then poisons GroupNorm statistics and every later layer:
// it clamps (tests/test_gpu_parity_configs.py::test_fp16_range_guard).  Same instruction count as the
// plain conversion (F2FP with the .SATFINITE modifier).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace wdno {

__device__ __forceinline__ __half2 h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));  // first source -> upper half
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ __half2 h2_sat(float2 v) { return h2_sat(v.x, v.y); }
__device__ __forceinline__ __half h_sat(float v) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return *reinterpret_cast<__half*>(&r);
}

}  // namespace wdno
