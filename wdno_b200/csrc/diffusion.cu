// Fused per-step diffusion algebra on the fp32 state [B, F, C, H, W] (Burgers: F = 1):
//   eps -> x0 (clamp) -> re-derived eps -> DDIM / DDPM update + noise -> condition re-imposition, one HBM pass.
// Reference: smoke/ddpm/diffusion_2d.py:689-699,723-754,769-785,851-933,970-976 ;
//            burgers/ddpm_burgers/diffusion_1d.py:172-182,205-258,276-307,376-460,520-527.
// All per-step scalars are read from DEVICE memory so a whole sampling step can live in one CUDA graph.
#include <cuda_runtime.h>
#include <math.h>
#include <algorithm>
#include <stdint.h>

#include "common.cuh"

namespace wdno {

struct CondProgram {
  wdno_cond_op op[WDNO_MAX_COND_OPS];
  int n;
};

struct StateDims {
  int B, F, C, H, W;
};

// later ops override earlier ones, exactly like the reference's sequence of in-place slice assignments
__device__ __forceinline__ float apply_conditions(const CondProgram& cp, float v, int b, int f, int c, int y, int x) {
#pragma unroll 1
  for (int k = 0; k < cp.n; ++k) {
    const wdno_cond_op& o = cp.op[k];
    if (f >= o.f0 && f < o.f1 && c >= o.c0 && c < o.c1 && y >= o.y0 && y < o.y1 && x >= o.x0 && x < o.x1) {
      if (o.src == nullptr) {
        v = 0.f;
      } else {
        v = o.src[b * o.sb + (f - o.of) * o.sf + (c - o.oc) * o.sc + (y - o.oy) * o.sy + (x - o.ox) * o.sx];
      }
    }
  }
  return v;
}

// 4 consecutive x of one row (W % 4 == 0): the (f, c, y) box tests are shared, only the x range is per element
__device__ __forceinline__ void apply_conditions4(const CondProgram& cp, float (&v)[4], int b, int f, int c, int y, int x0) {
#pragma unroll 1
  for (int k = 0; k < cp.n; ++k) {
    const wdno_cond_op& o = cp.op[k];
    if (f >= o.f0 && f < o.f1 && c >= o.c0 && c < o.c1 && y >= o.y0 && y < o.y1 && x0 + 3 >= o.x0 && x0 < o.x1) {
      const float* row = (o.src == nullptr) ? nullptr : o.src + (b * o.sb + (f - o.of) * o.sf + (c - o.oc) * o.sc + (y - o.oy) * o.sy);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = x0 + j;
        if (x >= o.x0 && x < o.x1) v[j] = (row == nullptr) ? 0.f : row[(x - o.ox) * o.sx];
      }
    }
  }
}

// 32-bit decode of a float4 index (total < 2^31)
__device__ __forceinline__ void decode4(unsigned i4, const StateDims& d, int& b, int& f, int& c, int& y, int& x) {
  const unsigned w4 = static_cast<unsigned>(d.W) >> 2;
  unsigned r = i4 / w4;
  x = static_cast<int>(i4 - r * w4) * 4;
  unsigned r2 = r / d.H;
  y = static_cast<int>(r - r2 * d.H);
  r = r2 / d.C;
  c = static_cast<int>(r2 - r * d.C);
  r2 = r / d.F;
  f = static_cast<int>(r - r2 * d.F);
  b = static_cast<int>(r2);
}

__device__ __forceinline__ void decode(size_t i, const StateDims& d, int& b, int& f, int& c, int& y, int& x) {
  x = static_cast<int>(i % d.W);
  size_t r = i / d.W;
  y = static_cast<int>(r % d.H);
  r /= d.H;
  c = static_cast<int>(r % d.C);
  r /= d.C;
  f = static_cast<int>(r % d.F);
  b = static_cast<int>(r / d.F);
}

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.f), 1.f); }

// coef: [0] sqrt_recip_alphas_cumprod[t]  [1] sqrt_recipm1_alphas_cumprod[t]  [2] sqrt(alpha_next)  [3] c  [4] sigma
//       [5] last-step flag (time_next < 0)  [6] guidance scale (eps += gscale * g)
__global__ void ddim_step_kernel(float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                 const float* __restrict__ g, const float* __restrict__ coef, CondProgram cp, StateDims d,
                                 size_t total, int cond_mode /*0 never, 1 except last step, 2 always*/) {
  const float sr = coef[0], srm1 = coef[1], san = coef[2], cc = coef[3], sigma = coef[4];
  const bool last = coef[5] != 0.f;
  const float gs = coef[6];
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float xv = x[i];
    float e = eps[i];
    // model_predictions: x0 = clamp(sr*x - srm1*eps); [guidance]; x0 = clamp(...) again; eps' = (sr*x - x0)/srm1
    if (g != nullptr) e = __fadd_rn(e, __fmul_rn(gs, g[i]));
    // explicit _rn intrinsics: same roundings as the reference's chain of separate torch ops (no FMA contraction)
    const float srx = __fmul_rn(sr, xv);
    const float x0 = clamp1(__fsub_rn(srx, __fmul_rn(srm1, e)));
    const float e2 = __fdiv_rn(__fsub_rn(srx, x0), srm1);
    float v;
    if (last) {
      v = x0;
    } else {
      v = __fadd_rn(__fmul_rn(x0, san), __fmul_rn(cc, e2));
      if (noise != nullptr) v = __fadd_rn(v, __fmul_rn(sigma, noise[i]));
    }
    if (cond_mode == 2 || (cond_mode == 1 && !last)) {
      int b, f, c, y, xx;
      decode(i, d, b, f, c, y, xx);
      v = apply_conditions(cp, v, b, f, c, y, xx);
    }
    x[i] = v;
  }
}

// float4 variant of the two step kernels (W % 4 == 0, 16-byte aligned tensors, total < 2^31): same per-element arithmetic
template <bool DDPM>
__global__ void __launch_bounds__(256) step4_kernel(float* __restrict__ x, const float* __restrict__ eps,
                                                    const float* __restrict__ noise, const float* __restrict__ g,
                                                    const float* __restrict__ coef, CondProgram cp, StateDims d, unsigned total4,
                                                    int cond_mode) {
  const float sr = coef[0], srm1 = coef[1], k2 = coef[2], k3 = coef[3], k4 = coef[4];
  const bool last = !DDPM && coef[5] != 0.f;
  const float gs = coef[6];
  const bool do_cond = DDPM ? (cond_mode != 0) : (cond_mode == 2 || (cond_mode == 1 && !last));
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const float4 xv4 = reinterpret_cast<const float4*>(x)[i];
    const float4 e4 = __ldg(reinterpret_cast<const float4*>(eps) + i);
    float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = n4;
    if (noise != nullptr) n4 = __ldg(reinterpret_cast<const float4*>(noise) + i);
    if (g != nullptr) g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    const float xa[4] = {xv4.x, xv4.y, xv4.z, xv4.w}, ea[4] = {e4.x, e4.y, e4.z, e4.w};
    const float na[4] = {n4.x, n4.y, n4.z, n4.w}, ga[4] = {g4.x, g4.y, g4.z, g4.w};
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float e = ea[j];
      if (g != nullptr) e = __fadd_rn(e, __fmul_rn(gs, ga[j]));
      const float srx = __fmul_rn(sr, xa[j]);
      const float x0 = clamp1(__fsub_rn(srx, __fmul_rn(srm1, e)));
      if (DDPM) {
        v[j] = __fadd_rn(__fmul_rn(k2, x0), __fmul_rn(k3, xa[j]));
        if (noise != nullptr && k4 != 0.f) v[j] = __fadd_rn(v[j], __fmul_rn(k4, na[j]));
      } else if (last) {
        v[j] = x0;
      } else {
        const float e2 = __fdiv_rn(__fsub_rn(srx, x0), srm1);
        v[j] = __fadd_rn(__fmul_rn(x0, k2), __fmul_rn(k3, e2));
        if (noise != nullptr) v[j] = __fadd_rn(v[j], __fmul_rn(k4, na[j]));
      }
    }
    if (do_cond) {
      int b, f, c, y, xx;
      decode4(i, d, b, f, c, y, xx);
      apply_conditions4(cp, v, b, f, c, y, xx);
    }
    reinterpret_cast<float4*>(x)[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

static bool vec4_ok(const void* a, const void* b, const void* c, const void* d, int W, size_t total) {
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return (W % 4 == 0) && total < (1ull << 31) && al(a) && al(b) && al(c) && al(d);
}

// coef: [0] sqrt_recip  [1] sqrt_recipm1  [2] posterior_mean_coef1  [3] posterior_mean_coef2
//       [4] exp(0.5*posterior_log_variance_clipped) (0 when t == 0)   [6] guidance scale
__global__ void ddpm_step_kernel(float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                 const float* __restrict__ g, const float* __restrict__ coef, CondProgram cp, StateDims d,
                                 size_t total, int cond_mode) {
  const float sr = coef[0], srm1 = coef[1], c1 = coef[2], c2 = coef[3], sd = coef[4];
  const float gs = coef[6];
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float xv = x[i];
    float e = eps[i];
    if (g != nullptr) e = __fadd_rn(e, __fmul_rn(gs, g[i]));
    const float x0 = clamp1(__fsub_rn(__fmul_rn(sr, xv), __fmul_rn(srm1, e)));
    float v = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xv));
    if (noise != nullptr && sd != 0.f) v = __fadd_rn(v, __fmul_rn(sd, noise[i]));
    if (cond_mode != 0) {
      int b, f, c, y, xx;
      decode(i, d, b, f, c, y, xx);
      v = apply_conditions(cp, v, b, f, c, y, xx);
    }
    x[i] = v;
  }
}

__global__ void apply_cond_kernel(float* __restrict__ x, CondProgram cp, StateDims d, size_t total) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int b, f, c, y, xx;
    decode(i, d, b, f, c, y, xx);
    x[i] = apply_conditions(cp, x[i], b, f, c, y, xx);
  }
}

// x0 = clamp(sr*x - srm1*eps) for the guidance callback (design_fn / nablaJ operate on x0)
__global__ void predict_x0_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ coef,
                                  float* __restrict__ x0, size_t total, int clip) {
  const float sr = coef[0], srm1 = coef[1];
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = __fsub_rn(__fmul_rn(sr, x[i]), __fmul_rn(srm1, eps[i]));
    x0[i] = clip ? clamp1(v) : v;
  }
}

// x_t = sqrt_ac[t_b] * x0 + sqrt_1mac[t_b] * noise   (per-sample t)
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const float* __restrict__ sa,
                                const float* __restrict__ s1a, const long long* __restrict__ t, float* __restrict__ out,
                                size_t per_sample, size_t total) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const long long tb = t[i / per_sample];
    out[i] = __fadd_rn(__fmul_rn(sa[tb], x0[i]), __fmul_rn(s1a[tb], noise[i]));
  }
}

// acc[b] += sum_i (pred - target)^2 * w[c]  (double atomics, one per block per sample)
__global__ void mse_weighted_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                    const float* __restrict__ w, int w_len, StateDims d, double* __restrict__ acc) {
  __shared__ double red[32];
  const int b = blockIdx.y;
  const size_t per = static_cast<size_t>(d.F) * d.C * d.H * d.W;
  const size_t hw = static_cast<size_t>(d.H) * d.W;
  double s = 0.0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < per;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>((i / hw) % d.C);
    const float df = pred[b * per + i] - target[b * per + i];
    const float ww = (w == nullptr) ? 1.f : (w_len == 1 ? w[0] : w[c]);
    s += static_cast<double>(df * df * ww);
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
    if (threadIdx.x == 0) atomicAdd(acc + b, v);
  }
}

// One block.  Copies the per-step scalars of step *step into fixed device slots, then advances *step.
__global__ void step_begin_kernel(int* __restrict__ step, const float* __restrict__ time_table,
                                  const float* __restrict__ coef_table, float* __restrict__ time_out,
                                  float* __restrict__ coef_out, int B, int n_steps) {
  int s = *step;
  if (s >= n_steps) s = n_steps - 1;
  const float t = time_table[s];
  for (int i = threadIdx.x; i < B; i += blockDim.x) time_out[i] = t;
  if (threadIdx.x < 8) coef_out[threadIdx.x] = coef_table[s * 8 + threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) *step = s + 1;
}

static int make_program(const wdno_cond_op* ops, int n_ops, CondProgram* cp) {
  if (n_ops < 0 || n_ops > WDNO_MAX_COND_OPS || (n_ops > 0 && !ops)) return set_error(WDNO_E_INVALID, "too many condition ops");
  cp->n = n_ops;
  for (int i = 0; i < n_ops; ++i) cp->op[i] = ops[i];
  return WDNO_OK;
}

static int ew_grid(size_t total) {
  size_t g = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  return static_cast<int>(g < cap ? g : cap);
}

}  // namespace wdno

using namespace wdno;

extern "C" int wdno_ddim_step(float* x, const float* eps, const float* noise, const float* guidance, const float* coef_dev,
                              const wdno_cond_op* ops_host, int n_ops, int B, int F, int C, int H, int W, int cond_mode,
                              void* stream) {
  if (!x || !eps || !coef_dev || B < 1 || F < 1 || C < 1 || H < 1 || W < 1) return set_error(WDNO_E_INVALID, "ddim_step: bad arguments");
  CondProgram cp;
  int rc = make_program(ops_host, n_ops, &cp);
  if (rc) return rc;
  StateDims d{B, F, C, H, W};
  const size_t total = static_cast<size_t>(B) * F * C * H * W;
  if (vec4_ok(x, eps, noise, guidance, W, total))
    step4_kernel<false><<<ew_grid(total / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, eps, noise, guidance, coef_dev, cp, d,
                                                                                        static_cast<unsigned>(total / 4), cond_mode);
  else
    ddim_step_kernel<<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, eps, noise, guidance, coef_dev, cp, d,
                                                                                    total, cond_mode);
  return check_launch("ddim_step");
}

extern "C" int wdno_ddpm_step(float* x, const float* eps, const float* noise, const float* guidance, const float* coef_dev,
                              const wdno_cond_op* ops_host, int n_ops, int B, int F, int C, int H, int W, int cond_mode,
                              void* stream) {
  if (!x || !eps || !coef_dev || B < 1 || F < 1 || C < 1 || H < 1 || W < 1) return set_error(WDNO_E_INVALID, "ddpm_step: bad arguments");
  CondProgram cp;
  int rc = make_program(ops_host, n_ops, &cp);
  if (rc) return rc;
  StateDims d{B, F, C, H, W};
  const size_t total = static_cast<size_t>(B) * F * C * H * W;
  if (vec4_ok(x, eps, noise, guidance, W, total))
    step4_kernel<true><<<ew_grid(total / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, eps, noise, guidance, coef_dev, cp, d,
                                                                                       static_cast<unsigned>(total / 4), cond_mode);
  else
    ddpm_step_kernel<<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, eps, noise, guidance, coef_dev, cp, d,
                                                                                    total, cond_mode);
  return check_launch("ddpm_step");
}

extern "C" int wdno_apply_conditions(float* x, const wdno_cond_op* ops_host, int n_ops, int B, int F, int C, int H, int W,
                                     void* stream) {
  if (!x || B < 1 || F < 1 || C < 1 || H < 1 || W < 1) return set_error(WDNO_E_INVALID, "apply_conditions: bad arguments");
  CondProgram cp;
  int rc = make_program(ops_host, n_ops, &cp);
  if (rc) return rc;
  StateDims d{B, F, C, H, W};
  const size_t total = static_cast<size_t>(B) * F * C * H * W;
  apply_cond_kernel<<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, cp, d, total);
  return check_launch("apply_conditions");
}

extern "C" int wdno_predict_x0(const float* x, const float* eps, const float* coef_dev, float* x0, int64_t total, int clip,
                               void* stream) {
  if (!x || !eps || !coef_dev || !x0 || total < 1) return set_error(WDNO_E_INVALID, "predict_x0: bad arguments");
  predict_x0_kernel<<<ew_grid(static_cast<size_t>(total)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, eps, coef_dev, x0, static_cast<size_t>(total), clip);
  return check_launch("predict_x0");
}

extern "C" int wdno_q_sample(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac,
                             const int64_t* t, float* out, int B, int64_t per_sample, void* stream) {
  if (!x0 || !noise || !sqrt_ac || !sqrt_1mac || !t || !out || B < 1 || per_sample < 1)
    return set_error(WDNO_E_INVALID, "q_sample: bad arguments");
  const size_t total = static_cast<size_t>(B) * per_sample;
  q_sample_kernel<<<ew_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x0, noise, sqrt_ac, sqrt_1mac, reinterpret_cast<const long long*>(t), out, static_cast<size_t>(per_sample), total);
  return check_launch("q_sample");
}

extern "C" int wdno_mse_weighted(const float* pred, const float* target, const float* w, int w_len, int B, int F, int C,
                                 int H, int W, double* acc, void* stream) {
  if (!pred || !target || !acc || B < 1 || F < 1 || C < 1 || H < 1 || W < 1 || (w && w_len != 1 && w_len != C))
    return set_error(WDNO_E_INVALID, "mse_weighted: bad arguments");
  StateDims d{B, F, C, H, W};
  const size_t per = static_cast<size_t>(F) * C * H * W;
  dim3 grid(static_cast<unsigned>(std::min<size_t>((per + 255) / 256, 64)), B);
  mse_weighted_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, target, w, w_len, d, acc);
  return check_launch("mse_weighted");
}

extern "C" int wdno_step_begin(int* step_dev, const float* time_table, const float* coef_table, float* time_out,
                               float* coef_out, int B, int n_steps, void* stream) {
  if (!step_dev || !time_table || !coef_table || !time_out || !coef_out || B < 1 || n_steps < 1)
    return set_error(WDNO_E_INVALID, "step_begin: bad arguments");
  step_begin_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(step_dev, time_table, coef_table, time_out, coef_out, B,
                                                                     n_steps);
  return check_launch("step_begin");
}
