// HBM-bound kernels of the training step: GroupNorm+SiLU backward, channel-LayerNorm backward,
// gradient packing / accumulation, fused clip + Adam + EMA.  Reference: autograd through Block / ResnetBlock / LayerNorm
// (video_diffusion_pytorch_conv3d.py:165-230; unet.py:55-65,129-181) and Trainer.train (diffusion_2d.py:1277-1297).
// Activation gradients are fp16 channels-last in a scaled domain (include/wdno_b200.h); reductions run in fp32 per thread,
// double across threads.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "../../include/wdno_b200.h"
#include "common.cuh"
#include "cvt_sat.cuh"

namespace wdno {

namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 t = __half22float2(h[k]);
    f[2 * k] = t.x;
    f[2 * k + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  __half2* h = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = h2_sat(f[2 * k], f[2 * k + 1]);
  return o;
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
// d silu(z) / dz = s (1 + z (1 - s)), s = sigmoid(z)
__device__ __forceinline__ float dsilu(float z) {
  const float s = 1.0f / (1.0f + __expf(-z));
  return s * fmaf(z, 1.0f - s, 1.0f);
}

// ------------------------------------------------------------------ GroupNorm + SiLU backward: reduce
// block (256 threads) = RPB voxels x (C/8) lanes; grid = (voxel chunks, B).  Each thread keeps (S1, Sy) of its 8 channels.
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(const __half* __restrict__ dh, const __half* __restrict__ y,
                                                            const float* __restrict__ a, const float* __restrict__ c,
                                                            double* __restrict__ sums, int C, long long vox) {
  extern __shared__ float red[];  // [256][16]
  const int cpv = C >> 3;
  const int b = blockIdx.y;
  const int lane_c = threadIdx.x % cpv, row0 = threadIdx.x / cpv, rpb = blockDim.x / cpv;
  float av[8], cv[8], s1[8], sy[8];
  load8(a + static_cast<size_t>(b) * C + lane_c * 8, av);
  load8(c + static_cast<size_t>(b) * C + lane_c * 8, cv);
#pragma unroll
  for (int k = 0; k < 8; ++k) { s1[k] = 0.f; sy[k] = 0.f; }
  const size_t base = static_cast<size_t>(b) * vox;
  if (row0 < rpb) {
    for (long long v = static_cast<long long>(blockIdx.x) * rpb + row0; v < vox; v += static_cast<long long>(gridDim.x) * rpb) {
      const size_t i = (base + v) * cpv + lane_c;
      float fy[8], fd[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), fy);
      unpack8(__ldg(reinterpret_cast<const uint4*>(dh) + i), fd);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float dz = fd[k] * dsilu(fmaf(av[k], fy[k], cv[k]));
        s1[k] += dz;
        sy[k] = fmaf(dz, fy[k], sy[k]);
      }
    }
  }
  float* mine = red + threadIdx.x * 16;
#pragma unroll
  for (int k = 0; k < 8; ++k) { mine[k] = s1[k]; mine[8 + k] = sy[k]; }
  __syncthreads();
  // thread t < cpv*16 sums column (lane t / 16, entry t % 16) over the rows
  for (int t = threadIdx.x; t < cpv * 16; t += blockDim.x) {
    const int lc = t >> 4, e = t & 15;
    double acc = 0.0;
    for (int r = 0; r < rpb; ++r) acc += static_cast<double>(red[(r * cpv + lc) * 16 + e]);
    const int ch = lc * 8 + (e & 7);
    atomicAdd(sums + (static_cast<size_t>(b) * C + ch) * 2 + (e >> 3), acc);
  }
}

// ------------------------------------------------------------------ finalize: one block per sample
__global__ void gn_bwd_finalize_kernel(const double* __restrict__ sums, const double* __restrict__ stats,
                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                       const float* __restrict__ ss, int ss_stride, float* __restrict__ d_gamma,
                                       float* __restrict__ d_beta, float* __restrict__ d_ss, int dss_stride,
                                       float* __restrict__ k1, float* __restrict__ k0, int C, int G, double count, float eps,
                                       float scale) {
  extern __shared__ double sh[];  // [2 * G]: M1, M2 accumulators
  const int b = blockIdx.x;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < 2 * G; g += blockDim.x) sh[g] = 0.0;
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const int g = ch / cpg;
    const double S = stats[(static_cast<size_t>(b) * G + g) * 2], SS = stats[(static_cast<size_t>(b) * G + g) * 2 + 1];
    const double mean = S / count;
    double var = SS / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
    const double S1 = sums[(static_cast<size_t>(b) * C + ch) * 2], Sy = sums[(static_cast<size_t>(b) * C + ch) * 2 + 1];
    const double A = rstd * (Sy - mean * S1), Bc = S1;
    const double sc = (ss != nullptr) ? static_cast<double>(ss[static_cast<size_t>(b) * ss_stride + ch]) : 0.0;
    const double gp = static_cast<double>(gamma[ch]) * (1.0 + sc);
    atomicAdd(&sh[g], gp * Bc);
    atomicAdd(&sh[G + g], gp * A);
    atomicAdd(d_gamma + ch, static_cast<float>(A * (1.0 + sc) * scale));
    atomicAdd(d_beta + ch, static_cast<float>(Bc * (1.0 + sc) * scale));
    if (d_ss != nullptr) {
      d_ss[static_cast<size_t>(b) * dss_stride + ch] = static_cast<float>((A * gamma[ch] + Bc * beta[ch]) * scale);
      d_ss[static_cast<size_t>(b) * dss_stride + C + ch] = static_cast<float>(Bc * scale);
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const double S = stats[(static_cast<size_t>(b) * G + g) * 2], SS = stats[(static_cast<size_t>(b) * G + g) * 2 + 1];
    const double mean = S / count;
    double var = SS / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
    const double M1 = sh[g] / count, M2 = sh[G + g] / count;
    k1[static_cast<size_t>(b) * G + g] = static_cast<float>(-rstd * rstd * M2);
    k0[static_cast<size_t>(b) * G + g] = static_cast<float>(-rstd * M1 + rstd * rstd * M2 * mean);
  }
}

// ------------------------------------------------------------------ apply: dy = a*dz + k1*y + k0 (+ add)
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const __half* __restrict__ dh, const __half* __restrict__ y,
                                                           const float* __restrict__ a, const float* __restrict__ c,
                                                           const float* __restrict__ k1, const float* __restrict__ k0,
                                                           const __half* __restrict__ add, __half* __restrict__ dy, int C, int G,
                                                           size_t vox, size_t total_chunks) {
  const int cpv = C >> 3, cpg = C / G;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total_chunks;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t v = i / cpv;
    const int ch = static_cast<int>(i - v * cpv) * 8;
    const int b = static_cast<int>(v / vox);
    float av[8], cv[8], fy[8], fd[8], o[8];
    load8(a + static_cast<size_t>(b) * C + ch, av);
    load8(c + static_cast<size_t>(b) * C + ch, cv);
    unpack8(__ldg(reinterpret_cast<const uint4*>(y) + i), fy);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dh) + i), fd);
    const int g = ch / cpg;  // cpg is a multiple of 8: the 8 channels share a group
    const float kk1 = __ldg(k1 + static_cast<size_t>(b) * G + g), kk0 = __ldg(k0 + static_cast<size_t>(b) * G + g);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float dz = fd[k] * dsilu(fmaf(av[k], fy[k], cv[k]));
      o[k] = fmaf(av[k], dz, fmaf(kk1, fy[k], kk0));
    }
    if (add != nullptr) {
      float fa[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(add) + i), fa);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += fa[k];
    }
    reinterpret_cast<uint4*>(dy)[i] = pack8(o);
  }
}

// ------------------------------------------------------------------ d_eps fp32 [B,F,C,H,W] -> fp16 [B,F,H,W,cp] * mul
__global__ void pack_grad_kernel(const float* __restrict__ g, __half* __restrict__ out, int C, int HW, int cp, float mul,
                                 size_t total) {
  // one thread per (b*f, pixel, channel pair)
  const int pairs = cp >> 1;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int pr = static_cast<int>(i % pairs);
    const size_t t = i / pairs;
    const int px = static_cast<int>(t % HW);
    const size_t bf = t / HW;
    const int c0 = 2 * pr;
    const float v0 = (c0 < C) ? g[(bf * C + c0) * HW + px] * mul : 0.f;
    const float v1 = (c0 + 1 < C) ? g[(bf * C + c0 + 1) * HW + px] * mul : 0.f;
    reinterpret_cast<__half2*>(out)[i] = h2_sat(v0, v1);
  }
}

__global__ void add_f16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, size_t n16) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n16;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float fa[8], fb[8];
    unpack8(__ldg(a + i), fa);
    unpack8(__ldg(b + i), fb);
#pragma unroll
    for (int k = 0; k < 8; ++k) fa[k] += fb[k];
    out[i] = pack8(fa);
  }
}

// nearest x2 up-sampling of the two inner spatial axes (nn.Upsample(scale_factor=2), unet.py:35-39) and its adjoint (2x2 sum):
// fp16 [N][H][W][C] <-> [N][2H][2W][C]; one thread per 16-byte chunk of the LOW-resolution tensor
__global__ void up2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int cpv, size_t total) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cpv);
    size_t t = i / cpv;
    const int xw = static_cast<int>(t % W);
    t /= W;
    const int yh = static_cast<int>(t % H);
    const size_t n = t / H;
    const uint4 v = __ldg(x + i);
    const size_t o = ((n * 2 * H + 2 * yh) * 2 * W + 2 * xw) * cpv + c;
    y[o] = v;
    y[o + cpv] = v;
    y[o + static_cast<size_t>(2 * W) * cpv] = v;
    y[o + static_cast<size_t>(2 * W) * cpv + cpv] = v;
  }
}
__global__ void sumpool2x2_kernel(const uint4* __restrict__ y, uint4* __restrict__ x, int H, int W, int cpv, size_t total) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cpv);
    size_t t = i / cpv;
    const int xw = static_cast<int>(t % W);
    t /= W;
    const int yh = static_cast<int>(t % H);
    const size_t n = t / H;
    const size_t o = ((n * 2 * H + 2 * yh) * 2 * W + 2 * xw) * cpv + c;
    float a[8], b[8];
    unpack8(__ldg(y + o), a);
    unpack8(__ldg(y + o + cpv), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += b[k];
    unpack8(__ldg(y + o + static_cast<size_t>(2 * W) * cpv), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += b[k];
    unpack8(__ldg(y + o + static_cast<size_t>(2 * W) * cpv + cpv), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += b[k];
    x[i] = pack8(a);
  }
}

// ------------------------------------------------------------------ channel LayerNorm backward
// LPV lanes per voxel, each lane CPL chunks of 8 channels (chunk j of lane l = channels (j*LPV + l)*8 ..).  A thread keeps
// its channels for the whole grid-stride loop, so d_gamma accumulates in registers.
template <int LPV, int CPL>
__global__ void __launch_bounds__(256) chan_ln_bwd_kernel(const __half* __restrict__ x, const __half* __restrict__ dy,
                                                          const float* __restrict__ gamma, const __half* __restrict__ add,
                                                          __half* __restrict__ dx, float* __restrict__ d_gamma, size_t nvox,
                                                          float eps, float scale) {
  constexpr int C = LPV * CPL * 8;
  constexpr int VPB = 256 / LPV;
  __shared__ float red[256 * 8];
  const int l = threadIdx.x % LPV, r = threadIdx.x / LPV;
  float gm[CPL][8], dg[CPL][8];
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    load8(gamma + (j * LPV + l) * 8, gm[j]);
#pragma unroll
    for (int k = 0; k < 8; ++k) dg[j][k] = 0.f;
  }
  // the trip count is uniform over the block (warp shuffles below use the full mask): voxels past the end are computed on
  // zeros and not stored
  for (size_t v0 = static_cast<size_t>(blockIdx.x) * VPB; v0 < nvox; v0 += static_cast<size_t>(gridDim.x) * VPB) {
    const size_t v = v0 + r;
    const bool valid = v < nvox;
    float fx[CPL][8], fd[CPL][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const size_t i = v * (C / 8) + j * LPV + l;
      uint4 rx = make_uint4(0u, 0u, 0u, 0u), rd = make_uint4(0u, 0u, 0u, 0u);
      if (valid) {
        rx = __ldg(reinterpret_cast<const uint4*>(x) + i);
        rd = __ldg(reinterpret_cast<const uint4*>(dy) + i);
      }
      unpack8(rx, fx[j]);
      unpack8(rd, fd[j]);
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += fx[j][k];
    }
#pragma unroll
    for (int sh = LPV / 2; sh > 0; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
    const float mean = sum * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        fx[j][k] -= mean;
        sq = fmaf(fx[j][k], fx[j][k], sq);
      }
#pragma unroll
    for (int sh = LPV / 2; sh > 0; sh >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, sh);
    const float rstd = rsqrtf(sq * (1.0f / C) + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int j = 0; j < CPL; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        fx[j][k] *= rstd;                      // xhat
        dg[j][k] = fmaf(fd[j][k], fx[j][k], dg[j][k]);
        fd[j][k] *= gm[j][k];                  // g * dy
        m1 += fd[j][k];
        m2 = fmaf(fd[j][k], fx[j][k], m2);
      }
#pragma unroll
    for (int sh = LPV / 2; sh > 0; sh >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, sh);
      m2 += __shfl_xor_sync(0xffffffffu, m2, sh);
    }
    m1 *= (1.0f / C);
    m2 *= (1.0f / C);
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const size_t i = v * (C / 8) + j * LPV + l;
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = rstd * (fd[j][k] - m1 - fx[j][k] * m2);
      if (valid) {
        if (add != nullptr) {
          float fa[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(add) + i), fa);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] += fa[k];
        }
        reinterpret_cast<uint4*>(dx)[i] = pack8(o);
      }
    }
  }
  // d_gamma: reduce over the block's voxel rows, then one atomic per channel
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = dg[j][k];
    __syncthreads();
    for (int t = threadIdx.x; t < LPV * 8; t += 256) {
      const int ll = t >> 3, k = t & 7;
      float acc = 0.f;
      for (int rr = 0; rr < VPB; ++rr) acc += red[(rr * LPV + ll) * 8 + k];
      atomicAdd(d_gamma + (j * LPV + ll) * 8 + k, acc * scale);
    }
  }
}

// ------------------------------------------------------------------ optimiser
__global__ void sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ out) {
  double acc = 0.0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = g[i];
    acc += static_cast<double>(v) * v;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, sh[0]);
}

// torch.nn.utils.clip_grad_norm_(max_norm): g *= min(1, max_norm / (norm + 1e-6)); torch.optim.Adam (no weight decay,
// no amsgrad): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps);
// ema_mode 0: none; 1: ema = p (warm-up copy); 2: ema += (1 - decay) * (p - ema)   (ema_pytorch: lerp)
__global__ void adam_clip_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                     float* __restrict__ v, float* __restrict__ ema, size_t n, const double* __restrict__ sumsq,
                                     float max_norm, float lr, float b1, float b2, float eps, float bc1, float bc2,
                                     float ema_decay, int ema_mode) {
  float clip = 1.0f;
  if (sumsq != nullptr && max_norm > 0.f) {
    const float norm = static_cast<float>(sqrt(*sumsq));
    const float cc = max_norm / (norm + 1e-6f);
    clip = cc < 1.0f ? cc : 1.0f;
  }
  const float step = lr / bc1, rs2 = rsqrtf(bc2);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * clip;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float pi = p[i] - step * mi / (sqrtf(vi) * rs2 + eps);
    p[i] = pi;
    if (ema_mode == 1) ema[i] = pi;
    else if (ema_mode == 2) ema[i] += (1.0f - ema_decay) * (pi - ema[i]);
  }
}

int ew_grid_(size_t total) {
  size_t g = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  return static_cast<int>(std::max<size_t>(1, g < cap ? g : cap));
}

}  // namespace

}  // namespace wdno

using namespace wdno;

extern "C" int wdno_gn_bwd_reduce(const void* dh, const void* y, const float* a, const float* c, double* sums, int B, int C,
                                  int64_t vox, void* stream) {
  if (!dh || !y || !a || !c || !sums || B < 1 || C < 8 || (C & 7) || C > 2048 || vox < 1)
    return set_error(WDNO_E_INVALID, "gn_bwd_reduce: bad arguments");
  const int cpv = C >> 3;
  if (cpv > 256) return set_error(WDNO_E_INVALID, "gn_bwd_reduce: C too large");
  const int rpb = 256 / cpv;
  long long gx = (vox + rpb - 1) / rpb;
  const long long cap = std::max<long long>(1, static_cast<long long>(num_sms()) * 8 / B);
  if (gx > cap) gx = cap;
  gn_bwd_reduce_kernel<<<dim3(static_cast<unsigned>(gx), B), 256, 256 * 16 * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(dh), static_cast<const __half*>(y), a, c, sums, C, vox);
  return check_launch("gn_bwd_reduce");
}

extern "C" int wdno_gn_bwd_finalize(const double* sums, const double* stats, const float* gamma, const float* beta,
                                    const float* ss, int ss_stride, float* d_gamma, float* d_beta, float* d_ss, int dss_stride,
                                    float* k1, float* k0, int B, int C, int G, double count, float eps, float scale, void* stream) {
  if (!sums || !stats || !gamma || !beta || !d_gamma || !d_beta || !k1 || !k0 || B < 1 || C < 1 || G < 1 || (C % G) || count <= 0)
    return set_error(WDNO_E_INVALID, "gn_bwd_finalize: bad arguments");
  gn_bwd_finalize_kernel<<<B, 256, 2 * G * sizeof(double), static_cast<cudaStream_t>(stream)>>>(
      sums, stats, gamma, beta, ss, ss_stride, d_gamma, d_beta, d_ss, dss_stride, k1, k0, C, G, count, eps, scale);
  return check_launch("gn_bwd_finalize");
}

extern "C" int wdno_gn_bwd_apply(const void* dh, const void* y, const float* a, const float* c, const float* k1, const float* k0,
                                 const void* add, void* dy, int B, int C, int G, int64_t vox, void* stream) {
  if (!dh || !y || !a || !c || !k1 || !k0 || !dy || B < 1 || C < 8 || (C & 7) || G < 1 || (C % G) || ((C / G) & 7) || vox < 1)
    return set_error(WDNO_E_INVALID, "gn_bwd_apply: bad arguments (channels per group must be a multiple of 8)");
  const size_t chunks = static_cast<size_t>(B) * vox * (C >> 3);
  gn_bwd_apply_kernel<<<ew_grid_(chunks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(dh), static_cast<const __half*>(y), a, c, k1, k0, static_cast<const __half*>(add),
      static_cast<__half*>(dy), C, G, static_cast<size_t>(vox), chunks);
  return check_launch("gn_bwd_apply");
}

extern "C" int wdno_pack_grad_f16(const float* g, void* out, int B, int F, int C, int H, int W, int cp, float mul, void* stream) {
  if (!g || !out || B < 1 || F < 1 || C < 1 || H < 1 || W < 1 || cp < C || (cp & 7))
    return set_error(WDNO_E_INVALID, "pack_grad_f16: bad arguments");
  const size_t total = static_cast<size_t>(B) * F * H * W * (cp >> 1);
  pack_grad_kernel<<<ew_grid_(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, static_cast<__half*>(out), C, H * W, cp, mul, total);
  return check_launch("pack_grad_f16");
}

extern "C" int wdno_add_f16(const void* a, const void* b, void* out, int64_t n, void* stream) {
  if (!a || !b || !out || n < 8 || (n & 7)) return set_error(WDNO_E_INVALID, "add_f16: n must be a positive multiple of 8");
  const size_t n16 = static_cast<size_t>(n) >> 3;
  add_f16_kernel<<<ew_grid_(n16), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(a), static_cast<const uint4*>(b),
                                                                            static_cast<uint4*>(out), n16);
  return check_launch("add_f16");
}

extern "C" int wdno_chan_layernorm_bwd(const void* x, const void* dy, const float* gamma, const void* add, void* dx,
                                       float* d_gamma, int64_t n_vox, int C, float eps, float scale, void* stream) {
  if (!x || !dy || !gamma || !dx || !d_gamma || n_vox < 1) return set_error(WDNO_E_INVALID, "chan_layernorm_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define WDNO_LNB(LPV, CPL)                                                                                              \
  {                                                                                                                     \
    const size_t vpb = 256 / LPV;                                                                                       \
    size_t g = (static_cast<size_t>(n_vox) + vpb - 1) / vpb;                                                            \
    const size_t cap = static_cast<size_t>(num_sms()) * 4;                                                              \
    if (g > cap) g = cap;                                                                                               \
    chan_ln_bwd_kernel<LPV, CPL><<<static_cast<unsigned>(g), 256, 0, st>>>(                                             \
        static_cast<const __half*>(x), static_cast<const __half*>(dy), gamma, static_cast<const __half*>(add),          \
        static_cast<__half*>(dx), d_gamma, static_cast<size_t>(n_vox), eps, scale);                                     \
  }
  switch (C) {
    case 32: WDNO_LNB(4, 1); break;
    case 64: WDNO_LNB(8, 1); break;
    case 128: WDNO_LNB(16, 1); break;
    case 256: WDNO_LNB(32, 1); break;
    case 512: WDNO_LNB(32, 2); break;
    case 1024: WDNO_LNB(32, 4); break;
    default: return set_error(WDNO_E_INVALID, "chan_layernorm_bwd: C must be 32/64/128/256/512/1024");
  }
#undef WDNO_LNB
  return check_launch("chan_layernorm_bwd");
}

extern "C" int wdno_sumsq(const float* g, int64_t n, double* out, void* stream) {
  if (!g || !out || n < 1) return set_error(WDNO_E_INVALID, "sumsq: bad arguments");
  sumsq_kernel<<<ew_grid_(static_cast<size_t>(n)), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, static_cast<size_t>(n), out);
  return check_launch("sumsq");
}

extern "C" int wdno_adam_clip_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n, const double* sumsq,
                                  float max_norm, float lr, float beta1, float beta2, float eps, float bc1, float bc2,
                                  float ema_decay, int ema_mode, void* stream) {
  if (!p || !g || !m || !v || n < 1 || (ema_mode != 0 && !ema) || bc1 <= 0.f || bc2 <= 0.f)
    return set_error(WDNO_E_INVALID, "adam_clip_ema: bad arguments");
  adam_clip_ema_kernel<<<ew_grid_(static_cast<size_t>(n)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, ema, static_cast<size_t>(n), sumsq, max_norm, lr, beta1, beta2, eps, bc1, bc2, ema_decay, ema_mode);
  return check_launch("adam_clip_ema");
}

extern "C" int wdno_upsample2x_f16(const void* x, void* y, int64_t N, int H, int W, int C, void* stream) {
  if (!x || !y || N < 1 || H < 1 || W < 1 || C < 8 || (C & 7)) return set_error(WDNO_E_INVALID, "upsample2x_f16: bad arguments");
  const size_t total = static_cast<size_t>(N) * H * W * (C >> 3);
  up2x_kernel<<<ew_grid_(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), H, W,
                                                                          C >> 3, total);
  return check_launch("upsample2x_f16");
}

extern "C" int wdno_sumpool2x2_f16(const void* y, void* x, int64_t N, int H, int W, int C, void* stream) {
  if (!x || !y || N < 1 || H < 1 || W < 1 || C < 8 || (C & 7)) return set_error(WDNO_E_INVALID, "sumpool2x2_f16: bad arguments");
  const size_t total = static_cast<size_t>(N) * H * W * (C >> 3);
  sumpool2x2_kernel<<<ew_grid_(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(y), static_cast<uint4*>(x),
                                                                                H, W, C >> 3, total);
  return check_launch("sumpool2x2_f16");
}
