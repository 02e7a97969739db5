// Separable DWT / IDWT passes along one axis of an fp32 tensor viewed as [outer][N][inner].
// One generic analysis kernel and one synthesis kernel cover every transform the reference calls
// (pytorch_wavelets afb1d/sfb1d in modes 'zero' and 'periodization', ptwt 0.1.6 wavedec3/waverec3 'zero';
//  SURVEY.md Appendix A.1-A.3; call sites inference_2d.py:37-46,141-147,178-186,220-254,
//  eval_ddpm_burgers.py:134-136,188-194, test_util.py:186-203) and their adjoints (guidance gradients):
//   analysis : out[i] = sum_k X(2i + k - off) * t[k]          X = zero-extended or periodic (odd N: last sample repeated)
//   synthesis: y[m]   = sum_{k: (m+off-k) even} C((m+off-k)/2) * t[k]   C = zero-extended or periodic
// HBM-bound: every thread produces one lo/hi pair (analysis) or one sample (synthesis); threads run along `inner`
// (coalesced) or, for the innermost axis, along the output index.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace wdno {

struct Taps {
  float t0[WDNO_MAX_TAPS];
  float t1[WDNO_MAX_TAPS];
  int L;
};

__global__ void dwt_analysis_kernel(const float* __restrict__ x, float* __restrict__ lo, float* __restrict__ hi, long long outer,
                                    int N, long long inner, int nout, long long x_ostride, long long lo_ostride, long long hi_ostride,
                                    Taps tp, int off, int periodic) {
  const long long total = outer * nout * inner;
  const int Np = N + (N & 1);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long in_i = idx % inner;
    const long long r = idx / inner;
    const int i = static_cast<int>(r % nout);
    const long long o = r / nout;
    const float* xb = x + o * x_ostride + in_i;
    float a0 = 0.f, a1 = 0.f;
    const int j0 = 2 * i - off;
#pragma unroll 1
    for (int k = 0; k < tp.L; ++k) {
      int j = j0 + k;
      float v;
      if (periodic) {
        j %= Np;
        if (j < 0) j += Np;
        if (j >= N) j = N - 1;
        v = xb[static_cast<long long>(j) * inner];
      } else {
        v = (j >= 0 && j < N) ? xb[static_cast<long long>(j) * inner] : 0.f;
      }
      a0 = fmaf(v, tp.t0[k], a0);
      a1 = fmaf(v, tp.t1[k], a1);
    }
    const long long oi = static_cast<long long>(i) * inner + in_i;
    lo[o * lo_ostride + oi] = a0;
    hi[o * hi_ostride + oi] = a1;
  }
}

__global__ void dwt_synthesis_kernel(const float* __restrict__ lo, const float* __restrict__ hi, float* __restrict__ y,
                                     long long outer, int n, long long inner, int Nout, long long lo_ostride,
                                     long long hi_ostride, long long y_ostride, Taps tp, int off, int periodic) {
  const long long total = outer * Nout * inner;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long in_i = idx % inner;
    const long long r = idx / inner;
    const int m = static_cast<int>(r % Nout);
    const long long o = r / Nout;
    const float* lb = lo + o * lo_ostride + in_i;
    const float* hb = hi + o * hi_ostride + in_i;
    float acc = 0.f;
    const int j = m + off;
#pragma unroll 1
    for (int k = (j & 1); k < tp.L; k += 2) {
      int i = (j - k) / 2;  // j-k even; may be negative
      if ((j - k) < 0) i = -((k - j) / 2);
      if (periodic) {
        i %= n;
        if (i < 0) i += n;
      } else if (i < 0 || i >= n) {
        continue;
      }
      acc = fmaf(lb[static_cast<long long>(i) * inner], tp.t0[k], acc);
      acc = fmaf(hb[static_cast<long long>(i) * inner], tp.t1[k], acc);
    }
    y[o * y_ostride + static_cast<long long>(m) * inner + in_i] = acc;
  }
}

// ---------------------------------------------------------------- fast paths: compile-time tap count, 32-bit indices, float4 along
// `inner` (axes other than the innermost one: threads run along the contiguous inner index, 16 bytes each, every tap
// load is a coalesced 512-byte warp request) -- the generic kernels above walk a runtime tap loop with 64-bit div/mod.
__device__ __forceinline__ float4 fma4(float4 v, float t, float4 a) {
  return make_float4(fmaf(v.x, t, a.x), fmaf(v.y, t, a.y), fmaf(v.z, t, a.z), fmaf(v.w, t, a.w));
}

template <int L, bool PERIODIC>
__global__ void __launch_bounds__(256) dwt_analysis_v4_kernel(const float* __restrict__ x, float* __restrict__ lo, float* __restrict__ hi,
                                                              unsigned total, int N, unsigned inner4, int nout, long long x_ostride,
                                                              long long lo_ostride, long long hi_ostride, Taps tp, int off) {
  const int Np = N + (N & 1);
  const size_t inner = static_cast<size_t>(inner4) * 4;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned in4 = idx % inner4;
    const unsigned r = idx / inner4;
    const int i = static_cast<int>(r % static_cast<unsigned>(nout));
    const unsigned o = r / static_cast<unsigned>(nout);
    const float* xb = x + o * x_ostride + in4 * 4;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    const int j0 = 2 * i - off;
    float4 v[L];
#pragma unroll
    for (int k = 0; k < L; ++k) {
      int j = j0 + k;
      bool ok = true;
      if (PERIODIC) {
        j %= Np;
        if (j < 0) j += Np;
        if (j >= N) j = N - 1;
      } else {
        ok = (j >= 0) && (j < N);
      }
      v[k] = ok ? __ldg(reinterpret_cast<const float4*>(xb + static_cast<size_t>(j) * inner)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < L; ++k) {
      a0 = fma4(v[k], tp.t0[k], a0);
      a1 = fma4(v[k], tp.t1[k], a1);
    }
    const size_t oi = static_cast<size_t>(i) * inner + in4 * 4;
    *reinterpret_cast<float4*>(lo + o * lo_ostride + oi) = a0;
    *reinterpret_cast<float4*>(hi + o * hi_ostride + oi) = a1;
  }
}

template <int L, bool PERIODIC>
__global__ void __launch_bounds__(256) dwt_synthesis_v4_kernel(const float* __restrict__ lo, const float* __restrict__ hi,
                                                               float* __restrict__ y, unsigned total, int n, unsigned inner4, int Nout,
                                                               long long lo_ostride, long long hi_ostride, long long y_ostride, Taps tp,
                                                               int off) {
  const size_t inner = static_cast<size_t>(inner4) * 4;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned in4 = idx % inner4;
    const unsigned r = idx / inner4;
    const int m = static_cast<int>(r % static_cast<unsigned>(Nout));
    const unsigned o = r / static_cast<unsigned>(Nout);
    const float* lb = lo + o * lo_ostride + in4 * 4;
    const float* hb = hi + o * hi_ostride + in4 * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int j = m + off;
    const int par = j & 1;
#pragma unroll
    for (int kk = 0; kk < (L + 1) / 2; ++kk) {
      const int k = 2 * kk + par;       // taps with the parity of j
      if (k < L) {
        int i = (j - k) >> 1;           // j - k is even: arithmetic shift is exact for negatives too
        bool ok = true;
        if (PERIODIC) {
          i %= n;
          if (i < 0) i += n;
        } else {
          ok = (i >= 0) && (i < n);
        }
        if (ok) {
          acc = fma4(__ldg(reinterpret_cast<const float4*>(lb + static_cast<size_t>(i) * inner)), tp.t0[k], acc);
          acc = fma4(__ldg(reinterpret_cast<const float4*>(hb + static_cast<size_t>(i) * inner)), tp.t1[k], acc);
        }
      }
    }
    *reinterpret_cast<float4*>(y + o * y_ostride + static_cast<size_t>(m) * inner + in4 * 4) = acc;
  }
}

// scalar variants (innermost axis, or unaligned views): compile-time taps, 32-bit indices; consecutive threads take
// consecutive `inner` / output indices, so a warp's tap loads cover one contiguous window
template <int L, bool PERIODIC>
__global__ void __launch_bounds__(256) dwt_analysis_t_kernel(const float* __restrict__ x, float* __restrict__ lo, float* __restrict__ hi,
                                                             unsigned total, int N, unsigned inner, int nout, long long x_ostride,
                                                             long long lo_ostride, long long hi_ostride, Taps tp, int off) {
  const int Np = N + (N & 1);
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned in_i = idx % inner;
    const unsigned r = idx / inner;
    const int i = static_cast<int>(r % static_cast<unsigned>(nout));
    const unsigned o = r / static_cast<unsigned>(nout);
    const float* xb = x + o * x_ostride + in_i;
    float a0 = 0.f, a1 = 0.f;
    const int j0 = 2 * i - off;
    float v[L];
#pragma unroll
    for (int k = 0; k < L; ++k) {
      int j = j0 + k;
      bool ok = true;
      if (PERIODIC) {
        j %= Np;
        if (j < 0) j += Np;
        if (j >= N) j = N - 1;
      } else {
        ok = (j >= 0) && (j < N);
      }
      v[k] = ok ? __ldg(xb + static_cast<size_t>(j) * inner) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < L; ++k) {
      a0 = fmaf(v[k], tp.t0[k], a0);
      a1 = fmaf(v[k], tp.t1[k], a1);
    }
    const size_t oi = static_cast<size_t>(i) * inner + in_i;
    lo[o * lo_ostride + oi] = a0;
    hi[o * hi_ostride + oi] = a1;
  }
}

template <int L, bool PERIODIC>
__global__ void __launch_bounds__(256) dwt_synthesis_t_kernel(const float* __restrict__ lo, const float* __restrict__ hi,
                                                              float* __restrict__ y, unsigned total, int n, unsigned inner, int Nout,
                                                              long long lo_ostride, long long hi_ostride, long long y_ostride, Taps tp,
                                                              int off) {
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const unsigned in_i = idx % inner;
    const unsigned r = idx / inner;
    const int m = static_cast<int>(r % static_cast<unsigned>(Nout));
    const unsigned o = r / static_cast<unsigned>(Nout);
    const float* lb = lo + o * lo_ostride + in_i;
    const float* hb = hi + o * hi_ostride + in_i;
    float acc = 0.f;
    const int j = m + off;
    const int par = j & 1;
#pragma unroll
    for (int kk = 0; kk < (L + 1) / 2; ++kk) {
      const int k = 2 * kk + par;
      if (k < L) {
        int i = (j - k) >> 1;
        bool ok = true;
        if (PERIODIC) {
          i %= n;
          if (i < 0) i += n;
        } else {
          ok = (i >= 0) && (i < n);
        }
        if (ok) {
          acc = fmaf(__ldg(lb + static_cast<size_t>(i) * inner), tp.t0[k], acc);
          acc = fmaf(__ldg(hb + static_cast<size_t>(i) * inner), tp.t1[k], acc);
        }
      }
    }
    y[o * y_ostride + static_cast<size_t>(m) * inner + in_i] = acc;
  }
}

static bool v4_ok(long long total4, long long inner, long long s0, long long s1, long long s2, const void* p0, const void* p1,
                  const void* p2) {
  return inner >= 4 && (inner % 4) == 0 && (s0 % 4) == 0 && (s1 % 4) == 0 && (s2 % 4) == 0 && total4 < (1ll << 31) &&
         ((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2)) & 15) == 0;
}

static int fill_taps(Taps* tp, const float* t0, const float* t1, int L) {
  if (!t0 || !t1 || L < 2 || L > WDNO_MAX_TAPS) return set_error(WDNO_E_INVALID, "dwt: 2 <= L <= WDNO_MAX_TAPS taps required");
  tp->L = L;
  for (int k = 0; k < WDNO_MAX_TAPS; ++k) {
    tp->t0[k] = k < L ? t0[k] : 0.f;
    tp->t1[k] = k < L ? t1[k] : 0.f;
  }
  return WDNO_OK;
}

static int grid_for(long long total) {
  long long g = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? g : cap);
}

}  // namespace wdno

using namespace wdno;

extern "C" int wdno_dwt_analysis_axis(const float* x, float* lo, float* hi, int64_t outer, int N, int64_t inner, int nout,
                                      int64_t x_ostride, int64_t lo_ostride, int64_t hi_ostride, const float* taps_lo_host,
                                      const float* taps_hi_host, int L, int off, int periodic, void* stream) {
  if (!x || !lo || !hi || outer < 1 || N < 1 || inner < 1 || nout < 1) return set_error(WDNO_E_INVALID, "dwt_analysis: bad arguments");
  Taps tp;
  int rc = fill_taps(&tp, taps_lo_host, taps_hi_host, L);
  if (rc) return rc;
  const long long total = static_cast<long long>(outer) * nout * inner;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (v4_ok(total / 4, inner, x_ostride, lo_ostride, hi_ostride, x, lo, hi) && (L == 6 || L == 10 || L == 2)) {
    const unsigned t4 = static_cast<unsigned>(total / 4), i4 = static_cast<unsigned>(inner / 4);
    const int grid = grid_for(total / 4);
#define WDNO_DWT_A(LL, PP) dwt_analysis_v4_kernel<LL, PP><<<grid, 256, 0, st>>>(x, lo, hi, t4, N, i4, nout, x_ostride, lo_ostride, hi_ostride, tp, off)
    if (L == 6) { if (periodic) WDNO_DWT_A(6, true); else WDNO_DWT_A(6, false); }
    else if (L == 10) { if (periodic) WDNO_DWT_A(10, true); else WDNO_DWT_A(10, false); }
    else { if (periodic) WDNO_DWT_A(2, true); else WDNO_DWT_A(2, false); }
#undef WDNO_DWT_A
    return check_launch("dwt_analysis");
  }
  if (total < (1ll << 31) && inner < (1ll << 31) && (L == 6 || L == 10 || L == 2)) {
    const unsigned tt = static_cast<unsigned>(total), ii = static_cast<unsigned>(inner);
    const int grid = grid_for(total);
#define WDNO_DWT_A(LL, PP) dwt_analysis_t_kernel<LL, PP><<<grid, 256, 0, st>>>(x, lo, hi, tt, N, ii, nout, x_ostride, lo_ostride, hi_ostride, tp, off)
    if (L == 6) { if (periodic) WDNO_DWT_A(6, true); else WDNO_DWT_A(6, false); }
    else if (L == 10) { if (periodic) WDNO_DWT_A(10, true); else WDNO_DWT_A(10, false); }
    else { if (periodic) WDNO_DWT_A(2, true); else WDNO_DWT_A(2, false); }
#undef WDNO_DWT_A
    return check_launch("dwt_analysis");
  }
  dwt_analysis_kernel<<<grid_for(total), 256, 0, st>>>(x, lo, hi, outer, N, inner, nout, x_ostride, lo_ostride, hi_ostride, tp, off,
                                                      periodic);
  return check_launch("dwt_analysis");
}

extern "C" int wdno_dwt_synthesis_axis(const float* lo, const float* hi, float* y, int64_t outer, int n, int64_t inner,
                                       int Nout, int64_t lo_ostride, int64_t hi_ostride, int64_t y_ostride, const float* taps_lo_host,
                                       const float* taps_hi_host, int L, int off, int periodic, void* stream) {
  if (!lo || !hi || !y || outer < 1 || n < 1 || inner < 1 || Nout < 1) return set_error(WDNO_E_INVALID, "dwt_synthesis: bad arguments");
  Taps tp;
  int rc = fill_taps(&tp, taps_lo_host, taps_hi_host, L);
  if (rc) return rc;
  const long long total = static_cast<long long>(outer) * Nout * inner;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (v4_ok(total / 4, inner, lo_ostride, hi_ostride, y_ostride, lo, hi, y) && (L == 6 || L == 10 || L == 2)) {
    const unsigned t4 = static_cast<unsigned>(total / 4), i4 = static_cast<unsigned>(inner / 4);
    const int grid = grid_for(total / 4);
#define WDNO_DWT_S(LL, PP) dwt_synthesis_v4_kernel<LL, PP><<<grid, 256, 0, st>>>(lo, hi, y, t4, n, i4, Nout, lo_ostride, hi_ostride, y_ostride, tp, off)
    if (L == 6) { if (periodic) WDNO_DWT_S(6, true); else WDNO_DWT_S(6, false); }
    else if (L == 10) { if (periodic) WDNO_DWT_S(10, true); else WDNO_DWT_S(10, false); }
    else { if (periodic) WDNO_DWT_S(2, true); else WDNO_DWT_S(2, false); }
#undef WDNO_DWT_S
    return check_launch("dwt_synthesis");
  }
  if (total < (1ll << 31) && inner < (1ll << 31) && (L == 6 || L == 10 || L == 2)) {
    const unsigned tt = static_cast<unsigned>(total), ii = static_cast<unsigned>(inner);
    const int grid = grid_for(total);
#define WDNO_DWT_S(LL, PP) dwt_synthesis_t_kernel<LL, PP><<<grid, 256, 0, st>>>(lo, hi, y, tt, n, ii, Nout, lo_ostride, hi_ostride, y_ostride, tp, off)
    if (L == 6) { if (periodic) WDNO_DWT_S(6, true); else WDNO_DWT_S(6, false); }
    else if (L == 10) { if (periodic) WDNO_DWT_S(10, true); else WDNO_DWT_S(10, false); }
    else { if (periodic) WDNO_DWT_S(2, true); else WDNO_DWT_S(2, false); }
#undef WDNO_DWT_S
    return check_launch("dwt_synthesis");
  }
  dwt_synthesis_kernel<<<grid_for(total), 256, 0, st>>>(lo, hi, y, outer, n, inner, Nout, lo_ostride, hi_ostride, y_ostride, tp, off,
                                                       periodic);
  return check_launch("dwt_synthesis");
}
