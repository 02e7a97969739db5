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
  dwt_analysis_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, lo, hi, outer, N, inner, nout,
                                                                                     x_ostride, lo_ostride, hi_ostride, tp, off, periodic);
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
  dwt_synthesis_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(lo, hi, y, outer, n, inner, Nout,
                                                                                      lo_ostride, hi_ostride, y_ostride, tp, off, periodic);
  return check_launch("dwt_synthesis");
}
