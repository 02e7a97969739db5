// Fused single-pass 3-D DWT / IDWT ('zero' mode, level 1): ptwt.wavedec3 / waverec3 of the smoke experiment
// (inference_2d.py:41,141,184,220,250 ; wave_trans_2d.py:129-149 ) and their adjoints, which the
// guided sampler runs EVERY step (design_fn differentiates through waverec3).
//
// The separable form (dwt.cu) needs 7 launches per transform and moves ~2.3x the algorithmic bytes through L2.  Here one
// CTA owns an output tile with the full W extent and keeps all three passes in shared memory:
//   synthesis: 8 coefficient bands [7 x 7 x nw] -> W pass -> 4 x [7 x 7 x Nw] -> H pass -> 2 x [7 x 8 x Nw] -> D pass
//              -> y tile [8 x 8 x Nw]                         (d, h coefficient halos re-read from L2 by the neighbours)
//   analysis : x tile [12 x 12 x Nw] -> W pass -> 2 x [12 x 12 x nw] -> H pass -> 4 x [12 x 4 x nw] -> D pass
//              -> 8 bands [4 x 4 x nw]                        (extents for L = 6)
// with the same per-axis formulas as dwt.cu:
//   analysis : out[i] = sum_k X(2 i + k - off) t[k]                     X zero outside [0, N)
//   synthesis: y[m]   = sum_{k : (m + off - k) even} C((m + off - k)/2) t[k]    C zero outside [0, n)
// Band b = 4 d + 2 h + w (0 = low-pass 'a', 1 = high-pass 'd'): aaa, aad, ada, add, daa, dad, dda, ddd (ptwt key order).
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace wdno {

namespace {

struct T3 {
  float t0[10];
  float t1[10];
};

struct P3 {
  const float* band_in[8];   // synthesis: coefficient bands
  float* band_out[8];        // analysis: coefficient bands
  const float* x;            // analysis input
  float* y;                  // synthesis output
  long long band_bstride, sig_bstride;   // batch strides (elements)
  int nd, nh, nw;            // coefficient extents
  int Nd, Nh, Nw;            // signal extents
  int off;
  T3 tp;
};

// tiles sized for two CTAs per SM (~103 KB / ~76 KB of shared memory at W = 64): the phases of one CTA overlap the other's
constexpr int kSTd = 8, kSTh = 8;    // synthesis: signal tile (d, h); full w
constexpr int kATd = 4, kATh = 4;    // analysis: coefficient tile (d, h); full w

// 4-byte asynchronous copy with zero fill: the tile loads are fire-and-forget (a register-staged load + store serialises
// on the load latency in every iteration of the in-order warp)
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, bool ok) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  const uint32_t n = ok ? 4u : 0u;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ int floordiv2(int a) { return a >> 1; }

// ---------------------------------------------------------------- synthesis
template <int L>
__global__ void __launch_bounds__(512) synth3d_kernel(const P3 p) {
  extern __shared__ float sm[];
  constexpr int CD = kSTd / 2 + (L + 1) / 2, CH = kSTh / 2 + (L + 1) / 2;   // coefficient rows per tile (upper bound)
  const int b = blockIdx.z;
  const int d0 = blockIdx.y * kSTd, h0 = blockIdx.x * kSTh;
  const int id0 = floordiv2(d0 + p.off - (L - 1)), ih0 = floordiv2(h0 + p.off - (L - 1));
  const int nw = p.nw, Nw = p.Nw;
  float* S0 = sm;                               // [8][CD][CH][nw]
  float* SW = S0 + 8 * CD * CH * nw;            // [4][CD][CH][Nw]
  float* SH = S0;                               // [2][CD][kSTh][Nw]   (S0 is dead after the W pass)
  // ---- load the 8 band tiles (zero outside the band)
  // (row = warp-strided, column = lane-strided loops: no per-element div / mod by run-time extents)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < 8 * CD * CH; r += nwarps) {
    const int ih = r % CH, id = (r / CH) % CD, bd = r / (CH * CD);
    const int gd = id0 + id, gh = ih0 + ih;
    const bool ok = gd >= 0 && gd < p.nd && gh >= 0 && gh < p.nh;
    const float* src = p.band_in[bd] + b * p.band_bstride + (static_cast<long long>(ok ? gd : 0) * p.nh + (ok ? gh : 0)) * nw;
    float* dst = S0 + static_cast<size_t>(r) * nw;
    for (int iw = lane; iw < nw; iw += 32) cp_async4(dst + iw, src + iw, ok);
  }
  cp_async_wait_all();
  __syncthreads();
  // ---- W pass: pairs (b even = low, b odd = high) -> 4 planes
  for (int r = warp; r < 4 * CD * CH; r += nwarps) {   // r = (c, id, ih) flattened
    const int c = r / (CD * CH), rem = r - c * (CD * CH);
    const float* lo = S0 + (static_cast<size_t>(2 * c) * CD * CH + rem) * nw;
    const float* hi = lo + static_cast<size_t>(CD) * CH * nw;
    float* dst = SW + static_cast<size_t>(r) * Nw;
    for (int m = lane; m < Nw; m += 32) {
      const int j = m + p.off, par = j & 1;
      float acc = 0.f;
#pragma unroll
      for (int kk = 0; kk < (L + 1) / 2; ++kk) {
        const int k = 2 * kk + par;
        const int i = (j - k) >> 1;
        if (k < L && i >= 0 && i < nw) {
          acc = fmaf(lo[i], p.tp.t0[k], acc);
          acc = fmaf(hi[i], p.tp.t1[k], acc);
        }
      }
      dst[m] = acc;
    }
  }
  __syncthreads();
  // ---- H pass: (c = 2 e) low, (c = 2 e + 1) high -> 2 planes of kSTh rows
  for (int r = warp; r < 2 * CD * kSTh; r += nwarps) {   // r = (dbd, id, mh)
    const int mh = r % kSTh, id = (r / kSTh) % CD, dbd = r / (kSTh * CD);
    const float* lo0 = SW + ((static_cast<size_t>(2 * dbd) * CD + id) * CH) * Nw;
    const float* hi0 = lo0 + static_cast<size_t>(CD) * CH * Nw;
    const int j = h0 + mh + p.off, par = j & 1;
    float* dst = SH + static_cast<size_t>(r) * Nw;
    for (int m = lane; m < Nw; m += 32) {
      float acc = 0.f;
#pragma unroll
      for (int kk = 0; kk < (L + 1) / 2; ++kk) {
        const int k = 2 * kk + par;
        if (k < L) {
          const int i = ((j - k) >> 1) - ih0;   // inside [0, CH) by construction; out-of-band rows were zero-filled
          acc = fmaf(lo0[static_cast<size_t>(i) * Nw + m], p.tp.t0[k], acc);
          acc = fmaf(hi0[static_cast<size_t>(i) * Nw + m], p.tp.t1[k], acc);
        }
      }
      dst[m] = acc;
    }
  }
  __syncthreads();
  // ---- D pass -> global
  for (int r = warp; r < kSTd * kSTh; r += nwarps) {
    const int mh = r % kSTh, md = r / kSTh;
    const int gd = d0 + md, gh = h0 + mh;
    if (gd >= p.Nd || gh >= p.Nh) continue;
    const float* lo0 = SH + static_cast<size_t>(mh) * Nw;                        // [0][id][mh][m]
    const float* hi0 = lo0 + static_cast<size_t>(CD) * kSTh * Nw;                 // [1][id][mh][m]
    const int j = gd + p.off, par = j & 1;
    float* dst = p.y + b * p.sig_bstride + (static_cast<long long>(gd) * p.Nh + gh) * Nw;
    for (int m = lane; m < Nw; m += 32) {
      float acc = 0.f;
#pragma unroll
      for (int kk = 0; kk < (L + 1) / 2; ++kk) {
        const int k = 2 * kk + par;
        if (k < L) {
          const int i = ((j - k) >> 1) - id0;
          acc = fmaf(lo0[static_cast<size_t>(i) * kSTh * Nw + m], p.tp.t0[k], acc);
          acc = fmaf(hi0[static_cast<size_t>(i) * kSTh * Nw + m], p.tp.t1[k], acc);
        }
      }
      dst[m] = acc;
    }
  }
}

// ---------------------------------------------------------------- analysis
template <int L>
__global__ void __launch_bounds__(512) ana3d_kernel(const P3 p) {
  extern __shared__ float sm[];
  constexpr int XD = 2 * kATd + L - 2, XH = 2 * kATh + L - 2;   // signal rows needed per tile
  const int b = blockIdx.z;
  const int i_d0 = blockIdx.y * kATd, i_h0 = blockIdx.x * kATh;
  const int xd0 = 2 * i_d0 - p.off, xh0 = 2 * i_h0 - p.off;
  const int nw = p.nw, Nw = p.Nw;
  float* X = sm;                                // [XD][XH][Nw]
  float* AW = X + XD * XH * Nw;                 // [2][XD][XH][nw]
  float* AH = X;                                // [4][XD][kATh][nw]  (X is dead after the W pass)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < XD * XH; r += nwarps) {
    const int xh = r % XH, xd = r / XH;
    const int gd = xd0 + xd, gh = xh0 + xh;
    const bool ok = gd >= 0 && gd < p.Nd && gh >= 0 && gh < p.Nh;
    const float* src = p.x + b * p.sig_bstride + (static_cast<long long>(ok ? gd : 0) * p.Nh + (ok ? gh : 0)) * Nw;
    float* dst = X + static_cast<size_t>(r) * Nw;
    for (int w = lane; w < Nw; w += 32) cp_async4(dst + w, src + w, ok);
  }
  cp_async_wait_all();
  __syncthreads();
  // ---- W pass: -> (low, high) planes [XD][XH][nw]
  for (int r = warp; r < XD * XH; r += nwarps) {
    const float* row = X + static_cast<size_t>(r) * Nw;
    float* d0p = AW + static_cast<size_t>(r) * nw;
    float* d1p = d0p + static_cast<size_t>(XD) * XH * nw;
    for (int i = lane; i < nw; i += 32) {
      const int j0 = 2 * i - p.off;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int k = 0; k < L; ++k) {
        const int j = j0 + k;
        const float v = (j >= 0 && j < Nw) ? row[j] : 0.f;
        a0 = fmaf(v, p.tp.t0[k], a0);
        a1 = fmaf(v, p.tp.t1[k], a1);
      }
      d0p[i] = a0;
      d1p[i] = a1;
    }
  }
  __syncthreads();
  // ---- H pass: plane wb (w band) -> planes 2*hb + wb, rows kATh
  for (int r = warp; r < 2 * XD * kATh; r += nwarps) {   // r = (wb, xd, ih)
    const int ih = r % kATh, xd = (r / kATh) % XD, wb = r / (kATh * XD);
    const float* col0 = AW + ((static_cast<size_t>(wb) * XD + xd) * XH + 2 * ih) * nw;   // local row 2 ih + k
    const size_t o = (static_cast<size_t>(xd) * kATh + ih) * nw;
    float* lo_p = AH + (static_cast<size_t>(0 * 2 + wb) * XD) * kATh * nw + o;   // h low
    float* hi_p = AH + (static_cast<size_t>(1 * 2 + wb) * XD) * kATh * nw + o;   // h high
    for (int i = lane; i < nw; i += 32) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int k = 0; k < L; ++k) {
        const float v = col0[static_cast<size_t>(k) * nw + i];
        a0 = fmaf(v, p.tp.t0[k], a0);
        a1 = fmaf(v, p.tp.t1[k], a1);
      }
      lo_p[i] = a0;
      hi_p[i] = a1;
    }
  }
  __syncthreads();
  // ---- D pass -> 8 bands
  for (int r = warp; r < 4 * kATd * kATh; r += nwarps) {   // r = (hw, id, ih), hw = 2 hb + wb
    const int ih = r % kATh, id = (r / kATh) % kATd, hw = r / (kATh * kATd);
    const int gd = i_d0 + id, gh = i_h0 + ih;
    if (gd >= p.nd || gh >= p.nh) continue;
    const float* col0 = AH + ((static_cast<size_t>(hw) * XD + 2 * id) * kATh + ih) * nw;
    const long long o = b * p.band_bstride + (static_cast<long long>(gd) * p.nh + gh) * nw;
    float* lo_p = p.band_out[hw] + o;        // d low : band hw
    float* hi_p = p.band_out[4 + hw] + o;    // d high: band 4 + hw
    for (int i = lane; i < nw; i += 32) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int k = 0; k < L; ++k) {
        const float v = col0[static_cast<size_t>(k) * kATh * nw + i];
        a0 = fmaf(v, p.tp.t0[k], a0);
        a1 = fmaf(v, p.tp.t1[k], a1);
      }
      lo_p[i] = a0;
      hi_p[i] = a1;
    }
  }
}

size_t synth_smem(int L, int nw, int Nw) {
  const int CD = kSTd / 2 + (L + 1) / 2, CH = kSTh / 2 + (L + 1) / 2;
  return sizeof(float) * (static_cast<size_t>(8) * CD * CH * nw + static_cast<size_t>(4) * CD * CH * Nw);
}
size_t ana_smem(int L, int nw, int Nw) {
  const int XD = 2 * kATd + L - 2, XH = 2 * kATh + L - 2;
  return sizeof(float) * (static_cast<size_t>(XD) * XH * Nw + static_cast<size_t>(2) * XD * XH * nw);
}

}  // namespace

}  // namespace wdno

using namespace wdno;

/* returns 1 if the fused kernels cover (L, nw, Nw), else 0 */
extern "C" int wdno_dwt3d_supported(int L, int nw, int Nw) {
  if (L != 2 && L != 6 && L != 10) return 0;
  if (nw < 1 || Nw < 1) return 0;
  const int CD = kSTd / 2 + (L + 1) / 2, CH = kSTh / 2 + (L + 1) / 2;
  // SH [2][CD][kSTh][Nw] reuses the S0 region, AH [4][XD][kATh][nw] the X region
  if (static_cast<size_t>(2) * CD * kSTh * Nw > static_cast<size_t>(8) * CD * CH * nw) return 0;
  const int XD = 2 * kATd + L - 2, XH = 2 * kATh + L - 2;
  if (static_cast<size_t>(4) * XD * kATh * nw > static_cast<size_t>(XD) * XH * Nw) return 0;
  return (synth_smem(L, nw, Nw) <= 227 * 1024 && ana_smem(L, nw, Nw) <= 227 * 1024) ? 1 : 0;
}

static int fill(P3& p, const float* t0, const float* t1, int L) {
  if (!t0 || !t1) return set_error(WDNO_E_INVALID, "dwt3d: taps missing");
  for (int k = 0; k < 10; ++k) {
    p.tp.t0[k] = k < L ? t0[k] : 0.f;
    p.tp.t1[k] = k < L ? t1[k] : 0.f;
  }
  return WDNO_OK;
}

extern "C" int wdno_dwt3d_synthesis(const float* const* bands8, int64_t band_bstride, float* y, int64_t B, int nd, int nh, int nw,
                                    int Nd, int Nh, int Nw, const float* taps_lo_host, const float* taps_hi_host, int L, int off,
                                    void* stream) {
  if (!bands8 || !y || B < 1 || B > 65535 || !wdno_dwt3d_supported(L, nw, Nw))
    return set_error(WDNO_E_INVALID, "dwt3d_synthesis: unsupported shape / taps");
  P3 p = {};
  for (int i = 0; i < 8; ++i) {
    if (!bands8[i]) return set_error(WDNO_E_INVALID, "dwt3d_synthesis: null band");
    p.band_in[i] = bands8[i];
  }
  int rc = fill(p, taps_lo_host, taps_hi_host, L);
  if (rc) return rc;
  rc = launch_syn3d_stream(bands8, band_bstride, y, B, nd, nh, nw, Nd, Nh, Nw, taps_lo_host, taps_hi_host, L, off,
                           static_cast<cudaStream_t>(stream));
  if (rc != 0) return rc < 0 ? rc : WDNO_OK;   // 0: outside the streaming kernel's envelope -> tile kernel below
  p.y = y;
  p.band_bstride = band_bstride;
  p.sig_bstride = static_cast<long long>(Nd) * Nh * Nw;
  p.nd = nd; p.nh = nh; p.nw = nw; p.Nd = Nd; p.Nh = Nh; p.Nw = Nw; p.off = off;
  const size_t smem = synth_smem(L, nw, Nw);
  dim3 grid((Nh + kSTh - 1) / kSTh, (Nd + kSTd - 1) / kSTd, static_cast<unsigned>(B));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define WDNO_S3(LL)                                                                                                   \
  {                                                                                                                   \
    static bool cfg = false;                                                                                          \
    if (!cfg) {                                                                                                       \
      cudaError_t e = cudaFuncSetAttribute(synth3d_kernel<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      if (e != cudaSuccess) return set_cuda_error(e, "dwt3d_synthesis: cudaFuncSetAttribute");                        \
      cfg = true;                                                                                                     \
    }                                                                                                                 \
    synth3d_kernel<LL><<<grid, 512, smem, st>>>(p);                                                                   \
  }
  if (L == 6) WDNO_S3(6) else if (L == 10) WDNO_S3(10) else WDNO_S3(2)
#undef WDNO_S3
  return check_launch("dwt3d_synthesis");
}

extern "C" int wdno_dwt3d_analysis(const float* x, float* const* bands8, int64_t band_bstride, int64_t B, int Nd, int Nh, int Nw,
                                   int nd, int nh, int nw, const float* taps_lo_host, const float* taps_hi_host, int L, int off,
                                   void* stream) {
  if (!bands8 || !x || B < 1 || B > 65535 || !wdno_dwt3d_supported(L, nw, Nw))
    return set_error(WDNO_E_INVALID, "dwt3d_analysis: unsupported shape / taps");
  P3 p = {};
  for (int i = 0; i < 8; ++i) {
    if (!bands8[i]) return set_error(WDNO_E_INVALID, "dwt3d_analysis: null band");
    p.band_out[i] = bands8[i];
  }
  int rc = fill(p, taps_lo_host, taps_hi_host, L);
  if (rc) return rc;
  rc = launch_ana3d_stream(x, bands8, band_bstride, B, Nd, Nh, Nw, nd, nh, nw, taps_lo_host, taps_hi_host, L, off,
                           static_cast<cudaStream_t>(stream));
  if (rc != 0) return rc < 0 ? rc : WDNO_OK;   // 0: outside the streaming kernel's envelope -> tile kernel below
  p.x = x;
  p.band_bstride = band_bstride;
  p.sig_bstride = static_cast<long long>(Nd) * Nh * Nw;
  p.nd = nd; p.nh = nh; p.nw = nw; p.Nd = Nd; p.Nh = Nh; p.Nw = Nw; p.off = off;
  const size_t smem = ana_smem(L, nw, Nw);
  dim3 grid((nh + kATh - 1) / kATh, (nd + kATd - 1) / kATd, static_cast<unsigned>(B));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define WDNO_A3(LL)                                                                                                 \
  {                                                                                                                 \
    static bool cfg = false;                                                                                        \
    if (!cfg) {                                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(ana3d_kernel<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
      if (e != cudaSuccess) return set_cuda_error(e, "dwt3d_analysis: cudaFuncSetAttribute");                       \
      cfg = true;                                                                                                   \
    }                                                                                                               \
    ana3d_kernel<LL><<<grid, 512, smem, st>>>(p);                                                                   \
  }
  if (L == 6) WDNO_A3(6) else if (L == 10) WDNO_A3(10) else WDNO_A3(2)
#undef WDNO_A3
  return check_launch("dwt3d_analysis");
}
