// Fused one-level 2-D DWT / IDWT of a batch of images: pytorch_wavelets DWTForward / DWTInverse (J levels = J calls), modes
// 'zero' and 'periodization' (call sites eval_ddpm_burgers.py:134-136,188-194,
// test_util.py:200-203, inference_2d.py:178-180,244-246, data_burgers_1d.py:66-68, wave_trans.py:103-108,
// wave_trans_2d.py:135-137).  The separable form (dwt.cu) needs three launches and two intermediate tensors per level
// (240 KB of traffic for a 156 KB transform of one Burgers sample, and launch-bound at 41 x 60 coefficient planes); here one
// CTA owns (image, strip of output rows), stages its input rows once in shared memory WITH the boundary extension applied
// (zero padding, or periodic wrap with the last sample repeated for odd lengths), runs the row pass into shared memory and
// the column pass out of it, and writes each sub-band as one contiguous run.
//   analysis : out[i] = sum_k X(2 i + k - off) t[k]                       X zero-extended | periodic (odd N: edge repeated)
//   synthesis: y[m]   = sum_{k : (m + off - k) even} C((m + off - k)/2) t[k]     C zero-extended | periodic
// Same per-axis formulas and summation order as dwt.cu: results are bit-identical to the separable path.
// Band order (LL | LH, HL, HH) = (lo_w lo_h | lo_w hi_h, hi_w lo_h, hi_w hi_h) (Appendix A.1).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace wdno {

namespace {

struct D2 {
  const float* in[4];        // analysis: in[0] = x ; synthesis: ll, lh, hl, hh
  float* out[4];             // analysis: ll, lh, hl, hh ; synthesis: out[0] = y
  long long in_istride[4];   // image strides (elements)
  long long out_istride[4];
  int H, W;                  // signal extents
  int nh, nw;                // coefficient extents
  int offh, offw, periodic;
  int T;                     // output rows per CTA
  int R, CS;                 // staged rows, staged row stride (multiple of 4 floats)
  int v16;                   // rows of the input(s) are 16-byte aligned: stage with 16-byte cp.async
  float t0[WDNO_MAX_TAPS], t1[WDNO_MAX_TAPS];
};

// signal index an analysis tap reads, or -1 for a zero (any j)
__device__ __forceinline__ int map_analysis(int j, int N, int periodic) {
  if (periodic) {
    const int Np = N + (N & 1);
    j %= Np;
    if (j < 0) j += Np;
    return j >= N ? N - 1 : j;
  }
  return (j >= 0 && j < N) ? j : -1;
}
// the same for -Np < j < 2 Np (one fold; no integer division): the per-element column maps
__device__ __forceinline__ int map_analysis1(int j, int N, int periodic) {
  if (periodic) {
    const int Np = N + (N & 1);
    if (j < 0) j += Np;
    else if (j >= Np) j -= Np;
    return j >= N ? N - 1 : j;
  }
  return (j >= 0 && j < N) ? j : -1;
}

// coefficient index a synthesis tap reads, or -1 for a zero (any i) / for -n < i < 2 n
__device__ __forceinline__ int map_synthesis(int i, int n, int periodic) {
  if (periodic) {
    i %= n;
    return i < 0 ? i + n : i;
  }
  return (i >= 0 && i < n) ? i : -1;
}
__device__ __forceinline__ int map_synthesis1(int i, int n, int periodic) {
  if (periodic) return i < 0 ? i + n : (i >= n ? i - n : i);
  return (i >= 0 && i < n) ? i : -1;
}

__device__ __forceinline__ void cpa4(float* dst_smem, const float* src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa16(float* dst_smem, const float* src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------- analysis
// smem: X [R][CS] (R = 2 T + L - 2 extended rows, column c <-> signal column c - offw, extension applied),
//       LO / HI [R][nw] (row pass).  Staging is fire-and-forget (cp.async: 16-byte chunks for the interior of a row when the
//       rows are 16-byte aligned, 4-byte copies for the extension columns), so a CTA has its whole tile in flight at once.
// tap masks: bit k of M0 / M1 set = taps t0[k] / t1[k] may be non-zero (all bits set = the plain kernel); products with exact
// zeros are compiled out together with the loads that only fed them (bior2.4's analysis high-pass has 3 non-zero taps of 10)
template <unsigned M>
__device__ __forceinline__ float fm(int k, float a, float t, float acc) {
  return ((M >> k) & 1u) ? fmaf(a, t, acc) : acc;
}

template <int L, unsigned M0, unsigned M1>
__global__ void __launch_bounds__(256) dwt2d_analysis_kernel(const D2 p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const long long img = blockIdx.y;
  const int ih0 = blockIdx.x * p.T;
  const int R = p.R, CS = p.CS, nw = p.nw, W = p.W, offw = p.offw;
  float* X = sm;
  float* LO = X + R * CS;
  float* HI = LO + R * nw;
  const float* x = p.in[0] + img * p.in_istride[0];
  for (int r = warp; r < R; r += nwarps) {
    const int gr = map_analysis(2 * ih0 - p.offh + r, p.H, p.periodic);
    float* dst = X + r * CS;
    if (gr < 0) {
      for (int c = lane; c < CS; c += 32) dst[c] = 0.f;
      continue;
    }
    const float* row = x + static_cast<long long>(gr) * W;
    if (p.v16) {
      for (int q = lane; 4 * q < W; q += 32) cpa16(dst + offw + 4 * q, row + 4 * q);
    } else {
      for (int c = lane; c < W; c += 32) cpa4(dst + offw + c, row + c);
    }
    for (int e = lane; e < CS - W; e += 32) {          // extension columns: left of the signal, then right of it
      const int c = e < offw ? e : e + W;
      const int gc = map_analysis1(c - offw, W, p.periodic);
      if (gc >= 0) cpa4(dst + c, row + gc);
      else dst[c] = 0.f;
    }
  }
  cpa_wait_all();
  __syncthreads();
  for (int r = warp; r < R; r += nwarps) {
    const float* row = X + r * CS;
    for (int i = lane; i < nw; i += 32) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int k = 0; k < L; k += 2) {
        const float2 v = *reinterpret_cast<const float2*>(row + 2 * i + k);
        a0 = fm<M0>(k, v.x, p.t0[k], a0);
        a1 = fm<M1>(k, v.x, p.t1[k], a1);
        a0 = fm<M0>(k + 1, v.y, p.t0[k + 1], a0);
        a1 = fm<M1>(k + 1, v.y, p.t1[k + 1], a1);
      }
      LO[r * nw + i] = a0;
      HI[r * nw + i] = a1;
    }
  }
  __syncthreads();
  const int rows = min(p.T, p.nh - ih0);
  for (int ihl = warp; ihl < rows; ihl += nwarps) {
    const long long o = static_cast<long long>(ih0 + ihl) * nw;
    for (int iw = lane; iw < nw; iw += 32) {
      const float* lo = LO + (2 * ihl) * nw + iw;
      const float* hi = HI + (2 * ihl) * nw + iw;
      float ll = 0.f, lh = 0.f, hl = 0.f, hh = 0.f;
#pragma unroll
      for (int k = 0; k < L; ++k) {
        const float a = lo[k * nw], b = hi[k * nw];
        ll = fm<M0>(k, a, p.t0[k], ll);
        lh = fm<M1>(k, a, p.t1[k], lh);
        hl = fm<M0>(k, b, p.t0[k], hl);
        hh = fm<M1>(k, b, p.t1[k], hh);
      }
      p.out[0][img * p.out_istride[0] + o + iw] = ll;
      p.out[1][img * p.out_istride[1] + o + iw] = lh;
      p.out[2][img * p.out_istride[2] + o + iw] = hl;
      p.out[3][img * p.out_istride[3] + o + iw] = hh;
    }
  }
}

// ---------------------------------------------------------------- synthesis
// (The first form of this kernel evaluated the periodic wrap / validity map of every coefficient once per TAP in its row pass:
// 212 instructions per output, profiles/r1e_dwt.md; measured 41.3 us vs 34.6 us for this form on the Burgers shapes,
// profiles/r2_dwt.md, and was removed.)  The column pass writes its result WITH the extension along W
// already applied: LOX / HIX [T][WE], entry e <-> coefficient column e - P, P = (L/2 - 1) - offw/2, WE = W/2 + L/2 - 1 (the map is
// evaluated once per staged column instead of once per tap), so the row pass of output pair q reads entries q + L/2 - 1 - kk
// at compile-time offsets, without maps or branches.  Tap masks as in the analysis kernel.  Same sums in the same order.
template <int L, unsigned M0, unsigned M1>
__global__ void __launch_bounds__(256) dwt2d_synthesis_x_kernel(const D2 p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int Hh = L / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const long long img = blockIdx.y;
  const int m0 = blockIdx.x * p.T;
  const int R = p.R, nw = p.nw, CS = p.CS, W = p.W;
  const int QW = W >> 1, P = (Hh - 1) - (p.offw >> 1);
  const int WE = QW + Hh - 1, WS = (WE + 1) & ~1;          // staged columns, row stride of LOX / HIX
  float* Cc = sm;
  float* LOX = Cc + 4 * R * CS;
  float* HIX = LOX + p.T * WS;
  const int jlo = m0 + p.offh - (L - 1);
  const int i0 = jlo >= 0 ? jlo / 2 : -((1 - jlo) / 2);
  for (int rr = warp; rr < 4 * R; rr += nwarps) {
    const int bd = rr / R, r = rr - bd * R;
    const int gi = map_synthesis(i0 + r, p.nh, p.periodic);
    float* dst = Cc + rr * CS;
    if (gi < 0) {
      for (int c = lane; c < nw; c += 32) dst[c] = 0.f;
      continue;
    }
    const float* row = p.in[bd] + img * p.in_istride[bd] + static_cast<long long>(gi) * nw;
    if (p.v16) {
      for (int q = lane; 4 * q < nw; q += 32) cpa16(dst + 4 * q, row + 4 * q);
    } else {
      for (int c = lane; c < nw; c += 32) cpa4(dst + c, row + c);
    }
  }
  cpa_wait_all();
  __syncthreads();
  const int rows = min(p.T, p.H - m0);
  for (int ml = warp; ml < rows; ml += nwarps) {
    const int j = m0 + ml + p.offh;
    const bool odd = j & 1;
    for (int e = lane; e < WE; e += 32) {
      const int gi = map_synthesis1(e - P, nw, p.periodic);
      float a = 0.f, b = 0.f;
      if (gi >= 0) {
#pragma unroll
        for (int kk = 0; kk < Hh; ++kk) {
          const int r = ((j - 2 * kk) >> 1) - i0;
          const float* c = Cc + r * CS + gi;
          if (((M0 >> (2 * kk)) & 3u) != 0u) {        // either tap of the pair may be non-zero
            const float k0 = odd ? p.t0[2 * kk + 1] : p.t0[2 * kk];
            a = fmaf(c[0], k0, a);
            b = fmaf(c[2 * R * CS], k0, b);
          }
          if (((M1 >> (2 * kk)) & 3u) != 0u) {
            const float k1 = odd ? p.t1[2 * kk + 1] : p.t1[2 * kk];
            a = fmaf(c[R * CS], k1, a);
            b = fmaf(c[3 * R * CS], k1, b);
          }
        }
      }
      LOX[ml * WS + e] = a;
      HIX[ml * WS + e] = b;
    }
  }
  __syncthreads();
  float* y = p.out[0] + img * p.out_istride[0] + static_cast<long long>(m0) * W;
  for (int ml = warp; ml < rows; ml += nwarps) {
    const float* lo = LOX + ml * WS + (Hh - 1);
    const float* hi = HIX + ml * WS + (Hh - 1);
    for (int q = lane; q < QW; q += 32) {
      float e = 0.f, o = 0.f;
#pragma unroll
      for (int kk = 0; kk < Hh; ++kk) {
        const float cl = lo[q - kk], ch = hi[q - kk];
        e = fm<M0>(2 * kk, cl, p.t0[2 * kk], e);
        e = fm<M1>(2 * kk, ch, p.t1[2 * kk], e);
        o = fm<M0>(2 * kk + 1, cl, p.t0[2 * kk + 1], o);
        o = fm<M1>(2 * kk + 1, ch, p.t1[2 * kk + 1], o);
      }
      *reinterpret_cast<float2*>(y + ml * W + 2 * q) = make_float2(e, o);
    }
  }
}

int fill(D2& p, const float* t0, const float* t1, int L) {
  if (!t0 || !t1 || L < 2 || L > WDNO_MAX_TAPS || (L & 1)) return set_error(WDNO_E_INVALID, "dwt2d: taps missing / odd length");
  for (int k = 0; k < WDNO_MAX_TAPS; ++k) {
    p.t0[k] = k < L ? t0[k] : 0.f;
    p.t1[k] = k < L ? t1[k] : 0.f;
  }
  return WDNO_OK;
}

constexpr size_t kSmemBudget = 56 * 1024;   // four CTAs per SM (taller images are cut into row strips)

template <typename K>
int launch2d(K kernel, size_t& cfg, dim3 grid, size_t smem, cudaStream_t st, const D2& p, const char* where) {
  if (smem > cfg) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget));
    if (e != cudaSuccess) return set_cuda_error(e, where);
    cfg = kSmemBudget;
  }
  kernel<<<grid, 256, smem, st>>>(p);
  return check_launch(where);
}

unsigned nz_mask2(const float* t, int L) {
  unsigned m = 0;
  for (int k = 0; k < L; ++k)
    if (t[k] != 0.f) m |= 1u << k;
  return m;
}

}  // namespace

}  // namespace wdno

using namespace wdno;

/* 1 if the fused 2-D kernels cover the geometry (tap count 2 / 6 / 10, a strip of at least one output row fits shared memory,
 * the periodic wrap needs at most one fold: extents >= L), else 0 (use the per-axis entry points) */
extern "C" int wdno_dwt2d_supported(int L, int H, int W, int nh, int nw, int periodic) {
  if (L != 2 && L != 6 && L != 10) return 0;
  if (H < 1 || W < 1 || nh < 1 || nw < 1) return 0;
  if (periodic && (H < L || W < L || nh < L || nw < L)) return 0;   // column maps fold at most once
  const size_t cs = static_cast<size_t>((2 * nw + L - 2 + 3) & ~3);
  const size_t ana = sizeof(float) * (static_cast<size_t>(L) * cs + 2ull * L * nw);              // T = 1
  const size_t syn = sizeof(float) * (4ull * (L / 2 + 2) * ((nw + 3) & ~3) + 2ull * 2 * nw);     // T = 2
  return (ana <= kSmemBudget && syn <= kSmemBudget) ? 1 : 0;
}

extern "C" int wdno_dwt2d_analysis(const float* x, float* const* bands4, const int64_t* band_istride4, int64_t n_img, int H, int W,
                                   int nh, int nw, const float* taps_lo_host, const float* taps_hi_host, int L, int offh, int offw,
                                   int periodic, void* stream) {
  if (!x || !bands4 || !band_istride4 || n_img < 1 || n_img > 65535 || !wdno_dwt2d_supported(L, H, W, nh, nw, periodic))
    return set_error(WDNO_E_INVALID, "dwt2d_analysis: unsupported shape / taps");
  D2 p = {};
  int rc = fill(p, taps_lo_host, taps_hi_host, L);
  if (rc) return rc;
  p.in[0] = x;
  p.in_istride[0] = static_cast<long long>(H) * W;
  for (int i = 0; i < 4; ++i) {
    if (!bands4[i]) return set_error(WDNO_E_INVALID, "dwt2d_analysis: null band");
    p.out[i] = bands4[i];
    p.out_istride[i] = band_istride4[i];
  }
  p.H = H; p.W = W; p.nh = nh; p.nw = nw; p.offh = offh; p.offw = offw; p.periodic = periodic;
  p.CS = (2 * nw + L - 2 + 3) & ~3;
  if (p.CS < offw + W) p.CS = (offw + W + 3) & ~3;
  p.v16 = (!(W & 3) && !(offw & 3) && !(reinterpret_cast<uintptr_t>(x) & 15)) ? 1 : 0;
  // largest strip that fits: (2 T + L - 2) (CS + 2 nw) floats
  int T = nh;
  while (T > 1 && sizeof(float) * static_cast<size_t>(2 * T + L - 2) * (p.CS + 2 * nw) > kSmemBudget) T = (T + 1) / 2;
  const int strips = (nh + T - 1) / T;
  T = (nh + strips - 1) / strips;
  p.T = T;
  p.R = 2 * T + L - 2;
  const size_t smem = sizeof(float) * static_cast<size_t>(p.R) * (p.CS + 2 * nw);
  dim3 grid(strips, static_cast<unsigned>(n_img));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static size_t c2 = 0, c6 = 0, c10 = 0, m6 = 0, m10 = 0;
  {   // tap-mask variants for the filter pairs the reference uses (reversed bior1.3 / bior2.4 analysis taps: 2 / 3 non-zero high-pass taps)
    const unsigned nz1 = nz_mask2(taps_hi_host, L);
    if (L == 6 && !(nz1 & ~0x0Cu)) return launch2d(dwt2d_analysis_kernel<6, 0x3Fu, 0x0Cu>, m6, grid, smem, st, p, "dwt2d_analysis");
    if (L == 10 && !(nz1 & ~0x70u)) return launch2d(dwt2d_analysis_kernel<10, 0x3FFu, 0x70u>, m10, grid, smem, st, p, "dwt2d_analysis");
  }
  if (L == 6) return launch2d(dwt2d_analysis_kernel<6, 0x3Fu, 0x3Fu>, c6, grid, smem, st, p, "dwt2d_analysis");
  if (L == 10) return launch2d(dwt2d_analysis_kernel<10, 0x3FFu, 0x3FFu>, c10, grid, smem, st, p, "dwt2d_analysis");
  return launch2d(dwt2d_analysis_kernel<2, 0x3u, 0x3u>, c2, grid, smem, st, p, "dwt2d_analysis");
}

extern "C" int wdno_dwt2d_synthesis(const float* const* bands4, const int64_t* band_istride4, float* y, int64_t n_img, int nh, int nw,
                                    int H, int W, const float* taps_lo_host, const float* taps_hi_host, int L, int offh, int offw,
                                    int periodic, void* stream) {
  if (!y || !bands4 || !band_istride4 || n_img < 1 || n_img > 65535 || !wdno_dwt2d_supported(L, H, W, nh, nw, periodic) || (W & 1) ||
      (offw & 1) || (reinterpret_cast<uintptr_t>(y) & 7))
    return set_error(WDNO_E_INVALID, "dwt2d_synthesis: unsupported shape / taps (W and offw must be even)");
  D2 p = {};
  int rc = fill(p, taps_lo_host, taps_hi_host, L);
  if (rc) return rc;
  for (int i = 0; i < 4; ++i) {
    if (!bands4[i]) return set_error(WDNO_E_INVALID, "dwt2d_synthesis: null band");
    p.in[i] = bands4[i];
    p.in_istride[i] = band_istride4[i];
  }
  p.out[0] = y;
  p.out_istride[0] = static_cast<long long>(H) * W;
  p.H = H; p.W = W; p.nh = nh; p.nw = nw; p.offh = offh; p.offw = offw; p.periodic = periodic;
  // T output rows read coefficient rows floor((m0 + offh - L + 1) / 2) .. floor((m0 + T - 1 + offh) / 2): at most T / 2 + L / 2 + 1
  int T = H;
  p.CS = (nw + 3) & ~3;
  p.v16 = !(nw & 3) ? 1 : 0;
  for (int i = 0; i < 4; ++i)
    if ((reinterpret_cast<uintptr_t>(bands4[i]) & 15) || (band_istride4[i] & 3)) p.v16 = 0;
  const size_t ws = static_cast<size_t>(((W >> 1) + L / 2 - 1 + 1) & ~1);   // column-pass row stride (extension staged)
  auto need = [&](int t) { return sizeof(float) * (4ull * (t / 2 + L / 2 + 1) * p.CS + 2ull * t * ws); };
  while (T > 2 && need(T) > kSmemBudget) T = (T + 1) / 2;
  const int strips = (H + T - 1) / T;
  T = (H + strips - 1) / strips;
  p.T = T;
  p.R = T / 2 + L / 2 + 1;
  const size_t smem = need(T);
  dim3 grid(strips, static_cast<unsigned>(n_img));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static size_t x2 = 0, x6 = 0, x10 = 0, m6 = 0, m10 = 0;
  {   // tap masks for the reference's reconstruction filters
    const unsigned nz0 = nz_mask2(taps_lo_host, L);
    if (L == 6 && !(nz0 & ~0x0Cu)) return launch2d(dwt2d_synthesis_x_kernel<6, 0x0Cu, 0x3Fu>, m6, grid, smem, st, p, "dwt2d_synthesis");
    if (L == 10 && !(nz0 & ~0x38u)) return launch2d(dwt2d_synthesis_x_kernel<10, 0x38u, 0x3FFu>, m10, grid, smem, st, p, "dwt2d_synthesis");
    if (L == 6) return launch2d(dwt2d_synthesis_x_kernel<6, 0x3Fu, 0x3Fu>, x6, grid, smem, st, p, "dwt2d_synthesis");
    if (L == 10) return launch2d(dwt2d_synthesis_x_kernel<10, 0x3FFu, 0x3FFu>, x10, grid, smem, st, p, "dwt2d_synthesis");
    return launch2d(dwt2d_synthesis_x_kernel<2, 0x3u, 0x3u>, x2, grid, smem, st, p, "dwt2d_synthesis");
  }
}
