// Streaming form of the fused 3-D DWT / IDWT ('zero' mode, level 1; ptwt.wavedec3 / waverec3, SURVEY.md Appendix A.3, call
// sites smoke/inference_2d.py:41,141,184,220,250 ; wave_trans_2d.py:129-149) -- the HBM-bound form of dwt3d.cu.
//
// dwt3d.cu keeps the three passes of a (d, h) tile in shared memory: three block barriers around passes whose 34-wide rows
// leave half of every second warp iteration idle, ~230 us for 95 MB (5 % of the HBM rate).  Here NO intermediate ever
// reaches shared memory.  A CTA owns (sample b, a strip of output rows, a chunk of output planes) and streams the planes of
// its input through a cp.async ring (16-byte chunks for the signal, 8-byte chunks for the 34-wide coefficient rows); a thread
// owns one (h, w) position, does the W and H passes of each arriving plane in registers and carries the D pass as a
// sliding window:
//   analysis : thread (ih, iw): plane xd -> 6 x float2 x 3 loads -> (w-low, w-high) per row -> 4 (h, w)-band values;
//              window of L planes x 4 values; every second plane emits the 8 bands of coefficient plane id.
//   synthesis: thread (qh, qw) owns the 2 x 2 x 2 signal block (2 qd + {0,1}, 2 qh + {0,1}, 2 qw + {0,1}):
//              coefficient plane id -> 8 bands x (L/2 x L/2) loads -> W pass -> H pass -> 2 d-bands x 2 x 2 values;
//              window of L/2 planes; every plane emits 8 signal values (float2 stores).
// Global traffic: every input byte is read once per CTA that needs it (row / plane halos come from L2), every output byte
// is written once, coalesced: the flattened (ih, iw) index of a strip is a contiguous run of the band plane.
// Per-axis formulas and summation order are those of dwt3d.cu / dwt.cu (results are bit-identical to the tile kernels):
//   analysis : out[i] = sum_k X(2 i + k - off) t[k]                     X zero outside [0, N)
//   synthesis: y[2 q + par] = sum_j C(q + L/2 - 1 - j) t[2 j + par]     (off = L - 2)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace wdno {

namespace {

struct S3 {
  const float* band_in[8];
  float* band_out[8];
  const float* x;
  float* y;
  long long band_bstride, sig_bstride;
  int nd, nh, nw, Nd, Nh, Nw;
  int off;
  int TH, TD;        // output rows / output planes per CTA (coefficient units for analysis, signal pairs for synthesis)
  int R, RS;         // tile rows, tile row stride (floats)
  float t0[10], t1[10];
};

constexpr int kStages = 4;

__device__ __forceinline__ void cpa16(float* dst_smem, const float* src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa8(float* dst_smem, const float* src) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- analysis
// grid (h strips, d chunks, B); block >= TH * nw threads.  Tile of one signal plane: R = 2 TH + L - 2 rows of RS floats,
// column c of the tile = signal column c - off (left / right pads and rows outside [0, Nh) stay zero from the initial fill).
template <int L>
__global__ void __launch_bounds__(640) ana3d_stream_kernel(const S3 p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b = blockIdx.z;
  const int ih0 = blockIdx.x * p.TH, id0 = blockIdx.y * p.TD;
  const int id1 = min(id0 + p.TD, p.nd);
  const int nw = p.nw, Nw = p.Nw, R = p.R, RS = p.RS;
  const int PT = R * RS;
  for (int i = tid; i < kStages * PT; i += nthr) sm[i] = 0.f;
  // rows of the tile that exist in the signal: tile row r <-> signal row xr0 + r
  const int xr0 = 2 * ih0 - p.off;
  const int r_lo = max(0, -xr0), r_hi = min(R, p.Nh - xr0);
  const int nq = Nw >> 2;                                   // 16-byte chunks per row
  const int nchunk = max(0, r_hi - r_lo) * nq;
  const float* xb = p.x + b * p.sig_bstride + static_cast<long long>(xr0 + r_lo) * Nw;
  const int NP = 2 * (id1 - id0) + L - 2;                   // planes this CTA streams
  const int gd0 = 2 * id0 - p.off;                          // signal plane of step 0
  // this thread's position
  const int pos = tid < p.TH * nw ? tid : 0;
  const int ihl = pos / nw, iw = pos - ihl * nw;
  const int ih = ih0 + ihl;
  const bool live = tid < p.TH * nw && ih < p.nh;
  const int tbase = (2 * ihl) * RS + 2 * iw;
  __syncthreads();

  auto issue = [&](int s) {
    const int gd = gd0 + s;
    if (s < NP && gd >= 0 && gd < p.Nd) {
      float* dst = sm + (s % kStages) * PT + r_lo * RS + p.off;
      const float* src = xb + static_cast<long long>(gd) * p.Nh * Nw;
      for (int c = tid; c < nchunk; c += nthr) {
        const int r = c / nq, q = c - r * nq;
        cpa16(dst + r * RS + 4 * q, src + static_cast<long long>(r) * Nw + 4 * q);
      }
    }
    cpa_commit();
  };
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) issue(s);

  float win[4][L];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < L; ++k) win[c][k] = 0.f;

  for (int s = 0; s < NP; ++s) {
    cpa_wait<kStages - 2>();
    __syncthreads();              // plane s has landed for everyone; everyone is done with plane s - 1 (its slot is refilled next)
    issue(s + kStages - 1);
    const int gd = gd0 + s;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    if (gd >= 0 && gd < p.Nd) {
      const float* tp = sm + (s % kStages) * PT + tbase;
#pragma unroll
      for (int kh = 0; kh < L; ++kh) {
        float v[L];
#pragma unroll
        for (int j = 0; j < L / 2; ++j) {
          const float2 f = *reinterpret_cast<const float2*>(tp + kh * RS + 2 * j);
          v[2 * j] = f.x;
          v[2 * j + 1] = f.y;
        }
        float wl = 0.f, wh = 0.f;
#pragma unroll
        for (int k = 0; k < L; ++k) {
          wl = fmaf(v[k], p.t0[k], wl);
          wh = fmaf(v[k], p.t1[k], wh);
        }
        a[0] = fmaf(wl, p.t0[kh], a[0]);   // hw = 2 hb + wb
        a[1] = fmaf(wh, p.t0[kh], a[1]);
        a[2] = fmaf(wl, p.t1[kh], a[2]);
        a[3] = fmaf(wh, p.t1[kh], a[3]);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int k = 0; k < L - 1; ++k) win[c][k] = win[c][k + 1];
      win[c][L - 1] = a[c];
    }
    const int e = s - (L - 1);
    if (e >= 0 && !(e & 1) && live) {   // window = planes 2 m .. 2 m + L - 1 of the chunk: coefficient plane id0 + m
      const int gid = id0 + (e >> 1);
      const long long o = b * p.band_bstride + (static_cast<long long>(gid) * p.nh + ih) * nw + iw;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float lo = 0.f, hi = 0.f;
#pragma unroll
        for (int k = 0; k < L; ++k) {
          lo = fmaf(win[c][k], p.t0[k], lo);
          hi = fmaf(win[c][k], p.t1[k], hi);
        }
        p.band_out[c][o] = lo;
        p.band_out[4 + c][o] = hi;
      }
    }
  }
  cpa_wait<0>();
}

// ---------------------------------------------------------------- synthesis
// grid (h strips, d chunks, B); block >= TH * (Nw / 2) threads.  Tile of one coefficient plane: 8 bands x R = TH + L/2 - 1
// rows x nw floats (a flat copy of the contiguous run of each band plane; rows at or beyond nh stay zero).
template <int L>
__global__ void __launch_bounds__(640) syn3d_stream_kernel(const S3 p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int H = L / 2;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b = blockIdx.z;
  const int qh0 = blockIdx.x * p.TH, qd0 = blockIdx.y * p.TD;
  const int QD = (p.Nd + 1) >> 1, QW = p.Nw >> 1;
  const int qd1 = min(qd0 + p.TD, QD);
  const int nw = p.nw, R = p.R;
  const int BT = R * nw, PT = 8 * BT;
  for (int i = tid; i < kStages * PT; i += nthr) sm[i] = 0.f;
  const int r_hi = min(R, p.nh - qh0);
  const int nchunk = max(0, r_hi) * (nw >> 1);              // 8-byte chunks per band
  const long long run0 = b * p.band_bstride + static_cast<long long>(qh0) * nw;
  const int NP = (qd1 - qd0) + H - 1;
  const int pos = tid < p.TH * QW ? tid : 0;
  const int qhl = pos / QW, qw = pos - qhl * QW;
  const int qh = qh0 + qhl;
  const bool live = tid < p.TH * QW && 2 * qh < p.Nh;
  const int tbase = qhl * nw + qw;
  __syncthreads();

  auto issue = [&](int s) {
    const int gid = qd0 + s;
    if (s < NP && gid < p.nd) {
      float* dst = sm + (s % kStages) * PT;
      const long long o = run0 + static_cast<long long>(gid) * p.nh * nw;
#pragma unroll
      for (int bd = 0; bd < 8; ++bd) {
        const float* src = p.band_in[bd] + o;
        for (int c = tid; c < nchunk; c += nthr) cpa8(dst + bd * BT + 2 * c, src + 2 * c);
      }
    }
    cpa_commit();
  };
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) issue(s);

  float win[H][8];   // [plane][4 db + 2 hpar + wpar]
#pragma unroll
  for (int k = 0; k < H; ++k)
#pragma unroll
    for (int c = 0; c < 8; ++c) win[k][c] = 0.f;

  for (int s = 0; s < NP; ++s) {
    cpa_wait<kStages - 2>();
    __syncthreads();
    issue(s + kStages - 1);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (qd0 + s < p.nd) {
      const float* tp = sm + (s % kStages) * PT + tbase;
#pragma unroll
      for (int j = 0; j < H; ++j) {           // h tap pair j <-> coefficient row qh + H - 1 - j
        const int r = H - 1 - j;
#pragma unroll
        for (int db = 0; db < 2; ++db) {
          float w[2][2];                       // [h band][w parity]: W pass of bands (db, hb, low) + (db, hb, high)
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            const float* lo = tp + (4 * db + 2 * hb) * BT + r * nw;
            const float* hi = lo + BT;
            float cl[H], ch[H];
#pragma unroll
            for (int i = 0; i < H; ++i) {
              cl[i] = lo[i];
              ch[i] = hi[i];
            }
            float e = 0.f, o = 0.f;
#pragma unroll
            for (int jw = 0; jw < H; ++jw) {
              e = fmaf(cl[H - 1 - jw], p.t0[2 * jw], e);
              e = fmaf(ch[H - 1 - jw], p.t1[2 * jw], e);
              o = fmaf(cl[H - 1 - jw], p.t0[2 * jw + 1], o);
              o = fmaf(ch[H - 1 - jw], p.t1[2 * jw + 1], o);
            }
            w[hb][0] = e;
            w[hb][1] = o;
          }
#pragma unroll
          for (int hp = 0; hp < 2; ++hp)
#pragma unroll
            for (int wp = 0; wp < 2; ++wp) {
              float acc = a[4 * db + 2 * hp + wp];
              acc = fmaf(w[0][wp], p.t0[2 * j + hp], acc);
              acc = fmaf(w[1][wp], p.t1[2 * j + hp], acc);
              a[4 * db + 2 * hp + wp] = acc;
            }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < H - 1; ++k)
#pragma unroll
      for (int c = 0; c < 8; ++c) win[k][c] = win[k + 1][c];
#pragma unroll
    for (int c = 0; c < 8; ++c) win[H - 1][c] = a[c];
    const int e = s - (H - 1);
    if (e >= 0 && live) {                      // window = coefficient planes qd .. qd + H - 1
      const int qd = qd0 + e;
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {
        const int gd = 2 * qd + dp;
        if (gd >= p.Nd) continue;
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
          const int gh = 2 * qh + hp;
          if (gh >= p.Nh) continue;
          float2 out;
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int j = 0; j < H; ++j) {
            acc0 = fmaf(win[H - 1 - j][2 * hp + 0], p.t0[2 * j + dp], acc0);
            acc0 = fmaf(win[H - 1 - j][4 + 2 * hp + 0], p.t1[2 * j + dp], acc0);
            acc1 = fmaf(win[H - 1 - j][2 * hp + 1], p.t0[2 * j + dp], acc1);
            acc1 = fmaf(win[H - 1 - j][4 + 2 * hp + 1], p.t1[2 * j + dp], acc1);
          }
          out.x = acc0;
          out.y = acc1;
          *reinterpret_cast<float2*>(p.y + b * p.sig_bstride + (static_cast<long long>(gd) * p.Nh + gh) * p.Nw + 2 * qw) = out;
        }
      }
    }
  }
  cpa_wait<0>();
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

bool stream_enabled() {
  static const bool on = [] { const char* e = getenv("WDNO_DWT3D_STREAM"); return !(e && e[0] == '0'); }();
  return on;
}

// rows per CTA so that the block has about `target` threads, strips balanced over `n` rows
int pick_rows(int n, int per_row, int target) {
  int t = target / per_row;
  if (t < 1) t = 1;
  if (t > n) t = n;
  const int strips = (n + t - 1) / t;
  return (n + strips - 1) / strips;
}

// planes per CTA: as few chunks as give >= 2 CTAs per SM (halo planes are re-read from L2), at least 4 planes per chunk
int pick_planes(int n, long long ctas_per_chunk) {
  const long long want = 2LL * num_sms();
  int chunks = static_cast<int>((want + ctas_per_chunk - 1) / ctas_per_chunk);
  if (chunks < 1) chunks = 1;
  const int max_chunks = n >= 4 ? n / 4 : 1;
  if (chunks > max_chunks) chunks = max_chunks;
  return (n + chunks - 1) / chunks;
}

template <typename K>
int prep(K kernel, size_t smem, const char* where) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  return e == cudaSuccess ? WDNO_OK : set_cuda_error(e, where);
}

}  // namespace

// returns 1 if the streaming kernel was launched, 0 if the shape / alignment is outside its envelope (caller falls back to the
// tile kernel), < 0 on error
int launch_ana3d_stream(const float* x, float* const* bands8, long long band_bstride, long long B, int Nd, int Nh, int Nw, int nd,
                        int nh, int nw, const float* t0, const float* t1, int L, int off, cudaStream_t st) {
  if (!stream_enabled()) return 0;
  if (L != 2 && L != 6 && L != 10) return 0;
  if (off != L - 2 || (off & 3) || (Nw & 3) || (reinterpret_cast<uintptr_t>(x) & 15)) return 0;
  if (nw > 640 || B > 65535) return 0;
  S3 p = {};
  for (int i = 0; i < 8; ++i) p.band_out[i] = bands8[i];
  for (int k = 0; k < 10; ++k) {
    p.t0[k] = k < L ? t0[k] : 0.f;
    p.t1[k] = k < L ? t1[k] : 0.f;
  }
  p.x = x;
  p.band_bstride = band_bstride;
  p.sig_bstride = static_cast<long long>(Nd) * Nh * Nw;
  p.nd = nd; p.nh = nh; p.nw = nw; p.Nd = Nd; p.Nh = Nh; p.Nw = Nw; p.off = off;
  static const int th_env = env_int("WDNO_DWT3D_ATH", 0), td_env = env_int("WDNO_DWT3D_ATD", 0);
  p.TH = th_env > 0 ? (th_env < nh ? th_env : nh) : pick_rows(nh, nw, 320);
  if (p.TH * nw > 640) p.TH = 640 / nw;
  const int strips = (nh + p.TH - 1) / p.TH;
  p.TD = td_env > 0 ? (td_env < nd ? td_env : nd) : pick_planes(nd, static_cast<long long>(strips) * B);
  p.R = 2 * p.TH + L - 2;
  p.RS = ((2 * nw + L - 2 > off + Nw ? 2 * nw + L - 2 : off + Nw) + 3) & ~3;
  const size_t smem = sizeof(float) * kStages * static_cast<size_t>(p.R) * p.RS;
  if (smem > 200 * 1024) return 0;
  const int block = ((p.TH * nw + 31) / 32) * 32;
  dim3 grid(strips, (nd + p.TD - 1) / p.TD, static_cast<unsigned>(B));
  if (grid.y > 65535) return 0;
  int rc;
#define WDNO_AS(LL)                                                                  \
  {                                                                                  \
    static size_t cfg = 0;                                                           \
    if (smem > cfg) {                                                                \
      if ((rc = prep(ana3d_stream_kernel<LL>, smem, "dwt3d_analysis(stream)"))) return rc; \
      cfg = smem;                                                                    \
    }                                                                                \
    ana3d_stream_kernel<LL><<<grid, block, smem, st>>>(p);                           \
  }
  if (L == 6) WDNO_AS(6) else if (L == 10) WDNO_AS(10) else WDNO_AS(2)
#undef WDNO_AS
  rc = check_launch("dwt3d_analysis(stream)");
  return rc ? rc : 1;
}

int launch_syn3d_stream(const float* const* bands8, long long band_bstride, float* y, long long B, int nd, int nh, int nw, int Nd,
                        int Nh, int Nw, const float* t0, const float* t1, int L, int off, cudaStream_t st) {
  if (!stream_enabled()) return 0;
  if (L != 2 && L != 6 && L != 10) return 0;
  if (off != L - 2 || (Nw & 1) || (nw & 1) || (band_bstride & 1) || (reinterpret_cast<uintptr_t>(y) & 7)) return 0;
  for (int i = 0; i < 8; ++i)
    if (reinterpret_cast<uintptr_t>(bands8[i]) & 7) return 0;
  // every coefficient an output needs must exist: N <= 2 n - L + 2 per axis
  if (Nd > 2 * nd - L + 2 || Nh > 2 * nh - L + 2 || Nw > 2 * nw - L + 2) return 0;
  const int QW = Nw >> 1, QH = (Nh + 1) >> 1, QD = (Nd + 1) >> 1;
  if (QW > 640 || B > 65535) return 0;
  S3 p = {};
  for (int i = 0; i < 8; ++i) p.band_in[i] = bands8[i];
  for (int k = 0; k < 10; ++k) {
    p.t0[k] = k < L ? t0[k] : 0.f;
    p.t1[k] = k < L ? t1[k] : 0.f;
  }
  p.y = y;
  p.band_bstride = band_bstride;
  p.sig_bstride = static_cast<long long>(Nd) * Nh * Nw;
  p.nd = nd; p.nh = nh; p.nw = nw; p.Nd = Nd; p.Nh = Nh; p.Nw = Nw; p.off = off;
  static const int th_env = env_int("WDNO_DWT3D_STH", 0), td_env = env_int("WDNO_DWT3D_STD", 0);
  p.TH = th_env > 0 ? (th_env < QH ? th_env : QH) : pick_rows(QH, QW, 256);
  if (p.TH * QW > 640) p.TH = 640 / QW;
  const int strips = (QH + p.TH - 1) / p.TH;
  p.TD = td_env > 0 ? (td_env < QD ? td_env : QD) : pick_planes(QD, static_cast<long long>(strips) * B);
  p.R = p.TH + L / 2 - 1;
  p.RS = nw;
  const size_t smem = sizeof(float) * kStages * 8 * static_cast<size_t>(p.R) * nw;
  if (smem > 200 * 1024) return 0;
  const int block = ((p.TH * QW + 31) / 32) * 32;
  dim3 grid(strips, (QD + p.TD - 1) / p.TD, static_cast<unsigned>(B));
  if (grid.y > 65535) return 0;
  int rc;
#define WDNO_SS(LL)                                                                  \
  {                                                                                  \
    static size_t cfg = 0;                                                           \
    if (smem > cfg) {                                                                \
      if ((rc = prep(syn3d_stream_kernel<LL>, smem, "dwt3d_synthesis(stream)"))) return rc; \
      cfg = smem;                                                                    \
    }                                                                                \
    syn3d_stream_kernel<LL><<<grid, block, smem, st>>>(p);                           \
  }
  if (L == 6) WDNO_SS(6) else if (L == 10) WDNO_SS(10) else WDNO_SS(2)
#undef WDNO_SS
  rc = check_launch("dwt3d_synthesis(stream)");
  return rc ? rc : 1;
}

}  // namespace wdno
