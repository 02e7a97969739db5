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
  int R, RS;         // tile rows; analysis: tile row stride, synthesis: band tile size (floats)
  int nq_shift;      // analysis: log2(Nw / 4) or -1
  int v16;           // synthesis: band runs are 16-byte aligned
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

// keep a kernel parameter in a vector register (opaque move, so the compiler cannot re-materialise it from the constant bank).
// Used for the taps; it did NOT remove the LDCU traffic of these kernels (8 % of the instructions, profiles/r1e_dwt.md): those
// loads are the other parameters (extents, band pointers) read inside the plane loop -- next candidates for the same treatment.
__device__ __forceinline__ float pin_reg(float v) {
  float r;
  asm volatile("mov.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// tap masks: bit k of M0 / M1 set = taps t0[k] / t1[k] may be non-zero.  bior1.3 has two non-zero taps of six in its
// high-pass analysis filter (and in the low-pass reconstruction filter); skipping the products with exact zeros leaves the
// results unchanged (acc + 0 * v) and lets the compiler drop the loads and whole row passes that only fed them.
template <unsigned M>
__device__ __forceinline__ float fm(int k, float a, float t, float acc) {
  return ((M >> k) & 1u) ? fmaf(a, t, acc) : acc;
}

// ---------------------------------------------------------------- analysis
// grid (h strips, d chunks, B); block >= (TH / NR) * nw threads; a thread owns NR consecutive output rows at one column
// (their 6-row input windows overlap: L + 2 (NR - 1) row passes serve NR outputs).  Tile of one signal plane:
// R = 2 TH + L - 2 rows of RS floats, tile column c = signal column c - off (pads and rows outside [0, Nh) stay zero).
// CRS: compile-time tile row stride (0 = run-time p.RS): the 18 float2 loads of a plane become one base register + immediates
template <int L, int NR, unsigned M0, unsigned M1, int CRS>
__global__ void __launch_bounds__(640) ana3d_stream_kernel(const S3 p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int KR = L + 2 * (NR - 1);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b = blockIdx.z;
  const int ih0 = blockIdx.x * p.TH, id0 = blockIdx.y * p.TD;
  const int id1 = min(id0 + p.TD, p.nd);
  const int nw = p.nw, Nw = p.Nw, R = p.R, RS = CRS ? CRS : p.RS;
  float t0[L], t1[L];                                       // taps in registers (the masked-out ones are dead)
#pragma unroll
  for (int k = 0; k < L; ++k) {
    t0[k] = pin_reg(p.t0[k]);
    t1[k] = pin_reg(p.t1[k]);
  }
  const int PT = R * RS;
  for (int i = tid; i < kStages * PT; i += nthr) sm[i] = 0.f;
  const int xr0 = 2 * ih0 - p.off;                          // tile row r <-> signal row xr0 + r
  const int r_lo = max(0, -xr0), r_hi = min(R, p.Nh - xr0);
  const int nq = Nw >> 2;                                   // 16-byte chunks per row
  const int nchunk = max(0, r_hi - r_lo) * nq;
  const float* xb = p.x + b * p.sig_bstride + static_cast<long long>(xr0 + r_lo) * Nw;
  const int NP = 2 * (id1 - id0) + L - 2;                   // planes this CTA streams (even)
  const int gd0 = 2 * id0 - p.off;                          // signal plane of step 0
  const int ngrp = p.TH / NR;
  const int pos = tid < ngrp * nw ? tid : 0;
  const int ihg = pos / nw, iw = pos - ihg * nw;
  const int ihb = ih0 + NR * ihg;                           // first of this thread's NR output rows
  const bool mine = tid < ngrp * nw;
  const int tbase = (2 * NR * ihg) * RS + 2 * iw;
  __syncthreads();

  auto issue = [&](int s) {
    const int gd = gd0 + s;
    if (s < NP && gd >= 0 && gd < p.Nd) {
      float* dst = sm + (s % kStages) * PT + r_lo * RS + p.off;
      const float* src = xb + static_cast<long long>(gd) * p.Nh * Nw;
      if (p.nq_shift >= 0) {
        for (int c = tid; c < nchunk; c += nthr) {
          const int r = c >> p.nq_shift, q = c & (nq - 1);
          cpa16(dst + r * RS + 4 * q, src + r * Nw + 4 * q);
        }
      } else {
        for (int c = tid; c < nchunk; c += nthr) {
          const int r = c / nq, q = c - r * nq;
          cpa16(dst + r * RS + 4 * q, src + r * Nw + 4 * q);
        }
      }
    }
    cpa_commit();
  };
  // W + H passes of one plane for this thread's NR rows: a[r][2 hb + wb]
  auto plane = [&](int s, float (&a)[NR][4]) {
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) a[r][c] = 0.f;
    const int gd = gd0 + s;
    if (gd >= 0 && gd < p.Nd) {
      const float* tp = sm + (s % kStages) * PT + tbase;
#pragma unroll
      for (int kh = 0; kh < KR; ++kh) {
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
          wl = fm<M0>(k, v[k], t0[k], wl);
          wh = fm<M1>(k, v[k], t1[k], wh);
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int k = kh - 2 * r;
          if (k >= 0 && k < L) {
            a[r][0] = fm<M0>(k, wl, t0[k], a[r][0]);
            a[r][1] = fm<M0>(k, wh, t0[k], a[r][1]);
            a[r][2] = fm<M1>(k, wl, t1[k], a[r][2]);
            a[r][3] = fm<M1>(k, wh, t1[k], a[r][3]);
          }
        }
      }
    }
  };
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) issue(s);

  float win[NR][4][L];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < L; ++k) win[r][c][k] = 0.f;

  for (int sp = 0; 2 * sp < NP; ++sp) {   // two planes per iteration: the window moves by 2, one coefficient plane comes out
    float a0[NR][4], a1[NR][4];
    cpa_wait<kStages - 2>();
    __syncthreads();              // plane s has landed for everyone; everyone is done with plane s - 1 (its slot is refilled next)
    issue(2 * sp + kStages - 1);
    plane(2 * sp, a0);
    cpa_wait<kStages - 2>();
    __syncthreads();
    issue(2 * sp + kStages);
    plane(2 * sp + 1, a1);
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int k = 0; k + 2 < L; ++k) win[r][c][k] = win[r][c][k + 2];
        win[r][c][L - 2] = a0[r][c];
        win[r][c][L - 1] = a1[r][c];
      }
    const int m = sp + 1 - L / 2;  // window = planes 2 m .. 2 m + L - 1 of the chunk: coefficient plane id0 + m
    if (m >= 0 && mine) {
      const int gid = id0 + m;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        if (ihb + r >= p.nh) continue;
        const long long o = b * p.band_bstride + (static_cast<long long>(gid) * p.nh + ihb + r) * nw + iw;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float lo = 0.f, hi = 0.f;
#pragma unroll
          for (int k = 0; k < L; ++k) {
            lo = fm<M0>(k, win[r][c][k], t0[k], lo);
            hi = fm<M1>(k, win[r][c][k], t1[k], hi);
          }
          p.band_out[c][o] = lo;
          p.band_out[4 + c][o] = hi;
        }
      }
    }
  }
  cpa_wait<0>();
}

// ---------------------------------------------------------------- synthesis
// grid (h strips, d chunks, B); block >= TH * (Nw / 2) threads.  Tile of one coefficient plane: 8 bands x R = TH + L/2 - 1
// rows x nw floats (a flat copy of the contiguous run of each band plane; rows at or beyond nh stay zero).
// CNW / CBT: compile-time coefficient row length and band tile size (0 = run-time): every LDS offset becomes an immediate
template <int L, unsigned M0, unsigned M1, int CNW, int CBT>
__global__ void __launch_bounds__(640) syn3d_stream_kernel(const S3 p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int H = L / 2;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b = blockIdx.z;
  const int qh0 = blockIdx.x * p.TH, qd0 = blockIdx.y * p.TD;
  const int QD = (p.Nd + 1) >> 1, QW = p.Nw >> 1;
  const int qd1 = min(qd0 + p.TD, QD);
  const int nw = CNW ? CNW : p.nw, R = p.R;
  const int BT = CBT ? CBT : p.RS, PT = 8 * BT;             // band tile (R nw rounded up to 16 bytes), plane tile
  float t0[L], t1[L];
#pragma unroll
  for (int k = 0; k < L; ++k) {
    t0[k] = pin_reg(p.t0[k]);
    t1[k] = pin_reg(p.t1[k]);
  }
  for (int i = tid; i < kStages * PT; i += nthr) sm[i] = 0.f;
  const int r_hi = min(R, p.nh - qh0);
  const int nfl = max(0, r_hi) * nw;                        // floats per band run (even)
  const long long run0 = b * p.band_bstride + static_cast<long long>(qh0) * nw;
  const int NP = (qd1 - qd0) + H - 1;
  const int pos = tid < p.TH * QW ? tid : 0;
  const int qhl = pos / QW, qw = pos - qhl * QW;
  const int qh = qh0 + qhl;
  const bool live = tid < p.TH * QW && 2 * qh < p.Nh;
  const int tbase = qhl * nw + qw;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  __syncthreads();

  auto issue = [&](int s) {
    const int gid = qd0 + s;
    if (s < NP && gid < p.nd) {
      float* dst = sm + (s % kStages) * PT;
      const long long o = run0 + static_cast<long long>(gid) * p.nh * nw;
      if (p.v16) {                                          // a warp per band, 16-byte chunks (+ one 8-byte tail)
        const int n16 = nfl >> 2;
        for (int bd = warp; bd < 8; bd += nwarps) {
          const float* src = p.band_in[bd] + o;
          float* d = dst + bd * BT;
          for (int c = lane; c < n16; c += 32) cpa16(d + 4 * c, src + 4 * c);
          if ((nfl & 2) && lane == 0) cpa8(d + 4 * n16, src + 4 * n16);
        }
      } else {
#pragma unroll
        for (int bd = 0; bd < 8; ++bd) {
          const float* src = p.band_in[bd] + o;
          for (int c = tid; 2 * c < nfl; c += nthr) cpa8(dst + bd * BT + 2 * c, src + 2 * c);
        }
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
              e = fm<M0>(2 * jw, cl[H - 1 - jw], t0[2 * jw], e);
              e = fm<M1>(2 * jw, ch[H - 1 - jw], t1[2 * jw], e);
              o = fm<M0>(2 * jw + 1, cl[H - 1 - jw], t0[2 * jw + 1], o);
              o = fm<M1>(2 * jw + 1, ch[H - 1 - jw], t1[2 * jw + 1], o);
            }
            w[hb][0] = e;
            w[hb][1] = o;
          }
#pragma unroll
          for (int hp = 0; hp < 2; ++hp)
#pragma unroll
            for (int wp = 0; wp < 2; ++wp) {
              float acc = a[4 * db + 2 * hp + wp];
              acc = fm<M0>(2 * j + hp, w[0][wp], t0[2 * j + hp], acc);
              acc = fm<M1>(2 * j + hp, w[1][wp], t1[2 * j + hp], acc);
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
            acc0 = fm<M0>(2 * j + dp, win[H - 1 - j][2 * hp + 0], t0[2 * j + dp], acc0);
            acc0 = fm<M1>(2 * j + dp, win[H - 1 - j][4 + 2 * hp + 0], t1[2 * j + dp], acc0);
            acc1 = fm<M0>(2 * j + dp, win[H - 1 - j][2 * hp + 1], t0[2 * j + dp], acc1);
            acc1 = fm<M1>(2 * j + dp, win[H - 1 - j][4 + 2 * hp + 1], t1[2 * j + dp], acc1);
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

unsigned nz_mask(const float* t, int L) {
  unsigned m = 0;
  for (int k = 0; k < L; ++k)
    if (t[k] != 0.f) m |= 1u << k;
  return m;
}

void fill_taps(S3& p, const float* t0, const float* t1, int L) {
  for (int k = 0; k < 10; ++k) {
    p.t0[k] = k < L ? t0[k] : 0.f;
    p.t1[k] = k < L ? t1[k] : 0.f;
  }
}

template <typename K>
int launch(K kernel, size_t& cfg, dim3 grid, int block, size_t smem, cudaStream_t st, const S3& p, const char* where) {
  if (smem > cfg) {
    int rc = prep(kernel, smem, where);
    if (rc) return rc;
    cfg = smem;
  }
  kernel<<<grid, block, smem, st>>>(p);
  return check_launch(where);
}

template <int L, int NR>
int launch_ana(unsigned nz0, unsigned nz1, dim3 grid, int block, size_t smem, cudaStream_t st, const S3& p) {
  const char* where = "dwt3d_analysis(stream)";
  constexpr unsigned F = (1u << L) - 1u, MID = 3u << (L / 2 - 1);   // all taps / the two centre taps
  static size_t c0 = 0, c1 = 0, c2 = 0;
  if constexpr (L == 6 && NR == 1) {
    static size_t d1 = 0, d2 = 0, e1 = 0, e2 = 0;
    // the two smoke geometries (64- and 128-wide planes) with compile-time row strides
    if (p.RS == 72 && !(nz1 & ~MID)) return launch(ana3d_stream_kernel<L, NR, F, MID, 72>, d1, grid, block, smem, st, p, where);
    if (p.RS == 72 && !(nz0 & ~MID)) return launch(ana3d_stream_kernel<L, NR, MID, F, 72>, d2, grid, block, smem, st, p, where);
    if (p.RS == 136 && !(nz1 & ~MID)) return launch(ana3d_stream_kernel<L, NR, F, MID, 136>, e1, grid, block, smem, st, p, where);
    if (p.RS == 136 && !(nz0 & ~MID)) return launch(ana3d_stream_kernel<L, NR, MID, F, 136>, e2, grid, block, smem, st, p, where);
  }
  if constexpr (L == 6) {
    if (!(nz1 & ~MID)) return launch(ana3d_stream_kernel<L, NR, F, MID, 0>, c1, grid, block, smem, st, p, where);
    if (!(nz0 & ~MID)) return launch(ana3d_stream_kernel<L, NR, MID, F, 0>, c2, grid, block, smem, st, p, where);
  }
  return launch(ana3d_stream_kernel<L, NR, F, F, 0>, c0, grid, block, smem, st, p, where);
}

template <int L>
int launch_syn(unsigned nz0, unsigned nz1, dim3 grid, int block, size_t smem, cudaStream_t st, const S3& p) {
  const char* where = "dwt3d_synthesis(stream)";
  constexpr unsigned F = (1u << L) - 1u, MID = 3u << (L / 2 - 1);
  static size_t c0 = 0, c1 = 0, c2 = 0;
  if constexpr (L == 6) {
    static size_t d1 = 0, d2 = 0, e1 = 0, e2 = 0;
    // the two smoke geometries (34- and 66-wide coefficient rows at the default strip heights)
    if (p.nw == 34 && p.RS == 340 && !(nz1 & ~MID)) return launch(syn3d_stream_kernel<L, F, MID, 34, 340>, d1, grid, block, smem, st, p, where);
    if (p.nw == 34 && p.RS == 340 && !(nz0 & ~MID)) return launch(syn3d_stream_kernel<L, MID, F, 34, 340>, d2, grid, block, smem, st, p, where);
    if (p.nw == 66 && p.RS == 396 && !(nz1 & ~MID)) return launch(syn3d_stream_kernel<L, F, MID, 66, 396>, e1, grid, block, smem, st, p, where);
    if (p.nw == 66 && p.RS == 396 && !(nz0 & ~MID)) return launch(syn3d_stream_kernel<L, MID, F, 66, 396>, e2, grid, block, smem, st, p, where);
    if (!(nz1 & ~MID)) return launch(syn3d_stream_kernel<L, F, MID, 0, 0>, c1, grid, block, smem, st, p, where);
    if (!(nz0 & ~MID)) return launch(syn3d_stream_kernel<L, MID, F, 0, 0>, c2, grid, block, smem, st, p, where);
  }
  return launch(syn3d_stream_kernel<L, F, F, 0, 0>, c0, grid, block, smem, st, p, where);
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
  fill_taps(p, t0, t1, L);
  p.x = x;
  p.band_bstride = band_bstride;
  p.sig_bstride = static_cast<long long>(Nd) * Nh * Nw;
  p.nd = nd; p.nh = nh; p.nw = nw; p.Nd = Nd; p.Nh = Nh; p.Nw = Nw; p.off = off;
  static const int th_env = env_int("WDNO_DWT3D_ATH", 0), td_env = env_int("WDNO_DWT3D_ATD", 0);
  static const int nr_env = env_int("WDNO_DWT3D_ANR", 1), thr_env = env_int("WDNO_DWT3D_ATHREADS", 256);
  const int NR = (nr_env == 2 && nh >= 2 && L <= 6) ? 2 : 1;  // output rows per thread (2: measured 7 % slower, fewer warps)
  p.TH = th_env > 0 ? (th_env < nh ? th_env : nh) : pick_rows(nh, (nw + NR - 1) / NR, thr_env);
  p.TH = ((p.TH + NR - 1) / NR) * NR;
  while (p.TH > NR && (p.TH / NR) * nw > 640) p.TH -= NR;
  if ((p.TH / NR) * nw > 640) return 0;
  const int strips = (nh + p.TH - 1) / p.TH;
  p.TD = td_env > 0 ? (td_env < nd ? td_env : nd) : pick_planes(nd, static_cast<long long>(strips) * B);
  p.R = 2 * p.TH + L - 2;
  p.RS = ((2 * nw + L - 2 > off + Nw ? 2 * nw + L - 2 : off + Nw) + 3) & ~3;
  const int nq = Nw >> 2;
  p.nq_shift = -1;
  for (int sh = 0; sh < 16; ++sh)
    if ((1 << sh) == nq) p.nq_shift = sh;
  const size_t smem = sizeof(float) * kStages * static_cast<size_t>(p.R) * p.RS;
  if (smem > 200 * 1024) return 0;
  const int block = (((p.TH / NR) * nw + 31) / 32) * 32;
  dim3 grid(strips, (nd + p.TD - 1) / p.TD, static_cast<unsigned>(B));
  if (grid.y > 65535) return 0;
  const unsigned nz0 = nz_mask(t0, L), nz1 = nz_mask(t1, L);
  int rc;
  if (L == 6) rc = NR == 2 ? launch_ana<6, 2>(nz0, nz1, grid, block, smem, st, p) : launch_ana<6, 1>(nz0, nz1, grid, block, smem, st, p);
  else if (L == 10) rc = NR == 2 ? launch_ana<10, 2>(nz0, nz1, grid, block, smem, st, p) : launch_ana<10, 1>(nz0, nz1, grid, block, smem, st, p);
  else rc = NR == 2 ? launch_ana<2, 2>(nz0, nz1, grid, block, smem, st, p) : launch_ana<2, 1>(nz0, nz1, grid, block, smem, st, p);
  return rc ? rc : 1;
}

int launch_syn3d_stream(const float* const* bands8, long long band_bstride, float* y, long long B, int nd, int nh, int nw, int Nd,
                        int Nh, int Nw, const float* t0, const float* t1, int L, int off, cudaStream_t st) {
  if (!stream_enabled()) return 0;
  if (L != 2 && L != 6 && L != 10) return 0;
  if (off != L - 2 || (Nw & 1) || (nw & 1) || (band_bstride & 1) || (reinterpret_cast<uintptr_t>(y) & 7)) return 0;
  bool a16 = !(band_bstride & 3) && !((static_cast<long long>(nh) * nw) & 3);
  for (int i = 0; i < 8; ++i) {
    if (reinterpret_cast<uintptr_t>(bands8[i]) & 7) return 0;
    if (reinterpret_cast<uintptr_t>(bands8[i]) & 15) a16 = false;
  }
  // every coefficient an output needs must exist: N <= 2 n - L + 2 per axis
  if (Nd > 2 * nd - L + 2 || Nh > 2 * nh - L + 2 || Nw > 2 * nw - L + 2) return 0;
  const int QW = Nw >> 1, QH = (Nh + 1) >> 1, QD = (Nd + 1) >> 1;
  if (QW > 640 || B > 65535) return 0;
  S3 p = {};
  for (int i = 0; i < 8; ++i) p.band_in[i] = bands8[i];
  fill_taps(p, t0, t1, L);
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
  p.RS = (p.R * nw + 3) & ~3;                               // band tile, 16-byte multiple
  p.v16 = (a16 && !((static_cast<long long>(p.TH) * nw) & 3)) ? 1 : 0;   // every strip's run starts on a 16-byte boundary
  const size_t smem = sizeof(float) * kStages * 8 * static_cast<size_t>(p.RS);
  if (smem > 200 * 1024) return 0;
  const int block = ((p.TH * QW + 31) / 32) * 32;
  dim3 grid(strips, (QD + p.TD - 1) / p.TD, static_cast<unsigned>(B));
  if (grid.y > 65535) return 0;
  const unsigned nz0 = nz_mask(t0, L), nz1 = nz_mask(t1, L);
  int rc;
  if (L == 6) rc = launch_syn<6>(nz0, nz1, grid, block, smem, st, p);
  else if (L == 10) rc = launch_syn<10>(nz0, nz1, grid, block, smem, st, p);
  else rc = launch_syn<2>(nz0, nz1, grid, block, smem, st, p);
  return rc ? rc : 1;
}

}  // namespace wdno
