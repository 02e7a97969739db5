// SpatialLinearAttention block of the smoke U-Net (reference conv3d.py:165-184, 232-258) with every product on tcgen05:
//   y = x + to_out( ctx^T softmax_d(q) * scale ),  ctx = softmax_n(k) v^T,  (q, k, v) = to_qkv(LayerNorm(x))
// Same three-launch structure and the same workspace as the mma.sync form in attn_fused.cu (la1 -> la_mid -> la2), but the
// two big kernels are rebuilt around ONE idea: put the axis a softmax runs over on the TMEM COLUMN axis, so that the thread
// that owns an accumulator row (tcgen05.ld 32x32b: lane = row) finds its whole softmax in its own registers -- no
// shuffles, no fragment bookkeeping, and the instruction count per element drops to the FFMA + MUFU.EX2 floor.
//   la1_tc : K^T[(h,d) x pixel] = Wk xn^T and V^T[(h,e) x pixel] = Wv xn^T  (M = 128 weight rows, N = 128 pixels, K = C);
//            row thread: running max over the pixels, ek = exp(k - max) -> fp16, written as 16-byte chunks of 8 pixels =
//            the UMMA canonical K-major layout (K = pixels) of the NEXT product  S[(h,d) x (h',e)] = ek v^T  (all head pairs
//            in one M = N = 128 MMA; the thread keeps the 32 columns of its own head) merged into a register-resident
//            (max, sum, S) per row (each of the two warps of a lane quarter keeps 16 of the 32 columns).  k, v, ek never leave the SM.
//   la2_tc : Q[pixel x (h,d)] = xn Wq^T (M = 128 pixels): row thread = pixel, per-head softmax over 32 columns in registers
//            -> fp16 A operand -> Y[pixel x C] = q M_img^T (M_img = scale * W_out blockdiag(ctx^T) from la_mid) -> warp-local
//            transpose through shared memory -> + bias + residual -> coalesced 64-byte row segments.
// Both kernels are persistent (one CTA per SM walks a contiguous range of 128-pixel tiles) and warp-specialised:
//   warps 0-3 LayerNorm producers (16 channels per lane, next tile prefetched in registers, output = UMMA operand tiles),
//   warps 4-11 accumulator-row consumers: TWO warps per TMEM lane quarter (= warp % 4), each owning one half of the columns
//   (ncu of the one-warp-per-quarter version: that warp ran at IPC 0.18, a pure latency chain), warp 12 = MMA issuer; mbarrier
//   rings between them, so the LayerNorm of tile i+1/i+2, the MMAs of tile i+1 and the exponentials of tile i overlap.
//   Floor: 128 x 128 MUFU.EX2 per tile = 1024 cycles per SM sub-partition; the tensor pipe needs 1024 (la1) / 640 (la2) cycles per tile and runs beside it.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/wdno_b200.h"
#include "common.cuh"
#include "cvt_sat.cuh"
#include "ptx.cuh"

namespace wdno {

namespace {

constexpr int kHid = 128;
constexpr int kThreads = 416;                    // warps 0-3: LayerNorm, 4-11: accumulator rows (2 per lane quarter), 12: MMA issue
constexpr int kMmaWarp = 12;
constexpr int kRowThreads = 256;
constexpr int kPartFloats = 64 + 32 * 32;        // per (image, part, head): max[32], sum[32], S[32][32]  (= kLaPart)
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo16) {
  // K-major, no swizzle: start >> 4 | (LBO >> 4) << 16 | (SBO = 128 B >> 4) << 32 | version 1 << 46
  return static_cast<uint64_t>(((saddr >> 4) & 0x3FFFu) | (lbo16 << 16)) | (static_cast<uint64_t>(8u | (1u << 14)) << 32);
}
__device__ __forceinline__ float ex2(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __half2 h = h2_sat(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

struct Tile {
  int img, p0, nv;
};
// walks the tiles of a CTA's range in order (one division at construction)
struct TileIter {
  int img, tin, tpi, n;
  __device__ __forceinline__ TileIter(long long g, int tpi_, int n_) : tpi(tpi_), n(n_) {
    img = static_cast<int>(g / tpi_);
    tin = static_cast<int>(g - static_cast<long long>(img) * tpi_);
  }
  __device__ __forceinline__ Tile get() const {
    Tile t;
    t.img = img;
    t.p0 = tin * 128;
    t.nv = min(128, n - t.p0);
    return t;
  }
  __device__ __forceinline__ void next() {
    if (++tin == tpi) {
      tin = 0;
      ++img;
    }
  }
};

// barrier block (all kernels): 8-byte slots
struct Bars {
  uint64_t xn_full[2], xn_empty[2];
  uint64_t a_full[2], a_empty[2];     // la1: K^T accumulator buffers;      la2: Q accumulator buffers
  uint64_t b_full[2], b_empty[2];     // la1: [0] V^T accumulator;          la2: Y accumulator buffers
  uint64_t op_full[2], op_empty[2];   // la1: [0] ek/v operand tiles ready; la2: softmax(q) operand tiles
  uint64_t s_full, s_empty;           // la1: S accumulator
  uint64_t m_empty[2];                // la2: per-image M buffers
  uint32_t tmem_base;
};
constexpr int kBarBytes = 256;
static_assert(sizeof(Bars) <= kBarBytes, "barrier block");

// ------------------------------------------------------------------ LayerNorm producers (128 threads)
// LPR = C/16 lanes per pixel row (16 channels each), 128/LPR rows per pass, LPR passes per tile, processed BATCH passes at a
// time so that their shuffle / rsqrt latency chains interleave; the raw values of the NEXT tile are fetched into the registers
// a batch has just consumed.  Output: xn[buf] = [C/8 chunks][128 rows][8 channels] fp16 (UMMA K-major, K = channels) of
// (x - mean) * rstd -- the LayerNorm gain is folded into the weight operands on the host; rows >= nv are zeros.
// MLOAD (la2): on an image change the image's M tile is copied next to it.
template <int C, bool MLOAD>
__device__ __forceinline__ void ln_role(const __half* __restrict__ x, uint8_t* xn_base, Bars* bars, long long t0,
                                        long long t1, int tpi, int n, float eps, int tid, const __half* __restrict__ mcanon,
                                        uint8_t* m_base) {
  constexpr int LPR = C / 16;                      // lanes per row
  constexpr int RPP = 128 / LPR;                   // rows per pass
  constexpr int PASSES = LPR;
  constexpr int BATCH = (C == 64) ? 4 : 2;
  constexpr int XNB = 128 * C * 2;                 // bytes per xn buffer
  const int q = tid % LPR, rsub = tid / LPR;
  uint4 cur[PASSES][2];
  auto load = [&](const Tile& t, int ps) {
    const int row = RPP * ps + rsub;
    const bool ok = row < t.nv;                    // nv = 0 past the end of the range
    const uint4* src = reinterpret_cast<const uint4*>(x + (static_cast<size_t>(t.img) * n + t.p0 + row) * C + q * 16);
    cur[ps][0] = ok ? __ldg(src) : make_uint4(0u, 0u, 0u, 0u);
    cur[ps][1] = ok ? __ldg(src + 1) : make_uint4(0u, 0u, 0u, 0u);
  };
  Tile tn = {0, 0, 0};
  TileIter it(t0, tpi, n);
  if (t0 < t1) tn = it.get();
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps) load(tn, ps);
  int cur_img = -1, kimg = -1;
  int j = 0;
  for (long long g = t0; g < t1; ++g, ++j) {
    const int buf = j & 1;
    const Tile t = tn;
    tn.nv = 0;
    it.next();
    if (g + 1 < t1) tn = it.get();
    ptx::mbar_wait(&bars->xn_empty[buf], ((j >> 1) & 1) ^ 1);
    if constexpr (MLOAD) {
      if (t.img != cur_img) {
        cur_img = t.img;
        ++kimg;
        const int mb = kimg & 1;
        ptx::mbar_wait(&bars->m_empty[mb], ((kimg >> 1) & 1) ^ 1);
        const uint4* src = reinterpret_cast<const uint4*>(mcanon + static_cast<size_t>(t.img) * C * kHid);
        const uint32_t dst = ptx::smem_u32(m_base + mb * (C * kHid * 2));
#pragma unroll
        for (int i = 0; i < C / 8; ++i) ptx::cp_async16_zfill(dst + (i * 128 + tid) * 16, src + i * 128 + tid, 16u);
        ptx::cp_async_commit();
      }
    }
    uint8_t* xb = xn_base + buf * XNB;
#pragma unroll
    for (int pb = 0; pb < PASSES; pb += BATCH) {
      // one-pass statistics (sum, sum of squares of the fp16 inputs in fp32; var = E[x^2] - mean^2, clamped): one FFMA per
      // element instead of a centring pass -- the producers share their issue slots with the row warps
      float f[BATCH][16], sum[BATCH], sq[BATCH];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        float sm[4] = {0.f, 0.f, 0.f, 0.f}, sqp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const __half2* hh = reinterpret_cast<const __half2*>(&cur[pb + b][v]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 tt = __half22float2(hh[i]);
            f[b][v * 8 + 2 * i] = tt.x;
            f[b][v * 8 + 2 * i + 1] = tt.y;
            sm[i] += tt.x + tt.y;
            sqp[i] = fmaf(tt.x, tt.x, sqp[i]);
            sqp[i] = fmaf(tt.y, tt.y, sqp[i]);
          }
        }
        sum[b] = (sm[0] + sm[1]) + (sm[2] + sm[3]);
        sq[b] = (sqp[0] + sqp[1]) + (sqp[2] + sqp[3]);
        load(tn, pb + b);                           // registers of this pass are free: next tile's rows
      }
#pragma unroll
      for (int o = 1; o < LPR; o <<= 1)
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          sum[b] += __shfl_xor_sync(0xffffffffu, sum[b], o);
          sq[b] += __shfl_xor_sync(0xffffffffu, sq[b], o);
        }
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int row = RPP * (pb + b) + rsub;
        const float mean = sum[b] * (1.0f / C);
        const float var = fmaxf(fmaf(-mean, mean, sq[b] * (1.0f / C)), 0.f);
        const float rstd = (row < t.nv) ? rsqrtf(var + eps) : 0.f;
        const float sh = -mean * rstd;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          uint4 ov;
          uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = pack2(fmaf(f[b][v * 8 + 2 * i], rstd, sh), fmaf(f[b][v * 8 + 2 * i + 1], rstd, sh));
          *reinterpret_cast<uint4*>(xb + ((q * 2 + v) * 128 + row) * 16) = ov;
        }
      }
    }
    if constexpr (MLOAD) ptx::cp_async_wait<0>();
    ptx::fence_proxy_async_smem();
    ptx::mbar_arrive(&bars->xn_full[buf]);
  }
}

// named barrier of the two row warps that share a TMEM lane quarter
__device__ __forceinline__ void pair_sync(int quarter) { asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory"); }

// ------------------------------------------------------------------ la1_tc
// shared memory: bars | exchange fp32 [2][2][128] | Wkv [C/8][256][8] | xn x2 | ek x2 [16][128][8] | vT [16][128][8]
template <int C>
struct La1Map {
  static constexpr int oExch = kBarBytes;
  static constexpr int oW = oExch + 2 * 2 * 128 * 4;
  static constexpr int oXn = oW + 256 * C * 2;
  static constexpr int oEk = oXn + 2 * 128 * C * 2;
  static constexpr int oV = oEk + 2 * 128 * 128 * 2;
  static constexpr int total = oV + 128 * 128 * 2;
};

template <int C>
__global__ void __launch_bounds__(kThreads, 1) la1_tc_kernel(const __half* __restrict__ x,
                                                             const uint4* __restrict__ wkv, float* __restrict__ part, int n, int tpi,
                                                             long long n_tiles, int nparts, float eps) {
  extern __shared__ __align__(128) uint8_t smem[];
  using Map = La1Map<C>;
  Bars* bars = reinterpret_cast<Bars*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s0 = ptx::smem_u32(smem);
  pdl_trigger();
  // one-time setup: weights (plan constants), barriers, TMEM
  for (int i = tid; i < 256 * C / 8; i += kThreads) reinterpret_cast<uint4*>(smem + Map::oW)[i] = __ldg(wkv + i);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->xn_full[i], 128);
      ptx::mbar_init(&bars->xn_empty[i], 1);
      ptx::mbar_init(&bars->a_full[i], 1);
      ptx::mbar_init(&bars->a_empty[i], kRowThreads);
      ptx::mbar_init(&bars->b_full[i], 1);
      ptx::mbar_init(&bars->b_empty[i], kRowThreads);
      ptx::mbar_init(&bars->op_full[i], kRowThreads);
      ptx::mbar_init(&bars->op_empty[i], 1);
      ptx::mbar_init(&bars->m_empty[i], 1);
    }
    ptx::mbar_init(&bars->s_full, 1);
    ptx::mbar_init(&bars->s_empty, kRowThreads);
    ptx::fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(&bars->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const long long t0 = (n_tiles * blockIdx.x) / gridDim.x, t1 = (n_tiles * (blockIdx.x + 1)) / gridDim.x;
  const int T = static_cast<int>(t1 - t0);
  pdl_wait();

  if (warp < 4) {
    ln_role<C, false>(x, smem + Map::oXn, bars, t0, t1, tpi, n, eps, tid, nullptr, nullptr);
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- MMA issue
    // TMEM columns: K^T buffers 0..127 / 128..255, V^T 256..383, S 384..511
    const uint32_t id128 = ptx::make_idesc_f16(128, 0);
    auto issue_kv = [&](int buf, uint32_t w_off, uint32_t dcol) {
      // D[128 weight rows x 128 pixels] = W[rows x C] xn^T
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        const uint64_t ad = desc(s0 + Map::oW + w_off + ks * 2 * 256 * 16, 256u);
        const uint64_t bd = desc(s0 + Map::oXn + buf * (128 * C * 2) + ks * 2 * 128 * 16, 128u);
        ptx::tc_mma_f16(tmem + dcol, ad, bd, id128, ks > 0 ? 1u : 0u);
      }
    };
    auto issue_s = [&](int eb) {
      // S = ek v^T : K = 128 pixels
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t ad = desc(s0 + Map::oEk + eb * (128 * 128 * 2) + ks * 2 * 128 * 16, 128u);
        const uint64_t bd = desc(s0 + Map::oV + ks * 2 * 128 * 16, 128u);
        ptx::tc_mma_f16(tmem + 384u, ad, bd, id128, ks > 0 ? 1u : 0u);
      }
    };
    if (T > 0) {
      ptx::mbar_wait(&bars->xn_full[0], 0);
      ptx::tc_fence_after();
      if (lane == 0) {
        issue_kv(0, 0u, 0u);
        ptx::tc_commit(&bars->a_full[0]);
      }
      __syncwarp();
    }
    for (int j = 0; j < T; ++j) {
      // V^T(j): the row threads have drained V^T(j-1)
      ptx::mbar_wait(&bars->b_empty[0], (j & 1) ^ 1);
      ptx::tc_fence_after();
      if (lane == 0) {
        issue_kv(j & 1, 128u * 16u, 256u);
        ptx::tc_commit(&bars->b_full[0]);
        ptx::tc_commit(&bars->xn_empty[j & 1]);
      }
      __syncwarp();
      if (j >= 1) {
        ptx::mbar_wait(&bars->op_full[0], (j - 1) & 1);
        ptx::mbar_wait(&bars->s_empty, ((j - 1) & 1) ^ 1);
        ptx::tc_fence_after();
        if (lane == 0) {
          issue_s((j - 1) & 1);
          ptx::tc_commit(&bars->s_full);
        }
        __syncwarp();
      }
      if (j + 1 < T) {
        const int nb = (j + 1) & 1;
        ptx::mbar_wait(&bars->xn_full[nb], ((j + 1) >> 1) & 1);
        ptx::mbar_wait(&bars->a_empty[nb], (((j + 1) >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        if (lane == 0) {
          issue_kv(nb, 0u, static_cast<uint32_t>(nb) * 128u);
          ptx::tc_commit(&bars->a_full[nb]);
        }
        __syncwarp();
      }
    }
    if (T > 0) {
      ptx::mbar_wait(&bars->op_full[0], (T - 1) & 1);
      ptx::mbar_wait(&bars->s_empty, ((T - 1) & 1) ^ 1);
      ptx::tc_fence_after();
      if (lane == 0) {
        issue_s((T - 1) & 1);
        ptx::tc_commit(&bars->s_full);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- accumulator rows
    // thread = (head h, d) for K^T / S, (h, e) for V^T; the warps qd and qd + 4 of a lane quarter own pixel columns
    // [0,64) / [64,128) of K^T and V^T and columns e [0,16) / [16,32) of the S block of their head
    const int wi = warp - 4, qd = wi & 3, hf = wi >> 2, row = qd * 32 + lane;
    const uint32_t lane_t = tmem + (static_cast<uint32_t>(qd * 32) << 16);
    float* exch = reinterpret_cast<float*>(smem + Map::oExch);   // [parity][half][row]
    const int c0 = 64 * hf;
    float m_run = -INFINITY, z = 0.f, m_S = -INFINITY, m_ref_prev = -INFINITY;
    float S[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) S[e] = 0.f;
    int cur_img = -1, pend_img = -1;
    float pend_m = 0.f, pend_z = 0.f;

    auto flush = [&](int img, float m, float zz, int slot) {
      // z of the two column halves is combined here; `slot` = the exchange slot the latest tile did NOT use (its previous
      // reader finished before that tile's pair barrier), the second barrier keeps the next tile from overwriting it early
      float* ex = exch + (slot * 2) * 128;
      ex[hf * 128 + row] = zz;
      pair_sync(qd);
      const float ztot = zz + ex[(1 - hf) * 128 + row];
      pair_sync(qd);
      const bool first = static_cast<long long>(img) * tpi >= t0;
      const int pi = (nparts == 2 && !first) ? 1 : 0;
      float* dst = part + ((static_cast<size_t>(img) * nparts + pi) * 4 + qd) * kPartFloats;
      if (hf == 0) {
        dst[lane] = m;
        dst[32 + lane] = ztot;
      }
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        *reinterpret_cast<float4*>(dst + 64 + lane * 32 + 16 * hf + e) = make_float4(S[e], S[e + 1], S[e + 2], S[e + 3]);
      if (nparts == 2 && first && static_cast<long long>(img + 1) * tpi <= t1) {
        float* d1 = dst + 4 * kPartFloats;           // the image ends inside this range: its second part is empty
        if (hf == 0) {
          d1[lane] = -INFINITY;
          d1[32 + lane] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(d1 + 64 + lane * 32 + 16 * hf + e) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto merge = [&](int jj) {
      // S(jj) (relative to m_ref_prev) into the running S (relative to m_S)
      uint32_t r[16];
      ptx::mbar_wait(&bars->s_full, jj & 1);
      ptx::tc_fence_after();
      ptx::tmem_ld16(lane_t + 384u + static_cast<uint32_t>(qd * 32 + 16 * hf), r);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars->s_empty);
      const float fs = ex2((m_S - m_ref_prev) * kLog2e);   // m_S = -inf -> 0 ; equal -> 1
#pragma unroll
      for (int e = 0; e < 16; ++e) S[e] = fmaf(S[e], fs, __uint_as_float(r[e]));
      m_S = m_ref_prev;
    };

    TileIter it(t0, tpi, n);
    for (int j = 0; j < T; ++j, it.next()) {
      const Tile t = it.get();
      if (t.img != cur_img) {
        if (cur_img >= 0) {
          pend_img = cur_img;
          pend_m = m_run;
          pend_z = z;
        }
        cur_img = t.img;
        m_run = -INFINITY;
        z = 0.f;
      }
      const int buf = j & 1;
      const uint32_t kcol = lane_t + static_cast<uint32_t>(buf * 128 + c0);
      ptx::mbar_wait(&bars->a_full[buf], (j >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t ra[32], rb[32];
      ptx::tmem_ld32(kcol, ra);
      ptx::tmem_ld32(kcol + 32u, rb);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars->a_empty[buf]);           // this half of the K^T row now lives in registers
      // ---- max over the tile's valid pixels (own 64 columns, four chains; the other half through shared memory), then
      //      ek = exp(k - max) -> fp16 chunks of 8 pixels (K-major A operand of the S product).  Two instantiations: the
      //      column masks exist only in the one that handles an image's last, partial tile.
      float* ex = exch + (buf * 2) * 128;
      uint8_t* ekb = smem + Map::oEk + buf * (128 * 128 * 2);
      float m_new = 0.f;
      auto tile_body = [&](auto full_c) {
        constexpr bool FULL = decltype(full_c)::value;
        float mxs[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float va = (FULL || c0 + c < t.nv) ? __uint_as_float(ra[c]) : -INFINITY;
          const float vb2 = (FULL || c0 + 32 + c < t.nv) ? __uint_as_float(rb[c]) : -INFINITY;
          mxs[c & 3] = fmaxf(mxs[c & 3], fmaxf(va, vb2));
        }
        const float mx = fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3]));
        ex[hf * 128 + row] = mx;
        pair_sync(qd);
        m_new = fmaxf(m_run, fmaxf(mx, ex[(1 - hf) * 128 + row]));
        z *= ex2((m_run - m_new) * kLog2e);
        m_run = m_new;
        const float nb = -m_new * kLog2e;
        float zp[4] = {0.f, 0.f, 0.f, 0.f};
        auto emit = [&](const uint32_t (&r)[32], int cb) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 ov;
            uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = ch * 8 + 2 * i;
              float e0 = ex2(fmaf(__uint_as_float(r[c]), kLog2e, nb));
              float e1 = ex2(fmaf(__uint_as_float(r[c + 1]), kLog2e, nb));
              if (!FULL) {
                if (cb + c >= t.nv) e0 = 0.f;
                if (cb + c + 1 >= t.nv) e1 = 0.f;
              }
              zp[i] += e0 + e1;
              o[i] = pack2(e0, e1);
            }
            *reinterpret_cast<uint4*>(ekb + (((cb >> 3) + ch) * 128 + row) * 16) = ov;
          }
        };
        emit(ra, c0);
        emit(rb, c0 + 32);
        z += (zp[0] + zp[1]) + (zp[2] + zp[3]);
      };
      if (t.nv == 128) tile_body(std::true_type{});
      else tile_body(std::false_type{});
      // ---- S of the previous tile
      if (j >= 1) {
        merge(j - 1);
        if (pend_img >= 0) {
          flush(pend_img, pend_m, pend_z, 1 - buf);
          pend_img = -1;
#pragma unroll
          for (int e = 0; e < 16; ++e) S[e] = 0.f;
          m_S = -INFINITY;
        }
      }
      m_ref_prev = m_new;
      // ---- v^T row -> fp16 chunks of 8 pixels (K-major B operand)
      ptx::mbar_wait(&bars->b_full[0], j & 1);
      ptx::tc_fence_after();
      const uint32_t vcol = lane_t + 256u + static_cast<uint32_t>(c0);
      ptx::tmem_ld32(vcol, ra);
      ptx::tmem_ld32(vcol + 32u, rb);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars->b_empty[0]);
      uint8_t* vb = smem + Map::oV;
      auto emit_v = [&](const uint32_t (&r)[32], int cb) {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint4 ov;
          uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = pack2(__uint_as_float(r[ch * 8 + 2 * i]), __uint_as_float(r[ch * 8 + 2 * i + 1]));
          *reinterpret_cast<uint4*>(vb + (((cb >> 3) + ch) * 128 + row) * 16) = ov;
        }
      };
      emit_v(ra, c0);
      emit_v(rb, c0 + 32);
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars->op_full[0]);
    }
    if (T > 0) {
      merge(T - 1);
      flush(cur_img, m_run, z, 1 - ((T - 1) & 1));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------ la2_tc
// shared memory: bars | bias[C] | Wq [C/8][128][8] | M x2 [16][C][8] | xn x2 | qs x NQ [16][128][8] | stage 8 x [32][80 B]
template <int C, int NQ>
struct La2Map {
  static constexpr int oBias = kBarBytes;
  static constexpr int oWq = oBias + C * 4;
  static constexpr int oM = oWq + 128 * C * 2;
  static constexpr int oXn = oM + 2 * C * kHid * 2;
  static constexpr int oQs = oXn + 2 * 128 * C * 2;
  static constexpr int oStage = oQs + NQ * 128 * kHid * 2;
  static constexpr int kPitch = 80;                         // 32 channels at a time + 16 B pad
  static constexpr int total = oStage + 8 * 32 * kPitch;
};

template <int C, int NQ>
__global__ void __launch_bounds__(kThreads, 1) la2_tc_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                             const uint4* __restrict__ wq,
                                                             const __half* __restrict__ mcanon, const float* __restrict__ bias, int n,
                                                             int tpi, long long n_tiles, float eps) {
  extern __shared__ __align__(128) uint8_t smem[];
  using Map = La2Map<C, NQ>;
  Bars* bars = reinterpret_cast<Bars*>(smem);
  float* sbias = reinterpret_cast<float*>(smem + Map::oBias);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s0 = ptx::smem_u32(smem);
  pdl_trigger();
  for (int i = tid; i < 128 * C / 8; i += kThreads) reinterpret_cast<uint4*>(smem + Map::oWq)[i] = __ldg(wq + i);
  for (int i = tid; i < C; i += kThreads) sbias[i] = bias != nullptr ? __ldg(bias + i) : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->xn_full[i], 128);
      ptx::mbar_init(&bars->xn_empty[i], 1);
      ptx::mbar_init(&bars->a_full[i], 1);
      ptx::mbar_init(&bars->a_empty[i], kRowThreads);
      ptx::mbar_init(&bars->b_full[i], 1);
      ptx::mbar_init(&bars->b_empty[i], kRowThreads);
      ptx::mbar_init(&bars->op_full[i], kRowThreads);
      ptx::mbar_init(&bars->op_empty[i], 1);
      ptx::mbar_init(&bars->m_empty[i], 1);
    }
    ptx::mbar_init(&bars->s_full, 1);
    ptx::mbar_init(&bars->s_empty, kRowThreads);
    ptx::fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(&bars->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const long long t0 = (n_tiles * blockIdx.x) / gridDim.x, t1 = (n_tiles * (blockIdx.x + 1)) / gridDim.x;
  const int T = static_cast<int>(t1 - t0);
  pdl_wait();

  if (warp < 4) {
    ln_role<C, true>(x, smem + Map::oXn, bars, t0, t1, tpi, n, eps, tid, mcanon, smem + Map::oM);
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- MMA issue
    // TMEM columns: Q buffers 0..127 / 128..255, Y buffers 256.. / 384..
    const uint32_t id128 = ptx::make_idesc_f16(128, 0), idC = ptx::make_idesc_f16(C, 0);
    auto issue_q = [&](int buf) {
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        const uint64_t ad = desc(s0 + Map::oXn + buf * (128 * C * 2) + ks * 2 * 128 * 16, 128u);
        const uint64_t bd = desc(s0 + Map::oWq + ks * 2 * 128 * 16, 128u);
        ptx::tc_mma_f16(tmem + static_cast<uint32_t>(buf) * 128u, ad, bd, id128, ks > 0 ? 1u : 0u);
      }
    };
    if (T > 0) {
      ptx::mbar_wait(&bars->xn_full[0], 0);
      ptx::tc_fence_after();
      if (lane == 0) {
        issue_q(0);
        ptx::tc_commit(&bars->a_full[0]);
        ptx::tc_commit(&bars->xn_empty[0]);
      }
      __syncwarp();
    }
    int cur_img = -1, kimg = -1;
    TileIter it(t0, tpi, n);
    for (int j = 0; j < T; ++j) {
      if (j + 1 < T) {
        const int nb = (j + 1) & 1;
        ptx::mbar_wait(&bars->xn_full[nb], ((j + 1) >> 1) & 1);
        ptx::mbar_wait(&bars->a_empty[nb], (((j + 1) >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        if (lane == 0) {
          issue_q(nb);
          ptx::tc_commit(&bars->a_full[nb]);
          ptx::tc_commit(&bars->xn_empty[nb]);
        }
        __syncwarp();
      }
      if (it.img != cur_img) {
        cur_img = it.img;
        ++kimg;
      }
      it.next();
      const bool last_of_img = (j + 1 == T) || (it.img != cur_img);
      const int qb = j % NQ, yb = j & 1, mb = kimg & 1;
      ptx::mbar_wait(&bars->op_full[qb], (j / NQ) & 1);
      ptx::mbar_wait(&bars->b_empty[yb], ((j >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      if (lane == 0) {
        // Y[128 pixels x C] = softmax(q) M_img^T : K = 128 hidden channels
#pragma unroll
        for (int ks = 0; ks < kHid / 16; ++ks) {
          const uint64_t ad = desc(s0 + Map::oQs + qb * (128 * kHid * 2) + ks * 2 * 128 * 16, 128u);
          const uint64_t bd = desc(s0 + Map::oM + mb * (C * kHid * 2) + ks * 2 * C * 16, static_cast<uint32_t>(C));
          ptx::tc_mma_f16(tmem + 256u + static_cast<uint32_t>(yb) * 128u, ad, bd, idC, ks > 0 ? 1u : 0u);
        }
        ptx::tc_commit(&bars->b_full[yb]);
        ptx::tc_commit(&bars->op_empty[qb]);
        if (last_of_img) ptx::tc_commit(&bars->m_empty[mb]);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- accumulator rows: thread = pixel; the warps qd and qd + 4 of a
    // lane quarter own heads {0,1} / {2,3} of Q and the lower / upper half of the output channels
    const int wi = warp - 4, qd = wi & 3, hf = wi >> 2, row = qd * 32 + lane;
    const uint32_t lane_t = tmem + (static_cast<uint32_t>(qd * 32) << 16);
    uint8_t* stage = smem + Map::oStage + wi * 32 * Map::kPitch;
    constexpr int NSUB = C / 64;               // 32-channel blocks of this warp's half of the output channels
    // y epilogue: 4 16-byte chunks per 32-channel block of a row, 8 rows per warp instruction
    const int ch = lane & 3, rsub = lane >> 2;
    float b8[NSUB][8];
#pragma unroll
    for (int sb = 0; sb < NSUB; ++sb)
#pragma unroll
      for (int i = 0; i < 8; ++i) b8[sb][i] = sbias[hf * (C / 2) + sb * 32 + ch * 8 + i];

    auto yepi = [&](int jj, const Tile& t) {
      const int yb = jj & 1;
      const size_t base = (static_cast<size_t>(t.img) * n + t.p0 + qd * 32) * C + hf * (C / 2) + ch * 8;
      // residual rows: independent of the MMA, fetched before waiting for it
      uint4 xr[NSUB][4];
#pragma unroll
      for (int sb = 0; sb < NSUB; ++sb)
#pragma unroll
        for (int it4 = 0; it4 < 4; ++it4) {
          const int rr = it4 * 8 + rsub;
          xr[sb][it4] = (qd * 32 + rr < t.nv) ? __ldg(reinterpret_cast<const uint4*>(x + base + static_cast<size_t>(rr) * C + sb * 32))
                                              : make_uint4(0u, 0u, 0u, 0u);
        }
      ptx::mbar_wait(&bars->b_full[yb], (jj >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t ycol = lane_t + 256u + static_cast<uint32_t>(yb * 128 + hf * (C / 2));
      uint32_t r[NSUB][32];
#pragma unroll
      for (int sb = 0; sb < NSUB; ++sb) ptx::tmem_ld32(ycol + sb * 32, r[sb]);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars->b_empty[yb]);
#pragma unroll
      for (int sb = 0; sb < NSUB; ++sb) {
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint4 ov;
          uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = pack2(__uint_as_float(r[sb][c8 * 8 + 2 * i]), __uint_as_float(r[sb][c8 * 8 + 2 * i + 1]));
          *reinterpret_cast<uint4*>(stage + lane * Map::kPitch + c8 * 16) = ov;
        }
        __syncwarp();
        // coalesced pass: 8 rows (64-byte segments) per instruction
#pragma unroll
        for (int it4 = 0; it4 < 4; ++it4) {
          const int rr = it4 * 8 + rsub;
          if (qd * 32 + rr < t.nv) {
            const uint4 sv = *reinterpret_cast<const uint4*>(stage + rr * Map::kPitch + ch * 16);
            const __half2* sh = reinterpret_cast<const __half2*>(&sv);
            const __half2* xh = reinterpret_cast<const __half2*>(&xr[sb][it4]);
            uint4 ov;
            __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 a = __half22float2(sh[i]), b = __half22float2(xh[i]);
              oh[i] = h2_sat(a.x + b8[sb][2 * i] + b.x, a.y + b8[sb][2 * i + 1] + b.y);
            }
            *reinterpret_cast<uint4*>(y + base + static_cast<size_t>(rr) * C + sb * 32) = ov;
          }
        }
        __syncwarp();
      }
    };

    Tile prev = {0, 0, 0};
    TileIter it(t0, tpi, n);
    for (int j = 0; j < T; ++j, it.next()) {
      const Tile t = it.get();
      const int buf = j & 1, qb = j % NQ;
      ptx::mbar_wait(&bars->a_full[buf], (j >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t qcol = lane_t + static_cast<uint32_t>(buf * 128 + hf * 64);
      uint32_t ra[32], rb[32];
      ptx::tmem_ld32(qcol, ra);
      ptx::tmem_ld32(qcol + 32u, rb);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bars->a_empty[buf]);           // the two heads of this row now live in registers
      ptx::mbar_wait(&bars->op_empty[qb], ((j / NQ) & 1) ^ 1);
      uint8_t* qs = smem + Map::oQs + qb * (128 * kHid * 2);
      auto head = [&](const uint32_t (&r)[32], int h) {
        // softmax over the 32 dims of head h (scale is folded into M)
        float mx4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) mx4[c] = __uint_as_float(r[c]);
#pragma unroll
        for (int c = 4; c < 32; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(r[c]));
        const float nb = -fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * kLog2e;
        float e[32];
        float sp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          e[c] = ex2(fmaf(__uint_as_float(r[c]), kLog2e, nb));
          sp[c & 3] += e[c];
        }
        const float inv = __fdividef(1.0f, (sp[0] + sp[1]) + (sp[2] + sp[3]));
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint4 ov;
          uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = pack2(e[c8 * 8 + 2 * i] * inv, e[c8 * 8 + 2 * i + 1] * inv);
          *reinterpret_cast<uint4*>(qs + ((h * 4 + c8) * 128 + row) * 16) = ov;
        }
      };
      head(ra, 2 * hf);
      head(rb, 2 * hf + 1);
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars->op_full[qb]);
      if (j >= 1) yepi(j - 1, prev);
      prev = t;
    }
    if (T > 0) yepi(T - 1, prev);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// tiles per image / grid: every CTA's contiguous tile range must be at least one image long, so that an image is split over
// at most two CTAs (two partials per image)
struct TcGeom {
  int tpi, grid, nparts;
  long long n_tiles;
};
bool tc_geom(long long n_img, int n_pos, int C, TcGeom* g) {
  if (C != 64 && C != 128) return false;
  if (n_img < 1 || n_pos < 1 || n_img > (1 << 20)) return false;
  const int tpi = (n_pos + 127) / 128;
  if (tpi > 4096) return false;
  const long long cap = num_sms();
  g->tpi = tpi;
  g->n_tiles = n_img * tpi;
  g->grid = static_cast<int>(n_img < cap ? n_img : cap);
  g->nparts = tpi > 1 ? 2 : 1;
  return true;
}

template <int C, int NQ>
int launch_tc(const __half* x, __half* y, const uint4* wq, const uint4* wkv, const float* wout, const float* bias,
              void* work, int n_img, int n_pos, float scale, float eps, const TcGeom& g, cudaStream_t st) {
  float* part = static_cast<float*>(work);
  __half* mcanon = reinterpret_cast<__half*>(part + static_cast<size_t>(n_img) * g.nparts * 4 * kPartFloats);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(la1_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, La1Map<C>::total);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la2_tc_kernel<C, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, La2Map<C, NQ>::total);
    if (e != cudaSuccess) return set_cuda_error(e, "linattn_block_tc: cudaFuncSetAttribute");
    configured = true;
  }
  cudaError_t le = launch_pdl(la1_tc_kernel<C>, dim3(g.grid), dim3(kThreads), static_cast<size_t>(La1Map<C>::total), st, x, wkv, part,
                              n_pos, g.tpi, g.n_tiles, g.nparts, eps);
  if (le != cudaSuccess) return set_cuda_error(le, "linattn_block_tc: la1 launch");
  const int rc = launch_la_mid(part, wout, mcanon, C, g.nparts, scale, 1, n_img, st);
  if (rc != WDNO_OK) return rc;
  le = launch_pdl(la2_tc_kernel<C, NQ>, dim3(g.grid), dim3(kThreads), static_cast<size_t>(La2Map<C, NQ>::total), st, x, y, wq,
                  static_cast<const __half*>(mcanon), bias, n_pos, g.tpi, g.n_tiles, eps);
  if (le != cudaSuccess) return set_cuda_error(le, "linattn_block_tc: la2 launch");
  return check_launch("linattn_block_tc");
}

static_assert(La1Map<64>::total <= 227 * 1024 && La1Map<128>::total <= 227 * 1024, "la1_tc shared-memory plan");
static_assert(La2Map<64, 2>::total <= 227 * 1024 && La2Map<128, 1>::total <= 227 * 1024, "la2_tc shared-memory plan");

}  // namespace

}  // namespace wdno

extern "C" int wdno_linattn_tc_supported(int64_t n_img, int n_pos, int C) {
  wdno::TcGeom g;
  return wdno::tc_geom(n_img, n_pos, C, &g) ? 1 : 0;
}

extern "C" int wdno_linattn_block_tc(const void* x, void* y, const void* wq_canon, const void* wkv_canon,
                                     const float* wout, const float* bias, void* work, int64_t n_img, int n_pos, int C, float scale,
                                     float eps, void* stream) {
  using namespace wdno;
  if (!x || !y || !wq_canon || !wkv_canon || !wout || !work) return set_error(WDNO_E_INVALID, "linattn_block_tc: bad arguments");
  TcGeom g;
  if (!tc_geom(n_img, n_pos, C, &g)) return set_error(WDNO_E_INVALID, "linattn_block_tc: shape outside the envelope (use wdno_linattn_block)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* xi = static_cast<const __half*>(x);
  __half* yo = static_cast<__half*>(y);
  const uint4* wq = static_cast<const uint4*>(wq_canon);
  const uint4* wkv = static_cast<const uint4*>(wkv_canon);
  if (C == 64) return launch_tc<64, 2>(xi, yo, wq, wkv, wout, bias, work, static_cast<int>(n_img), n_pos, scale, eps, g, st);
  return launch_tc<128, 1>(xi, yo, wq, wkv, wout, bias, work, static_cast<int>(n_img), n_pos, scale, eps, g, st);
}
