// Temporal attention block of the smoke U-Net at C = 64 (the four full-resolution instances: 14 % of a DDIM step) with the
// projections on tcgen05:   y = x + to_out(softmax(rot(q) rot(k)^T + rel_pos_bias) v),  (q, k, v) = to_qkv(LayerNorm(x))
// reference: Residual(PreNorm(dim, EinopsToAndFrom('b c f h w', 'b (h w) f c', Attention))), conv3d.py:165-184, 262-353, 383.
//
// Why: in the warp-per-pixel mma.sync kernel (attn_fused.cu) every warp streams all 64 KB of weight fragments through the
// shared-memory port for ONE pixel's 32 token rows -- 0.034 B/FLOP against a 128 B/clk port caps it at half the tensor rate,
// and 82 % of its MMAs are the two projections.  Here a CTA owns tiles of 4 pixels x 32 token rows = one M = 128 UMMA tile:
//   raw tokens (cp.async, next tile prefetched) -> LayerNorm by the lane that owns the token -> A operand in the UMMA
//   canonical K-major layout -> ONE thread issues  QKV[128 x 384] = A[128 x 64] Wqkv^T  (tcgen05.mma, weights resident in
//   shared memory for the CTA's lifetime, accumulators in TMEM columns 0..383) -> each warp pulls ITS pixel's q / k / v rows
//   out of TMEM (lane = token), applies scale + rotary, and runs the 24 x 24 attention of the four heads on mma.sync from a
//   private fp16 staging tile (no block barrier) -> O[128 x 128] written straight into the canonical A layout ->
//   Y[128 x 64] = O Wout^T (tcgen05.mma, TMEM columns 384..447) -> + residual from the raw tile -> 128-byte row stores.
// HBM traffic: one read + one write of the fp16 residual stream (the algorithmic minimum); q / k / v / o never leave the SM.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/wdno_b200.h"
#include "common.cuh"
#include "cvt_sat.cuh"
#include "mma_sync.cuh"
#include "ptx.cuh"

namespace wdno {

namespace {

constexpr int C = 64, kHid = 128, kQkv = 384;
constexpr int kThreads = 512;                    // 16 warps: warp = 4 * head + pixel slot (TMEM lane quarter = warp % 4)
constexpr int kRawPitch = C * 2 + 16;            // 144 B per raw token row
constexpr int kStRow = 64;                       // bytes per staging row: 32 halfs, 16-byte chunks XOR-swizzled by (row >> 1) & 3
constexpr int BS = 40, RS = 17;                  // bias (fp16) / rotary (float2: 136 B rows, lane-strided reads are conflict-free) table pitches
// shared-memory map (bytes)
constexpr int oBar = 0;                                   // 2 mbarriers + tmem base
constexpr int oWq = 128;                                  // [8][384][8] fp16
constexpr int oWo = oWq + kQkv * C * 2;                   // [16][64][8] fp16
constexpr int oAO = oWo + C * kHid * 2;                   // A [8][128][8] (16 KB) aliased by O [16][128][8] (32 KB)
constexpr int oRaw = oAO + 128 * kHid * 2;                // [128][144 B]: next tile's raw tokens (LayerNorm input)
constexpr int oStage = oRaw + 128 * kRawPitch;            // 16 warps x 3 x [32][64 B]
constexpr int oBias = oStage + 16 * 3 * 32 * kStRow;      // [4][32][BS] fp16
constexpr int oRot = oBias + 4 * 32 * BS * 2;             // [32][RS] float2
constexpr int oGamma = oRot + 32 * RS * 8;                // [C] fp32
constexpr int kSmem = oGamma + C * 4;
static_assert(kSmem <= 227 * 1024, "tattn_tc shared-memory plan");

// byte offset of 16-byte chunk `c` of staging row `r`
__device__ __forceinline__ uint32_t st_off(int r, int c) { return static_cast<uint32_t>(r * kStRow + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo16) {
  // K-major, no swizzle: start >> 4 | (LBO >> 4) << 16 | (SBO = 128 B >> 4) << 32 | version 1 << 46
  return static_cast<uint64_t>(((saddr >> 4) & 0x3FFFu) | (lbo16 << 16)) | (static_cast<uint64_t>(8u | (1u << 14)) << 32);
}

__global__ void __launch_bounds__(kThreads, 1) tattn_tc_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                               const float* __restrict__ gamma, const uint4* __restrict__ wq,
                                                               const uint4* __restrict__ wo, const float* __restrict__ bias,
                                                               const float* __restrict__ rot_cos, const float* __restrict__ rot_sin,
                                                               long long n_pix, long long hw, int n, float scale, float eps) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + oBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + oBar + 32);
  __half* sbias = reinterpret_cast<__half*>(smem + oBias);
  float2* scs = reinterpret_cast<float2*>(smem + oRot);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const uint32_t s0 = ptx::smem_u32(smem);
  pdl_trigger();

  // ---------------------------------------------------------------- one-time setup: weights, tables, TMEM, barriers
  for (int i = tid; i < kQkv * C / 8; i += kThreads) reinterpret_cast<uint4*>(smem + oWq)[i] = __ldg(wq + i);
  for (int i = tid; i < C * kHid / 8; i += kThreads) reinterpret_cast<uint4*>(smem + oWo)[i] = __ldg(wo + i);
  for (int i = tid; i < 4 * 32 * 32; i += kThreads) {
    const int hh = i >> 10, r = (i >> 5) & 31, c = i & 31;
    float b = 0.f;
    if (c >= n) b = -INFINITY;
    else if (r < n && bias != nullptr) b = __ldg(bias + (static_cast<size_t>(hh) * n + r) * n + c);
    sbias[(hh * 32 + r) * BS + c] = __float2half(b);
  }
  for (int i = tid; i < 32 * 16; i += kThreads) {
    const int f = i >> 4;
    scs[f * RS + (i & 15)] = make_float2((rot_cos != nullptr && f < n) ? __ldg(rot_cos + f * 16 + (i & 15)) : 1.0f,
                                         (rot_sin != nullptr && f < n) ? __ldg(rot_sin + f * 16 + (i & 15)) : 0.0f);
  }
  for (int i = tid; i < 128 * kHid / 8; i += kThreads) reinterpret_cast<uint4*>(smem + oAO)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 16 * 3 * 32 * kStRow / 16; i += kThreads) reinterpret_cast<uint4*>(smem + oStage)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  float* sgm = reinterpret_cast<float*>(smem + oGamma);
  for (int j = tid; j < C; j += kThreads) sgm[j] = __ldg(gamma + j);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int slot = warp & 3, head = warp >> 2;            // pixel slot (= TMEM lane quarter) and head of this warp
  const uint32_t lane_t = tmem + (static_cast<uint32_t>(slot * 32) << 16);
  pdl_wait();

  const long long n_tiles = (n_pix + 3) >> 2;
  const int row = slot * 32 + lane;                       // this thread's row of the M = 128 tile (token = lane)
  const bool tok = lane < n;
  // staging tiles of this warp: q', k', v rows of ITS head
  uint8_t* Qs = smem + oStage + warp * 3 * 32 * kStRow;
  uint8_t* Ks = Qs + 32 * kStRow;
  uint8_t* Vs = Ks + 32 * kStRow;
  const uint32_t qs_s = ptx::smem_u32(Qs), ks_s = ptx::smem_u32(Ks), vs_s = ptx::smem_u32(Vs);

  // raw tokens / LayerNorm: warp (slot, head) owns tokens 8 head .. 8 head + 7 of its pixel slot, 4 lanes per token (16 channels
  // each) -- the thread that fetched a piece normalises it, so no block barrier sits between the copy and the LayerNorm
  const int ln_t = head * 8 + (lane >> 2), ln_q = lane & 3;
  const bool ln_tok = ln_t < n;
  const int ln_row = slot * 32 + ln_t;
  auto prefetch = [&](long long tile) {
    const long long pix = tile * 4 + slot;
    if (tile < n_tiles && pix < n_pix && ln_tok) {
      const long long bimg = pix / hw, pin = pix - bimg * hw;
      const __half* src = x + ((bimg * n + ln_t) * hw + pin) * C + ln_q * 16;
      const uint32_t dst = s0 + oRaw + ln_row * kRawPitch + ln_q * 32;
      ptx::cp_async16_zfill(dst, src, 16u);
      ptx::cp_async16_zfill(dst + 16, src + 8, 16u);
    }
    ptx::cp_async_commit();
  };

  uint32_t ph = 0;
  prefetch(blockIdx.x);
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long pix = tile * 4 + slot;
    const bool live = tok && pix < n_pix;
    // ---- LayerNorm: 4 lanes per token -> A operand [C/8][128][8]
    {
      ptx::cp_async_wait<0>();                            // this thread's own piece has landed
      const bool ln_live = ln_tok && pix < n_pix;
      const uint8_t* rawp = smem + oRaw + ln_row * kRawPitch + ln_q * 32;
      float f[16];
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ln_live) v = *reinterpret_cast<const uint4*>(rawp + c * 16);
        const __half2* hh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(hh[j]);
          f[c * 8 + 2 * j] = t.x;
          f[c * 8 + 2 * j + 1] = t.y;
          sum += t.x + t.y;
        }
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float mean = sum * (1.0f / C);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        f[j] -= mean;
        sq = fmaf(f[j], f[j], sq);
      }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      const float rstd = ln_live ? rsqrtf(sq * (1.0f / C) + eps) : 0.f;
      if (ln_t < 32) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 ov;
          uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 gg = *reinterpret_cast<const float2*>(sgm + ln_q * 16 + c * 8 + 2 * j);
            o[j] = pack_h2(f[c * 8 + 2 * j] * rstd * gg.x, f[c * 8 + 2 * j + 1] * rstd * gg.y);
          }
          *reinterpret_cast<uint4*>(smem + oAO + ((ln_q * 2 + c) * 128 + ln_row) * 16) = ov;
        }
      }
      prefetch(tile + gridDim.x);                         // the raw piece is consumed: fetch the next tile's behind the MMAs / attention
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    // ---- QKV = A Wqkv^T : M = 128, N = 256 + 128, K = 64 (4 k-steps)
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t id256 = ptx::make_idesc_f16(256, 0), id128 = ptx::make_idesc_f16(128, 0);
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        const uint64_t ad = desc(s0 + oAO + ks * 2 * 128 * 16, 128u);
        const uint64_t bd = desc(s0 + oWq + ks * 2 * kQkv * 16, static_cast<uint32_t>(kQkv));
        ptx::tc_mma_f16(tmem, ad, bd, id256, ks > 0 ? 1u : 0u);
        ptx::tc_mma_f16(tmem + 256u, ad, bd + 256u, id128, ks > 0 ? 1u : 0u);   // rows 256..383 of Wqkv: + 256 x 16 B
      }
      ptx::tc_commit(&bars[0]);
    }
    ptx::mbar_wait(&bars[0], ph);
    ptx::tc_fence_after();

    // ---- attention of (pixel slot, head) = this warp: all warp-local, 16 warps in flight
    {
      const int h = head;
      {
        // one TMEM round trip for q, k, v of this head (lane = token), then scale + rotary and the fp16 staging rows
        uint32_t rq[32], rk[32], rv[32];
        ptx::tmem_ld32(lane_t + static_cast<uint32_t>(h * 32), rq);
        ptx::tmem_ld32(lane_t + static_cast<uint32_t>(kHid + h * 32), rk);
        ptx::tmem_ld32(lane_t + static_cast<uint32_t>(2 * kHid + h * 32), rv);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 oq, ok, ov;
          uint32_t* pq = reinterpret_cast<uint32_t*>(&oq);
          uint32_t* pk = reinterpret_cast<uint32_t*>(&ok);
          uint32_t* pv = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 cs = scs[lane * RS + c * 4 + j];
            const float a = __uint_as_float(rq[c * 8 + 2 * j]) * scale, b = __uint_as_float(rq[c * 8 + 2 * j + 1]) * scale;
            pq[j] = pack_h2(a * cs.x - b * cs.y, b * cs.x + a * cs.y);
            const float d = __uint_as_float(rk[c * 8 + 2 * j]), e = __uint_as_float(rk[c * 8 + 2 * j + 1]);
            pk[j] = pack_h2(d * cs.x - e * cs.y, e * cs.x + d * cs.y);
            pv[j] = pack_h2(__uint_as_float(rv[c * 8 + 2 * j]), __uint_as_float(rv[c * 8 + 2 * j + 1]));
          }
          *reinterpret_cast<uint4*>(Qs + st_off(lane, c)) = oq;
          *reinterpret_cast<uint4*>(Ks + st_off(lane, c)) = ok;
          *reinterpret_cast<uint4*>(Vs + st_off(lane, c)) = ov;
        }
      }
      __syncwarp();
      // S = Q K^T (+ bias, keys >= n masked by the table), softmax over the keys
      uint32_t qa[2][2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)   // A frags: rows = queries mt*16 + (lane & 15), dim chunk 2 ks + (lane >> 4)
          ldsm_x4(qs_s + st_off(mt * 16 + (lane & 15), 2 * ks + (lane >> 4)), qa[mt][ks][0], qa[mt][ks][1], qa[mt][ks][2], qa[mt][ks][3]);
      float sfr[2][4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t b0, b1, b2, b3;         // B frags of K: keys j*8 + (lane & 7), dim chunk lane >> 3
        ldsm_x4(ks_s + st_off(j * 8 + (lane & 7), lane >> 3), b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
          for (int c = 0; c < 4; ++c) sfr[mt][j][c] = 0.f;
          mma16816(sfr[mt][j], qa[mt][0], b0, b1);
          mma16816(sfr[mt][j], qa[mt][1], b2, b3);
        }
      }
      uint32_t pa[2][2][4];
      float inv[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const __half* brow = sbias + (h * 32 + 16 * mt + g + 8 * r) * BS + 2 * q;
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 bv = __half22float2(*reinterpret_cast<const __half2*>(brow + 8 * j));
            sfr[mt][j][2 * r] += bv.x;
            sfr[mt][j][2 * r + 1] += bv.y;
            mx = fmaxf(mx, fmaxf(sfr[mt][j][2 * r], sfr[mt][j][2 * r + 1]));
          }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float p0 = __expf(sfr[mt][j][2 * r] - mx), p1 = __expf(sfr[mt][j][2 * r + 1] - mx);
            sfr[mt][j][2 * r] = p0;
            sfr[mt][j][2 * r + 1] = p1;
            sum += p0 + p1;
          }
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          sum += __shfl_xor_sync(0xffffffffu, sum, 2);
          inv[mt][r] = 1.0f / sum;
        }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          pa[mt][kk][0] = pack_h2(sfr[mt][2 * kk][0], sfr[mt][2 * kk][1]);
          pa[mt][kk][1] = pack_h2(sfr[mt][2 * kk][2], sfr[mt][2 * kk][3]);
          pa[mt][kk][2] = pack_h2(sfr[mt][2 * kk + 1][0], sfr[mt][2 * kk + 1][1]);
          pa[mt][kk][3] = pack_h2(sfr[mt][2 * kk + 1][2], sfr[mt][2 * kk + 1][3]);
        }
      }
      // O = P V (V^T fragments by ldmatrix.trans), normalised, written into the canonical A layout of the output projection
#pragma unroll
      for (int np = 0; np < 2; ++np) {      // pairs of 8-wide dim tiles
        float ofr[2][2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) ofr[i][j][c] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          uint32_t b0, b1, b2, b3;         // keys kk*16 + (lane & 7) + 8 ((lane >> 3) & 1), dim chunk 2 np + (lane >> 4)
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                       : "r"(vs_s + st_off(kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * np + (lane >> 4))));
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma16816(ofr[mt][0], pa[mt][kk], b0, b1);
            mma16816(ofr[mt][1], pa[mt][kk], b2, b3);
          }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int chunk = h * 4 + np * 2 + j;       // 8-channel chunk of the 128 hidden channels
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int rr = slot * 32 + 16 * mt + g + 8 * r;
              *reinterpret_cast<uint32_t*>(smem + oAO + (chunk * 128 + rr) * 16 + q * 4) =
                  pack_h2(ofr[mt][j][2 * r] * inv[mt][r], ofr[mt][j][2 * r + 1] * inv[mt][r]);
            }
          }
      }
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    // ---- Y = O Wout^T : M = 128, N = 64, K = 128 (8 k-steps), TMEM columns 384..447
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t id64 = ptx::make_idesc_f16(64, 0);
#pragma unroll
      for (int ks = 0; ks < kHid / 16; ++ks) {
        const uint64_t ad = desc(s0 + oAO + ks * 2 * 128 * 16, 128u);
        const uint64_t bd = desc(s0 + oWo + ks * 2 * C * 16, static_cast<uint32_t>(C));
        ptx::tc_mma_f16(tmem + 384u, ad, bd, id64, ks > 0 ? 1u : 0u);
      }
      ptx::tc_commit(&bars[1]);
    }
    // the residual (L2-resident: this tile was read a few microseconds ago) is fetched while the MMA runs
    uint4 xres[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
    size_t yoff = 0;
    if (live) {
      const long long bimg = pix / hw, pin = pix - bimg * hw;
      yoff = static_cast<size_t>(((bimg * n + lane) * hw + pin) * C + head * 16);
      xres[0] = __ldg(reinterpret_cast<const uint4*>(x + yoff));
      xres[1] = __ldg(reinterpret_cast<const uint4*>(x + yoff) + 1);
    }
    ptx::mbar_wait(&bars[1], ph);
    ptx::tc_fence_after();
    ph ^= 1u;
    // ---- y = Y + x : warp (slot, head) stores columns 16 head .. 16 head + 15 of its pixel's token rows (residual re-read: L2)
    {
      uint32_t r0[16];
      ptx::tmem_ld16(lane_t + 384u + static_cast<uint32_t>(head * 16), r0);
      ptx::tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const __half2* xh = reinterpret_cast<const __half2*>(&xres[c]);
          uint4 ov;
          __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 xr = __half22float2(xh[j]);
            oh[j] = h2_sat(__uint_as_float(r0[c * 8 + 2 * j]) + xr.x, __uint_as_float(r0[c * 8 + 2 * j + 1]) + xr.y);
          }
          *(reinterpret_cast<uint4*>(y + yoff) + c) = ov;
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();   // TMEM and the A / O region are free for the next tile
  }
  ptx::cp_async_wait<0>();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

}  // namespace wdno

extern "C" int wdno_tattn_block_tc(const void* x, void* y, const float* gamma, const void* wqkv_canon, const void* wout_canon,
                                   const float* bias, const float* rot_cos, const float* rot_sin, int64_t n_samples, int n_frames,
                                   int64_t hw, int C, float scale, float eps, void* stream) {
  using namespace wdno;
  if (!x || !y || !gamma || !wqkv_canon || !wout_canon || n_samples < 1 || hw < 1)
    return set_error(WDNO_E_INVALID, "tattn_block_tc: bad arguments");
  if (C != 64) return set_error(WDNO_E_INVALID, "tattn_block_tc: built for C = 64 (use wdno_tattn_block otherwise)");
  if (n_frames < 1 || n_frames > 32) return set_error(WDNO_E_INVALID, "tattn_block_tc: frames must be in [1,32]");
  if ((rot_cos == nullptr) != (rot_sin == nullptr)) return set_error(WDNO_E_INVALID, "tattn_block_tc: rotary tables must both be given");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tattn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_cuda_error(e, "tattn_block_tc: cudaFuncSetAttribute");
    configured = true;
  }
  const long long n_pix = n_samples * hw;
  const long long tiles = (n_pix + 3) / 4;
  const long long cap = num_sms();
  const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
  cudaError_t le = launch_pdl(tattn_tc_kernel, dim3(grid), dim3(kThreads), static_cast<size_t>(kSmem), static_cast<cudaStream_t>(stream),
                              static_cast<const __half*>(x), static_cast<__half*>(y), gamma, static_cast<const uint4*>(wqkv_canon),
                              static_cast<const uint4*>(wout_canon), bias, rot_cos, rot_sin, n_pix, static_cast<long long>(hw), n_frames,
                              scale, eps);
  if (le != cudaSuccess) return set_cuda_error(le, "tattn_block_tc: launch");
  return check_launch("tattn_block_tc");
}
