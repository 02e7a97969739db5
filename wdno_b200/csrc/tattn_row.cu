// Temporal attention block of the smoke U-Net at C = 64 (the full-resolution instances) with EVERY product on tcgen05 and every
// softmax in the registers of the thread that owns the accumulator row:
//   y = x + to_out(softmax(rot(q) rot(k)^T * scale + rel_pos_bias) v),  (q, k, v) = to_qkv(LayerNorm(x))
// reference: Residual(PreNorm(dim, EinopsToAndFrom('b c f h w', 'b (h w) f c', Attention))), conv3d.py:165-184, 262-353, 383.
//
// tattn_tc.cu put the two projections on tcgen05 and kept the 24 x 24 attention of a (pixel, head) on mma.sync in one warp:
// 4 450 warp instructions per tile and SM sub-partition, most of them fragment shuffling, with block barriers between the
// phases (ncu: issue slots 34 % busy, 330 us per full-resolution launch).  Here a tile is 4 pixels x 32 token slots = the 128
// rows of a UMMA, and the attention itself is made of 128-row products:
//   QKV[128 x 384] = xn Wqkv^T                      -> row thread (lane = token): rotary, fp16 operand tiles
//   S_h[128 x 128] = Q_h K_h^T  (K = 32 dims)       -> all four pixels' keys at once; a row keeps the 32 columns of ITS pixel:
//                                                       + bias, max, exp, sum in registers (no shuffles), P_h -> fp16 operand
//   O_h,s[128 x 32] = P_h V_h,s (K = 32 keys of pixel slot s, V read MN-major) -> a row keeps the block of its slot, x 1/sum
//   Y[128 x 64]    = O Wout^T                       -> + residual -> 64-byte row segments
// The cross-pixel blocks of S and O are wasted MMA work (4x on two small products: 2 700 tensor cycles per tile in all), the
// price of turning ~1 900 fragment instructions per row into ~900 straight-line ones.
// Warp roles as in linattn_tc.cu: warps 0-3 LayerNorm producers (the next tile's rows prefetched in registers), warps 4-11 row
// warps (two per TMEM lane quarter, heads {0,1} / {2,3}), warp 12 issues the MMAs; mbarriers between them, no block barrier in
// the loop.  TMEM (512 columns): QKV 0..383 -> S_h at 128 h -> O_h,s at 128 h + 32 s -> Y at 448..511, so the next tile's QKV
// product runs while this tile's output rows are stored.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/wdno_b200.h"
#include "common.cuh"
#include "cvt_sat.cuh"
#include "ptx.cuh"

namespace wdno {

namespace {

constexpr int C = 64, kHid = 128, kQkv = 384;
constexpr int kThreads = 416, kMmaWarp = 12;
constexpr int BS = 40, RS = 17;                  // bias (fp16, 80-byte rows: LDS.128 conflict-free) / rotary (float2) table pitches
constexpr float kLog2e = 1.4426950408889634f;

struct Bars {
  uint64_t xn_full[2], xn_empty[2];
  uint64_t qkv_full, staged, o_full, y_full;
  uint64_t s_full[4], p_full[4], pv_full[4];
  uint32_t tmem_base;
};
constexpr int oBar = 0;
constexpr int oWq = 256;                                  // [8][384][8] fp16 (LayerNorm gain folded in)
constexpr int oWo = oWq + kQkv * C * 2;                   // [16][64][8]
constexpr int oXn = oWo + C * kHid * 2;                   // 2 x [8][128][8]
constexpr int oQ = oXn + 2 * 128 * C * 2;                 // [4 heads][4 chunks][128][8]; aliased by O [16][128][8]
constexpr int oK = oQ + 128 * kHid * 2;                   // [4][4][128][8]; head h aliased by P_h [4 key chunks][128][8]
constexpr int oV = oK + 128 * kHid * 2;                   // [4][4 dim chunks][128 keys][8 dims]  (MN-major B operand)
constexpr int oBias = oV + 128 * kHid * 2;                // [4][32][BS] fp16
constexpr int oRot = oBias + 4 * 32 * BS * 2;             // [32][RS] float2 (cos, sin); scale * log2 e is folded into the q rows of Wqkv
constexpr int kSmem = oRot + 32 * RS * 8;
static_assert(sizeof(Bars) <= 256 && kSmem <= 227 * 1024, "tattn_row shared-memory plan");

__device__ __forceinline__ uint64_t desc_k(uint32_t saddr, uint32_t lbo16) {
  // K-major, no swizzle: start >> 4 | (LBO >> 4) << 16 | (SBO = 128 B >> 4) << 32 | version 1 << 46
  return static_cast<uint64_t>(((saddr >> 4) & 0x3FFFu) | (lbo16 << 16)) | (static_cast<uint64_t>(8u | (1u << 14)) << 32);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t sbo16) {
  // MN-major, no swizzle: LBO = 128 B between 8-position K blocks, SBO = stride between 8-channel MN chunks (wgrad_tc.cu)
  return static_cast<uint64_t>(((saddr >> 4) & 0x3FFFu) | (8u << 16)) | (static_cast<uint64_t>((sbo16 & 0x3FFFu) | (1u << 14)) << 32);
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

__global__ void __launch_bounds__(kThreads, 1) tattn_row_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                                const uint4* __restrict__ wq, const uint4* __restrict__ wo,
                                                                const float* __restrict__ bias, const float* __restrict__ rot_cos,
                                                                const float* __restrict__ rot_sin, int n_pix, int hw, int n,
                                                                float scale, float eps) {
  extern __shared__ __align__(128) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + oBar);
  __half* sbias = reinterpret_cast<__half*>(smem + oBias);
  float2* rot = reinterpret_cast<float2*>(smem + oRot);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s0 = ptx::smem_u32(smem);
  pdl_trigger();

  // ---------------------------------------------------------------- one-time setup
  for (int i = tid; i < kQkv * C / 8; i += kThreads) reinterpret_cast<uint4*>(smem + oWq)[i] = __ldg(wq + i);
  for (int i = tid; i < C * kHid / 8; i += kThreads) reinterpret_cast<uint4*>(smem + oWo)[i] = __ldg(wo + i);
  for (int i = tid; i < 4 * 32 * 32; i += kThreads) {
    const int hh = i >> 10, r = (i >> 5) & 31, c = i & 31;
    float b = 0.f;
    if (c >= n) b = -INFINITY;
    else if (r < n && bias != nullptr) b = __ldg(bias + (static_cast<size_t>(hh) * n + r) * n + c);
    sbias[(hh * 32 + r) * BS + c] = __float2half(b * kLog2e);   // the softmax runs in base 2
  }
  for (int i = tid; i < 32 * 16; i += kThreads) {
    const int f = i >> 4;
    const float cs = (rot_cos != nullptr && f < n) ? __ldg(rot_cos + f * 16 + (i & 15)) : 1.0f;
    const float sn = (rot_sin != nullptr && f < n) ? __ldg(rot_sin + f * 16 + (i & 15)) : 0.0f;
    rot[f * RS + (i & 15)] = make_float2(cs, sn);
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->xn_full[i], 128);
      ptx::mbar_init(&bars->xn_empty[i], 1);
    }
    ptx::mbar_init(&bars->qkv_full, 1);
    ptx::mbar_init(&bars->staged, 256);
    ptx::mbar_init(&bars->o_full, 256);
    ptx::mbar_init(&bars->y_full, 1);
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&bars->s_full[i], 1);
      ptx::mbar_init(&bars->p_full[i], 128);
      ptx::mbar_init(&bars->pv_full[i], 1);
    }
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
  const int n_tiles = (n_pix + 3) >> 2;
  const int T = (n_tiles > static_cast<int>(blockIdx.x)) ? (n_tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
  pdl_wait();

  if (warp < 4) {
    // ---------------------------------------------------------------- LayerNorm producers: pass = pixel slot, 4 lanes per token
    const int q = tid & 3, tok = tid >> 2;
    const bool tok_ok = tok < n;
    uint4 cur[4][2];
    auto load = [&](int tile, int ps) {
      const int pix = tile * 4 + ps;
      const bool ok = tok_ok && tile < n_tiles && pix < n_pix;
      const int bimg = ok ? pix / hw : 0;
      const int pin = pix - bimg * hw;
      const uint4* src = reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(bimg) * n + tok) * hw + pin) * C + q * 16);
      cur[ps][0] = ok ? __ldg(src) : make_uint4(0u, 0u, 0u, 0u);
      cur[ps][1] = ok ? __ldg(src + 1) : make_uint4(0u, 0u, 0u, 0u);
    };
#pragma unroll
    for (int ps = 0; ps < 4; ++ps) load(blockIdx.x, ps);
    int tile = blockIdx.x;
    for (int j = 0; j < T; ++j, tile += gridDim.x) {
      const int buf = j & 1;
      ptx::mbar_wait(&bars->xn_empty[buf], ((j >> 1) & 1) ^ 1);
      uint8_t* xb = smem + oXn + buf * (128 * C * 2);
      float f[4][16], sum[4], sq[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float sm[4] = {0.f, 0.f, 0.f, 0.f}, sqp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const __half2* hh = reinterpret_cast<const __half2*>(&cur[b][v]);
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
        load(tile + gridDim.x, b);                     // registers of this pass are free: next tile's rows
      }
#pragma unroll
      for (int o = 1; o < 4; o <<= 1)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          sum[b] += __shfl_xor_sync(0xffffffffu, sum[b], o);
          sq[b] += __shfl_xor_sync(0xffffffffu, sq[b], o);
        }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int row = 32 * b + tok;
        const bool live = tok_ok && tile * 4 + b < n_pix;
        const float mean = sum[b] * (1.0f / C);
        const float var = fmaxf(fmaf(-mean, mean, sq[b] * (1.0f / C)), 0.f);
        const float rstd = live ? rsqrtf(var + eps) : 0.f;
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
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars->xn_full[buf]);
    }
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- MMA issue
    const uint32_t id256 = ptx::make_idesc_f16(256, 0), id128 = ptx::make_idesc_f16(128, 0), id64 = ptx::make_idesc_f16(64, 0);
    const uint32_t id32mn = ptx::make_idesc_f16(32, 0) | (1u << 16);     // B MN-major (V: dims contiguous)
    auto issue_qkv = [&](int buf) {
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks) {
        const uint64_t ad = desc_k(s0 + oXn + buf * (128 * C * 2) + ks * 2 * 128 * 16, 128u);
        const uint64_t bd = desc_k(s0 + oWq + ks * 2 * kQkv * 16, static_cast<uint32_t>(kQkv));
        ptx::tc_mma_f16(tmem, ad, bd, id256, ks > 0 ? 1u : 0u);
        ptx::tc_mma_f16(tmem + 256u, ad, bd + 256u, id128, ks > 0 ? 1u : 0u);   // rows 256..383 of Wqkv: + 256 x 16 B
      }
    };
    if (T > 0) {
      ptx::mbar_wait(&bars->xn_full[0], 0);
      ptx::tc_fence_after();
      if (lane == 0) {
        issue_qkv(0);
        ptx::tc_commit(&bars->qkv_full);
        ptx::tc_commit(&bars->xn_empty[0]);
      }
      __syncwarp();
    }
    for (int j = 0; j < T; ++j) {
      const uint32_t par = j & 1;
      // S_h = Q_h K_h^T for all heads once every q / k / v row has left TMEM
      ptx::mbar_wait(&bars->staged, par);
      ptx::tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int hi = 0; hi < 4; ++hi) {
          const int h = (hi & 1) * 2 + (hi >> 1);          // 0, 2, 1, 3: the first head of either row-warp group first
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t ad = desc_k(s0 + oQ + (h * 4 + ks * 2) * (128 * 16), 128u);
            const uint64_t bd = desc_k(s0 + oK + (h * 4 + ks * 2) * (128 * 16), 128u);
            ptx::tc_mma_f16(tmem + static_cast<uint32_t>(128 * h), ad, bd, id128, ks > 0 ? 1u : 0u);
          }
          ptx::tc_commit(&bars->s_full[h]);
        }
      }
      __syncwarp();
      // O_h,s = P_h V_h,s per pixel slot: K = that slot's 32 keys
#pragma unroll 1
      for (int hi = 0; hi < 4; ++hi) {
        const int h = (hi & 1) * 2 + (hi >> 1);
        ptx::mbar_wait(&bars->p_full[h], par);
        ptx::tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t ad = desc_k(s0 + oK + (h * 4 + ks * 2) * (128 * 16), 128u);                       // P_h: key chunks 2 ks, 2 ks + 1
              const uint64_t bd = desc_mn(s0 + oV + h * 4 * (128 * 16) + (32 * s + 16 * ks) * 16, 128u);       // V_h: keys 32 s + 16 ks ..
              ptx::tc_mma_f16(tmem + static_cast<uint32_t>(128 * h + 32 * s), ad, bd, id32mn, ks > 0 ? 1u : 0u);
            }
          ptx::tc_commit(&bars->pv_full[h]);
        }
        __syncwarp();
      }
      // Y = O Wout^T
      ptx::mbar_wait(&bars->o_full, par);
      ptx::tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < kHid / 16; ++ks) {
          const uint64_t ad = desc_k(s0 + oQ + ks * 2 * 128 * 16, 128u);
          const uint64_t bd = desc_k(s0 + oWo + ks * 2 * C * 16, static_cast<uint32_t>(C));
          ptx::tc_mma_f16(tmem + 448u, ad, bd, id64, ks > 0 ? 1u : 0u);
        }
        ptx::tc_commit(&bars->y_full);
      }
      __syncwarp();
      if (j + 1 < T) {
        const int nb = (j + 1) & 1;
        ptx::mbar_wait(&bars->xn_full[nb], ((j + 1) >> 1) & 1);
        ptx::tc_fence_after();
        if (lane == 0) {
          issue_qkv(nb);
          ptx::tc_commit(&bars->qkv_full);
          ptx::tc_commit(&bars->xn_empty[nb]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- row warps: thread = (pixel slot qd, token = lane), heads 2 hf, 2 hf + 1
    const int wi = warp - 4, qd = wi & 3, hf = wi >> 2, row = qd * 32 + lane;
    const uint32_t lane_t = tmem + (static_cast<uint32_t>(qd * 32) << 16);
    const bool tok_ok = lane < n;
    // pixel of this thread's slot, advanced by 4 * gridDim per tile without divisions
    int pix = static_cast<int>(blockIdx.x) * 4 + qd;
    int bimg = pix / hw, pin = pix - bimg * hw;
    const int step = 4 * static_cast<int>(gridDim.x);
    for (int j = 0; j < T; ++j) {
      const uint32_t par = j & 1;
      const bool live = tok_ok && pix < n_pix;
      const size_t yoff = ((static_cast<size_t>(bimg) * n + lane) * hw + pin) * C + hf * 32;
      // ---- q / k / v rows of the two heads: rotary -> operand tiles
      ptx::mbar_wait(&bars->qkv_full, par);
      ptx::tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * hf + hh;
        uint32_t rq[32], rk[32];
        ptx::tmem_ld32(lane_t + static_cast<uint32_t>(h * 32), rq);
        ptx::tmem_ld32(lane_t + static_cast<uint32_t>(kHid + h * 32), rk);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 oq, ok;
          uint32_t* pq = reinterpret_cast<uint32_t*>(&oq);
          uint32_t* pk = reinterpret_cast<uint32_t*>(&ok);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 cs = rot[lane * RS + c * 4 + i];
            const float a = __uint_as_float(rq[c * 8 + 2 * i]), b = __uint_as_float(rq[c * 8 + 2 * i + 1]);
            pq[i] = pack2(fmaf(a, cs.x, -b * cs.y), fmaf(b, cs.x, a * cs.y));
            const float d = __uint_as_float(rk[c * 8 + 2 * i]), e = __uint_as_float(rk[c * 8 + 2 * i + 1]);
            pk[i] = pack2(fmaf(d, cs.x, -e * cs.y), fmaf(e, cs.x, d * cs.y));
          }
          const int off = ((h * 4 + c) * 128 + row) * 16;
          *reinterpret_cast<uint4*>(smem + oQ + off) = oq;
          *reinterpret_cast<uint4*>(smem + oK + off) = ok;
        }
        ptx::tmem_ld32(lane_t + static_cast<uint32_t>(2 * kHid + h * 32), rq);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 ov;
          uint32_t* pv = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int i = 0; i < 4; ++i) pv[i] = pack2(__uint_as_float(rq[c * 8 + 2 * i]), __uint_as_float(rq[c * 8 + 2 * i + 1]));
          *reinterpret_cast<uint4*>(smem + oV + ((h * 4 + c) * 128 + row) * 16) = ov;
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars->staged);
      // ---- softmax of the row over its pixel's keys (base 2: scale and log2 e are folded into q and the bias table)
      float inv[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * hf + hh;
        uint32_t r[32];
        ptx::mbar_wait(&bars->s_full[h], par);
        ptx::tc_fence_after();
        ptx::tmem_ld32(lane_t + static_cast<uint32_t>(128 * h + 32 * qd), r);
        const uint4* brow = reinterpret_cast<const uint4*>(sbias + (h * 32 + lane) * BS);
        uint4 bv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) bv[c] = brow[c];
        ptx::tmem_ld_wait();
        float sv[32];
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const __half2* bh = reinterpret_cast<const __half2*>(&bv[c]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 bb = __half22float2(bh[i]);
            sv[c * 8 + 2 * i] = __uint_as_float(r[c * 8 + 2 * i]) + bb.x;
            sv[c * 8 + 2 * i + 1] = __uint_as_float(r[c * 8 + 2 * i + 1]) + bb.y;
            mx4[i] = fmaxf(mx4[i], fmaxf(sv[c * 8 + 2 * i], sv[c * 8 + 2 * i + 1]));
          }
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        float sp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 ov;
          uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float e0 = ex2(sv[c * 8 + 2 * i] - mx), e1 = ex2(sv[c * 8 + 2 * i + 1] - mx);
            sp[i] += e0 + e1;
            o[i] = pack2(e0, e1);
          }
          *reinterpret_cast<uint4*>(smem + oK + ((h * 4 + c) * 128 + row) * 16) = ov;   // P_h over K_h (S_h is complete)
        }
        inv[hh] = __fdividef(1.0f, (sp[0] + sp[1]) + (sp[2] + sp[3]));
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bars->p_full[h]);
      }
      // ---- O rows of the slot's block, normalised -> A operand of the output projection (over Q)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * hf + hh;
        uint32_t r[32];
        ptx::mbar_wait(&bars->pv_full[h], par);
        ptx::tc_fence_after();
        ptx::tmem_ld32(lane_t + static_cast<uint32_t>(128 * h + 32 * qd), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 ov;
          uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            o[i] = pack2(__uint_as_float(r[c * 8 + 2 * i]) * inv[hh], __uint_as_float(r[c * 8 + 2 * i + 1]) * inv[hh]);
          *reinterpret_cast<uint4*>(smem + oQ + ((h * 4 + c) * 128 + row) * 16) = ov;
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&bars->o_full);
      // ---- y = Y + x : this thread's 32 channels of its token row (the residual is fetched behind the output projection)
      uint4 xres[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) xres[c] = live ? __ldg(reinterpret_cast<const uint4*>(x + yoff) + c) : make_uint4(0u, 0u, 0u, 0u);
      {
        uint32_t r[32];
        ptx::mbar_wait(&bars->y_full, par);
        ptx::tc_fence_after();
        ptx::tmem_ld32(lane_t + 448u + static_cast<uint32_t>(hf * 32), r);
        ptx::tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const __half2* xh = reinterpret_cast<const __half2*>(&xres[c]);
            uint4 ov;
            __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 xr = __half22float2(xh[i]);
              oh[i] = h2_sat(__uint_as_float(r[c * 8 + 2 * i]) + xr.x, __uint_as_float(r[c * 8 + 2 * i + 1]) + xr.y);
            }
            *(reinterpret_cast<uint4*>(y + yoff) + c) = ov;
          }
        }
      }
      pix += step;
      pin += step;
      while (pin >= hw) {
        pin -= hw;
        ++bimg;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

}  // namespace wdno

// `scale` is NOT applied here: the caller folds scale * log2(e) into the q rows of wqkv_canon (include/wdno_b200.h)
extern "C" int wdno_tattn_block_row(const void* x, void* y, const void* wqkv_canon, const void* wout_canon, const float* bias,
                                    const float* rot_cos, const float* rot_sin, int64_t n_samples, int n_frames, int64_t hw, int C,
                                    float scale, float eps, void* stream) {
  using namespace wdno;
  if (!x || !y || !wqkv_canon || !wout_canon || n_samples < 1 || hw < 1) return set_error(WDNO_E_INVALID, "tattn_block_row: bad arguments");
  if (C != 64) return set_error(WDNO_E_INVALID, "tattn_block_row: built for C = 64 (use wdno_tattn_block otherwise)");
  if (n_frames < 1 || n_frames > 32) return set_error(WDNO_E_INVALID, "tattn_block_row: frames must be in [1,32]");
  if ((rot_cos == nullptr) != (rot_sin == nullptr)) return set_error(WDNO_E_INVALID, "tattn_block_row: rotary tables must both be given");
  if (n_samples * hw > (1LL << 30) || hw > (1LL << 30)) return set_error(WDNO_E_INVALID, "tattn_block_row: too many pixels");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tattn_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_cuda_error(e, "tattn_block_row: cudaFuncSetAttribute");
    configured = true;
  }
  const int n_pix = static_cast<int>(n_samples * hw);
  const int tiles = (n_pix + 3) / 4;
  const int cap = num_sms();
  const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
  cudaError_t le = launch_pdl(tattn_row_kernel, dim3(grid), dim3(kThreads), static_cast<size_t>(kSmem), static_cast<cudaStream_t>(stream),
                              static_cast<const __half*>(x), static_cast<__half*>(y), static_cast<const uint4*>(wqkv_canon),
                              static_cast<const uint4*>(wout_canon), bias, rot_cos, rot_sin, n_pix, static_cast<int>(hw), n_frames, scale,
                              eps);
  if (le != cudaSuccess) return set_cuda_error(le, "tattn_block_row: launch");
  return check_launch("tattn_block_row");
}
