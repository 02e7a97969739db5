// tap-GEMM: persistent, warp-specialised tcgen05 implicit-GEMM for sm_100a.
//
// One kernel serves every dense contraction of the two WDNO U-Nets (see include/wdno_b200.h).
// Design:
//   * activations are fp16 channels-last; a CTA owns a "unit" = (sample b, group of ZT output planes,
//     PT*128 consecutive padded-row positions, one N-chunk of output channels);
//   * A operand: producer warps copy a haloed slab of every needed input plane into shared memory ONCE per
//     K-set (KC channels) in the UMMA no-swizzle K-major layout [KC/8][position][8 x fp16]; every filter tap is
//     then just a descriptor whose start address is shifted by (ky*Wp+kx)*16 bytes -- no im2col traffic.
//     GroupNorm-apply + SiLU (a*x+c -> silu) is fused into that copy;
//   * B operand: weights are pre-packed per (N-chunk, K-set, tap) tile and streamed with 1-D bulk async
//     copies (TMA engine) through an mbarrier ring; one tile feeds ZT*PT accumulators;
//   * D: up to 4 fp32 accumulators of 128 x N in TMEM (double buffered when they fit in 512 columns);
//   * epilogue warps: tcgen05.ld -> +bias (+residual) -> GroupNorm partial sums -> fp16 / fp32 stores.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wdno_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace wdno {

constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 4;
constexpr int kMmaWarp = kEpiWarps;       // warp 4
constexpr int kBWarp = kEpiWarps + 1;     // warp 5
constexpr int kFirstProdWarp = kEpiWarps + 2;
constexpr int kThreads = (kEpiWarps + 2 + kProdWarps) * 32;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kMaxSlots = 12;
constexpr int kMaxBStages = 8;
constexpr int kBarBytes = 512;

struct Bars {
  uint64_t slab_full[kMaxSlots];
  uint64_t slab_empty[kMaxSlots];
  uint64_t b_full[kMaxBStages];
  uint64_t b_empty[kMaxBStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= kBarBytes, "barrier block too large");

struct UnitCoord {
  int b, zg, pt, nc;
};

__device__ __forceinline__ UnitCoord decode_unit(int u, int n_chunks, int ptiles, int zgroups) {
  UnitCoord c;
  c.nc = u % n_chunks;
  int r = u / n_chunks;
  c.pt = r % ptiles;
  r /= ptiles;
  c.zg = r % zgroups;
  c.b = r / zgroups;
  return c;
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

__global__ void __launch_bounds__(kThreads, 1) tapgemm_kernel(const wdno_tapgemm_params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem);
  uint8_t* slab_base = smem + kBarBytes;
  const int CH = p.KC >> 3;  // 16-byte chunks per position
  const uint32_t lbo_a = static_cast<uint32_t>(p.S_pad) * 16u;
  const uint32_t slot_bytes = static_cast<uint32_t>(CH) * lbo_a;
  uint8_t* b_base = slab_base + ((static_cast<size_t>(p.NSLOT) * slot_bytes + 127) & ~static_cast<size_t>(127));
  const uint32_t lbo_b = static_cast<uint32_t>(p.N) * 16u;
  const uint32_t btile_bytes = static_cast<uint32_t>(p.N) * static_cast<uint32_t>(p.KC) * 2u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int NACC = p.ZT * p.PT;
  const int NPAD = (p.N <= 64) ? 64 : 128;
  const int NBUF = (NACC * NPAD * 2 <= 512) ? 2 : 1;
  const int P = p.ZT + p.KD - 1;  // input planes per K-set
  const int positions = p.H * p.Wp;
  const int ptiles = (positions + 128 * p.PT - 1) / (128 * p.PT);
  const int zgroups = (p.D + p.ZT - 1) / p.ZT;
  const int n_units = p.B * zgroups * ptiles * p.n_chunks;

  // ---------------------------------------------------------------- setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.NSLOT; ++i) {
      ptx::mbar_init(&bars->slab_full[i], kProdWarps);
      ptx::mbar_init(&bars->slab_empty[i], 1);
    }
    for (int i = 0; i < p.NBST; ++i) {
      ptx::mbar_init(&bars->b_full[i], 1);
      ptx::mbar_init(&bars->b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->acc_full[i], 1);
      ptx::mbar_init(&bars->acc_empty[i], kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(&bars->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp >= kFirstProdWarp) {
    // ============================================================ A producers
    const int ptid = threadIdx.x - kFirstProdWarp * 32;
    const int S = 128 * p.PT + p.maxshift;  // positions needed per plane
    const int items = S * CH;
    const int ch_shift = (CH == 8) ? 3 : (CH == 4) ? 2 : 1;
    uint32_t g = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const UnitCoord uc = decode_unit(u, p.n_chunks, ptiles, zgroups);
      const wdno_nchunk ci = p.chunks[uc.nc];
      const int o0 = uc.pt * 128 * p.PT;
      const int z0 = uc.zg * p.ZT;
      for (int si = 0; si < ci.set_count; ++si) {
        const wdno_kset st = p.sets[ci.set_begin + si];
        const __half* src = static_cast<const __half*>(p.src[st.src]);
        const int csrc = p.src_c[st.src];
        const float* ca = p.coef_a[st.src];
        const float* cc = p.coef_c[st.src];
        for (int j = 0; j < P; ++j, ++g) {
          const int slot = g % p.NSLOT;
          const uint32_t par = (g / p.NSLOT) & 1u;
          ptx::mbar_wait(&bars->slab_empty[slot], par ^ 1u);
          uint8_t* dst = slab_base + static_cast<size_t>(slot) * slot_bytes;
          const int zi = z0 - p.pz + j;
          const bool zok = (zi >= 0) && (zi < p.D);
          // source plane geometry
          int Hs = p.H, Ws = p.W;
          if (p.src_mode == 1) { Hs = 2 * p.H; Ws = 2 * p.W; }
          if (p.src_mode == 2) { Hs = p.H >> 1; Ws = p.W >> 1; }
          const size_t plane_off = (static_cast<size_t>(uc.b) * p.D + (zok ? zi : 0)) * Hs * Ws;
          for (int it0 = ptid; it0 < items; it0 += kProdThreads * 4) {
            uint4 v[4];
            int dsto[4];
            int chn[4];
            bool ok[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int it = it0 + k * kProdThreads;
              ok[k] = false;
              dsto[k] = -1;
              chn[k] = 0;
              v[k] = make_uint4(0u, 0u, 0u, 0u);
              if (it < items) {
                const int s = it >> ch_shift;
                const int c = it & (CH - 1);
                dsto[k] = (c * p.S_pad + s) * 16;
                const int q = o0 + s;
                const int yp = q / p.Wp;
                const int xp = q - yp * p.Wp;
                const int y = yp - p.py;
                const int x = xp - p.px;
                if (zok && y >= 0 && y < p.H && x >= 0 && x < p.W) {
                  int ys = y, xs = x;
                  if (p.src_mode == 1) { ys = 2 * y + st.ph_y; xs = 2 * x + st.ph_x; }
                  if (p.src_mode == 2) { ys = y >> 1; xs = x >> 1; }
                  chn[k] = st.ch_off + c * 8;
                  const __half* ptr = src + (plane_off + static_cast<size_t>(ys) * Ws + xs) * csrc + chn[k];
                  v[k] = __ldg(reinterpret_cast<const uint4*>(ptr));
                  ok[k] = true;
                }
              }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (dsto[k] < 0) continue;
              if (ca != nullptr && ok[k]) {
                const float* pa = ca + static_cast<size_t>(uc.b) * csrc + chn[k];
                const float* pc = cc + static_cast<size_t>(uc.b) * csrc + chn[k];
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(pa));
                const float4 a1 = __ldg(reinterpret_cast<const float4*>(pa + 4));
                const float4 c0 = __ldg(reinterpret_cast<const float4*>(pc));
                const float4 c1 = __ldg(reinterpret_cast<const float4*>(pc + 4));
                __half2* h = reinterpret_cast<__half2*>(&v[k]);
                float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
                float2 f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
                f0.x = silu_f(fmaf(a0.x, f0.x, c0.x)); f0.y = silu_f(fmaf(a0.y, f0.y, c0.y));
                f1.x = silu_f(fmaf(a0.z, f1.x, c0.z)); f1.y = silu_f(fmaf(a0.w, f1.y, c0.w));
                f2.x = silu_f(fmaf(a1.x, f2.x, c1.x)); f2.y = silu_f(fmaf(a1.y, f2.y, c1.y));
                f3.x = silu_f(fmaf(a1.z, f3.x, c1.z)); f3.y = silu_f(fmaf(a1.w, f3.y, c1.w));
                h[0] = __float22half2_rn(f0); h[1] = __float22half2_rn(f1);
                h[2] = __float22half2_rn(f2); h[3] = __float22half2_rn(f3);
              }
              *reinterpret_cast<uint4*>(dst + dsto[k]) = v[k];
            }
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bars->slab_full[slot]);
        }
      }
    }
  } else if (warp == kBWarp) {
    // ============================================================ B (weight tile) producer
    // warp-uniform control flow, one elected lane issues the bulk copies
    uint32_t bst = 0, bph = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int nc = u % p.n_chunks;
      const wdno_nchunk ci = p.chunks[nc];
      const uint8_t* wsrc = static_cast<const uint8_t*>(p.wpacked) + static_cast<size_t>(ci.w_tile_off) * btile_bytes;
      for (int i = 0; i < ci.n_tiles; ++i) {
        ptx::mbar_wait(&bars->b_empty[bst], bph ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&bars->b_full[bst], btile_bytes);
          ptx::bulk_g2s(b_base + static_cast<size_t>(bst) * btile_bytes, wsrc + static_cast<size_t>(i) * btile_bytes,
                        btile_bytes, &bars->b_full[bst]);
        }
        __syncwarp();
        if (++bst == static_cast<uint32_t>(p.NBST)) { bst = 0; bph ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================================================ MMA issuer
    // The whole warp runs the warp-uniform control flow (so descriptors stay in uniform registers and no
    // per-operand waterfall loops are generated); one elected lane issues tcgen05.mma / tcgen05.commit.
    // Ring positions are tracked incrementally -- no integer division on this path.
    const uint32_t idesc = ptx::make_idesc_f16(p.N, 0);
    const uint32_t desc_hi = 8u | (1u << 14);                        // SBO = 128 B, descriptor version 1
    const uint32_t a_lo0 = (ptx::smem_u32(slab_base) >> 4) + (static_cast<uint32_t>(p.S_pad) << 16);  // LBO = S_pad*16
    const uint32_t b_lo0 = (ptx::smem_u32(b_base) >> 4) + (static_cast<uint32_t>(p.N) << 16);         // LBO = N*16
    const uint32_t slot_u = slot_bytes >> 4, btile_u = btile_bytes >> 4;
    const uint32_t a_kstep = 2u * static_cast<uint32_t>(p.S_pad), b_kstep = 2u * static_cast<uint32_t>(p.N);
    const int ksteps = p.KC >> 4;
    const uint32_t nslot = static_cast<uint32_t>(p.NSLOT);
    uint32_t s0 = 0, sph = 0;   // ring slot / phase of plane 0 of the current K-set
    uint32_t bst = 0, bph = 0;  // weight stage / phase
    uint32_t ucnt = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ucnt) {
      const int nc = u % p.n_chunks;
      const wdno_nchunk ci = p.chunks[nc];
      const uint32_t buf = (NBUF == 2) ? (ucnt & 1u) : 0u;
      const uint32_t aph = (NBUF == 2) ? ((ucnt >> 1) & 1u) : (ucnt & 1u);
      ptx::mbar_wait(&bars->acc_empty[buf], aph ^ 1u);
      ptx::tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * static_cast<uint32_t>(NACC * NPAD);
      uint32_t accum = 0u;
      for (int si = 0; si < ci.set_count; ++si) {
        const wdno_kset st = p.sets[ci.set_begin + si];
        for (int j = 0; j < p.ZT; ++j) {
          uint32_t s = s0 + j, ph = sph;
          if (s >= nslot) { s -= nslot; ph ^= 1u; }
          ptx::mbar_wait(&bars->slab_full[s], ph);
        }
        int cur_kz = 0;
        const wdno_tap* tp_ptr = p.taps + st.tap_begin;
        for (int t = 0; t < st.tap_count; ++t) {
          const wdno_tap tp = tp_ptr[t];
          while (cur_kz < tp.kz) {
            // plane cur_kz is dead: release it; plane cur_kz + ZT becomes needed
            uint32_t sd = s0 + cur_kz;
            if (sd >= nslot) sd -= nslot;
            if (ptx::elect_one()) ptx::tc_commit(&bars->slab_empty[sd]);
            __syncwarp();
            uint32_t sn = s0 + cur_kz + p.ZT, ph = sph;
            if (sn >= nslot) { sn -= nslot; ph ^= 1u; }
            ptx::mbar_wait(&bars->slab_full[sn], ph);
            ++cur_kz;
          }
          ptx::mbar_wait(&bars->b_full[bst], bph);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t b_lo = b_lo0 + bst * btile_u;
            for (int za = 0; za < p.ZT; ++za) {
              uint32_t sa = s0 + tp.kz + za;
              if (sa >= nslot) sa -= nslot;
              const uint32_t a_lo = a_lo0 + sa * slot_u + static_cast<uint32_t>(tp.shift);
              for (int pi = 0; pi < p.PT; ++pi) {
                const uint32_t dcol = acc0 + static_cast<uint32_t>((za * p.PT + pi) * NPAD);
                uint32_t al = a_lo + static_cast<uint32_t>(pi * 128), bl = b_lo;
                uint32_t acc_flag = accum;
                for (int k = 0; k < ksteps; ++k) {
                  const uint64_t ad = (static_cast<uint64_t>(desc_hi) << 32) | al;
                  const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | bl;
                  ptx::tc_mma_f16(dcol, ad, bd, idesc, acc_flag);
                  acc_flag = 1u;
                  al += a_kstep;
                  bl += b_kstep;
                }
              }
            }
            ptx::tc_commit(&bars->b_empty[bst]);
          }
          __syncwarp();
          accum = 1u;
          if (++bst == static_cast<uint32_t>(p.NBST)) { bst = 0; bph ^= 1u; }
        }
        if (ptx::elect_one()) {
          for (int j = cur_kz; j < P; ++j) {
            uint32_t sd = s0 + j;
            if (sd >= nslot) sd -= nslot;
            ptx::tc_commit(&bars->slab_empty[sd]);
          }
        }
        __syncwarp();
        s0 += static_cast<uint32_t>(P);
        if (s0 >= nslot) { s0 -= nslot; sph ^= 1u; }
      }
      if (ptx::elect_one()) ptx::tc_commit(&bars->acc_full[buf]);
      __syncwarp();
    }
  } else {
    // ============================================================ epilogue warps 0..3
    const int row = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int nblk = p.N >> 3;  // 8-column blocks, <= 16
    uint32_t ucnt = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ucnt) {
      const UnitCoord uc = decode_unit(u, p.n_chunks, ptiles, zgroups);
      const wdno_nchunk ci = p.chunks[uc.nc];
      const uint32_t buf = (NBUF == 2) ? (ucnt & 1u) : 0u;
      const uint32_t aph = (NBUF == 2) ? ((ucnt >> 1) & 1u) : (ucnt & 1u);
      const int o0 = uc.pt * 128 * p.PT;
      const int z0 = uc.zg * p.ZT;
      float s1[16], s2[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
      ptx::mbar_wait(&bars->acc_full[buf], aph);
      ptx::tc_fence_after();
      for (int a = 0; a < NACC; ++a) {
        const int za = a / p.PT, pi = a - za * p.PT;
        const int z = z0 + za;
        const int o = o0 + pi * 128 + row;
        const int y = o / p.Wp;
        const int x = o - y * p.Wp;
        const bool valid = (z < p.D) && (y < p.H) && (x < p.W);
        const uint32_t tcol = tmem_base + lane_base + static_cast<uint32_t>(buf * NACC * NPAD + a * NPAD);
        size_t obase = 0;
        if (p.out_mode == 0) {
          obase = (((static_cast<size_t>(uc.b) * p.D + z) * p.H + y) * p.W + x) * p.out_c + ci.out_ch_off;
        } else if (p.out_mode == 1) {
          obase = (((static_cast<size_t>(uc.b) * p.D + z) * (2 * p.H) + (2 * y + ci.ph_y)) * (2 * p.W) + (2 * x + ci.ph_x)) *
                      p.out_c + ci.out_ch_off;
        } else {
          obase = ((static_cast<size_t>(uc.b) * p.D + z) * p.out_c + ci.out_ch_off) * (static_cast<size_t>(p.H) * p.W) +
                  static_cast<size_t>(y) * p.W + x;
        }
        const int n16 = p.N >> 4;
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16) {
          if (c16 >= n16) break;
          uint32_t r[16];
          ptx::tmem_ld16(tcol + static_cast<uint32_t>(c16 * 16), r);
          ptx::tmem_ld_wait();
          if (!valid) continue;
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(r[i]);
          const int ncol0 = c16 * 16;
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (ncol0 + i < ci.n_valid) f[i] += __ldg(p.bias + ci.out_ch_off + ncol0 + i);
          }
          if (p.stats != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (ncol0 + i < ci.n_valid) {
                s1[c16 * 2 + (i >> 3)] += f[i];
                s2[c16 * 2 + (i >> 3)] += f[i] * f[i];
              }
            }
          }
          if (p.out_mode == 2) {
            float* o32 = static_cast<float*>(p.out);
            const size_t cs = static_cast<size_t>(p.H) * p.W;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (ncol0 + i < ci.n_valid) o32[obase + (ncol0 + i) * cs] = f[i];
          } else {
            __half* o16 = static_cast<__half*>(p.out);
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              if (ncol0 + hb * 8 >= ci.n_valid) continue;
              const size_t off = obase + ncol0 + hb * 8;
              if (p.resid != nullptr) {
                const uint4 rv = __ldg(reinterpret_cast<const uint4*>(static_cast<const __half*>(p.resid) + off));
                const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 rf = __half22float2(rh[i]);
                  f[hb * 8 + 2 * i] += rf.x;
                  f[hb * 8 + 2 * i + 1] += rf.y;
                }
              }
              uint4 ov;
              __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
              for (int i = 0; i < 4; ++i) oh[i] = __floats2half2_rn(f[hb * 8 + 2 * i], f[hb * 8 + 2 * i + 1]);
              *reinterpret_cast<uint4*>(o16 + off) = ov;
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars->acc_empty[buf]);
      if (p.stats != nullptr) {
        // 8-column blocks -> groups (cpg is a multiple of 8); warp-reduce, then one atomic pair per group
        float g1 = 0.f, g2 = 0.f;
        int cur_g = -1;
#pragma unroll
        for (int blk = 0; blk < 16; ++blk) {
          if (blk >= nblk || blk * 8 >= ci.n_valid) break;
          const int gi = (ci.out_ch_off + blk * 8) / p.cpg;
          if (gi != cur_g) {
            if (cur_g >= 0 && lane == 0) {
              double* sp = p.stats + (static_cast<size_t>(uc.b) * p.G + cur_g) * 2;
              atomicAdd(sp, static_cast<double>(g1));
              atomicAdd(sp + 1, static_cast<double>(g2));
            }
            g1 = 0.f;
            g2 = 0.f;
            cur_g = gi;
          }
          float a1 = s1[blk], a2 = s2[blk];
#pragma unroll
          for (int sh = 16; sh > 0; sh >>= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, sh);
            a2 += __shfl_xor_sync(0xffffffffu, a2, sh);
          }
          g1 += a1;
          g2 += a2;
        }
        if (cur_g >= 0 && lane == 0) {
          double* sp = p.stats + (static_cast<size_t>(uc.b) * p.G + cur_g) * 2;
          atomicAdd(sp, static_cast<double>(g1));
          atomicAdd(sp + 1, static_cast<double>(g2));
        }
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

static int64_t smem_bytes_of(const wdno_tapgemm_params* p) {
  const int64_t CH = p->KC / 8;
  const int64_t slot = CH * p->S_pad * 16;
  const int64_t slabs = (p->NSLOT * slot + 127) & ~static_cast<int64_t>(127);
  const int64_t bt = static_cast<int64_t>(p->N) * p->KC * 2;
  return kBarBytes + slabs + p->NBST * bt;
}

static int validate(const wdno_tapgemm_params* p) {
  if (!p) return set_error(WDNO_E_INVALID, "tapgemm: null params");
  if (p->KC != 16 && p->KC != 32 && p->KC != 64) return set_error(WDNO_E_INVALID, "tapgemm: KC must be 16/32/64");
  if (p->N < 16 || p->N > 128 || (p->N % 16)) return set_error(WDNO_E_INVALID, "tapgemm: N must be a multiple of 16 in [16,128]");
  if (p->ZT < 1 || p->PT < 1 || p->ZT * p->PT > 4) return set_error(WDNO_E_INVALID, "tapgemm: ZT*PT must be in [1,4]");
  if ((p->N > 64) && p->ZT * p->PT * 128 > 512) return set_error(WDNO_E_INVALID, "tapgemm: accumulators exceed TMEM");
  const int P = p->ZT + p->KD - 1;
  if (p->NSLOT < P || p->NSLOT > kMaxSlots) return set_error(WDNO_E_INVALID, "tapgemm: NSLOT must be in [ZT+KD-1, 12]");
  if (p->NBST < 2 || p->NBST > kMaxBStages) return set_error(WDNO_E_INVALID, "tapgemm: NBST must be in [2,8]");
  if (p->S_pad < 128 * p->PT + p->maxshift) return set_error(WDNO_E_INVALID, "tapgemm: S_pad smaller than slab");
  if (p->S_pad >= 16384) return set_error(WDNO_E_INVALID, "tapgemm: S_pad too large for descriptor");
  if (p->B < 1 || p->D < 1 || p->H < 1 || p->W < 1 || p->n_chunks < 1) return set_error(WDNO_E_INVALID, "tapgemm: empty problem");
  if (p->src_c[0] % 8 || (p->src[1] && p->src_c[1] % 8)) return set_error(WDNO_E_INVALID, "tapgemm: source channels must be multiples of 8");
  if (p->out_mode < 0 || p->out_mode > 2) return set_error(WDNO_E_INVALID, "tapgemm: bad out_mode");
  if (p->out_mode != 2 && (p->out_c % 8)) return set_error(WDNO_E_INVALID, "tapgemm: fp16 output channels must be a multiple of 8");
  if (p->stats && (p->cpg % 8 || p->cpg < 8)) return set_error(WDNO_E_INVALID, "tapgemm: cpg must be a multiple of 8");
  if (p->src_mode == 2 && ((p->H & 1) || (p->W & 1))) return set_error(WDNO_E_INVALID, "tapgemm: up2 needs even H,W");
  if (smem_bytes_of(p) > 227 * 1024) return set_error(WDNO_E_INVALID, "tapgemm: shared-memory plan exceeds 227 KB");
  if (p->grid < 1) return set_error(WDNO_E_INVALID, "tapgemm: grid must be >= 1");
  return WDNO_OK;
}

}  // namespace wdno

extern "C" int64_t wdno_tapgemm_smem_bytes(const wdno_tapgemm_params* p) {
  if (!p) return WDNO_E_INVALID;
  return wdno::smem_bytes_of(p);
}

extern "C" int wdno_tapgemm(const wdno_tapgemm_params* p, void* stream) {
  int rc = wdno::validate(p);
  if (rc != WDNO_OK) return rc;
  const int64_t smem = wdno::smem_bytes_of(p);
  static int64_t configured = -1;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(wdno::tapgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return wdno::set_cuda_error(e, "tapgemm: cudaFuncSetAttribute");
    configured = 227 * 1024;
  }
  wdno::tapgemm_kernel<<<p->grid, wdno::kThreads, smem, static_cast<cudaStream_t>(stream)>>>(*p);
  return wdno::check_launch("tapgemm");
}
