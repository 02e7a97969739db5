// tap-GEMM: persistent, warp-specialised tcgen05 implicit-GEMM for sm_100a.
//
// One kernel serves every dense contraction of the two WDNO U-Nets (see include/wdno_b200.h).
// Design ("tap-GEMM"):
//   * activations are fp16 channels-last; a CTA owns work items = (sample b, group of ZT output planes,
//     PT*128 consecutive padded-row positions) x N-chunks of output channels;
//   * A operand: producer warps copy a haloed slab of every needed input plane into shared memory ONCE per
//     K-set (KC channels) in the UMMA no-swizzle K-major layout [KC/8][position][8 x fp16]; every filter tap is
//     then just a descriptor whose start address is shifted by (ky*Wp+kx)*16 bytes -- no im2col traffic.
//     GroupNorm-apply + SiLU (a*x+c -> silu) is fused into that copy.  For 1x1 layers whose K-sets all fit in the
//     slab ring the slabs are loaded once and reused by every N-chunk (p.reuse);
//   * B operand: weights are pre-packed per (N-chunk, K-set, tap) tile and streamed TPS tiles at a time with 1-D
//     bulk async copies (TMA engine) through an mbarrier ring; one tile feeds ZT*PT accumulators;
//   * D: up to 4 fp32 accumulators of 128 x N in TMEM (double buffered when they fit in 512 columns);
//   * epilogue warps: tcgen05.ld -> +bias (+residual) -> GroupNorm partial sums -> fp16 / fp32 stores.
#include <cuda_fp16.h>
#include "cvt_sat.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wdno_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace wdno {

constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 8;
constexpr int kMmaWarp = kEpiWarps;       // warp 4
constexpr int kBWarp = kEpiWarps + 1;     // warp 5
constexpr int kFirstProdWarp = kEpiWarps + 2;
constexpr int kLoadWarps = 2;              // act mode only: dedicated cp.async issuers feeding the transforming producers
constexpr int kFirstLoadWarp = kFirstProdWarp + kProdWarps;
constexpr int kThreads = (kEpiWarps + 2 + kProdWarps + kLoadWarps) * 32;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kMaxSlots = 12;
constexpr int kMaxBStages = 4;
constexpr int kBarBytes = 512;
constexpr int kMaxTaps = 384;             // tap table copied to shared memory (7x7x7 = 343)
constexpr int kStageRow = 144;             // epilogue staging: 128 B of fp16 columns + 16 B pad (row's global offset)
constexpr int kStageWarp = 32 * kStageRow;
constexpr int kStageBytes = kEpiWarps * kStageWarp;
constexpr int kMaxBias = 1024;             // bias of the whole layer, staged in shared memory
constexpr int kMaxChunks = 32;             // N-chunk descriptors staged in shared memory
constexpr int kMaxSets = 128;              // K-set descriptors staged in shared memory
constexpr int kHdrBytes = kBarBytes + kMaxTaps * 8 + kStageBytes + kMaxBias * 4 + kMaxChunks * 40 + kMaxSets * 24;
static_assert(sizeof(wdno_kset) == 24, "wdno_kset layout");
static_assert(sizeof(wdno_nchunk) == 40, "wdno_nchunk layout");
constexpr int kLoadBatch = 8;             // independent 16-byte loads in flight per producer thread

#ifdef WDNO_PROF
__device__ unsigned long long g_prof[32];
// per-role stall accounting (debug builds only): g_prof[i] += cycles lane 0 of the role's first warp spent in region i
#define PROF_DECL long long _pacc[6] = {0, 0, 0, 0, 0, 0}; const long long _pstart = clock64()
#define PROF_REGION(i, ...) do { const long long _t = clock64(); __VA_ARGS__; _pacc[i] += clock64() - _t; } while (0)
#define PROF_COMMIT(base, cond) do { if (cond) { for (int _i = 0; _i < 6; ++_i) atomicAdd(&g_prof[(base) + _i], static_cast<unsigned long long>(_pacc[_i])); \
    atomicAdd(&g_prof[(base) + 6], static_cast<unsigned long long>(clock64() - _pstart)); } } while (0)
#else
#define PROF_DECL do {} while (0)
#define PROF_REGION(i, ...) do { __VA_ARGS__; } while (0)
#define PROF_COMMIT(base, cond) do {} while (0)
#endif

struct Bars {
  uint64_t slab_full[kMaxSlots];
  uint64_t slab_empty[kMaxSlots];
  uint64_t raw_full[kMaxSlots];
  uint64_t b_full[kMaxBStages];
  uint64_t b_empty[kMaxBStages];
  uint64_t b_peer[kMaxBStages];   // cluster mode, leader CTA: the peer's stage is free and its mbarrier armed
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= kBarBytes, "barrier block too large");

// work item w -> (b, zg, pt) and the N-chunk range it covers
struct Work {
  int b, zg, pt, nc0, nc1;
  int xs0;  // first tensor column of the work item's column strip (0 unless p.strips > 1)
};

__device__ __forceinline__ Work decode_work(int w, int n_chunks, int ptiles, int zgroups, int reuse, int strips = 1,
                                            int strip_w = 0) {
  Work k;
  int r = w;
  if (reuse) {
    k.nc0 = 0;
    k.nc1 = n_chunks;
  } else {
    k.nc0 = w % n_chunks;
    k.nc1 = k.nc0 + 1;
    r = w / n_chunks;
  }
  k.pt = r % ptiles;
  r /= ptiles;
  k.zg = r % zgroups;
  r /= zgroups;
  k.b = r;
  k.xs0 = 0;
  if (strips > 1) {
    k.b = r / strips;
    k.xs0 = (r - k.b * strips) * strip_w;
  }
  return k;
}

// i-th work item of this CTA, or -1.  Default: blockIdx.x + i * gridDim.x.  Cluster mode (p.cluster == 2, CTA pairs that
// share every weight tile through one multicast copy): the pair walks pair-items q = pair + i * n_pairs; both CTAs take the
// same N-chunk (q % n_chunks) of two neighbouring position / plane units, so they consume the weight stream in lock step.
__device__ __forceinline__ int work_at(const wdno_tapgemm_params& p, int i, int n_work) {
  if (p.cluster != 2) {
    const long long w = static_cast<long long>(blockIdx.x) + static_cast<long long>(i) * gridDim.x;
    return w < n_work ? static_cast<int>(w) : -1;
  }
  const long long q = static_cast<long long>(blockIdx.x >> 1) + static_cast<long long>(i) * (gridDim.x >> 1);
  if (2 * q >= n_work) return -1;
  const int chunk = static_cast<int>(q % p.n_chunks);
  const int rest = static_cast<int>(q / p.n_chunks) * 2 + static_cast<int>(blockIdx.x & 1u);
  return chunk + p.n_chunks * rest;
}

// silu(v) = v * sigmoid(v) = 0.5 v (1 + tanh(v/2)): one MUFU op (tanh.approx.f32, rel. error 2^-11, i.e. the fp16 rounding
// the value receives anyway) instead of ex2 + rcp + Newton steps -- the fused prologue recomputes halo positions, so its
// transcendental count is ~2.5x that of a stand-alone pass and must stay off the critical path
__device__ __forceinline__ float silu_f(float v) {
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// h = v / 2 -> silu(v) = h + h * tanh(h)
__device__ __forceinline__ float silu_half(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// ------------------------------------------------------------------ MMA issue
// All MMAs of one weight tile: ZT*PT accumulators x KS k-steps, fully unrolled, operands in registers.
template <int ZT, int PT, int KS>
__device__ __forceinline__ void issue_tile(uint32_t a_tap_lo, uint32_t s_kz, uint32_t nslot, uint32_t slot_u, uint32_t b_lo,
                                           uint32_t acc0, uint32_t npad, uint32_t a_kstep, uint32_t b_kstep, uint32_t idesc,
                                           uint32_t accum) {
  constexpr uint64_t kDescHi = static_cast<uint64_t>(8u | (1u << 14)) << 32;  // SBO = 128 B, descriptor version 1
#pragma unroll
  for (int za = 0; za < ZT; ++za) {
    uint32_t sa = s_kz + za;
    if (sa >= nslot) sa -= nslot;
    const uint32_t a_lo = a_tap_lo + sa * slot_u;
#pragma unroll
    for (int pi = 0; pi < PT; ++pi) {
      const uint32_t dcol = acc0 + static_cast<uint32_t>(za * PT + pi) * npad;
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const uint64_t ad = kDescHi | (a_lo + static_cast<uint32_t>(pi * 128) + static_cast<uint32_t>(k) * a_kstep);
        const uint64_t bd = kDescHi | (b_lo + static_cast<uint32_t>(k) * b_kstep);
        ptx::tc_mma_f16(dcol, ad, bd, idesc, (k == 0) ? accum : 1u);
      }
    }
  }
}

// The whole warp runs the warp-uniform control flow (waits, ring bookkeeping); one elected lane issues the MMAs of a
// whole weight stage (TPS taps of one kz group) and the tcgen05.commit that frees it.  No divisions on this path.
template <int ZT, int PT, int KS>
__device__ __forceinline__ void mma_role(const wdno_tapgemm_params& p, Bars* bars, const wdno_tap* s_taps,
                                         const wdno_nchunk* s_chunks, const wdno_kset* s_sets, uint32_t tmem_base,
                                         const uint8_t* slab_base, const uint8_t* b_base, int n_work, int ptiles, int zgroups) {
  constexpr int NACC = ZT * PT;
  const uint32_t N = static_cast<uint32_t>(p.N);
  const uint32_t npad = (N <= 64) ? 64u : 128u;
  const bool two_buf = NACC * npad * 2 <= 512;
  const int KD = p.KD, TPS = p.TPS, n_chunks = p.n_chunks, reuse = p.reuse;
  const int P = ZT + KD - 1;
  const uint32_t nslot = static_cast<uint32_t>(p.NSLOT), nbst = static_cast<uint32_t>(p.NBST);
  const uint32_t idesc = ptx::make_idesc_f16(p.N, 0);
  const uint32_t S_pad = static_cast<uint32_t>(p.S_pad);
  const uint32_t slot_u = (static_cast<uint32_t>(p.KC >> 3) * S_pad * 16u) >> 4;
  const uint32_t btile_u = (N * static_cast<uint32_t>(p.KC) * 2u) >> 4, bstage_u = btile_u * static_cast<uint32_t>(TPS);
  // (in a cluster launch the numeric shared address carries the CTA rank above bit 18: descriptors take the CTA-relative offset)
  const uint32_t a_lo0 = ((ptx::smem_u32(slab_base) & 0x3FFFFu) >> 4) + (S_pad << 16);  // start | LBO = S_pad*16 B
  const uint32_t b_lo0 = ((ptx::smem_u32(b_base) & 0x3FFFFu) >> 4) + (N << 16);         // start | LBO = N*16 B
  const uint32_t a_kstep = 2u * S_pad, b_kstep = 2u * N;
  uint32_t s0 = 0, sph = 0;   // ring slot / phase of plane 0 of the current K-set
  uint32_t bst = 0, bph = 0;  // weight stage / phase
  uint32_t acnt = 0;          // accumulator-buffer use counter
  PROF_DECL;
  for (int wi = 0, w; (w = work_at(p, wi, n_work)) >= 0; ++wi) {
    const int nc0 = reuse ? 0 : (w % n_chunks);
    const int nc1 = reuse ? n_chunks : nc0 + 1;
    const uint32_t su = s0, suph = sph;  // ring position at the start of this work item
    for (int nc = nc0; nc < nc1; ++nc, ++acnt) {
      const wdno_nchunk ci = s_chunks[nc];
      const bool first_pass = (nc == nc0), last_pass = (nc == nc1 - 1);
      if (reuse) { s0 = su; sph = suph; }
      const uint32_t buf = two_buf ? (acnt & 1u) : 0u;
      const uint32_t aph = two_buf ? ((acnt >> 1) & 1u) : (acnt & 1u);
      PROF_REGION(0, ptx::mbar_wait(&bars->acc_empty[buf], aph ^ 1u));
      ptx::tc_fence_after();
      const uint32_t acc0 = tmem_base + buf * static_cast<uint32_t>(NACC) * npad;
      uint32_t accum = 0u;
      for (int si = 0; si < ci.set_count; ++si) {
        const wdno_kset st = s_sets[ci.set_begin + si];
        const int groups_per_kz = (st.tap_count / KD) / TPS;
        if (first_pass) {
#pragma unroll
          for (int j = 0; j < ZT; ++j) {
            uint32_t s = s0 + j, ph = sph;
            if (s >= nslot) { s -= nslot; ph ^= 1u; }
            PROF_REGION(1, ptx::mbar_wait(&bars->slab_full[s], ph));
          }
          ptx::fence_proxy_async_smem();  // cp.async (generic proxy) slab writes -> tcgen05 (async proxy) reads
        }
        const wdno_tap* tp_ptr = s_taps + st.tap_begin;
        for (int kz = 0; kz < KD; ++kz) {
          uint32_t s_kz = s0 + kz;
          if (s_kz >= nslot) s_kz -= nslot;
          if (kz > 0) {
            // plane kz-1 is dead: release it; plane kz+ZT-1 becomes needed (KD > 1 never uses slab reuse)
            uint32_t sd = s0 + kz - 1;
            if (sd >= nslot) sd -= nslot;
            if (ptx::elect_one()) ptx::tc_commit(&bars->slab_empty[sd]);
            __syncwarp();
            uint32_t sn = s0 + kz + ZT - 1, ph = sph;
            if (sn >= nslot) { sn -= nslot; ph ^= 1u; }
            PROF_REGION(1, ptx::mbar_wait(&bars->slab_full[sn], ph));
            ptx::fence_proxy_async_smem();
          }
          for (int g = 0; g < groups_per_kz; ++g) {
            PROF_REGION(2, ptx::mbar_wait(&bars->b_full[bst], bph));
            ptx::tc_fence_after();
            PROF_REGION(3, if (ptx::elect_one()) {
              uint32_t b_lo = b_lo0 + bst * bstage_u;
              uint32_t af = accum;
#pragma unroll 1
              for (int i = 0; i < TPS; ++i) {
                const uint32_t shift = static_cast<uint32_t>(tp_ptr[i].shift);
                issue_tile<ZT, PT, KS>(a_lo0 + shift, s_kz, nslot, slot_u, b_lo, acc0, npad, a_kstep, b_kstep, idesc, af);
                af = 1u;
                b_lo += btile_u;
              }
              ptx::tc_commit(&bars->b_empty[bst]);
            });
            __syncwarp();
            accum = 1u;
            tp_ptr += TPS;
            if (++bst == nbst) { bst = 0; bph ^= 1u; }
          }
        }
        if (last_pass) {
          if (ptx::elect_one()) {
            for (int j = KD - 1; j < P; ++j) {
              uint32_t sd = s0 + j;
              if (sd >= nslot) sd -= nslot;
              ptx::tc_commit(&bars->slab_empty[sd]);
            }
          }
          __syncwarp();
        }
        s0 += static_cast<uint32_t>(P);
        if (s0 >= nslot) { s0 -= nslot; sph ^= 1u; }
      }
      if (ptx::elect_one()) ptx::tc_commit(&bars->acc_full[buf]);
      __syncwarp();
    }
  }
  PROF_COMMIT(0, (threadIdx.x & 31) == 0);
}

// ------------------------------------------------------------------ kz-stacked MMA issue (p.zstack)
// For 64-wide N-chunks of a KD x KH x KW convolution an N = 64 MMA is shared-memory-bandwidth bound (A and B are both
// re-read per MMA: 48 cycles instead of 32).  Here the weight tile of one in-plane tap holds all KD depth taps stacked
// along N ([kz][64] rows), and the ZT = 4 accumulators sit in TMEM in DESCENDING plane order, so that for input plane j
// ONE MMA with N = 64 * (number of valid kz) updates every output plane that plane contributes to (o = j - kz):
// A is read once for up to 4 taps.  All P = ZT + KD - 1 planes of a K-set stay resident; they are acquired plane by
// plane during the first in-plane tap and released plane by plane during the last one.
template <int KS, int KD>
__device__ __forceinline__ void mma_role_zstack(const wdno_tapgemm_params& p, Bars* bars, const wdno_tap* s_taps,
                                                const wdno_nchunk* s_chunks, const wdno_kset* s_sets, uint32_t tmem_base,
                                                const uint8_t* slab_base, const uint8_t* b_base, int n_work) {
  constexpr int ZT = 4;
  constexpr int P = ZT + KD - 1;
  constexpr uint64_t kDescHi = static_cast<uint64_t>(8u | (1u << 14)) << 32;  // SBO = 128 B, descriptor version 1
  constexpr uint32_t rows = 64u * KD;  // rows of one stacked weight tile
  const int TPS = p.TPS, n_chunks = p.n_chunks;
  const uint32_t nslot = static_cast<uint32_t>(p.NSLOT), nbst = static_cast<uint32_t>(p.NBST);
  const uint32_t S_pad = static_cast<uint32_t>(p.S_pad);
  const uint32_t slot_u = (static_cast<uint32_t>(p.KC >> 3) * S_pad * 16u) >> 4;
  const uint32_t btile_u = (rows * static_cast<uint32_t>(p.KC) * 2u) >> 4, bstage_u = btile_u * static_cast<uint32_t>(TPS);
  // (in a cluster launch the numeric shared address carries the CTA rank above bit 18: descriptors take the CTA-relative offset)
  const uint32_t a_lo0 = ((ptx::smem_u32(slab_base) & 0x3FFFFu) >> 4) + (S_pad << 16);  // start | LBO = S_pad*16 B
  const uint32_t b_lo0 = ((ptx::smem_u32(b_base) & 0x3FFFFu) >> 4) + (rows << 16);      // start | LBO = rows*16 B
  const uint32_t a_kstep = 2u * S_pad;
  constexpr uint32_t b_kstep = 2u * rows;
  const uint32_t id1 = ptx::make_idesc_f16(64, 0), id2 = ptx::make_idesc_f16(128, 0), id3 = ptx::make_idesc_f16(192, 0),
                 id4 = ptx::make_idesc_f16(256, 0);
  uint32_t s0 = 0, sph = 0, bst = 0, bph = 0, acnt = 0;
  PROF_DECL;
  // all MMAs of in-plane tap (a_tap, b_tap) for planes [j0, j1): compile-time kz ranges, one MMA per (plane, k-step)
  auto issue = [&](uint32_t a_tap, uint32_t b_tap, uint32_t acc0, int j0, int j1, bool init, bool release) {
    uint32_t sj = s0;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      constexpr int dummy = 0;
      (void)dummy;
      const int kz_lo = (j - (ZT - 1) > 0) ? j - (ZT - 1) : 0;
      const int kz_hi = (j < KD - 1) ? j : KD - 1;
      const int nkz = kz_hi - kz_lo + 1;
      const int o_hi = j - kz_lo;
      if (j >= j0 && j < j1) {
        const uint32_t a_lo = a_tap + sj * slot_u;
        const uint32_t b_lo = b_tap + static_cast<uint32_t>(kz_lo) * 64u;
        const uint32_t dcol = acc0 + static_cast<uint32_t>(ZT - 1 - o_hi) * 64u;
        const uint32_t idn = (nkz == 1) ? id1 : (nkz == 2) ? id2 : (nkz == 3) ? id3 : id4;
        if (init) {
          // very first k-step of the work item: one N = 64 MMA per (plane, kz); output plane o is first touched by
          // plane j = o through kz = 0, which therefore clears the accumulator
#pragma unroll
          for (int kz = kz_lo; kz <= kz_hi; ++kz)
            ptx::tc_mma_f16(acc0 + static_cast<uint32_t>(ZT - 1 - (j - kz)) * 64u, kDescHi | a_lo,
                            kDescHi | (b_tap + static_cast<uint32_t>(kz) * 64u), id1, (kz == 0) ? 0u : 1u);
        }
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          if (k == 0 && init) continue;
          ptx::tc_mma_f16(dcol, kDescHi | (a_lo + static_cast<uint32_t>(k) * a_kstep),
                          kDescHi | (b_lo + static_cast<uint32_t>(k) * b_kstep), idn, 1u);
        }
        if (release) ptx::tc_commit(&bars->slab_empty[sj]);
      }
      if (++sj == nslot) sj = 0;
    }
  };
  for (int wi = 0, w; (w = work_at(p, wi, n_work)) >= 0; ++wi, ++acnt) {
    const wdno_nchunk ci = s_chunks[w % n_chunks];
    const uint32_t buf = acnt & 1u, aph = (acnt >> 1) & 1u;
    PROF_REGION(0, ptx::mbar_wait(&bars->acc_empty[buf], aph ^ 1u));
    ptx::tc_fence_after();
    const uint32_t acc0 = tmem_base + buf * 256u;
    for (int si = 0; si < ci.set_count; ++si) {
      const wdno_kset st = s_sets[ci.set_begin + si];
      const int T = st.tap_count;  // in-plane taps
      const wdno_tap* tp_ptr = s_taps + st.tap_begin;
      int t = 0;
      for (int gi = 0; gi < T / TPS; ++gi) {
        PROF_REGION(2, ptx::mbar_wait(&bars->b_full[bst], bph));
        ptx::tc_fence_after();
        for (int i = 0; i < TPS; ++i, ++t) {
          const uint32_t a_tap = a_lo0 + static_cast<uint32_t>(tp_ptr[t].shift);
          const uint32_t b_tap = b_lo0 + bst * bstage_u + static_cast<uint32_t>(i) * btile_u;
          const bool last_t = (t == T - 1);
          if (t == 0) {
            // acquire the K-set's planes one by one so the MMAs start as soon as the first plane has landed
            for (int j = 0; j < P; ++j) {
              uint32_t sj = s0 + j, ph = sph;
              if (sj >= nslot) { sj -= nslot; ph ^= 1u; }
              PROF_REGION(1, ptx::mbar_wait(&bars->slab_full[sj], ph));
              ptx::fence_proxy_async_smem();  // cp.async (generic proxy) slab writes -> tcgen05 (async proxy) reads
              PROF_REGION(3, if (ptx::elect_one()) issue(a_tap, b_tap, acc0, j, j + 1, si == 0, last_t));
              __syncwarp();
            }
          } else {
            PROF_REGION(3, if (ptx::elect_one()) issue(a_tap, b_tap, acc0, 0, P, false, last_t));
            __syncwarp();
          }
        }
        if (ptx::elect_one()) ptx::tc_commit(&bars->b_empty[bst]);
        __syncwarp();
        if (++bst == nbst) { bst = 0; bph ^= 1u; }
      }
      s0 += static_cast<uint32_t>(P);
      if (s0 >= nslot) { s0 -= nslot; sph ^= 1u; }
    }
    if (ptx::elect_one()) ptx::tc_commit(&bars->acc_full[buf]);
    __syncwarp();
  }
  PROF_COMMIT(0, (threadIdx.x & 31) == 0);
}

__global__ void __launch_bounds__(kThreads, 1) tapgemm_kernel(const wdno_tapgemm_params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  pdl_trigger();
  Bars* bars = reinterpret_cast<Bars*>(smem);
  wdno_tap* s_taps = reinterpret_cast<wdno_tap*>(smem + kBarBytes);
  uint8_t* stage_base = smem + kBarBytes + kMaxTaps * 8;
  float* s_bias = reinterpret_cast<float*>(stage_base + kStageBytes);
  wdno_nchunk* s_chunks = reinterpret_cast<wdno_nchunk*>(stage_base + kStageBytes + kMaxBias * 4);
  wdno_kset* s_sets = reinterpret_cast<wdno_kset*>(stage_base + kStageBytes + kMaxBias * 4 + kMaxChunks * 40);
  uint8_t* slab_base = smem + kHdrBytes;
  const int CH = p.KC >> 3;  // 16-byte chunks per position
  const uint32_t lbo_a = static_cast<uint32_t>(p.S_pad) * 16u;
  const uint32_t slot_bytes = static_cast<uint32_t>(CH) * lbo_a;
  uint8_t* b_base = slab_base + ((static_cast<size_t>(p.NSLOT) * slot_bytes + 127) & ~static_cast<size_t>(127));
  const uint32_t btile_bytes = static_cast<uint32_t>(p.N) * static_cast<uint32_t>(p.KC) * 2u * (p.zstack ? p.KD : 1);
  const uint32_t bstage_bytes = btile_bytes * static_cast<uint32_t>(p.TPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int NACC = p.ZT * p.PT;
  const int NPAD = (p.N <= 64) ? 64 : 128;
  const int NBUF = (NACC * NPAD * 2 <= 512) ? 2 : 1;
  const int P = p.ZT + p.KD - 1;  // input planes per K-set
  const int positions = p.H * p.Wp;
  const int ptiles = (positions + 128 * p.PT - 1) / (128 * p.PT);
  const int zgroups = (p.D + p.ZT - 1) / p.ZT;
  const int n_work = p.B * p.strips * zgroups * ptiles * (p.reuse ? 1 : p.n_chunks);

  // ---------------------------------------------------------------- setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.NSLOT; ++i) {
      ptx::mbar_init(&bars->slab_full[i], kProdThreads);  // every producer thread arrives (directly or via cp.async)
      ptx::mbar_init(&bars->slab_empty[i], 1);
      ptx::mbar_init(&bars->raw_full[i], kLoadThreads);  // the loader threads' copies arrive by themselves
    }
    for (int i = 0; i < p.NBST; ++i) {
      ptx::mbar_init(&bars->b_full[i], 1);
      ptx::mbar_init(&bars->b_empty[i], 1);
      ptx::mbar_init(&bars->b_peer[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&bars->acc_full[i], 1);
      ptx::mbar_init(&bars->acc_empty[i], kEpiWarps);
    }
    ptx::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.n_taps; i += blockDim.x) s_taps[i] = p.taps[i];
  for (int i = threadIdx.x; i < p.n_chunks; i += blockDim.x) s_chunks[i] = p.chunks[i];
  for (int i = threadIdx.x; i < p.n_sets; i += blockDim.x) s_sets[i] = p.sets[i];
  if (p.bias != nullptr)
    for (int i = threadIdx.x; i < p.bias_len; i += blockDim.x) s_bias[i] = p.bias[i];
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(&bars->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (p.cluster == 2) ptx::cluster_sync();   // the peer's mbarriers exist before the first multicast copy / remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // everything above touched only plan constants (tables, bias) and on-chip state; activations, GroupNorm
  // coefficients, statistics and outputs are the predecessors' business from here on
  pdl_wait();

  if (warp >= kFirstLoadWarp) {
    // ============================================================ A loaders (GroupNorm+SiLU prologue only)
    // With a fused activation the 8 producer warps spend their time transforming; these two warps run ahead of them
    // and only issue the asynchronous copies of every plane (same slab layout), so copy latency, address arithmetic
    // and the transform overlap instead of alternating in one thread.
    const bool any_act = (p.coef_a[0] != nullptr) || (p.coef_a[1] != nullptr);
    if (any_act) {
      const int ltid = threadIdx.x - kFirstLoadWarp * 32;
      const int S = 128 * p.PT + p.maxshift;
      const int P = p.ZT + p.KD - 1;
      const int c = ltid & (CH - 1);
      const int ch_shift = (CH == 8) ? 3 : (CH == 4) ? 2 : 1;
      const int s_first = ltid >> ch_shift;
      const int SP = kLoadThreads >> ch_shift;
      const int step_y = SP / p.Wp, step_x = SP - step_y * p.Wp;
      const uint32_t wp_magic = 0xFFFFFFFFu / static_cast<uint32_t>(p.Wp) + 1u;
      int Hs = p.H, Ws = p.Wfull;
      if (p.src_mode == 1) { Hs = 2 * p.H; Ws = 2 * p.Wfull; }
      if (p.src_mode == 2) { Hs = p.H >> 1; Ws = p.Wfull >> 1; }
      uint32_t slot = 0, ph = 0;
      for (int wi = 0, w; (w = work_at(p, wi, n_work)) >= 0; ++wi) {
        const Work wk = decode_work(w, p.n_chunks, ptiles, zgroups, p.reuse, p.strips, p.W);
        const wdno_nchunk ci = s_chunks[wk.nc0];
        const int q0 = wk.pt * 128 * p.PT + s_first;
        const int yp0 = static_cast<int>(__umulhi(static_cast<uint32_t>(q0), wp_magic));
        const int xp0 = q0 - yp0 * p.Wp;
        for (int si = 0; si < ci.set_count; ++si) {
          const wdno_kset st = s_sets[ci.set_begin + si];
          const int csrc = p.src_c[st.src];
          for (int j = 0; j < P; ++j) {
            const int zi = wk.zg * p.ZT - p.pz + j;
            const bool zok = (zi >= 0) && (zi < p.D);
            const __half* plane = static_cast<const __half*>(p.src[st.src]) +
                                  ((static_cast<size_t>(wk.b) * p.D + (zok ? zi : 0)) * Hs * Ws + (st.ph_y * Ws + st.ph_x)) * csrc +
                                  st.ch_off + c * 8;
            const uint32_t dst = ptx::smem_u32(slab_base) + slot * slot_bytes + static_cast<uint32_t>(c) * lbo_a;
            ptx::mbar_wait(&bars->slab_empty[slot], ph ^ 1u);
            int yp = yp0, xp = xp0;
            for (int sq = s_first; sq < S; sq += SP) {
              const int y = yp - p.py, x = xp - p.px + wk.xs0;
              const bool ok = zok && y >= 0 && y < p.H && x >= 0 && x < p.Wfull;
              int so = y * Ws + x;
              if (p.src_mode == 1) so = 2 * y * Ws + 2 * x;
              if (p.src_mode == 2) so = (y >> 1) * Ws + (x >> 1);
              ptx::cp_async16_zfill(dst + static_cast<uint32_t>(sq) * 16u, ok ? plane + static_cast<size_t>(so) * csrc : plane,
                                    ok ? 16u : 0u);
              xp += step_x;
              yp += step_y;
              if (xp >= p.Wp) { xp -= p.Wp; ++yp; }
            }
            ptx::cp_async_mbar_arrive_noinc(&bars->raw_full[slot]);
            if (++slot == static_cast<uint32_t>(p.NSLOT)) { slot = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp >= kFirstProdWarp) {
    // ============================================================ A producers
    // thread -> fixed 16-byte chunk c of positions s_first, s_first+SP, ... of every plane.  The source offset of each of
    // the thread's positions depends only on the work item's position tile, so it is computed once per work item
    // (kIt cached offsets; longer slabs recompute the tail) and reused by all planes / K-sets of the item.
    // Every plane is fetched with asynchronous 16-byte copies (LDGSTS, zero fill outside the tensor).  Identity
    // prologue: the copies signal the slot's mbarrier themselves, so a thread never waits for data.  GroupNorm+SiLU
    // prologue: up to kAhead planes are kept in flight (cp.async groups); the oldest one is then transformed in
    // place in shared memory (each thread touches only the chunks it copied; kIt independent chains) and published.
    constexpr int kIt = 8;
    constexpr int kSkip = -2, kZero = -1;  // offset codes: beyond the slab / outside the tensor (zero fill)
    const int ptid = threadIdx.x - kFirstProdWarp * 32;
    const int S = 128 * p.PT + p.maxshift;  // positions needed per plane
    const int c = ptid & (CH - 1);
    const int ch_shift = (CH == 8) ? 3 : (CH == 4) ? 2 : 1;
    const int s_first = ptid >> ch_shift;
    const int SP = kProdThreads >> ch_shift;  // positions covered per sweep of all producer threads
    const int n_it = (S - s_first + SP - 1) / SP;  // this thread's positions per plane
    const uint32_t wp_magic = 0xFFFFFFFFu / static_cast<uint32_t>(p.Wp) + 1u;  // exact q / Wp for q * Wp < 2^32
    int Hs = p.H, Ws = p.Wfull;
    if (p.src_mode == 1) { Hs = 2 * p.H; Ws = 2 * p.Wfull; }
    if (p.src_mode == 2) { Hs = p.H >> 1; Ws = p.Wfull >> 1; }
    const bool any_act = (p.coef_a[0] != nullptr) || (p.coef_a[1] != nullptr);

    // source position (ys * Ws + xs, without the space-to-depth phase) of slab position index i, or a code
    // (xs0: first tensor column of the work item's column strip; the strip's halo columns are real neighbours)
    auto src_pos = [&](int q0, int xs0, int i) -> int {
      if (i >= n_it) return kSkip;
      const int q = q0 + s_first + i * SP;
      const int yp = static_cast<int>(__umulhi(static_cast<uint32_t>(q), wp_magic));
      const int y = yp - p.py, x = q - yp * p.Wp - p.px + xs0;
      if (y < 0 || y >= p.H || x < 0 || x >= p.Wfull) return kZero;
      if (p.src_mode == 1) return 2 * y * Ws + 2 * x;
      if (p.src_mode == 2) return (y >> 1) * Ws + (x >> 1);
      return y * Ws + x;
    };

    // cursor over the sequence of plane jobs (work item, K-set, plane) this CTA produces
    struct Cursor {
      int w, wi, si, j, nset, set_begin;
      uint32_t slot, ph;
      int b, z0, q0, xs0;
      int off[kIt];
    };
    auto load_work = [&](Cursor& cu) {
      if (cu.w < 0) return;
      const Work wk = decode_work(cu.w, p.n_chunks, ptiles, zgroups, p.reuse, p.strips, p.W);
      const wdno_nchunk ci = s_chunks[wk.nc0];  // with reuse every chunk shares chunk 0's K-sets
      cu.nset = ci.set_count;
      cu.set_begin = ci.set_begin;
      cu.b = wk.b;
      cu.z0 = wk.zg * p.ZT;
      cu.q0 = wk.pt * 128 * p.PT;
      cu.xs0 = wk.xs0;
#pragma unroll
      for (int i = 0; i < kIt; ++i) cu.off[i] = src_pos(cu.q0, cu.xs0, i);
    };
    auto advance = [&](Cursor& cu) {
      if (++cu.slot == static_cast<uint32_t>(p.NSLOT)) { cu.slot = 0; cu.ph ^= 1u; }
      if (++cu.j == P) {
        cu.j = 0;
        if (++cu.si == cu.nset) {
          cu.si = 0;
          cu.w = work_at(p, ++cu.wi, n_work);
          load_work(cu);
        }
      }
    };
    // issue the asynchronous copies of the cursor's plane (the slot must be free)
    auto issue = [&](const Cursor& cu) {
      const wdno_kset st = s_sets[cu.set_begin + cu.si];
      const int csrc = p.src_c[st.src];
      const int zi = cu.z0 - p.pz + cu.j;
      const bool zok = (zi >= 0) && (zi < p.D);
      const __half* plane = static_cast<const __half*>(p.src[st.src]) +
                            ((static_cast<size_t>(cu.b) * p.D + (zok ? zi : 0)) * Hs * Ws + (st.ph_y * Ws + st.ph_x)) * csrc +
                            st.ch_off + c * 8;
      const uint32_t dst = ptx::smem_u32(slab_base) + cu.slot * slot_bytes + static_cast<uint32_t>(c) * lbo_a +
                           static_cast<uint32_t>(s_first) * 16u;
#pragma unroll
      for (int i = 0; i < kIt; ++i) {
        const int o = cu.off[i];
        if (o != kSkip) {
          const bool ok = zok && (o >= 0);
          ptx::cp_async16_zfill(dst + static_cast<uint32_t>(i * SP) * 16u, ok ? plane + static_cast<size_t>(o) * csrc : plane,
                                ok ? 16u : 0u);
        }
      }
      for (int i = kIt; i < n_it; ++i) {  // long slabs (2-D layers with PT = 4): offsets recomputed
        const int o = src_pos(cu.q0, cu.xs0, i);
        const bool ok = zok && (o >= 0);
        ptx::cp_async16_zfill(dst + static_cast<uint32_t>(i * SP) * 16u, ok ? plane + static_cast<size_t>(o) * csrc : plane,
                              ok ? 16u : 0u);
      }
    };
    // in-place silu(a*x + c) on the chunks this thread copied (zero-filled positions stay zero)
    auto transform = [&](const Cursor& cu) {
      const wdno_kset st = s_sets[cu.set_begin + cu.si];
      if (p.coef_a[st.src] == nullptr) return;
      const int zi = cu.z0 - p.pz + cu.j;
      if (zi < 0 || zi >= p.D) return;
      const int csrc = p.src_c[st.src];
      const int chn = st.ch_off + c * 8;
      const size_t sample = p.fold ? static_cast<size_t>(cu.b) * p.D + zi : static_cast<size_t>(cu.b);
      const float* pa = p.coef_a[st.src] + sample * csrc + chn;
      const float* pc = p.coef_c[st.src] + sample * csrc + chn;
      // silu(v) = h + h * tanh(h) with h = v / 2 = (a/2) x + c/2
      float4 a0 = __ldg(reinterpret_cast<const float4*>(pa)), a1 = __ldg(reinterpret_cast<const float4*>(pa + 4));
      float4 c0 = __ldg(reinterpret_cast<const float4*>(pc)), c1 = __ldg(reinterpret_cast<const float4*>(pc + 4));
      a0.x *= 0.5f; a0.y *= 0.5f; a0.z *= 0.5f; a0.w *= 0.5f; a1.x *= 0.5f; a1.y *= 0.5f; a1.z *= 0.5f; a1.w *= 0.5f;
      c0.x *= 0.5f; c0.y *= 0.5f; c0.z *= 0.5f; c0.w *= 0.5f; c1.x *= 0.5f; c1.y *= 0.5f; c1.z *= 0.5f; c1.w *= 0.5f;
      uint8_t* dst = slab_base + static_cast<size_t>(cu.slot) * slot_bytes + static_cast<size_t>(c) * lbo_a +
                     static_cast<size_t>(s_first) * 16;
      auto act8 = [&](uint4& v) {
        __half2* h = reinterpret_cast<__half2*>(&v);
        float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
        float2 f2 = __half22float2(h[2]), f3 = __half22float2(h[3]);
        f0.x = silu_half(fmaf(a0.x, f0.x, c0.x)); f0.y = silu_half(fmaf(a0.y, f0.y, c0.y));
        f1.x = silu_half(fmaf(a0.z, f1.x, c0.z)); f1.y = silu_half(fmaf(a0.w, f1.y, c0.w));
        f2.x = silu_half(fmaf(a1.x, f2.x, c1.x)); f2.y = silu_half(fmaf(a1.y, f2.y, c1.y));
        f3.x = silu_half(fmaf(a1.z, f3.x, c1.z)); f3.y = silu_half(fmaf(a1.w, f3.y, c1.w));
        h[0] = wdno::h2_sat(f0); h[1] = wdno::h2_sat(f1);
        h[2] = wdno::h2_sat(f2); h[3] = wdno::h2_sat(f3);
      };
#pragma unroll
      for (int i0 = 0; i0 < kIt; i0 += 4) {  // 4 independent load -> math -> store chains in flight
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (cu.off[i0 + u] >= 0) v[u] = *reinterpret_cast<const uint4*>(dst + static_cast<size_t>((i0 + u) * SP) * 16);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (cu.off[i0 + u] >= 0) act8(v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (cu.off[i0 + u] >= 0) *reinterpret_cast<uint4*>(dst + static_cast<size_t>((i0 + u) * SP) * 16) = v[u];
      }
      for (int i = kIt; i < n_it; ++i) {
        if (src_pos(cu.q0, cu.xs0, i) >= 0) {
          uint4 v = *reinterpret_cast<const uint4*>(dst + static_cast<size_t>(i * SP) * 16);
          act8(v);
          *reinterpret_cast<uint4*>(dst + static_cast<size_t>(i * SP) * 16) = v;
        }
      }
    };

    Cursor iss;
    iss.wi = 0; iss.w = work_at(p, 0, n_work); iss.si = 0; iss.j = 0; iss.slot = 0; iss.ph = 0; iss.nset = 0; iss.set_begin = 0;
    iss.b = 0; iss.z0 = 0; iss.q0 = 0;
#pragma unroll
    for (int i = 0; i < kIt; ++i) iss.off[i] = kSkip;
    load_work(iss);
    PROF_DECL;
    if (!any_act) {
      while (iss.w >= 0) {
        PROF_REGION(0, ptx::mbar_wait(&bars->slab_empty[iss.slot], iss.ph ^ 1u));
        issue(iss);
        ptx::cp_async_mbar_arrive_noinc(&bars->slab_full[iss.slot]);
        advance(iss);
      }
    } else {
      // the loader warps fill the slot; transform it in place (each thread its own chunks) and publish
      while (iss.w >= 0) {
        PROF_REGION(0, ptx::mbar_wait(&bars->raw_full[iss.slot], iss.ph));
        PROF_REGION(3, transform(iss));
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bars->slab_full[iss.slot]);
        advance(iss);
      }
    }
    PROF_COMMIT(8, ptid == 0);
  } else if (warp == kBWarp) {
    // ============================================================ B (weight tile) producer
    // warp-uniform control flow, one elected lane issues the bulk copies (TPS tiles per stage)
    // Cluster mode (opt-in, WDNO_CLUSTER=1): all CTAs of a layer stream the SAME few hundred KB of weight tiles out of L2, again
    // for every work item; a CTA pair fetches each tile once -- the leader (cluster rank 0) issues one multicast copy that
    // lands at the same offset in both CTAs and completes the same-offset mbarrier in both.  The peer only arms its barrier and
    // tells the leader (remote arrive) that its stage is free.  Measured neutral (256 -> 256 at 10 x 10: 170.5 vs 171.9 us,
    // 64 -> 64 at 40 x 40: 44.6 vs 43.9 us): the weight stream out of L2 is NOT what holds these layers below the MMA rate.
    uint32_t bst = 0, bph = 0;
    const bool clustered = p.cluster == 2;
    const uint32_t crank = clustered ? ptx::cluster_ctarank() : 0u;
    PROF_DECL;
    for (int wi = 0, w; (w = work_at(p, wi, n_work)) >= 0; ++wi) {
      const Work wk = decode_work(w, p.n_chunks, ptiles, zgroups, p.reuse);
      for (int nc = wk.nc0; nc < wk.nc1; ++nc) {
        const wdno_nchunk ci = s_chunks[nc];
        const uint8_t* wsrc = static_cast<const uint8_t*>(p.wpacked) + static_cast<size_t>(ci.w_tile_off) * btile_bytes;
        for (int i = 0; i < ci.n_tiles; i += p.TPS) {  // the plan guarantees TPS | taps of every kz group
          const uint32_t cnt = static_cast<uint32_t>(p.TPS);
          PROF_REGION(0, ptx::mbar_wait(&bars->b_empty[bst], bph ^ 1u));
          if (!clustered) {
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&bars->b_full[bst], cnt * btile_bytes);
              ptx::bulk_g2s(b_base + static_cast<size_t>(bst) * bstage_bytes, wsrc + static_cast<size_t>(i) * btile_bytes,
                            cnt * btile_bytes, &bars->b_full[bst]);
            }
          } else if (crank != 0u) {
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&bars->b_full[bst], cnt * btile_bytes);
              ptx::mbar_arrive_remote(&bars->b_peer[bst], 0u);
            }
          } else {
            PROF_REGION(1, ptx::mbar_wait_cluster(&bars->b_peer[bst], bph));
            if (ptx::elect_one()) {
              ptx::mbar_arrive_expect_tx(&bars->b_full[bst], cnt * btile_bytes);
              ptx::bulk_g2s_multicast(b_base + static_cast<size_t>(bst) * bstage_bytes, wsrc + static_cast<size_t>(i) * btile_bytes,
                                      cnt * btile_bytes, &bars->b_full[bst], static_cast<uint16_t>(3));
            }
          }
          __syncwarp();
          if (++bst == static_cast<uint32_t>(p.NBST)) { bst = 0; bph ^= 1u; }
        }
      }
    }
    PROF_COMMIT(16, lane == 0);
  } else if (warp == kMmaWarp) {
    // ============================================================ MMA issuer (templated on the accumulator shape)
    const int ks = p.KC >> 4;
    if (p.zstack) {
#define WDNO_ZS(KS_, KD_) mma_role_zstack<KS_, KD_>(p, bars, s_taps, s_chunks, s_sets, tmem_base, slab_base, b_base, n_work)
      if (p.KD == 3) {
        if (ks == 1) WDNO_ZS(1, 3); else if (ks == 2) WDNO_ZS(2, 3); else WDNO_ZS(4, 3);
      } else {
        if (ks == 1) WDNO_ZS(1, 7); else if (ks == 2) WDNO_ZS(2, 7); else WDNO_ZS(4, 7);
      }
#undef WDNO_ZS
    } else
#define WDNO_MMA(ZT_, PT_) \
    if (ks == 1) mma_role<ZT_, PT_, 1>(p, bars, s_taps, s_chunks, s_sets, tmem_base, slab_base, b_base, n_work, ptiles, zgroups); \
    else if (ks == 2) mma_role<ZT_, PT_, 2>(p, bars, s_taps, s_chunks, s_sets, tmem_base, slab_base, b_base, n_work, ptiles, zgroups); \
    else mma_role<ZT_, PT_, 4>(p, bars, s_taps, s_chunks, s_sets, tmem_base, slab_base, b_base, n_work, ptiles, zgroups);
    {
      if (p.ZT == 4) { WDNO_MMA(4, 1) }
      else if (p.ZT == 2) { WDNO_MMA(2, 1) }
      else if (p.PT == 4) { WDNO_MMA(1, 4) }
      else { WDNO_MMA(1, 1) }
    }
#undef WDNO_MMA
  } else {
    // ============================================================ epilogue warps 0..3
    // fp16 outputs go through a per-warp staging tile (32 rows x 128 B, rows padded to 144 B; the pad holds the row's
    // global offset) so that global stores are coalesced: 8 lanes write one row's 128 contiguous bytes.
    const int row = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int nblk = p.N >> 3;  // 8-column blocks, <= 16
    const int n16 = p.N >> 4;
    uint8_t* stage_w = stage_base + warp * kStageWarp;
    uint8_t* stage_row = stage_w + lane * kStageRow;
    const uint32_t wp_magic = 0xFFFFFFFFu / static_cast<uint32_t>(p.Wp) + 1u;  // exact o / Wp for o * Wp < 2^32
    const long long plane_elems = static_cast<long long>(p.H) * p.Wfull * p.out_c * ((p.out_mode == 1) ? 4 : 1);
    const bool has_bias = p.bias != nullptr, has_stats = p.stats != nullptr;
    uint32_t acnt = 0;
    PROF_DECL;
    for (int wi = 0, w; (w = work_at(p, wi, n_work)) >= 0; ++wi) {
      const Work wk = decode_work(w, p.n_chunks, ptiles, zgroups, p.reuse, p.strips, p.W);
      const int o0 = wk.pt * 128 * p.PT;
      const int z0 = wk.zg * p.ZT;
      for (int nc = wk.nc0; nc < wk.nc1; ++nc, ++acnt) {
        const wdno_nchunk ci = s_chunks[nc];
        const uint32_t buf = (NBUF == 2) ? (acnt & 1u) : 0u;
        const uint32_t aph = (NBUF == 2) ? ((acnt >> 1) & 1u) : (acnt & 1u);
        float s1[16], s2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
        // GroupNorm partial sums of `sample`: 8-column blocks -> groups (cpg is a multiple of 8); warp-reduce, then one
        // atomic pair per group; the per-lane sums are cleared (batch-folded 2-D layers flush once per plane)
        auto flush_stats = [&](size_t sample) {
          float g1 = 0.f, g2 = 0.f;
          int cur_g = -1;
#pragma unroll
          for (int blk = 0; blk < 16; ++blk) {
            if (blk >= nblk || blk * 8 >= ci.n_valid) break;
            const int gi = (ci.out_ch_off + blk * 8) / p.cpg;
            if (gi != cur_g) {
              if (cur_g >= 0 && lane == 0) {
                double* sp = p.stats + (sample * p.G + cur_g) * 2;
                atomicAdd(sp, static_cast<double>(g1));
                atomicAdd(sp + 1, static_cast<double>(g2));
              }
              g1 = 0.f;
              g2 = 0.f;
              cur_g = gi;
            }
            float a1 = s1[blk], a2 = s2[blk];
            s1[blk] = 0.f;
            s2[blk] = 0.f;
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) {
              a1 += __shfl_xor_sync(0xffffffffu, a1, sh);
              a2 += __shfl_xor_sync(0xffffffffu, a2, sh);
            }
            g1 += a1;
            g2 += a2;
          }
          if (cur_g >= 0 && lane == 0) {
            double* sp = p.stats + (sample * p.G + cur_g) * 2;
            atomicAdd(sp, static_cast<double>(g1));
            atomicAdd(sp + 1, static_cast<double>(g2));
          }
        };
        const float* bias_s = s_bias + ci.out_ch_off;
        PROF_REGION(0, ptx::mbar_wait(&bars->acc_full[buf], aph));
        ptx::tc_fence_after();
        int a = 0;
        for (int pi = 0; pi < p.PT; ++pi) {
          // row geometry: shared by the ZT planes of this position tile
          const int o = o0 + pi * 128 + row;
          const int y = static_cast<int>(__umulhi(static_cast<uint32_t>(o), wp_magic));
          const int xl = o - y * p.Wp;   // column inside the strip's tap grid
          const int x = xl + wk.xs0;     // tensor column
          const bool valid_yx = (y < p.H) && (xl < p.W) && (x < p.Wfull);
          long long rbase;
          if (p.out_mode == 0) {
            rbase = ((static_cast<long long>(wk.b) * p.D * p.H + y) * p.Wfull + x) * p.out_c + ci.out_ch_off;
          } else if (p.out_mode == 1) {
            rbase = ((static_cast<long long>(wk.b) * p.D * (2 * p.H) + (2 * y + ci.ph_y)) * (2 * p.Wfull) + (2 * x + ci.ph_x)) * p.out_c +
                    ci.out_ch_off;
          } else {
            rbase = (static_cast<long long>(wk.b) * p.D * p.out_c + ci.out_ch_off) * (static_cast<long long>(p.H) * p.Wfull) +
                    static_cast<long long>(y) * p.Wfull + x;
          }
          for (int za = 0; za < p.ZT; ++za, ++a) {
            const int z = z0 + za;
            const bool valid = valid_yx && (z < p.D);
            const uint32_t tcol = tmem_base + lane_base + buf * static_cast<uint32_t>(NACC * NPAD) +
                                  static_cast<uint32_t>((p.zstack ? (p.ZT - 1 - za) : (za * p.PT + pi)) * NPAD);
            if (p.out_mode == 2) {
              // fp32 [B, D, C, H, W]: consecutive lanes are consecutive x -> already coalesced per channel
              const long long obase = rbase + static_cast<long long>(z) * p.out_c * (static_cast<long long>(p.H) * p.Wfull);
              float* o32 = static_cast<float*>(p.out);
              const size_t cs = static_cast<size_t>(p.H) * p.Wfull;
              for (int c16 = 0; c16 < n16; ++c16) {
                uint32_t r[16];
                ptx::tmem_ld16(tcol + static_cast<uint32_t>(c16 * 16), r);
                ptx::tmem_ld_wait();
                if (!valid) continue;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const int col = c16 * 16 + i;
                  if (col < ci.n_valid) o32[obase + col * cs] = __uint_as_float(r[i]) + (has_bias ? bias_s[col] : 0.f);
                }
              }
              continue;
            }
            const long long obase = rbase + static_cast<long long>(z) * plane_elems;
            *reinterpret_cast<long long*>(stage_row + 128) = valid ? obase : -1ll;
#pragma unroll
            for (int cg = 0; cg < 2; ++cg) {  // 64-column groups
              if (cg * 4 >= n16) break;
#pragma unroll
              for (int c2 = 0; c2 < 2; ++c2) {  // 32 columns per TMEM round trip
                const int c16a = cg * 4 + c2 * 2;
                if (c16a >= n16) break;
                uint32_t r[32];
                if (c16a + 1 < n16) {
                  ptx::tmem_ld32(tcol + static_cast<uint32_t>(c16a * 16), r);
                } else {
                  ptx::tmem_ld16(tcol + static_cast<uint32_t>(c16a * 16), *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
#pragma unroll
                  for (int i = 16; i < 32; ++i) r[i] = 0u;
                }
                ptx::tmem_ld_wait();
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
                if (has_bias) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 bv = *reinterpret_cast<const float4*>(bias_s + c16a * 16 + q * 4);
                    f[q * 4 + 0] += bv.x; f[q * 4 + 1] += bv.y; f[q * 4 + 2] += bv.z; f[q * 4 + 3] += bv.w;
                  }
                }
                if (has_stats && valid) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    s1[c16a * 2 + (i >> 3)] += f[i];
                    s2[c16a * 2 + (i >> 3)] = fmaf(f[i], f[i], s2[c16a * 2 + (i >> 3)]);
                  }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint4 ov;
                  __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                  for (int i = 0; i < 4; ++i) oh[i] = wdno::h2_sat(f[q * 8 + 2 * i], f[q * 8 + 2 * i + 1]);
                  *reinterpret_cast<uint4*>(stage_row + (c2 * 4 + q) * 16) = ov;
                }
              }
              // coalesced phase: item = (row, 16-byte chunk); 8 consecutive lanes cover one row's 128 bytes
              __syncwarp();
              __half* o16 = static_cast<__half*>(p.out);
              const __half* res16 = static_cast<const __half*>(p.resid);
              const int ck = lane & 7;
              const int colc = cg * 64 + ck * 8;
              const bool col_ok = colc < ci.n_valid;
              const uint8_t* srow = stage_w + (lane >> 3) * kStageRow;
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {  // 2 x 4 rows in flight per lane
                const uint8_t* sr = srow + hf * 16 * kStageRow;
                long long ob[4];
                uint4 v[4];
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                  ob[it] = *reinterpret_cast<const long long*>(sr + it * 4 * kStageRow + 128);
                  v[it] = *reinterpret_cast<const uint4*>(sr + it * 4 * kStageRow + ck * 16);
                  if (!col_ok) ob[it] = -1ll;
                }
                if (res16 != nullptr) {
                  uint4 rv[4];
#pragma unroll
                  for (int it = 0; it < 4; ++it)
                    if (ob[it] >= 0) rv[it] = __ldg(reinterpret_cast<const uint4*>(res16 + ob[it] + colc));
#pragma unroll
                  for (int it = 0; it < 4; ++it) {
                    if (ob[it] >= 0) {
                      __half2* vh = reinterpret_cast<__half2*>(&v[it]);
                      const __half2* rh = reinterpret_cast<const __half2*>(&rv[it]);
#pragma unroll
                      for (int i = 0; i < 4; ++i) {
                        const float2 a2 = __half22float2(vh[i]), b2 = __half22float2(rh[i]);
                        vh[i] = wdno::h2_sat(a2.x + b2.x, a2.y + b2.y);
                      }
                    }
                  }
                }
#pragma unroll
                for (int it = 0; it < 4; ++it)
                  if (ob[it] >= 0) *reinterpret_cast<uint4*>(o16 + ob[it] + colc) = v[it];
              }
              __syncwarp();
            }
            // batch folded into depth (2-D layers): every plane is a sample of its own
            if (has_stats && p.fold && z < p.D) flush_stats(static_cast<size_t>(wk.b) * p.D + z);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&bars->acc_empty[buf]);
        if (has_stats && !p.fold) flush_stats(static_cast<size_t>(wk.b));
      }
    }
    PROF_COMMIT(24, threadIdx.x == 0);
  }

  // ---------------------------------------------------------------- teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (p.cluster == 2) ptx::cluster_sync();   // neither CTA of a pair leaves while the other may still signal it
  if (warp == kMmaWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

static int64_t smem_bytes_of(const wdno_tapgemm_params* p) {
  const int64_t CH = p->KC / 8;
  const int64_t slot = CH * p->S_pad * 16;
  const int64_t slabs = (p->NSLOT * slot + 127) & ~static_cast<int64_t>(127);
  const int64_t bt = static_cast<int64_t>(p->N) * p->KC * 2 * (p->zstack ? p->KD : 1);
  return kHdrBytes + slabs + static_cast<int64_t>(p->NBST) * p->TPS * bt;
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
  if (p->n_taps < 1 || p->n_taps > kMaxTaps) return set_error(WDNO_E_INVALID, "tapgemm: n_taps must be in [1,384]");
  if (p->TPS < 1 || p->TPS > 64) return set_error(WDNO_E_INVALID, "tapgemm: TPS must be in [1,64]");
  if (p->ZT != 4 && p->ZT != 2 && p->ZT != 1) return set_error(WDNO_E_INVALID, "tapgemm: ZT must be 4, 2 or 1");
  if (p->ZT > 1 && p->PT != 1) return set_error(WDNO_E_INVALID, "tapgemm: PT must be 1 when ZT > 1");
  if (p->ZT == 1 && p->PT != 1 && p->PT != 4) return set_error(WDNO_E_INVALID, "tapgemm: PT must be 1 or 4 when ZT == 1");
  if (p->reuse && p->KD != 1) return set_error(WDNO_E_INVALID, "tapgemm: slab reuse needs KD == 1");
  if (p->S_pad < 128 * p->PT + p->maxshift) return set_error(WDNO_E_INVALID, "tapgemm: S_pad smaller than slab");
  if (p->S_pad >= 16384) return set_error(WDNO_E_INVALID, "tapgemm: S_pad too large for descriptor");
  if (p->B < 1 || p->D < 1 || p->H < 1 || p->W < 1 || p->n_chunks < 1) return set_error(WDNO_E_INVALID, "tapgemm: empty problem");
  if (p->src_c[0] % 8 || (p->src[1] && p->src_c[1] % 8)) return set_error(WDNO_E_INVALID, "tapgemm: source channels must be multiples of 8");
  if (p->out_mode < 0 || p->out_mode > 2) return set_error(WDNO_E_INVALID, "tapgemm: bad out_mode");
  if (p->out_mode != 2 && (p->out_c % 8)) return set_error(WDNO_E_INVALID, "tapgemm: fp16 output channels must be a multiple of 8");
  if (p->stats && (p->cpg % 8 || p->cpg < 8)) return set_error(WDNO_E_INVALID, "tapgemm: cpg must be a multiple of 8");
  if (p->src_mode == 2 && ((p->H & 1) || (p->W & 1))) return set_error(WDNO_E_INVALID, "tapgemm: up2 needs even H,W");
  if (p->fold && p->KD != 1) return set_error(WDNO_E_INVALID, "tapgemm: batch folding needs KD == 1");
  if (p->strips < 1 || p->Wfull < 1) return set_error(WDNO_E_INVALID, "tapgemm: strips / Wfull must be >= 1");
  if (p->strips == 1 && p->Wfull != p->W) return set_error(WDNO_E_INVALID, "tapgemm: Wfull must equal W without strips");
  if (p->strips > 1 && (p->src_mode != 0 || p->out_mode == 1 || p->Wp != p->W + 2 * p->px ||
                        static_cast<int64_t>(p->W) * p->strips < p->Wfull || static_cast<int64_t>(p->W) * (p->strips - 1) >= p->Wfull))
    return set_error(WDNO_E_INVALID, "tapgemm: strips need src_mode 0, out_mode 0/2, Wp = W + 2*px and strips*W covering Wfull");
  if (smem_bytes_of(p) > 227 * 1024) return set_error(WDNO_E_INVALID, "tapgemm: shared-memory plan exceeds 227 KB");
  if (p->grid < 1) return set_error(WDNO_E_INVALID, "tapgemm: grid must be >= 1");
  if (p->zstack && (p->ZT != 4 || p->PT != 1 || p->N != 64 || (p->KD != 3 && p->KD != 7) || p->reuse || p->NSLOT < p->ZT + p->KD - 1))
    return set_error(WDNO_E_INVALID, "tapgemm: zstack needs ZT=4, PT=1, N=64, KD in {3,7}, no slab reuse and NSLOT >= ZT+KD-1");
  if (p->n_chunks > kMaxChunks) return set_error(WDNO_E_INVALID, "tapgemm: more than 32 N-chunks");
  if (p->n_sets < 1 || p->n_sets > kMaxSets) return set_error(WDNO_E_INVALID, "tapgemm: n_sets must be in [1,128]");
  if (p->bias && (p->bias_len < 1 || p->bias_len > kMaxBias)) return set_error(WDNO_E_INVALID, "tapgemm: bias_len must be in [1,1024]");
  if (p->NBST > kMaxBStages) return set_error(WDNO_E_INVALID, "tapgemm: more than 4 weight stages");
  if (p->cluster != 0 && p->cluster != 1 && p->cluster != 2) return set_error(WDNO_E_INVALID, "tapgemm: cluster must be 0, 1 or 2");
  if (p->cluster == 2) {
    const long long positions = static_cast<long long>(p->H) * p->Wp;
    const long long ptiles = (positions + 128 * p->PT - 1) / (128 * p->PT);
    const long long units = static_cast<long long>(p->B) * p->strips * ((p->D + p->ZT - 1) / p->ZT) * ptiles;
    if (p->reuse || (units & 1) || (p->grid & 1))
      return set_error(WDNO_E_INVALID, "tapgemm: CTA pairs need an even grid, an even number of (sample, plane group, tile) units and no slab reuse");
  }
  return WDNO_OK;
}

}  // namespace wdno

#ifdef WDNO_PROF
// debug builds: read (and clear) the per-role stall counters; not part of the product ABI
extern "C" int wdno_tapgemm_prof_read(unsigned long long* out32_host) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out32_host, wdno::g_prof, sizeof(unsigned long long) * 32);
  unsigned long long z[32] = {0};
  cudaMemcpyToSymbol(wdno::g_prof, z, sizeof(z));
  return 0;
}
#endif

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
  cudaError_t le;
  if (p->cluster == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p->grid);
    cfg.blockDim = dim3(wdno::kThreads);
    cfg.dynamicSmemBytes = static_cast<size_t>(smem);
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = wdno::pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    le = cudaLaunchKernelEx(&cfg, wdno::tapgemm_kernel, *p);
  } else {
    le = wdno::launch_pdl(wdno::tapgemm_kernel, dim3(p->grid), dim3(wdno::kThreads), static_cast<size_t>(smem),
                          static_cast<cudaStream_t>(stream), *p);
  }
  if (le != cudaSuccess) return wdno::set_cuda_error(le, "tapgemm: launch");
  return wdno::check_launch("tapgemm");
}

// CTAs of 2-CTA clusters that can be resident at once with `smem_bytes` of dynamic shared memory each (the persistent grid of a
// cluster-mode plan must not exceed it: a pair that does not fit would run as a second wave)
extern "C" int wdno_tapgemm_max_cluster_ctas(int64_t smem_bytes) {
  if (smem_bytes < 0 || smem_bytes > 227 * 1024) return WDNO_E_INVALID;
  cudaError_t e = cudaFuncSetAttribute(wdno::tapgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return wdno::set_cuda_error(e, "tapgemm: cudaFuncSetAttribute");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * wdno::num_sms());
  cfg.blockDim = dim3(wdno::kThreads);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  e = cudaOccupancyMaxActiveClusters(&n, wdno::tapgemm_kernel, &cfg);
  if (e != cudaSuccess) return wdno::set_cuda_error(e, "tapgemm: cudaOccupancyMaxActiveClusters");
  return 2 * n;
}
