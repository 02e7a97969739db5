// Weight gradient of the 'same' convolutions on tcgen05 (training step; see wgrad.cu for the
// mma.sync form that still serves the strided / 1x1 / tiny layers):
//   dW[co][ci][tap] += scale * sum over positions p of  dY[p][co] * X[p + shift(tap)][ci]
// GEMM view: M = dY channels, N = X channels (one MMA per tap), K = positions.  Both operands are channels-last fp16, i.e.
// K-row / MN-contiguous: staged as [8-channel chunk][position][8] they ARE the UMMA canonical MN-major no-swizzle layout
// (core matrix = 8 positions x 16 B), so
//   * a filter tap is the X slab read through a descriptor whose start address is shifted by shift(tap) x 16 B -- no im2col;
//   * for 64-channel layers TWO adjacent dY planes (z, z + 1) are stacked along M: with the X plane z + kz - pz the upper 64
//     accumulator rows collect depth tap kz and the lower 64 rows depth tap kz - 1 (the pair moves through ALL planes, so
//     both sums are complete); wider layers put 128 output channels on M;
//   * up to 8 in-plane taps x 64 columns of fp32 accumulators live in TMEM for the whole K loop; a CTA owns one job
//     (depth-tap pair, in-plane tap group, channel tiles) x a slice of the positions, 8 producer warps feed a 3-stage cp.async
//     ring (zero fill outside the tensor), one thread issues the MMAs, the epilogue adds the accumulators into dW (fp32 red).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wdno_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace wdno {

namespace {

constexpr int kProdWarps = 8, kProdThreads = kProdWarps * 32;
constexpr int kThreads = 32 + kProdThreads;        // warp 0: MMA issuer; warps 1..8: producers; epilogue: warps 0..3
constexpr int kStagesMax = 3;
constexpr int kPos = 128;                           // positions per K item
constexpr int kATile = 16 * kPos * 16;              // 16 chunks x 128 positions x 16 B

struct Bars {
  uint64_t full[kStagesMax];
  uint64_t empty[kStagesMax];
  uint64_t done;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo16, uint32_t sbo16) {
  // no swizzle: start >> 4 | (LBO >> 4) << 16 | (SBO >> 4) << 32 | version 1 << 46
  return static_cast<uint64_t>(((saddr >> 4) & 0x3FFFu) | ((lbo16 & 0x3FFFu) << 16)) |
         (static_cast<uint64_t>((sbo16 & 0x3FFFu) | (1u << 14)) << 32);
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const wdno_wgrad_tc_params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem);
  uint8_t* stage0 = smem + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const wdno_wgrad_tc_job job = p.jobs[blockIdx.y];
  const int Wp = p.W + p.padw;
  const int nq = p.H * Wp;
  const int ptiles = (nq + kPos - 1) / kPos;
  const int S_rows = kPos + job.span;                       // X slab rows
  const int S_pad = (S_rows + 7) & ~7;
  const int xchunks = (p.nx + 7) >> 3;                      // 8-channel chunks of the X tile (N = nx columns)
  const uint32_t stage_bytes = static_cast<uint32_t>(kATile + xchunks * S_pad * 16);
  // planes: stacked pairs run z = -1 .. D - 1 (plane z on rows 0..63, plane z + 1 on rows 64..127); plain tiles z = 0 .. D - 1
  const int z_lo = p.stack ? -1 : 0;
  const int nz = p.D - z_lo;
  const long long n_items = static_cast<long long>(p.B) * nz * ptiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(&bars->full[i], kProdThreads);
      ptx::mbar_init(&bars->empty[i], 1);
    }
    ptx::mbar_init(&bars->done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&bars->tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp >= 1) {
    // ============================================================ producers
    const int ptid = threadIdx.x - 32;
    const int a_pos = ptid & (kPos - 1), a_half = ptid >> 7;     // A: position, upper / lower 8 chunks
    const __half* dy = static_cast<const __half*>(p.dy);
    const __half* x = static_cast<const __half*>(p.x);
    uint32_t st = 0, ph = 0;
    for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
      const int pt = static_cast<int>(it % ptiles);
      const long long r = it / ptiles;
      const int z = static_cast<int>(r % nz) + z_lo;
      const long long b = r / nz;
      const int q0 = pt * kPos;
      ptx::mbar_wait(&bars->empty[st], ph ^ 1u);
      const uint32_t sA = ptx::smem_u32(stage0) + st * stage_bytes;
      const uint32_t sX = sA + kATile;
      {
        // dY: this thread's position, 8 chunks = 64 channels (one 128-byte row of the source)
        const int q = q0 + a_pos;
        const int yy = q / Wp, xx = q - yy * Wp;
        const int zp = p.stack ? z + a_half : z;
        const int ch0 = p.stack ? 0 : job.m0 + a_half * 64;
        const bool ok = (yy < p.H) && (xx < p.W) && (zp >= 0) && (zp < p.D);
        const __half* src = dy + (((b * p.D + (ok ? zp : 0)) * p.H + (ok ? yy : 0)) * p.W + (ok ? xx : 0)) * static_cast<long long>(p.Cy) + ch0;
        const uint32_t dst = sA + static_cast<uint32_t>((a_half * 8 * kPos + a_pos) * 16);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const bool okc = ok && (ch0 + c * 8 < p.Cy);
          ptx::cp_async16_zfill(dst + static_cast<uint32_t>(c * kPos * 16), okc ? src + c * 8 : dy, okc ? 16u : 0u);
        }
      }
      {
        // X slab: rows r = X_lin[q0 + qmin + r]
        const int zx = z + job.kz - p.pz;
        const bool zok = (zx >= 0) && (zx < p.D);
        const __half* xpl = x + ((b * p.D + (zok ? zx : 0)) * p.H * p.W) * static_cast<long long>(p.Cx) + p.cx_off + job.n0;
        for (int rr = ptid; rr < S_rows; rr += kProdThreads) {
          const int q = q0 + job.qmin + rr;
          int yy = 0, xx = 0;
          bool ok = zok && (q >= 0);
          if (ok) {
            yy = q / Wp;
            xx = q - yy * Wp;
            ok = (yy < p.H) && (xx < p.W);
          }
          const __half* src = xpl + (static_cast<long long>(yy) * p.W + xx) * p.Cx;
          const uint32_t dst = sX + static_cast<uint32_t>(rr * 16);
          for (int c = 0; c < xchunks; ++c) {
            const bool okc = ok && (job.n0 + c * 8 < p.cx_n);
            ptx::cp_async16_zfill(dst + static_cast<uint32_t>(c * S_pad * 16), okc ? src + c * 8 : x, okc ? 16u : 0u);
          }
        }
      }
      ptx::cp_async_mbar_arrive_noinc(&bars->full[st]);
      if (++st == static_cast<uint32_t>(p.stages)) { st = 0; ph ^= 1u; }
    }
  } else {
    // ============================================================ MMA issuer (warp 0)
    const uint32_t idesc = ptx::make_idesc_f16(p.nx, 0) | (1u << 15) | (1u << 16);   // A and B MN-major
    const uint32_t a_lbo = p.swap_lbo_sbo ? static_cast<uint32_t>(kPos) : 8u;
    const uint32_t a_sbo = p.swap_lbo_sbo ? 8u : static_cast<uint32_t>(kPos);
    const uint32_t b_lbo = p.swap_lbo_sbo ? static_cast<uint32_t>(S_pad) : 8u;
    const uint32_t b_sbo = p.swap_lbo_sbo ? 8u : static_cast<uint32_t>(S_pad);
    uint32_t st = 0, ph = 0;
    bool first = true;
    for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
      ptx::mbar_wait(&bars->full[st], ph);
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t sA = ptx::smem_u32(stage0) + st * stage_bytes;
        const uint32_t sX = sA + kATile;
#pragma unroll 1
        for (int t = 0; t < job.n_taps; ++t) {
          const uint32_t dcol = tmem + static_cast<uint32_t>(t * p.ncols);
          const uint32_t xb = sX + static_cast<uint32_t>(job.shift[t]) * 16u;
#pragma unroll
          for (int ks = 0; ks < kPos / 16; ++ks) {
            const uint64_t ad = desc_mn(sA + static_cast<uint32_t>(ks * 256), a_lbo, a_sbo);
            const uint64_t bd = desc_mn(xb + static_cast<uint32_t>(ks * 256), b_lbo, b_sbo);
            ptx::tc_mma_f16(dcol, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
          }
        }
        ptx::tc_commit(&bars->empty[st]);
      }
      __syncwarp();
      first = false;
      if (++st == static_cast<uint32_t>(p.stages)) { st = 0; ph ^= 1u; }
    }
    if (ptx::elect_one()) ptx::tc_commit(&bars->done);
    __syncwarp();
  }

  // ============================================================ epilogue: warps 0..3 (TMEM lane quarter = warp)
  if (warp < 4) {
    if (n_items > static_cast<long long>(blockIdx.x)) {
      ptx::mbar_wait(&bars->done, 0);
      ptx::tc_fence_after();
      const int row = warp * 32 + lane;
      const int m = p.stack ? (row & 63) : job.m0 + row;
      const bool lower = p.stack && row >= 64;
      float* dw = p.dw;
      for (int t = 0; t < job.n_taps; ++t) {
        const long long tap = lower ? job.out_lo[t] : job.out_hi[t];
        for (int c0 = 0; c0 < p.nx; c0 += 16) {
          uint32_t r[16];
          ptx::tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(t * p.ncols + c0), r);
          ptx::tmem_ld_wait();
          if (tap >= 0 && m < p.m_valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int n = job.n0 + c0 + i;
              if (n < p.cx_n) atomicAdd(dw + (static_cast<long long>(m) * p.n_total + p.n_off + n) * p.t_total + tap, __uint_as_float(r[i]) * p.scale);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// column sums of dY (bias gradient): thread -> 8 channels of a position stripe
__global__ void __launch_bounds__(256) colsum_f16_kernel(const __half* __restrict__ dy, long long n_pos, int C, float* __restrict__ out,
                                                         float scale) {
  extern __shared__ float red[];
  const int cpv = C >> 3;
  const int lc = threadIdx.x % cpv, r0 = threadIdx.x / cpv, rpb = blockDim.x / cpv;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (r0 < rpb) {
    for (long long v = static_cast<long long>(blockIdx.x) * rpb + r0; v < n_pos; v += static_cast<long long>(gridDim.x) * rpb) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy) + v * cpv + lc);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 t = __half22float2(h[k]);
        acc[2 * k] += t.x;
        acc[2 * k + 1] += t.y;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
  __syncthreads();
  for (int t = threadIdx.x; t < cpv * 8; t += blockDim.x) {
    const int l = t >> 3, k = t & 7;
    float s = 0.f;
    for (int r = 0; r < rpb; ++r) s += red[(r * cpv + l) * 8 + k];
    atomicAdd(out + l * 8 + k, s * scale);
  }
}

}  // namespace

}  // namespace wdno

extern "C" int64_t wdno_wgrad_tc_smem_bytes(const wdno_wgrad_tc_params* p, int span) {
  if (!p) return WDNO_E_INVALID;
  const int S_pad = (wdno::kPos + span + 7) & ~7;
  return 128 + static_cast<int64_t>(p->stages) * (wdno::kATile + ((p->nx + 7) / 8) * S_pad * 16);
}

extern "C" int wdno_wgrad_tc(const wdno_wgrad_tc_params* p, int max_span, void* stream) {
  using namespace wdno;
  if (!p || !p->x || !p->dy || !p->dw || !p->jobs) return set_error(WDNO_E_INVALID, "wgrad_tc: null argument");
  if (p->B < 1 || p->D < 1 || p->H < 1 || p->W < 1 || p->n_jobs < 1 || p->split < 1) return set_error(WDNO_E_INVALID, "wgrad_tc: empty problem");
  if ((p->Cy & 7) || (p->Cx & 7) || (p->cx_off & 7)) return set_error(WDNO_E_INVALID, "wgrad_tc: channels must be multiples of 8");
  if (p->stack && p->Cy != 64) return set_error(WDNO_E_INVALID, "wgrad_tc: plane stacking is for 64 output channels");
  if (!p->stack && (p->Cy & 127)) return set_error(WDNO_E_INVALID, "wgrad_tc: unstacked tiles need a multiple of 128 output channels");
  if (p->nx < 16 || p->nx > 64 || (p->nx & 15) || p->ncols < p->nx || p->ncols * 8 > 512) return set_error(WDNO_E_INVALID, "wgrad_tc: bad N");
  if (p->stages < 2 || p->stages > kStagesMax) return set_error(WDNO_E_INVALID, "wgrad_tc: stages must be 2 or 3");
  const int64_t smem = wdno_wgrad_tc_smem_bytes(p, max_span);
  if (smem > 227 * 1024) return set_error(WDNO_E_INVALID, "wgrad_tc: shared-memory plan exceeds 227 KB");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_cuda_error(e, "wgrad_tc: cudaFuncSetAttribute");
    configured = true;
  }
  wgrad_tc_kernel<<<dim3(p->split, p->n_jobs), kThreads, static_cast<size_t>(smem), static_cast<cudaStream_t>(stream)>>>(*p);
  return check_launch("wgrad_tc");
}

extern "C" int wdno_colsum_f16(const void* dy, int64_t n_pos, int C, float* out, float scale, void* stream) {
  using namespace wdno;
  if (!dy || !out || n_pos < 1 || C < 8 || (C & 7) || C > 2048) return set_error(WDNO_E_INVALID, "colsum_f16: bad arguments");
  const int cpv = C >> 3, rpb = 256 / cpv;
  long long g = (n_pos + rpb - 1) / rpb;
  const long long cap = static_cast<long long>(num_sms()) * 4;
  if (g > cap) g = cap;
  colsum_f16_kernel<<<static_cast<unsigned>(g), 256, 256 * 8 * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(dy), n_pos, C, out, scale);
  return check_launch("colsum_f16");
}
