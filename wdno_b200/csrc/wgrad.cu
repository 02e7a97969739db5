// Weight gradient of every convolution / projection of the two U-Nets (training step):
//   dW[m][n][tap] += scale * sum over positions p of  dY[p][m] * X[p + shift(tap)][n]
// for Conv3d/Conv2d "same" (conv3d.py:192,393; unet.py:133,317), 1x1 (conv3d.py:216,471), the (1,4,4) stride-2 conv and its
// transposed twin (conv3d.py:159-163; one launch per sub-pixel phase, roles of X and dY swapped for the transposed one),
// Linear layers of the attention blocks (conv3d.py:291-292).  The reference gets these from cuDNN/cuBLAS through autograd
// (Trainer.train, diffusion_2d.py:1277-1284: loss.backward()).
//
// GEMM view: M = dY channels, N = X channels x taps, K = positions (B*D*H*W: up to 230k per sample batch).  Both operands
// are channels-last fp16, i.e. K-row / MN-contiguous: exactly what ldmatrix.trans turns into mma.sync fragments.
//   * positions are linearised over a padded row (Wp = W + max |dx|): one run of zero columns serves both neighbours, so
//     a tap is a ROW SHIFT of the staged X slab -- the slab of a 64-position chunk is loaded once for all taps of a group
//     (taps that share (dz, dy) and differ in dx), and the dY tile once for the whole group;
//   * a CTA owns (chunk range [split-K], tap group, 64x64 channel tile): 4 warps x (32 x 32) x TG taps of fp32 accumulators
//     in registers, 3-stage cp.async ring (zero fill outside the tensor), fp32 atomicAdd into dW at the end;
//   * the bias gradient (column sums of dY) rides along in the CTAs of tap group 0 / first N tile.
// HBM/L2 traffic: every tap group re-reads dY and X (they are L2-resident at training batch sizes: 6 x 9.8 MB per tensor).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wdno_b200.h"
#include "common.cuh"
#include "mma_sync.cuh"

namespace wdno {

namespace {

constexpr int kTG = WDNO_WGRAD_MAX_GROUP_TAPS;  // taps per group (register accumulators: 32 floats per tap per thread)
constexpr int kKC = 64;                         // positions per chunk
constexpr int kPitch = 64 + 8;                  // halfs per smem row (144 B: conflict-free ldmatrix)
constexpr int kMaxSpan = 8;                     // extra slab rows (max dx span of a group)
constexpr int kStages = 3;
constexpr int kThreads = 128;
constexpr int kDyBytes = kKC * kPitch * 2;
constexpr int kXBytes = (kKC + kMaxSpan) * kPitch * 2;
constexpr int kStageBytes = kDyBytes + kXBytes;

__device__ __forceinline__ void cp16z(uint32_t dst, const void* src, bool ok) {
  const uint32_t n = ok ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit_() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait_() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

template <int TG>
__global__ void __launch_bounds__(kThreads) wgrad_kernel(const wdno_wgrad_params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 1, wn = warp >> 1;
  const __half* const px_ = static_cast<const __half*>(p.x);
  const __half* const pdy_ = static_cast<const __half*>(p.dy);
  const wdno_wgrad_group grp = p.groups[blockIdx.y];
  const int n_mt = (p.Cy + 63) >> 6;
  const int mt_idx = blockIdx.z % n_mt, nt_idx = blockIdx.z / n_mt;
  const int m0 = mt_idx * 64, n0 = nt_idx * 64;
  const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem));

  const int Wp = p.W + p.padw;
  const int nq = p.H * Wp;
  const int cpp = (nq + kKC - 1) / kKC;            // chunks per plane
  const long long n_chunks = static_cast<long long>(p.B) * p.D * cpp;
  const int Hv = (p.Hs - p.ph_y + p.sy - 1) / p.sy, Wv = (p.Ws - p.ph_x + p.sx - 1) / p.sx;  // extents of the X view
  const int slab_rows = kKC + grp.span;
  const int q_shift = grp.dy * Wp + grp.dx_min;    // slab row r holds X_lin[q0 + q_shift + r]

  // ---- loader: one stage = dY tile [64 positions][64 channels of the M tile] + X slab [64 + span][64 channels of the N tile]
  // A thread copies chunk column c8 of rows r0, r0 + 16, ...: ONE division per tile locates row r0, the others follow by
  // adding 16 positions with a carry into the next padded row (the loader was 60 % of the instruction stream with a
  // division per copy: ncu, profiles/r2_wgrad.md).
  const int r0 = tid >> 3, c8 = tid & 7;
  const bool m_ok = (m0 + c8 * 8) < p.Cy, n_ok = (n0 + c8 * 8) < p.cx_n;
  const int step_y = 16 / Wp, step_x = 16 - step_y * Wp;   // 16 positions = step_y rows + step_x columns
  auto load = [&](long long chunk, int st) {
    const long long plane = chunk / cpp;
    const int q0 = static_cast<int>(chunk - plane * cpp) * kKC;
    const int z = static_cast<int>(plane % p.D);
    const long long b = plane / p.D;
    const uint32_t s_dy = sbase + st * kStageBytes + static_cast<uint32_t>((r0 * kPitch + c8 * 8) * 2);
    const uint32_t s_x = sbase + st * kStageBytes + kDyBytes + static_cast<uint32_t>((r0 * kPitch + c8 * 8) * 2);
    const __half* dyp = pdy_ + (plane * p.H * p.W) * p.Cy + m0 + c8 * 8;
    {
      const int q = q0 + r0;
      int y = q / Wp, x = q - y * Wp;
#pragma unroll
      for (int i = 0; i < kKC / 16; ++i) {
        const bool ok = m_ok && (y < p.H) && (x < p.W);
        cp16z(s_dy + static_cast<uint32_t>(i * 16 * kPitch * 2), ok ? dyp + (static_cast<long long>(y) * p.W + x) * p.Cy : pdy_, ok);
        x += step_x;
        y += step_y;
        if (x >= Wp) { x -= Wp; ++y; }
      }
    }
    const int zs = z + grp.dz;
    const bool zok = n_ok && zs >= 0 && zs < p.D;
    const __half* xp = px_ + ((b * p.D + (zok ? zs : 0)) * p.Hs * p.Ws) * static_cast<long long>(p.Cx) + p.cx_off + n0 + c8 * 8;
    {
      // slab row r holds X_lin[q0 + q_shift + r]; negative positions (above the plane) are zero
      const int q = q0 + q_shift + r0;
      int y = (q >= 0) ? q / Wp : -((-q + Wp - 1) / Wp);
      int x = q - y * Wp;
      for (int r = r0; r < slab_rows; r += 16) {
        const bool ok = zok && (y >= 0) && (y < Hv) && (x < Wv);
        const long long off = (static_cast<long long>(p.sy * y + p.ph_y) * p.Ws + (p.sx * x + p.ph_x)) * p.Cx;
        cp16z(s_x + static_cast<uint32_t>((r - r0) * kPitch * 2), ok ? xp + off : px_, ok);
        x += step_x;
        y += step_y;
        if (x >= Wp) { x -= Wp; ++y; }
      }
    }
  };

  float acc[TG][2][4][4];
#pragma unroll
  for (int t = 0; t < TG; ++t)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[t][i][j][c] = 0.f;
  const bool do_bias = (p.dbias != nullptr) && (blockIdx.y == 0) && (nt_idx == 0);
  float bsum = 0.f;

  // chunks of this CTA: c = blockIdx.x, blockIdx.x + gridDim.x, ...
  long long c_load = blockIdx.x;
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (c_load < n_chunks) load(c_load, s);
    cp_commit_();
    c_load += gridDim.x;
  }
  // per-lane ldmatrix offsets (bytes)
  const uint32_t a_off = static_cast<uint32_t>((((lane & 7) + ((lane >> 4) & 1) * 8) * kPitch + wm * 32 + ((lane >> 3) & 1) * 8) * 2);
  const uint32_t b_off = static_cast<uint32_t>((((lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + wn * 32 + ((lane >> 4) & 1) * 8) * 2);
  int it = 0;
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
    cp_wait_<kStages - 2>();
    __syncthreads();
    if (c_load < n_chunks) load(c_load, (it + kStages - 1) % kStages);
    cp_commit_();
    c_load += gridDim.x;
    const uint32_t s_dy = sbase + (it % kStages) * kStageBytes, s_x = s_dy + kDyBytes;
    if (do_bias) {
      // column sums of the dY tile: thread -> channel (tid & 63), half of the positions (tid >> 6)
      const __half* t = reinterpret_cast<const __half*>(smem + (it % kStages) * kStageBytes) + (tid >> 6) * 32 * kPitch + (tid & 63);
      float s = 0.f;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) s += __half2float(t[r * kPitch]);
      bsum += s;
    }
#pragma unroll
    for (int ks = 0; ks < kKC / 16; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ldsm_x4_trans(s_dy + a_off + static_cast<uint32_t>((ks * 16 * kPitch + mt * 16) * 2), a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
#pragma unroll
      for (int t = 0; t < TG; ++t) {
        if (t < grp.n) {
          const uint32_t xb = s_x + b_off + static_cast<uint32_t>(((ks * 16 + grp.dxo[t]) * kPitch) * 2);
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_trans(xb + static_cast<uint32_t>(np * 16 * 2), b0, b1, b2, b3);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              mma16816(acc[t][mt][2 * np], a[mt], b0, b1);
              mma16816(acc[t][mt][2 * np + 1], a[mt], b2, b3);
            }
          }
        }
      }
    }
  }
  cp_wait_<0>();

  // ---- epilogue: fp32 atomics into dW[(m * Ntot + n_off + n) * Ttot + tap]
  const int g = lane >> 2, q2 = (lane & 3) * 2;
#pragma unroll
  for (int t = 0; t < TG; ++t) {
    if (t < grp.n) {
      const long long tap = grp.out[t];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int m = m0 + wm * 32 + mt * 16 + g + ((cc >> 1) ? 8 : 0);
            const int n = n0 + wn * 32 + nt * 8 + q2 + (cc & 1);
            if (m < p.m_valid && n < p.cx_n)
              atomicAdd(p.dw + (static_cast<long long>(m) * p.n_total + p.n_off + n) * p.t_total + tap, acc[t][mt][nt][cc] * p.scale);
          }
    }
  }
  if (do_bias && (m0 + (tid & 63)) < p.m_valid) atomicAdd(p.dbias + m0 + (tid & 63), bsum * p.scale);
}

}  // namespace

}  // namespace wdno

extern "C" int wdno_wgrad(const wdno_wgrad_params* p, void* stream) {
  using namespace wdno;
  if (!p || !p->x || !p->dy || !p->dw || !p->groups) return set_error(WDNO_E_INVALID, "wgrad: null argument");
  if (p->B < 1 || p->D < 1 || p->H < 1 || p->W < 1 || p->n_groups < 1) return set_error(WDNO_E_INVALID, "wgrad: empty problem");
  if ((p->Cy & 7) || (p->Cx & 7) || (p->cx_off & 7) || p->cx_n < 1 || p->cx_off + p->cx_n > ((p->Cx + 7) & ~7))
    return set_error(WDNO_E_INVALID, "wgrad: channel counts / offsets must be multiples of 8");
  if (p->sy < 1 || p->sx < 1 || p->ph_y < 0 || p->ph_x < 0 || p->padw < 0 || p->padw > 7)
    return set_error(WDNO_E_INVALID, "wgrad: bad view / pad");
  if (p->split < 1) return set_error(WDNO_E_INVALID, "wgrad: split must be >= 1");
  if (p->m_valid < 1 || p->m_valid > p->Cy) return set_error(WDNO_E_INVALID, "wgrad: m_valid must be in [1, Cy]");
  const int n_mt = (p->Cy + 63) / 64, n_nt = (p->cx_n + 63) / 64;
  dim3 grid(p->split, p->n_groups, n_mt * n_nt);
  const size_t smem = static_cast<size_t>(kStages) * kStageBytes;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<kTG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e, "wgrad: cudaFuncSetAttribute");
    configured = true;
  }
  wgrad_kernel<kTG><<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(*p);
  return check_launch("wgrad");
}
