// 1x1 convolution / Linear over channels-last fp16 activations -- the HBM-bound layers of the U-Nets:
//   res_conv of a ResnetBlock (conv3d.py:216 ; unet.py:162), attention to_qkv / to_out (conv3d.py:291-292 ; unet.py:190-192,
//   233-234), the final 1x1 that emits eps in the reference's fp32 [B,F,C,H,W] layout (conv3d.py:471 ; unet.py:369).
// out[m, n] = sum_k concat(src0, src1)[m, k] * W[n, k] + bias[n] (+ resid[m, n])
//
// These layers move 2*(Cin + Cout) bytes per position for 2*Cin*Cout FLOPs (64..512 FLOP/B at most, usually far below the
// tensor ridge), so the design goal is HBM bandwidth, not tensor throughput: a 128 x BN output tile per CTA, K streamed
// in 32-channel slices through a 4-stage cp.async ring (16-byte copies, rows padded to 80 B so ldmatrix is conflict
// free), mma.sync.m16n8k16 on register fragments, two or three CTAs resident per SM to keep ~100 KB of loads in flight,
// and a shared-memory staged epilogue so that every global store is a full 16-byte (fp16) or 128-byte-per-warp (fp32
// planar) transaction.  The tcgen05 tap-GEMM handles these layers too (and still does when a GroupNorm prologue or
// statistics are fused), but its persistent one-CTA-per-SM pipeline is built for long K loops; here K is 2..48 slices.
#include <cuda_fp16.h>
#include "cvt_sat.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "mma_sync.cuh"

namespace wdno {

namespace {

constexpr int kBM = 128, kBK = 32, kStages = 4, kThreads = 256;
constexpr int kPitch = kBK + 8;  // halfs per smem row (80 B)

struct C1Params {
  const __half* src0;
  const __half* src1;
  const __half* w;      // [Npad][K] fp16, K contiguous
  const float* bias;    // [Npad] or null
  const __half* resid;  // [M][Cout] or null (fp16 output only)
  void* out;
  long long M;
  int c0, c1, K, Cout, out_mode;  // out_mode 0: fp16 [M][Cout]; 2: fp32 [M/HW][Cout][HW]
  long long HW;
};

__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool ok) {
  const uint32_t n = ok ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kThreads, (BN == 128) ? 2 : 3) conv1x1_kernel(const C1Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int WN = BN / 2;          // warp tile: 32 rows x WN columns (warps: 4 along M x 2 along N)
  constexpr int NT = WN / 8;          // n-tiles per warp
  constexpr int A_BYTES = kBM * kPitch * 2, W_BYTES = BN * kPitch * 2, STAGE = A_BYTES + W_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;
  const long long m0 = static_cast<long long>(blockIdx.x) * kBM;
  const int n0 = blockIdx.y * BN;
  const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  pdl_trigger();
  pdl_wait();

  const int KT = p.K / kBK;
  // loader geometry: a thread copies chunk `lc` (16 B) of rows lr, lr + 64 of A and of rows lr (, lr + 64) of W
  const int lr = tid >> 2, lc = tid & 3;
  auto load = [&](int kt, int st) {
    const int k0 = kt * kBK;
    const bool first = k0 < p.c0;
    const __half* sp = first ? p.src0 : p.src1;
    const int cs = first ? p.c0 : p.c1;
    const int kc = (first ? k0 : k0 - p.c0) + lc * 8;
    const uint32_t sa = sbase + st * STAGE + static_cast<uint32_t>((lr * kPitch + lc * 8) * 2);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long long m = m0 + lr + 64 * i;
      const bool ok = m < p.M;
      cp16(sa + static_cast<uint32_t>(i * 64 * kPitch * 2), sp + (ok ? m : 0) * cs + kc, ok);
    }
    const uint32_t sw = sbase + st * STAGE + A_BYTES + static_cast<uint32_t>((lr * kPitch + lc * 8) * 2);
#pragma unroll
    for (int i = 0; i < BN / 64; ++i)
      cp16(sw + static_cast<uint32_t>(i * 64 * kPitch * 2), p.w + static_cast<size_t>(n0 + lr + 64 * i) * p.K + k0 + lc * 8, true);
  };

  float acc[2][NT][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < KT) load(s, s);
    cp_commit();
  }
  const uint32_t a_off = static_cast<uint32_t>(((wm * 32 + (lane & 15)) * kPitch + 8 * (lane >> 4)) * 2);
  const uint32_t b_off = static_cast<uint32_t>(((wn * WN + (lane & 7) + 8 * (lane >> 4)) * kPitch + 8 * ((lane >> 3) & 1)) * 2);
  for (int kt = 0; kt < KT; ++kt) {
    cp_wait<kStages - 2>();
    __syncthreads();
    if (kt + kStages - 1 < KT) load(kt + kStages - 1, (kt + kStages - 1) % kStages);
    cp_commit();
    const uint32_t sa = sbase + (kt % kStages) * STAGE + a_off;
    const uint32_t sw = sbase + (kt % kStages) * STAGE + A_BYTES + b_off;
#pragma unroll
    for (int ks = 0; ks < kBK / 16; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ldsm_x4(sa + static_cast<uint32_t>((mt * 16 * kPitch + ks * 16) * 2), a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sw + static_cast<uint32_t>((np * 16 * kPitch + ks * 16) * 2), b0, b1, b2, b3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(acc[mt][2 * np], a[mt], b0, b1);
          mma16816(acc[mt][2 * np + 1], a[mt], b2, b3);
        }
      }
    }
  }
  cp_wait<0>();
  __syncthreads();  // every warp is done with the operand ring: reuse it as the output staging tile

  const int g = lane >> 2, q = lane & 3;
  if (p.out_mode == 0) {
    // fp16 [128][BN + 8] staging, then 16-byte row-chunk stores (+ residual)
    constexpr int OP = BN + 8;
    __half* so = reinterpret_cast<__half*>(smem);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = wn * WN + nt * 8 + 2 * q;
        const float b0 = p.bias ? __ldg(p.bias + n0 + col) : 0.f, b1 = p.bias ? __ldg(p.bias + n0 + col + 1) : 0.f;
        const int r = wm * 32 + mt * 16 + g;
        *reinterpret_cast<uint32_t*>(so + r * OP + col) = pack_h2(acc[mt][nt][0] + b0, acc[mt][nt][1] + b1);
        *reinterpret_cast<uint32_t*>(so + (r + 8) * OP + col) = pack_h2(acc[mt][nt][2] + b0, acc[mt][nt][3] + b1);
      }
    __syncthreads();
    constexpr int CPR = BN / 8;  // 16-byte chunks per tile row
    __half* o16 = static_cast<__half*>(p.out);
    for (int idx = tid; idx < kBM * CPR; idx += kThreads) {
      const int r = idx / CPR, ch = idx - r * CPR;
      const long long m = m0 + r;
      const int col = n0 + ch * 8;
      if (m < p.M && col < p.Cout) {
        uint4 v = *reinterpret_cast<const uint4*>(so + r * OP + ch * 8);
        const size_t off = static_cast<size_t>(m) * p.Cout + col;
        if (p.resid != nullptr) {
          const uint4 rv = __ldg(reinterpret_cast<const uint4*>(p.resid + off));
          __half2* vh = reinterpret_cast<__half2*>(&v);
          const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 a2 = __half22float2(vh[i]), b2 = __half22float2(rh[i]);
            vh[i] = wdno::h2_sat(a2.x + b2.x, a2.y + b2.y);
          }
        }
        *reinterpret_cast<uint4*>(o16 + off) = v;
      }
    }
  } else {
    // fp32 planar output [M/HW][Cout][HW]: stage transposed [BN][128 + 1] floats, then position-contiguous stores
    constexpr int OP = kBM + 1;
    float* so = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int col = wn * WN + nt * 8 + 2 * q;
        const float b0 = p.bias ? __ldg(p.bias + n0 + col) : 0.f, b1 = p.bias ? __ldg(p.bias + n0 + col + 1) : 0.f;
        const int r = wm * 32 + mt * 16 + g;
        so[col * OP + r] = acc[mt][nt][0] + b0;
        so[(col + 1) * OP + r] = acc[mt][nt][1] + b1;
        so[col * OP + r + 8] = acc[mt][nt][2] + b0;
        so[(col + 1) * OP + r + 8] = acc[mt][nt][3] + b1;
      }
    __syncthreads();
    float* o32 = static_cast<float*>(p.out);
    const int r = tid & (kBM - 1);
    const long long m = m0 + r;
    if (m < p.M) {
      const long long plane = m / p.HW, pos = m - plane * p.HW;
      for (int c = tid >> 7; c < BN; c += kThreads / kBM) {
        const int col = n0 + c;
        if (col < p.Cout) o32[(static_cast<size_t>(plane) * p.Cout + col) * p.HW + pos] = so[c * OP + r];
      }
    }
  }
}

template <int BN>
int launch(const C1Params& p, int npad, cudaStream_t st) {
  constexpr int stage = (kBM + BN) * kPitch * 2;
  constexpr int out16 = kBM * (BN + 8) * 2, out32 = BN * (kBM + 1) * 4;
  constexpr int smem = (kStages * stage > out32) ? ((kStages * stage > out16) ? kStages * stage : out16) : out32;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv1x1_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "conv1x1: cudaFuncSetAttribute");
    configured = true;
  }
  const long long mt = (p.M + kBM - 1) / kBM;
  if (mt > 2147483647LL) return set_error(WDNO_E_INVALID, "conv1x1: too many position tiles");
  dim3 grid(static_cast<unsigned>(mt), static_cast<unsigned>(npad / BN));
  cudaError_t le = launch_pdl(conv1x1_kernel<BN>, grid, dim3(kThreads), static_cast<size_t>(smem), st, p);
  if (le != cudaSuccess) return set_cuda_error(le, "conv1x1: launch");
  return check_launch("conv1x1");
}

}  // namespace

}  // namespace wdno

extern "C" int wdno_conv1x1(const void* src0, int c0, const void* src1, int c1, const void* w, int npad, const float* bias,
                            const void* resid, void* out, int64_t M, int cout, int out_mode, int64_t hw, void* stream) {
  using namespace wdno;
  if (!src0 || !w || !out || M < 1 || cout < 1) return set_error(WDNO_E_INVALID, "conv1x1: bad arguments");
  if (c0 < 32 || (c0 % 32) || c1 < 0 || (c1 % 32) || (c1 > 0 && !src1))
    return set_error(WDNO_E_INVALID, "conv1x1: source channels must be multiples of 32");
  if (npad < cout || (npad % 64)) return set_error(WDNO_E_INVALID, "conv1x1: npad must be a multiple of 64 covering cout");
  if (out_mode != 0 && out_mode != 2) return set_error(WDNO_E_INVALID, "conv1x1: out_mode must be 0 (fp16) or 2 (fp32 planar)");
  if (out_mode == 0 && (cout % 8)) return set_error(WDNO_E_INVALID, "conv1x1: fp16 output channels must be a multiple of 8");
  if (out_mode == 2 && (hw < 1 || resid)) return set_error(WDNO_E_INVALID, "conv1x1: planar output needs hw >= 1 and no residual");
  C1Params p;
  p.src0 = static_cast<const __half*>(src0);
  p.src1 = static_cast<const __half*>(src1);
  p.w = static_cast<const __half*>(w);
  p.bias = bias;
  p.resid = static_cast<const __half*>(resid);
  p.out = out;
  p.M = M;
  p.c0 = c0;
  p.c1 = c1;
  p.K = c0 + c1;
  p.Cout = cout;
  p.out_mode = out_mode;
  p.HW = hw > 0 ? hw : 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (npad % 128 == 0) return launch<128>(p, npad, st);
  return launch<64>(p, npad, st);
}
