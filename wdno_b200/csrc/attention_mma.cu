// Short-sequence softmax attention on tensor cores (n <= 32 tokens, 4 heads x 32): the temporal attention of the smoke
// U-Net (n = 24 frames per pixel; rotary on q,k; T5 relative-position bias).  reference: conv3d.py:277-353, 383.
// One block = 4 warps = the 4 heads of one sequence; a sequence's q|k|v rows (n x 384 fp16) are staged once into shared
// memory with coalesced 16-byte loads (scale + rotary applied on the way), S = QK^T and O = PV run as warp-level
// mma.sync m16n8k16 (fp16 in, fp32 accumulate) with ldmatrix fragments, softmax stays in registers (FA-2 layout trick:
// the S accumulator fragment is the A fragment of P), O goes back through shared memory for coalesced stores.
// 2 x 16 MMAs per (sequence, head) -- not worth a tcgen05/TMEM pipeline; HBM traffic (read 768 B, write 256 B per token)
// is what bounds it.
#include <cuda_fp16.h>
#include "cvt_sat.cuh"
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace wdno {

constexpr int kH = 4, kD = 32, kHid = 128, kQkv = 384;
constexpr int kRow = 40;  // halves per shared-memory row: 32 + 8 pad (80 B keeps 16-byte alignment, breaks bank conflicts)

struct SeqMap2 {
  long long inner, outerT, innerT, tokT;
};

__device__ __forceinline__ long long tok_of(const SeqMap2& m, long long s, int t) {
  return (s / m.inner) * m.outerT + (s % m.inner) * m.innerT + static_cast<long long>(t) * m.tokT;
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float x, float y) {
  const __half2 h = wdno::h2_sat(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(128) short_attn_mma_kernel(const __half* __restrict__ qkv, __half* __restrict__ out,
                                                            const float* __restrict__ bias, const float* __restrict__ rot_cos,
                                                            const float* __restrict__ rot_sin, SeqMap2 map, long long n_seq,
                                                            int n, float scale) {
  __shared__ __align__(16) __half sq[kH][32][kRow];
  __shared__ __align__(16) __half sk[kH][32][kRow];
  __shared__ __align__(16) __half sv[kH][32][kRow];
  const int tid = threadIdx.x, lane = tid & 31, h = tid >> 5;
  const int g = lane >> 2, q4 = lane & 3;

  // k / v rows >= n stay zero for the whole kernel (padded keys are masked with -inf and must not inject NaNs);
  // q rows >= n only feed S rows that are never stored
  for (int i = tid; i < kH * 32 * kRow; i += 128) {
    (&sq[0][0][0])[i] = __float2half(0.f);
    (&sk[0][0][0])[i] = __float2half(0.f);
    (&sv[0][0][0])[i] = __float2half(0.f);
  }
  // additive bias / key mask of this thread's S-fragment positions: rows 16mt+g+8r, cols 8nt+2q+e
  float bfrag[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = 16 * mt + g + 8 * (c >> 1), j = 8 * nt + 2 * q4 + (c & 1);
        float b = 0.f;
        if (j >= n) b = -INFINITY;
        else if (i < n && bias != nullptr) b = __ldg(bias + (static_cast<size_t>(h) * n + i) * n + j);
        bfrag[mt][nt][c] = b;
      }
  __syncthreads();

  const uint32_t q_base = static_cast<uint32_t>(__cvta_generic_to_shared(&sq[h][0][0]));
  const uint32_t k_base = static_cast<uint32_t>(__cvta_generic_to_shared(&sk[h][0][0]));
  const uint32_t v_base = static_cast<uint32_t>(__cvta_generic_to_shared(&sv[h][0][0]));
  // per-lane ldmatrix row addresses (bytes)
  const uint32_t a_off = static_cast<uint32_t>((((lane & 7) + 8 * ((lane >> 3) & 1)) * kRow + 8 * (lane >> 4)) * 2);
  const uint32_t b_off = static_cast<uint32_t>(((lane & 7) * kRow + 8 * ((lane >> 3) & 1)) * 2);
  const uint32_t v_off = static_cast<uint32_t>((((lane & 7) + 8 * ((lane >> 3) & 1)) * kRow) * 2);

  for (long long s = blockIdx.x; s < n_seq; s += gridDim.x) {
    // ---- stage q (scaled, rotated), k (rotated), v of every token: one 16-byte load per thread-item, coalesced rows
    for (int idx = tid; idx < n * (kQkv / 8); idx += 128) {
      const int f = idx / (kQkv / 8), c8 = idx - f * (kQkv / 8);
      const int col = c8 * 8, sect = col >> 7, hh = (col & 127) >> 5, d0 = col & 31;
      uint4 v = __ldg(reinterpret_cast<const uint4*>(qkv + tok_of(map, s, f) * kQkv + col));
      if (sect < 2) {
        __half2* hp = reinterpret_cast<__half2*>(&v);
        const float sc = (sect == 0) ? scale : 1.0f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 x = __half22float2(hp[e]);
          x.x *= sc;
          x.y *= sc;
          if (rot_cos != nullptr) {
            const int m = (d0 >> 1) + e;
            const float cs = __ldg(rot_cos + f * 16 + m), sn = __ldg(rot_sin + f * 16 + m);
            const float x0 = x.x, x1 = x.y;
            x.x = x0 * cs - x1 * sn;
            x.y = x1 * cs + x0 * sn;
          }
          hp[e] = wdno::h2_sat(x);
        }
      }
      __half* dst = (sect == 0) ? &sq[hh][f][d0] : (sect == 1) ? &sk[hh][f][d0] : &sv[hh][f][d0];
      *reinterpret_cast<uint4*>(dst) = v;
    }
    __syncthreads();

    // ---- S = Q K^T  (32 x 32, K = 32)
    float sfr[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) sfr[mt][nt][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ldsm_x4(q_base + a_off + static_cast<uint32_t>((mt * 16 * kRow + ks * 16) * 2), a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t b0, b1;
        ldsm_x2(k_base + b_off + static_cast<uint32_t>((nt * 8 * kRow + ks * 16) * 2), b0, b1);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma16816(sfr[mt][nt], a[mt], b0, b1);
      }
    }
    // ---- softmax over keys (row = 16mt + g + 8r lives in one quad)
    uint32_t pa[2][2][4];
    float inv[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          sfr[mt][nt][2 * r] += bfrag[mt][nt][2 * r];
          sfr[mt][nt][2 * r + 1] += bfrag[mt][nt][2 * r + 1];
          mx = fmaxf(mx, fmaxf(sfr[mt][nt][2 * r], sfr[mt][nt][2 * r + 1]));
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float p0 = __expf(sfr[mt][nt][2 * r] - mx), p1 = __expf(sfr[mt][nt][2 * r + 1] - mx);
          sfr[mt][nt][2 * r] = p0;
          sfr[mt][nt][2 * r + 1] = p1;
          sum += p0 + p1;
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        inv[mt][r] = 1.0f / sum;
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        pa[mt][ks][0] = pack_h2(sfr[mt][2 * ks][0], sfr[mt][2 * ks][1]);
        pa[mt][ks][1] = pack_h2(sfr[mt][2 * ks][2], sfr[mt][2 * ks][3]);
        pa[mt][ks][2] = pack_h2(sfr[mt][2 * ks + 1][0], sfr[mt][2 * ks + 1][1]);
        pa[mt][ks][3] = pack_h2(sfr[mt][2 * ks + 1][2], sfr[mt][2 * ks + 1][3]);
      }
    }
    // ---- O = P V  (32 x 32, K = 32 keys)
    float ofr[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) ofr[mt][nt][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t b0, b1;
        ldsm_x2_trans(v_base + v_off + static_cast<uint32_t>((ks * 16 * kRow + nt * 8) * 2), b0, b1);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma16816(ofr[mt][nt], pa[mt][ks], b0, b1);
      }
    // ---- O (normalised, fp16) -> this head's q region -> coalesced store
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int i = 16 * mt + g + 8 * r;
          *reinterpret_cast<uint32_t*>(&sq[h][i][8 * nt + 2 * q4]) =
              pack_h2(ofr[mt][nt][2 * r] * inv[mt][r], ofr[mt][nt][2 * r + 1] * inv[mt][r]);
        }
    __syncthreads();
    for (int idx = tid; idx < n * (kHid / 8); idx += 128) {
      const int f = idx >> 4, c = idx & 15;
      const uint4 v = *reinterpret_cast<const uint4*>(&sq[c >> 2][f][(c & 3) * 8]);
      *reinterpret_cast<uint4*>(out + tok_of(map, s, f) * kHid + c * 8) = v;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ medium sequences (32 < n <= 512), no bias / rotary
// The mid-level spatial attention (conv3d.py:450-452: n = H*W = 100 tokens at the base resolution, 400 in the
// super-resolution model) and the Burgers mid attention (unet.py:225-261, n = 64).  One block = one (sequence, head):
// q (scaled), k, v of the head are staged once in shared memory; each of the 4 warps takes 16-query tiles and runs an
// online-softmax loop over 32-key chunks -- S = Q K^T (8 MMAs), P V (8 MMAs), the S accumulator fragments being the A
// fragments of P.  The scalar fp32 kernel it replaces took 0.14 ms (n = 100) / 1.9 ms (n = 400) per step.
__global__ void __launch_bounds__(128) flash_attn_mma_kernel(const __half* __restrict__ qkv, __half* __restrict__ out,
                                                            SeqMap2 map, int n, int npad, float scale) {
  extern __shared__ __align__(16) __half fsm[];
  __half* sq = fsm;                       // [npad][kRow]
  __half* sk = sq + static_cast<size_t>(npad) * kRow;
  __half* sv = sk + static_cast<size_t>(npad) * kRow;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q4 = lane & 3;
  const long long s = blockIdx.x >> 2;
  const int h = blockIdx.x & 3;
  // ---- stage: item = (token, section, 16-byte chunk); rows >= n are zero
  for (int idx = tid; idx < npad * 12; idx += 128) {
    const int f = idx / 12, r = idx - f * 12, sect = r >> 2, c = r & 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (f < n) {
      v = __ldg(reinterpret_cast<const uint4*>(qkv + tok_of(map, s, f) * kQkv + sect * kHid + h * kD + c * 8));
      if (sect == 0) {
        __half2* hp = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 x = __half22float2(hp[e]);
          x.x *= scale;
          x.y *= scale;
          hp[e] = wdno::h2_sat(x);
        }
      }
    }
    __half* dst = ((sect == 0) ? sq : (sect == 1) ? sk : sv) + static_cast<size_t>(f) * kRow + c * 8;
    *reinterpret_cast<uint4*>(dst) = v;
  }
  __syncthreads();
  const uint32_t q_base = static_cast<uint32_t>(__cvta_generic_to_shared(sq));
  const uint32_t k_base = static_cast<uint32_t>(__cvta_generic_to_shared(sk));
  const uint32_t v_base = static_cast<uint32_t>(__cvta_generic_to_shared(sv));
  const uint32_t a_off = static_cast<uint32_t>((((lane & 7) + 8 * ((lane >> 3) & 1)) * kRow + 8 * (lane >> 4)) * 2);
  const uint32_t b_off = static_cast<uint32_t>(((lane & 7) * kRow + 8 * ((lane >> 3) & 1)) * 2);
  const uint32_t v_off = static_cast<uint32_t>((((lane & 7) + 8 * ((lane >> 3) & 1)) * kRow) * 2);
  const int mtiles = (n + 15) >> 4;
  for (int mt = warp; mt < mtiles; mt += 4) {
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      ldsm_x4(q_base + a_off + static_cast<uint32_t>((mt * 16 * kRow + ks * 16) * 2), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    float ofr[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) ofr[nt][c] = 0.f;
    float mrun[2] = {-INFINITY, -INFINITY}, lrun[2] = {0.f, 0.f};
    for (int j0 = 0; j0 < n; j0 += 32) {
      float sfr[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sfr[nt][c] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t b0, b1;
          ldsm_x2(k_base + b_off + static_cast<uint32_t>(((j0 + nt * 8) * kRow + ks * 16) * 2), b0, b1);
          mma16816(sfr[nt], qa[ks], b0, b1);
        }
      }
      uint32_t pa[2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float mx = mrun[r];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = j0 + 8 * nt + 2 * q4 + e;
            if (j >= n) sfr[nt][2 * r + e] = -INFINITY;
            mx = fmaxf(mx, sfr[nt][2 * r + e]);
          }
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float corr = __expf(mrun[r] - mx);  // first chunk: exp(-inf) = 0 and the accumulators are zero
        mrun[r] = mx;
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float p0 = __expf(sfr[nt][2 * r] - mx), p1 = __expf(sfr[nt][2 * r + 1] - mx);
          sfr[nt][2 * r] = p0;
          sfr[nt][2 * r + 1] = p1;
          sum += p0 + p1;
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        lrun[r] = lrun[r] * corr + sum;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          ofr[nt][2 * r] *= corr;
          ofr[nt][2 * r + 1] *= corr;
        }
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        pa[ks][0] = pack_h2(sfr[2 * ks][0], sfr[2 * ks][1]);
        pa[ks][1] = pack_h2(sfr[2 * ks][2], sfr[2 * ks][3]);
        pa[ks][2] = pack_h2(sfr[2 * ks + 1][0], sfr[2 * ks + 1][1]);
        pa[ks][3] = pack_h2(sfr[2 * ks + 1][2], sfr[2 * ks + 1][3]);
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          uint32_t b0, b1;
          ldsm_x2_trans(v_base + v_off + static_cast<uint32_t>(((j0 + ks * 16) * kRow + nt * 8) * 2), b0, b1);
          mma16816(ofr[nt], pa[ks], b0, b1);
        }
    }
    // ---- normalised O -> the tile's own q rows (dead) -> 64-byte row stores
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float inv = 1.0f / lrun[r];
      const int i = 16 * mt + g + 8 * r;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        *reinterpret_cast<uint32_t*>(sq + static_cast<size_t>(i) * kRow + 8 * nt + 2 * q4) = pack_h2(ofr[nt][2 * r] * inv, ofr[nt][2 * r + 1] * inv);
    }
    __syncwarp();
    for (int idx = lane; idx < 16 * 4; idx += 32) {
      const int f = 16 * mt + (idx >> 2), c = idx & 3;
      if (f < n)
        *reinterpret_cast<uint4*>(out + tok_of(map, s, f) * kHid + h * kD + c * 8) =
            *reinterpret_cast<const uint4*>(sq + static_cast<size_t>(f) * kRow + c * 8);
    }
  }
}

int launch_flash_attn_mma(const void* qkv, void* out, long long n_seq, int n_tok, long long inner, long long outerT,
                          long long innerT, long long tokT, float scale, cudaStream_t st) {
  SeqMap2 m{inner, outerT, innerT, tokT};
  const int npad = (n_tok + 31) / 32 * 32;
  const size_t smem = static_cast<size_t>(3) * npad * kRow * sizeof(__half);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e, "flash_attn_mma: cudaFuncSetAttribute");
    configured = smem;
  }
  if (n_seq * 4 > 2147483647LL) return set_error(WDNO_E_INVALID, "flash_attn_mma: too many sequences");
  flash_attn_mma_kernel<<<static_cast<unsigned>(n_seq * 4), 128, smem, st>>>(static_cast<const __half*>(qkv), static_cast<__half*>(out),
                                                                          m, n_tok, npad, scale);
  return check_launch("flash_attn_mma");
}

int launch_short_attn_mma(const void* qkv, void* out, const float* bias, const float* rot_cos, const float* rot_sin,
                          long long n_seq, int n_tok, long long inner, long long outerT, long long innerT, long long tokT,
                          float scale, cudaStream_t st) {
  SeqMap2 m{inner, outerT, innerT, tokT};
  const long long cap = static_cast<long long>(num_sms()) * 7;
  const unsigned grid = static_cast<unsigned>(n_seq < cap ? n_seq : cap);
  short_attn_mma_kernel<<<grid, 128, 0, st>>>(static_cast<const __half*>(qkv), static_cast<__half*>(out), bias, rot_cos, rot_sin,
                                              m, n_seq, n_tok, scale);
  return check_launch("short_attn_mma");
}

}  // namespace wdno
