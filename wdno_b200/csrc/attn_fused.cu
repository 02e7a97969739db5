// Fused attention blocks of the smoke U-Net: Residual(PreNorm(dim, attention)) in ONE pass over the residual stream.
//
//   linattn_block  -- SpatialLinearAttention block (reference conv3d.py:165-184, 232-258):
//        y = x + to_out( ctx^T softmax_d(q) * scale ) ,  ctx = softmax_n(k) v^T ,  (q,k,v) = to_qkv(LayerNorm(x))
//      la1: per image, per tile of 64 positions: LayerNorm -> k^T, v^T (tensor cores, weights as the A operand so the
//           results are already the fragments of the next product) -> exp(k - running max) -> S += ek v^T ; one partial
//           (max, sum, S) per (image, part, head) goes to a small workspace -- k and v never touch HBM;
//      la_mid: per image: merge the partials, ctx = S / z, and fold to_out into it:  M = scale * W_out blockdiag(ctx^T)
//           (C x 128), written in B-fragment order;
//      la2: per tile of 128 positions: LayerNorm -> q (tensor cores) -> softmax over each head's 32 dims in registers
//           -> y = q M^T + bias + x -> coalesced store.
//   HBM traffic: x is read twice and y written once (3 x 2C bytes per voxel) instead of the ~21 x 2C bytes of the
//   unfused LN / qkv GEMM / attention core / out GEMM chain.
#include <cuda_fp16.h>
#include "cvt_sat.cuh"
#include <stdlib.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"
#include "mma_sync.cuh"

namespace wdno {

constexpr int kLaHeads = 4, kLaDh = 32, kLaHid = 128;

// weight fragments: shared memory (staged once per block) when WS, else read-only global loads
template <bool WS>
__device__ __forceinline__ uint4 ldw(const uint4* p) {
  if constexpr (WS) return *p;
  else return __ldg(p);
}
__device__ __forceinline__ void stage_weights(uint4* dst, const uint4* __restrict__ src, int n_u4, int tid, int nthreads) {
  for (int i = tid; i < n_u4; i += nthreads) dst[i] = __ldg(src + i);
}
constexpr int kLaPart = 64 + 32 * 32;  // floats per (image, part, head): max[32], sum[32], S[32][32]

// ------------------------------------------------------------------ la1: context partials
// NPH = position halves per block: 2 -> 8 warps share a 64-position LayerNorm tile (two block barriers per tile);
// 1 -> 4 warps (one per head) per 32-position tile, twice as many blocks resident per SM -- the barriers couple fewer warps
template <int C, bool WS, int NPH>
__global__ void __launch_bounds__(128 * NPH, (NPH == 2) ? 2 : 4) la1_kernel(const __half* __restrict__ x, const float* __restrict__ gamma,
                                                                         const uint4* wkv, float* __restrict__ part, int n, int split,
                                                                         float eps) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  pdl_trigger();
  constexpr int TP = 32 * NPH, NT_ = 128 * NPH;   // positions per tile, threads per block
  __half* xn = reinterpret_cast<__half*>(smem_raw);  // [TP][C + 8]
  constexpr int XS = C + 8;
  constexpr int KS = C / 16;
  const int img = blockIdx.x / split, sp = blockIdx.x - img * split;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int h = warp & 3, ph = warp >> 2;
  const int tiles = (n + TP - 1) / TP;
  const int t0 = (tiles * sp) / split, t1 = (tiles * (sp + 1)) / split;
  if constexpr (WS) {
    uint4* wsm = reinterpret_cast<uint4*>(smem_raw + TP * XS * 2);
    stage_weights(wsm, wkv, 2 * 4 * 2 * KS * 32, threadIdx.x, NT_);
    wkv = wsm;  // visible after the first barrier of the tile loop
  }
  const uint4* wk = wkv + static_cast<size_t>((0 * 4 + h) * 2) * KS * 32 + lane;
  const uint4* wv = wkv + static_cast<size_t>((1 * 4 + h) * 2) * KS * 32 + lane;
  const __half* ximg = x + static_cast<size_t>(img) * n * C;
  const uint32_t xn_s = static_cast<uint32_t>(__cvta_generic_to_shared(xn));
  // ldmatrix lane address of the B operand (activations): row = position, 8-column blocks along channels
  const uint32_t b_off = static_cast<uint32_t>(((32 * ph + (lane & 7)) * XS + 8 * (lane >> 3)) * 2);

  float S[2][4][4];
  float mrun[4], z[4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) S[i][j][c] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { mrun[i] = -INFINITY; z[i] = 0.f; }

  pdl_wait();  // weights above are plan constants; x is the predecessor's output
  for (int t = t0; t < t1; ++t) {
    __syncthreads();
    const int p0 = t * TP;
    ln_tile<C, TP, NT_>(ximg + static_cast<size_t>(p0) * C, min(TP, n - p0), gamma, xn, eps);
    __syncthreads();
    // ---- k^T[d][p] = Wk_h xn^T
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
#pragma unroll
    for (int kp = 0; kp < C / 32; ++kp) {
      uint32_t bf[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        ldsm_x4(xn_s + b_off + static_cast<uint32_t>((nt * 8 * XS + kp * 32) * 2), bf[nt][0], bf[nt][1], bf[nt][2], bf[nt][3]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint4 a0 = ldw<WS>(wk + (mt * KS + 2 * kp) * 32), a1 = ldw<WS>(wk + (mt * KS + 2 * kp + 1) * 32);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma16816(acc[mt][nt], a0.x, a0.y, a0.z, a0.w, bf[nt][0], bf[nt][1]);
          mma16816(acc[mt][nt], a1.x, a1.y, a1.z, a1.w, bf[nt][2], bf[nt][3]);
        }
      }
    }
    // ---- running column max (over positions) per row d, rescale, ek = exp(k - max)
    const int pbase = p0 + 32 * ph + 2 * q;
    uint32_t ekA[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (pbase + 8 * nt + e < n) mx = fmaxf(mx, acc[mt][nt][2 * r + e]);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const int ri = mt * 2 + r;
        const float mnew = fmaxf(mrun[ri], mx);
        const float sc = (mnew == -INFINITY) ? 1.0f : __expf(mrun[ri] - mnew);
        mrun[ri] = mnew;
        z[ri] *= sc;
#pragma unroll
        for (int ne = 0; ne < 4; ++ne) {
          S[mt][ne][2 * r] *= sc;
          S[mt][ne][2 * r + 1] *= sc;
        }
        float zs = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float v = (pbase + 8 * nt + e < n) ? __expf(acc[mt][nt][2 * r + e] - mnew) : 0.f;
            acc[mt][nt][2 * r + e] = v;
            zs += v;
          }
        z[ri] += zs;
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        ekA[mt][kk][0] = pack_h2(acc[mt][2 * kk][0], acc[mt][2 * kk][1]);
        ekA[mt][kk][1] = pack_h2(acc[mt][2 * kk][2], acc[mt][2 * kk][3]);
        ekA[mt][kk][2] = pack_h2(acc[mt][2 * kk + 1][0], acc[mt][2 * kk + 1][1]);
        ekA[mt][kk][3] = pack_h2(acc[mt][2 * kk + 1][2], acc[mt][2 * kk + 1][3]);
      }
    }
    // ---- v^T[e][p] = Wv_h xn^T
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
#pragma unroll
    for (int kp = 0; kp < C / 32; ++kp) {
      uint32_t bf[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        ldsm_x4(xn_s + b_off + static_cast<uint32_t>((nt * 8 * XS + kp * 32) * 2), bf[nt][0], bf[nt][1], bf[nt][2], bf[nt][3]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint4 a0 = ldw<WS>(wv + (mt * KS + 2 * kp) * 32), a1 = ldw<WS>(wv + (mt * KS + 2 * kp + 1) * 32);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          mma16816(acc[mt][nt], a0.x, a0.y, a0.z, a0.w, bf[nt][0], bf[nt][1]);
          mma16816(acc[mt][nt], a1.x, a1.y, a1.z, a1.w, bf[nt][2], bf[nt][3]);
        }
      }
    }
    // ---- S[d][e] += sum_p ek[d][p] v^T[e][p]   (v^T accumulator fragments are the B fragments)
#pragma unroll
    for (int ne = 0; ne < 4; ++ne) {
      const int me = ne >> 1, rr = 2 * (ne & 1);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint32_t b0 = pack_h2(acc[me][2 * kk][rr], acc[me][2 * kk][rr + 1]);
        const uint32_t b1 = pack_h2(acc[me][2 * kk + 1][rr], acc[me][2 * kk + 1][rr + 1]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma16816(S[mt][ne], ekA[mt][kk], b0, b1);
      }
    }
  }
  // ---- partial of this (image, part = 2*sp + ph, head)
  float* dst = part + (static_cast<size_t>(img * (split * NPH) + sp * NPH + ph) * kLaHeads + h) * kLaPart;
#pragma unroll
  for (int ri = 0; ri < 4; ++ri) {
    float zz = z[ri];
    zz += __shfl_xor_sync(0xffffffffu, zz, 1);
    zz += __shfl_xor_sync(0xffffffffu, zz, 2);
    const int d = (ri >> 1) * 16 + g + 8 * (ri & 1);
    if (q == 0) {
      dst[d] = mrun[ri];
      dst[32 + d] = zz;
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ne = 0; ne < 4; ++ne)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int d = mt * 16 + g + 8 * r;
        *reinterpret_cast<float2*>(dst + 64 + d * 32 + ne * 8 + 2 * q) = make_float2(S[mt][ne][2 * r], S[mt][ne][2 * r + 1]);
      }
}

// ------------------------------------------------------------------ la_mid: merge partials, fold to_out into the context
// mpack[img]: M[c][j] = scale * sum_e Wout[c][32h + e] ctx_h[d][e]  (j = 32h + d) in B-fragment order
// [c/8][j/32][lane][8 halves] (see mma_sync.cuh).
__global__ void __launch_bounds__(256) la_mid_kernel(const float* __restrict__ part, const float* __restrict__ wout,
                                                     __half* __restrict__ mpack, int C, int nparts, float scale, int canon) {
  __shared__ float ctx[kLaHeads][kLaDh][kLaDh + 1];
  __shared__ float wts[4][kLaHid];  // per part: exp(m_i - m) * scale / z  for (h, d)
  pdl_trigger();
  pdl_wait();
  const int img = blockIdx.x;
  const float* pimg = part + static_cast<size_t>(img) * nparts * kLaHeads * kLaPart;
  if (threadIdx.x < kLaHid) {
    const int h = threadIdx.x >> 5, d = threadIdx.x & 31;
    const float* pb = pimg + static_cast<size_t>(h) * kLaPart;
    float mi[4], zi[4];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mi[i] = -INFINITY;
      zi[i] = 0.f;
      if (i < nparts) {
        mi[i] = pb[static_cast<size_t>(i) * kLaHeads * kLaPart + d];
        zi[i] = pb[static_cast<size_t>(i) * kLaHeads * kLaPart + 32 + d];
      }
      m = fmaxf(m, mi[i]);
    }
    float zt = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mi[i] = (i < nparts) ? __expf(mi[i] - m) : 0.f;
      zt = fmaf(zi[i], mi[i], zt);
    }
    const float inv = scale / zt;
#pragma unroll
    for (int i = 0; i < 4; ++i) wts[i][threadIdx.x] = mi[i] * inv;
  }
  __syncthreads();
  // ctx[h][d][e] = sum_i S_i[h][d][e] * w_i[h][d]   (coalesced over e)
  // all loads of four outputs are issued before the first use (the loop was one dependent global load per FMA: 45 % of the
  // kernel's stall samples sat on its first FFMA)
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) {
    float pv[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = threadIdx.x + 256 * (k4 * 4 + k);
      const int h = idx >> 10, de = idx & 1023;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        pv[k][i] = (i < nparts) ? __ldg(pimg + (static_cast<size_t>(i) * kLaHeads + h) * kLaPart + 64 + de) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = threadIdx.x + 256 * (k4 * 4 + k);
      const int h = idx >> 10, d = (idx >> 5) & 31, e = idx & 31;
      float sacc = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) sacc = fmaf(pv[k][i], wts[i][h * 32 + d], sacc);
      ctx[h][d][e] = sacc;
    }
  }
  __syncthreads();
  const int t = threadIdx.x & 127, chalf = threadIdx.x >> 7;
  const int h = t >> 5, d = t & 31;
  float cr[kLaDh];
#pragma unroll
  for (int e = 0; e < kLaDh; ++e) cr[e] = ctx[h][d][e];
  const int j = h * 32 + d;
  const int ks = j >> 4, kk = j & 15;
  const int reg = kk >> 3, qq = (kk & 7) >> 1, half = kk & 1;
  extern __shared__ __align__(16) __half msm[];  // the image's M in fragment order, then copied out coalesced
  const int c0 = chalf * (C / 2);
  // W_out rows reach the lanes through a per-warp staging tile: the warp (fixed head h) loads the 32 weights of 8 channels
  // with coalesced 128-byte requests (lane = e), every lane then reads them back as broadcast LDS.128 -- the previous
  // version issued 8 dependent uniform-address LDG.128 per channel and was latency bound (35..94 us per launch).
  __shared__ __align__(16) float wstage[8][8][32];
  float(*ws)[32] = wstage[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const float* wbase = wout + static_cast<size_t>(c0) * kLaHid + h * 32 + lane;
  float nxt[8];
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) nxt[cc] = __ldg(wbase + static_cast<size_t>(cc) * kLaHid);
  for (int cg = 0; cg < C / 2; cg += 8) {
    __syncwarp();
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) ws[cc][lane] = nxt[cc];
    __syncwarp();
    if (cg + 8 < C / 2) {
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) nxt[cc] = __ldg(wbase + static_cast<size_t>(cg + 8 + cc) * kLaHid);
    }
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      const int c = c0 + cg + cc;
      const float4* wr = reinterpret_cast<const float4*>(ws[cc]);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int e4 = 0; e4 < 8; ++e4) {
        const float4 w4 = wr[e4];
        s0 = fmaf(w4.x, cr[4 * e4], s0);
        s1 = fmaf(w4.y, cr[4 * e4 + 1], s1);
        s2 = fmaf(w4.z, cr[4 * e4 + 2], s2);
        s3 = fmaf(w4.w, cr[4 * e4 + 3], s3);
      }
      const int nt = c >> 3, gg = c & 7;
      const __half mv = wdno::h_sat((s0 + s1) + (s2 + s3));
      if (canon) msm[((j >> 3) * C + c) * 8 + (j & 7)] = mv;   // UMMA K-major B operand [128/8][C][8] (linattn_tc.cu)
      else msm[((nt * 4 + (ks >> 1)) * 32 + (gg * 4 + qq)) * 8 + ((ks & 1) * 2 + reg) * 2 + half] = mv;
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(mpack + static_cast<size_t>(img) * C * kLaHid);
  for (int i = threadIdx.x; i < C * kLaHid / 8; i += 256) dst[i] = reinterpret_cast<const uint4*>(msm)[i];
}

// ------------------------------------------------------------------ la2: q -> softmax_d -> y = q M^T + bias + x
// Persistent blocks walk over (image, 128-position tile) items; the raw x tile of the NEXT item is fetched with cp.async
// into the other half of a double buffer while the current one is normalised and multiplied, and also serves the
// residual add (x is read from HBM exactly once here).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int C, bool WS>
__global__ void __launch_bounds__(256, 2) la2_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                  const float* __restrict__ gamma, const uint4* wq,
                                                  const uint4* __restrict__ mpack, const float* __restrict__ bias, int n,
                                                  int tiles, int n_items, float eps) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int XS = C + 8;
  constexpr int LP = C / 8, RPP = 256 / LP, PASSES = 128 / RPP;
  __half* xn = reinterpret_cast<__half*>(smem_raw);             // [128][C + 8]
  __half* raw = xn + 128 * XS;                                  // [2][128][C]
  uint4* wsm = reinterpret_cast<uint4*>(raw + 2 * 128 * C);     // WS: Wq fragments, then the image's M fragments
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int ln_l = threadIdx.x % LP, ln_r0 = threadIdx.x / LP;
  constexpr int n_wq = 16 * (C / 32) * 32, n_m = (C / 8) * 4 * 32;
  pdl_trigger();
  if constexpr (WS) {
    stage_weights(wsm, wq, n_wq, threadIdx.x, 256);
    wq = wsm;
  }
  float gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gm[j] = __ldg(gamma + ln_l * 8 + j);
  const uint32_t raw_s = static_cast<uint32_t>(__cvta_generic_to_shared(raw));
  auto fetch = [&](int item, int buf) {
    const int img = item / tiles, tile = item - img * tiles;
    const int p0 = tile * 128;
    const int rows_valid = min(128, n - p0);
    const __half* xt = x + (static_cast<size_t>(img) * n + p0) * C;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int r = ln_r0 + ps * RPP;
      const bool ok = r < rows_valid;
      cp_async16(raw_s + static_cast<uint32_t>(((buf * 128 + r) * C + ln_l * 8) * 2), ok ? xt + static_cast<size_t>(r) * C + ln_l * 8 : xt,
                 ok ? 16u : 0u);
    }
    cp_async_commit_group();
  };
  const uint32_t xn_s = static_cast<uint32_t>(__cvta_generic_to_shared(xn));
  const uint32_t a_off = static_cast<uint32_t>(((16 * warp + (lane & 15)) * XS + 8 * (lane >> 4)) * 2);
  int cur_img = -1;
  int buf = 0;
  pdl_wait();  // Wq / gamma are plan constants; x and the per-image M come from predecessors
  if (static_cast<int>(blockIdx.x) < n_items) fetch(blockIdx.x, 0);

  for (int item = blockIdx.x; item < n_items; item += gridDim.x, buf ^= 1) {
    const int img = item / tiles, tile = item - img * tiles;
    const int p0 = tile * 128;
    const int rows_valid = min(128, n - p0);
    __half* yt = y + (static_cast<size_t>(img) * n + p0) * C;
    const __half* rawb = raw + buf * 128 * C;
    cp_async_wait_group<0>();
    __syncthreads();  // raw[buf] landed for everyone; all warps are done with the previous item (xn, raw[buf^1], M)
    if (item + static_cast<int>(gridDim.x) < n_items) fetch(item + gridDim.x, buf ^ 1);
    const uint4* mp_img = mpack + static_cast<size_t>(img) * n_m;
    if constexpr (WS) {
      if (img != cur_img) stage_weights(wsm + n_wq, mp_img, n_m, threadIdx.x, 256);
      mp_img = wsm + n_wq;
    }
    cur_img = img;
    // ---- LayerNorm raw[buf] -> xn
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int r = ln_r0 + ps * RPP;
      const uint4 rv = *reinterpret_cast<const uint4*>(rawb + r * C + ln_l * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&rv);
      float f[8];
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(h[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
        sum += t.x + t.y;
      }
#pragma unroll
      for (int sh = LP / 2; sh > 0; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
      const float mean = sum * (1.0f / C);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] -= mean;
        sq = fmaf(f[j], f[j], sq);
      }
#pragma unroll
      for (int sh = LP / 2; sh > 0; sh >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, sh);
      const float rstd = rsqrtf(sq * (1.0f / C) + eps);
      uint4 ov;
      uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = pack_h2(f[2 * j] * rstd * gm[2 * j], f[2 * j + 1] * rstd * gm[2 * j + 1]);
      if (r >= rows_valid) ov = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(xn + r * XS + ln_l * 8) = ov;
    }
    __syncthreads();
    if (warp * 16 >= rows_valid) continue;  // warp-uniform; the barriers above are reached by every warp each item

    // ---- q[16 rows][128] = xn Wq^T
    float acc[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
#pragma unroll
    for (int kp = 0; kp < C / 32; ++kp) {
      uint32_t a0[4], a1[4];
      ldsm_x4(xn_s + a_off + static_cast<uint32_t>(kp * 32 * 2), a0[0], a0[1], a0[2], a0[3]);
      ldsm_x4(xn_s + a_off + static_cast<uint32_t>((kp * 32 + 16) * 2), a1[0], a1[1], a1[2], a1[3]);
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        const uint4 b = ldw<WS>(wq + (nt * (C / 32) + kp) * 32 + lane);
        mma16816(acc[nt], a0, b.x, b.y);
        mma16816(acc[nt], a1, b.z, b.w);
      }
    }
    // ---- softmax over each head's 32 dims (rows g and g+8 of this warp's 16); scale is folded into M
    uint32_t aq[8][4];
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) mx = fmaxf(mx, fmaxf(acc[4 * hh + j][2 * r], acc[4 * hh + j][2 * r + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float e0 = __expf(acc[4 * hh + j][2 * r] - mx), e1 = __expf(acc[4 * hh + j][2 * r + 1] - mx);
          acc[4 * hh + j][2 * r] = e0;
          acc[4 * hh + j][2 * r + 1] = e1;
          sum += e0 + e1;
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[4 * hh + j][2 * r] *= inv;
          acc[4 * hh + j][2 * r + 1] *= inv;
        }
      }
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const int ks = 2 * hh + k2;
        aq[ks][0] = pack_h2(acc[2 * ks][0], acc[2 * ks][1]);
        aq[ks][1] = pack_h2(acc[2 * ks][2], acc[2 * ks][3]);
        aq[ks][2] = pack_h2(acc[2 * ks + 1][0], acc[2 * ks + 1][1]);
        aq[ks][3] = pack_h2(acc[2 * ks + 1][2], acc[2 * ks + 1][3]);
      }
    }
    // ---- y = q M^T + bias, 64 output channels at a time, written over this warp's own (dead) xn rows
    const uint4* mp = mp_img + lane;
    __syncwarp();
#pragma unroll 1
    for (int cc = 0; cc < C / 64; ++cc) {
      float yacc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) yacc[i][c] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint4 b = ldw<WS>(mp + ((cc * 8 + nt) * 4 + kp) * 32);
          mma16816(yacc[nt], aq[2 * kp], b.x, b.y);
          mma16816(yacc[nt], aq[2 * kp + 1], b.z, b.w);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = cc * 64 + nt * 8 + 2 * q;
        const float b0 = bias ? __ldg(bias + c) : 0.f, b1 = bias ? __ldg(bias + c + 1) : 0.f;
        *reinterpret_cast<uint32_t*>(xn + (16 * warp + g) * XS + c) = pack_h2(yacc[nt][0] + b0, yacc[nt][1] + b1);
        *reinterpret_cast<uint32_t*>(xn + (16 * warp + g + 8) * XS + c) = pack_h2(yacc[nt][2] + b0, yacc[nt][3] + b1);
      }
    }
    __syncwarp();
    // ---- + residual (from the raw tile in shared memory), coalesced store of this warp's 16 rows
    constexpr int CPR = C / 8;  // 16-byte chunks per row
#pragma unroll
    for (int it = 0; it < 16 * CPR / 32; ++it) {
      const int idx = it * 32 + lane;
      const int r = idx / CPR, ch = idx - r * CPR;
      const int row = 16 * warp + r;
      if (row < rows_valid) {
        uint4 v = *reinterpret_cast<const uint4*>(xn + row * XS + ch * 8);
        const uint4 rv = *reinterpret_cast<const uint4*>(rawb + row * C + ch * 8);
        __half2* vh = reinterpret_cast<__half2*>(&v);
        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a2 = __half22float2(vh[i]), b2 = __half22float2(rh[i]);
          vh[i] = wdno::h2_sat(a2.x + b2.x, a2.y + b2.y);
        }
        *(reinterpret_cast<uint4*>(yt + static_cast<size_t>(row) * C) + ch) = v;
      }
    }
  }
}

// ------------------------------------------------------------------ la2, one WARP per 16-position tile
// Same arithmetic as la2_kernel, but a warp owns its 16 rows end to end (private cp.async double buffer, private
// LayerNorm tile, weights through L1), so there is no block-level barrier at all: the block kernel synchronised 8 warps
// twice per 128-position item and ran at ~0.27 IPC per scheduler.
template <int C>
__global__ void __launch_bounds__(128, 4) la2_warp_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                          const float* __restrict__ gamma, const uint4* __restrict__ wq,
                                                          const uint4* __restrict__ mpack, const float* __restrict__ bias, int n,
                                                          int tiles, int n_items, float eps) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int XS = C + 8;
  constexpr int LP = C / 8, RPP = 32 / LP, PASSES = 16 / RPP;
  constexpr int n_m = (C / 8) * 4 * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  __half* xn = reinterpret_cast<__half*>(smem_raw) + warp * (16 * XS + 2 * 16 * C);   // [16][C + 8]
  __half* raw = xn + 16 * XS;                                                          // [2][16][C]
  const int ln_l = lane % LP, ln_r0 = lane / LP;
  pdl_trigger();
  float gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gm[j] = __ldg(gamma + ln_l * 8 + j);
  const uint32_t raw_s = static_cast<uint32_t>(__cvta_generic_to_shared(raw));
  auto fetch = [&](int item, int buf) {
    const int img = item / tiles, tile = item - img * tiles;
    const int p0 = tile * 16;
    const int rows_valid = min(16, n - p0);
    const __half* xt = x + (static_cast<size_t>(img) * n + p0) * C;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int r = ln_r0 + ps * RPP;
      const bool ok = r < rows_valid;
      cp_async16(raw_s + static_cast<uint32_t>(((buf * 16 + r) * C + ln_l * 8) * 2), ok ? xt + static_cast<size_t>(r) * C + ln_l * 8 : xt,
                 ok ? 16u : 0u);
    }
    cp_async_commit_group();
  };
  const uint32_t xn_s = static_cast<uint32_t>(__cvta_generic_to_shared(xn));
  const uint32_t a_off = static_cast<uint32_t>((((lane & 15)) * XS + 8 * (lane >> 4)) * 2);
  const int w0 = blockIdx.x * 4 + warp, wstride = gridDim.x * 4;
  int buf = 0;
  pdl_wait();
  if (w0 < n_items) fetch(w0, 0);
  for (int item = w0; item < n_items; item += wstride, buf ^= 1) {
    const int img = item / tiles, tile = item - img * tiles;
    const int p0 = tile * 16;
    const int rows_valid = min(16, n - p0);
    __half* yt = y + (static_cast<size_t>(img) * n + p0) * C;
    const __half* rawb = raw + buf * 16 * C;
    cp_async_wait_group<0>();
    __syncwarp();  // raw[buf] landed for the whole warp; the previous item's reads of xn / raw[buf^1] are done
    if (item + wstride < n_items) fetch(item + wstride, buf ^ 1);
    const uint4* mp_img = mpack + static_cast<size_t>(img) * n_m;
    // ---- LayerNorm raw[buf] -> xn
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int r = ln_r0 + ps * RPP;
      const uint4 rv = *reinterpret_cast<const uint4*>(rawb + r * C + ln_l * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&rv);
      float f[8];
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(h[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
        sum += t.x + t.y;
      }
#pragma unroll
      for (int sh = LP / 2; sh > 0; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
      const float mean = sum * (1.0f / C);
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] -= mean;
        sq = fmaf(f[j], f[j], sq);
      }
#pragma unroll
      for (int sh = LP / 2; sh > 0; sh >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, sh);
      const float rstd = rsqrtf(sq * (1.0f / C) + eps);
      uint4 ov;
      uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = pack_h2(f[2 * j] * rstd * gm[2 * j], f[2 * j + 1] * rstd * gm[2 * j + 1]);
      if (r >= rows_valid) ov = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(xn + r * XS + ln_l * 8) = ov;
    }
    __syncwarp();

    // ---- q[16 rows][128] = xn Wq^T
    float acc[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
#pragma unroll
    for (int kp = 0; kp < C / 32; ++kp) {
      uint32_t a0[4], a1[4];
      ldsm_x4(xn_s + a_off + static_cast<uint32_t>(kp * 32 * 2), a0[0], a0[1], a0[2], a0[3]);
      ldsm_x4(xn_s + a_off + static_cast<uint32_t>((kp * 32 + 16) * 2), a1[0], a1[1], a1[2], a1[3]);
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        const uint4 b = __ldg(wq + (nt * (C / 32) + kp) * 32 + lane);
        mma16816(acc[nt], a0, b.x, b.y);
        mma16816(acc[nt], a1, b.z, b.w);
      }
    }
    // ---- softmax over each head's 32 dims (rows g and g+8 of this warp's 16); scale is folded into M
    uint32_t aq[8][4];
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) mx = fmaxf(mx, fmaxf(acc[4 * hh + j][2 * r], acc[4 * hh + j][2 * r + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float e0 = __expf(acc[4 * hh + j][2 * r] - mx), e1 = __expf(acc[4 * hh + j][2 * r + 1] - mx);
          acc[4 * hh + j][2 * r] = e0;
          acc[4 * hh + j][2 * r + 1] = e1;
          sum += e0 + e1;
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[4 * hh + j][2 * r] *= inv;
          acc[4 * hh + j][2 * r + 1] *= inv;
        }
      }
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const int ks = 2 * hh + k2;
        aq[ks][0] = pack_h2(acc[2 * ks][0], acc[2 * ks][1]);
        aq[ks][1] = pack_h2(acc[2 * ks][2], acc[2 * ks][3]);
        aq[ks][2] = pack_h2(acc[2 * ks + 1][0], acc[2 * ks + 1][1]);
        aq[ks][3] = pack_h2(acc[2 * ks + 1][2], acc[2 * ks + 1][3]);
      }
    }
    // ---- y = q M^T + bias, 64 output channels at a time, written over this warp's own (dead) xn rows
    const uint4* mp = mp_img + lane;
    __syncwarp();
#pragma unroll 1
    for (int cc = 0; cc < C / 64; ++cc) {
      float yacc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) yacc[i][c] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint4 b = __ldg(mp + ((cc * 8 + nt) * 4 + kp) * 32);
          mma16816(yacc[nt], aq[2 * kp], b.x, b.y);
          mma16816(yacc[nt], aq[2 * kp + 1], b.z, b.w);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = cc * 64 + nt * 8 + 2 * q;
        const float b0 = bias ? __ldg(bias + c) : 0.f, b1 = bias ? __ldg(bias + c + 1) : 0.f;
        *reinterpret_cast<uint32_t*>(xn + (g) * XS + c) = pack_h2(yacc[nt][0] + b0, yacc[nt][1] + b1);
        *reinterpret_cast<uint32_t*>(xn + (g + 8) * XS + c) = pack_h2(yacc[nt][2] + b0, yacc[nt][3] + b1);
      }
    }
    __syncwarp();
    // ---- + residual (from the raw tile in shared memory), coalesced store of this warp's 16 rows
    constexpr int CPR = C / 8;  // 16-byte chunks per row
#pragma unroll
    for (int it = 0; it < 16 * CPR / 32; ++it) {
      const int idx = it * 32 + lane;
      const int r = idx / CPR, ch = idx - r * CPR;
      const int row = r;
      if (row < rows_valid) {
        uint4 v = *reinterpret_cast<const uint4*>(xn + row * XS + ch * 8);
        const uint4 rv = *reinterpret_cast<const uint4*>(rawb + row * C + ch * 8);
        __half2* vh = reinterpret_cast<__half2*>(&v);
        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a2 = __half22float2(vh[i]), b2 = __half22float2(rh[i]);
          vh[i] = wdno::h2_sat(a2.x + b2.x, a2.y + b2.y);
        }
        *(reinterpret_cast<uint4*>(yt + static_cast<size_t>(row) * C) + ch) = v;
      }
    }
  }
}

// ================================================================== temporal attention block
// Residual(PreNorm(dim, EinopsToAndFrom('b c f h w', 'b (h w) f c', Attention(dim, heads=4, dim_head=32, rotary))))
// reference conv3d.py:165-184, 262-353 (focus_present_mask all False), 74-112 (T5 relative position bias).
// One sequence = the n <= 24..32 frames of one pixel.  A block of 8 warps works on two adjacent pixels at a time:
//   LayerNorm -> shared xn[2][32][C+8] (rows >= n stay zero);
//   warp (s, h): Q = xn Wq_h^T (scale, rotary), K = xn Wk_h^T (rotary), V^T = Wv_h xn^T -- all three land directly in
//   the register fragments the next products need -- S = Q K^T + bias -> softmax -> O = P V -> shared obuf[s][32][128];
//   warp (s, c4): Y[:, c4*C/4 ...] = O Wout^T -> shared (over xn) -> + x -> coalesced store.
constexpr int kTaRows = 32;

template <int C, bool WS>
__global__ void __launch_bounds__(256, 2) tattn_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                        const float* __restrict__ gamma, const uint4* wqk,
                                                        const uint4* wv, const uint4* wo,
                                                        const float* __restrict__ bias, const float* __restrict__ rot_cos,
                                                        const float* __restrict__ rot_sin, long long n_pix, long long hw,
                                                        int n, float scale, float eps) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int XS = C + 8, OS = kLaHid + 8, KS = C / 16;
  __half* xn = reinterpret_cast<__half*>(smem_raw);                       // [2][32][XS]
  __half* obuf = xn + 2 * kTaRows * XS;                                    // [2][32][OS]
  // bias rows are 40 floats apart and (cos, sin) rows 20 float2 apart: the 8 rows a warp touches per load then fall
  // into distinct banks
  constexpr int BS = 40, RS = 20;
  float* sbias = reinterpret_cast<float*>(obuf + 2 * kTaRows * OS);        // [4][32][BS], -inf for keys >= n
  float2* scs = reinterpret_cast<float2*>(sbias + 4 * 32 * BS);            // [32][RS] (cos, sin)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int s = warp >> 2, h = warp & 3;
  pdl_trigger();

  for (int i = tid; i < 2 * kTaRows * XS / 8; i += 256) reinterpret_cast<uint4*>(xn)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 4 * 32 * 32; i += 256) {
    const int hh = i >> 10, r = (i >> 5) & 31, c = i & 31;
    float b = 0.f;
    if (c >= n) b = -INFINITY;
    else if (r < n && bias != nullptr) b = __ldg(bias + (static_cast<size_t>(hh) * n + r) * n + c);
    sbias[(hh * 32 + r) * BS + c] = b;
  }
  for (int i = tid; i < 32 * 16; i += 256) {
    const int f = i >> 4;
    scs[f * RS + (i & 15)] = make_float2((rot_cos != nullptr && f < n) ? __ldg(rot_cos + f * 16 + (i & 15)) : 1.0f,
                                         (rot_sin != nullptr && f < n) ? __ldg(rot_sin + f * 16 + (i & 15)) : 0.0f);
  }

  if constexpr (WS) {
    uint4* wsm = reinterpret_cast<uint4*>(scs + 32 * RS);
    constexpr int n_qk = 32 * (C / 32) * 32, n_v = 4 * 2 * KS * 32, n_o = (C / 8) * 4 * 32;
    stage_weights(wsm, wqk, n_qk, tid, 256);
    stage_weights(wsm + n_qk, wv, n_v, tid, 256);
    stage_weights(wsm + n_qk + n_v, wo, n_o, tid, 256);
    wqk = wsm;
    wv = wsm + n_qk;
    wo = wsm + n_qk + n_v;
  }
  const uint32_t xn_s = static_cast<uint32_t>(__cvta_generic_to_shared(xn + s * kTaRows * XS));
  const uint32_t ob_s = static_cast<uint32_t>(__cvta_generic_to_shared(obuf + s * kTaRows * OS));
  const uint32_t a_off = static_cast<uint32_t>((((lane & 15)) * XS + 8 * (lane >> 4)) * 2);   // A operand rows = tokens
  const uint32_t b_off = static_cast<uint32_t>((((lane & 7)) * XS + 8 * (lane >> 3)) * 2);    // B operand rows = tokens
  const uint32_t oa_off = static_cast<uint32_t>((((lane & 15)) * OS + 8 * (lane >> 4)) * 2);
  constexpr int LP = C / 8, RPP = 256 / LP, ROWS = 2 * 24;
  constexpr int PASSES = (2 * kTaRows + RPP - 1) / RPP;  // enough for n <= 32
  const long long n_pairs = (n_pix + 1) >> 1;

  // raw x of the NEXT pair is fetched into registers while the current pair is being computed
  const int ln_l = tid % LP, ln_r0 = tid / LP;
  float gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gm[j] = __ldg(gamma + ln_l * 8 + j);
  uint4 raw[PASSES];
  auto fetch = [&](long long pr) {
    const long long pix0 = pr * 2;
    const long long bimg = pix0 / hw, pin = pix0 - bimg * hw;
    const __half* xb = x + (static_cast<size_t>(bimg) * n * hw + pin) * C;
    const int nseq = (pix0 + 1 < n_pix) ? 2 : 1;
#pragma unroll
    for (int ps = 0; ps < PASSES; ++ps) {
      const int rr = ln_r0 + ps * RPP;
      const int f = rr >> 1, ss = rr & 1;
      raw[ps] = make_uint4(0u, 0u, 0u, 0u);
      if (f < n && ss < nseq) raw[ps] = __ldg(reinterpret_cast<const uint4*>(xb + (static_cast<size_t>(f) * hw + ss) * C) + ln_l);
    }
  };
  pdl_wait();  // tables / weights above are plan constants; x is the predecessor's output
  if (blockIdx.x < n_pairs) fetch(blockIdx.x);

  for (long long pr = blockIdx.x; pr < n_pairs; pr += gridDim.x) {
    const long long pix0 = pr * 2;
    const long long bimg = pix0 / hw, pin = pix0 - bimg * hw;  // both pixels of a pair lie in one sample when hw is even
    const __half* xb = x + (static_cast<size_t>(bimg) * n * hw + pin) * C;
    __half* yb = y + (static_cast<size_t>(bimg) * n * hw + pin) * C;
    const int nseq = (pix0 + 1 < n_pix) ? 2 : 1;
    __syncthreads();  // previous iteration's staging reads are done (also orders the table setup)
    // ---- LayerNorm of 2 x n tokens: row rr = 2 f + s
    {
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int rr = ln_r0 + ps * RPP;
        const int f = rr >> 1, ss = rr & 1;
        const __half2* hh = reinterpret_cast<const __half2*>(&raw[ps]);
        float fv[8];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(hh[j]);
          fv[2 * j] = t.x;
          fv[2 * j + 1] = t.y;
          sum += t.x + t.y;
        }
#pragma unroll
        for (int sh = LP / 2; sh > 0; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
        const float mean = sum * (1.0f / C);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          fv[j] -= mean;
          sq = fmaf(fv[j], fv[j], sq);
        }
#pragma unroll
        for (int sh = LP / 2; sh > 0; sh >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, sh);
        const float rstd = rsqrtf(sq * (1.0f / C) + eps);
        if (f < n) {
          uint4 ov;
          uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = pack_h2(fv[2 * j] * rstd * gm[2 * j], fv[2 * j + 1] * rstd * gm[2 * j + 1]);
          if (ss >= nseq) ov = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(xn + (ss * kTaRows + f) * XS + ln_l * 8) = ov;
        }
      }
    }
    __syncthreads();
    if (pr + gridDim.x < n_pairs) fetch(pr + gridDim.x);

    // ================= attention core of (sequence s, head h)
    {
      // ---- Q (scaled, rotated) -> A fragments ; K (rotated) -> B fragments
      uint32_t qa[2][2][4], kb[4][2][2];
#pragma unroll
      for (int sec = 0; sec < 2; ++sec) {
        float acc[2][4][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
        const uint4* wb = wqk + static_cast<size_t>((sec * 16 + 4 * h) * (C / 32)) * 32 + lane;
#pragma unroll
        for (int kp = 0; kp < C / 32; ++kp) {
          uint32_t a0[2][4], a1[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            ldsm_x4(xn_s + a_off + static_cast<uint32_t>((mt * 16 * XS + kp * 32) * 2), a0[mt][0], a0[mt][1], a0[mt][2], a0[mt][3]);
            ldsm_x4(xn_s + a_off + static_cast<uint32_t>((mt * 16 * XS + kp * 32 + 16) * 2), a1[mt][0], a1[mt][1], a1[mt][2], a1[mt][3]);
          }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const uint4 b = ldw<WS>(wb + (nt * (C / 32) + kp) * 32);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              mma16816(acc[mt][nt], a0[mt], b.x, b.y);
              mma16816(acc[mt][nt], a1[mt], b.z, b.w);
            }
          }
        }
        const float sc = (sec == 0) ? scale : 1.0f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int f = 16 * mt + g + 8 * r;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const float2 cs = scs[f * RS + 4 * nt + q];
              const float x0 = acc[mt][nt][2 * r] * sc, x1 = acc[mt][nt][2 * r + 1] * sc;
              acc[mt][nt][2 * r] = x0 * cs.x - x1 * cs.y;
              acc[mt][nt][2 * r + 1] = x1 * cs.x + x0 * cs.y;
            }
          }
        if (sec == 0) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              qa[mt][ks][0] = pack_h2(acc[mt][2 * ks][0], acc[mt][2 * ks][1]);
              qa[mt][ks][1] = pack_h2(acc[mt][2 * ks][2], acc[mt][2 * ks][3]);
              qa[mt][ks][2] = pack_h2(acc[mt][2 * ks + 1][0], acc[mt][2 * ks + 1][1]);
              qa[mt][ks][3] = pack_h2(acc[mt][2 * ks + 1][2], acc[mt][2 * ks + 1][3]);
            }
        } else {
          // key tile j (keys 8j..8j+7) = rows g + 8 (j & 1) of m-tile j >> 1
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const int rr = 2 * (j & 1);
              kb[j][ks][0] = pack_h2(acc[j >> 1][2 * ks][rr], acc[j >> 1][2 * ks][rr + 1]);
              kb[j][ks][1] = pack_h2(acc[j >> 1][2 * ks + 1][rr], acc[j >> 1][2 * ks + 1][rr + 1]);
            }
        }
      }
      // ---- S = Q K^T + bias (keys >= n masked by the table), softmax over keys
      float sfr[2][4][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int c = 0; c < 4; ++c) sfr[mt][j][c] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) mma16816(sfr[mt][j], qa[mt][ks], kb[j][ks][0], kb[j][ks][1]);
        }
      uint32_t pa[2][2][4];
      float inv[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float* brow = sbias + (h * 32 + 16 * mt + g + 8 * r) * BS + 2 * q;
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 bv = *reinterpret_cast<const float2*>(brow + 8 * j);
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
      // ---- V^T[d][key] = Wv_h xn^T  (its accumulator fragments are the B fragments of P V)
      float vt[2][4][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) vt[i][j][c] = 0.f;
      {
        const uint4* wa = wv + static_cast<size_t>(h * 2) * KS * 32 + lane;
#pragma unroll
        for (int kp = 0; kp < C / 32; ++kp) {
          uint32_t bf[4][4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            ldsm_x4(xn_s + b_off + static_cast<uint32_t>((nt * 8 * XS + kp * 32) * 2), bf[nt][0], bf[nt][1], bf[nt][2], bf[nt][3]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint4 a0 = ldw<WS>(wa + (mt * KS + 2 * kp) * 32), a1 = ldw<WS>(wa + (mt * KS + 2 * kp + 1) * 32);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              mma16816(vt[mt][nt], a0.x, a0.y, a0.z, a0.w, bf[nt][0], bf[nt][1]);
              mma16816(vt[mt][nt], a1.x, a1.y, a1.z, a1.w, bf[nt][2], bf[nt][3]);
            }
          }
        }
      }
      // ---- O = P V, normalised -> obuf[s][token][32 h + d]
      float ofr[2][4][4];
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        const int me = nd >> 1, rr = 2 * (nd & 1);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int c = 0; c < 4; ++c) ofr[mt][nd][c] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint32_t b0 = pack_h2(vt[me][2 * kk][rr], vt[me][2 * kk][rr + 1]);
          const uint32_t b1 = pack_h2(vt[me][2 * kk + 1][rr], vt[me][2 * kk + 1][rr + 1]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma16816(ofr[mt][nd], pa[mt][kk], b0, b1);
        }
      }
      __half* ob = obuf + s * kTaRows * OS + h * 32;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nd = 0; nd < 4; ++nd)
#pragma unroll
          for (int r = 0; r < 2; ++r)
            *reinterpret_cast<uint32_t*>(ob + (16 * mt + g + 8 * r) * OS + 8 * nd + 2 * q) =
                pack_h2(ofr[mt][nd][2 * r] * inv[mt][r], ofr[mt][nd][2 * r + 1] * inv[mt][r]);
    }
    __syncthreads();

    // ================= output projection: warp (s, c4 = h) -> channels [c4*C/4, (c4+1)*C/4)
    {
      constexpr int NT = C / 32;  // n-tiles per warp
      float yacc[2][NT][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) yacc[i][j][c] = 0.f;
      const uint4* wb = wo + static_cast<size_t>(h * NT) * 4 * 32 + lane;
#pragma unroll
      for (int kp = 0; kp < 4; ++kp) {
        uint32_t a0[2][4], a1[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          ldsm_x4(ob_s + oa_off + static_cast<uint32_t>((mt * 16 * OS + kp * 32) * 2), a0[mt][0], a0[mt][1], a0[mt][2], a0[mt][3]);
          ldsm_x4(ob_s + oa_off + static_cast<uint32_t>((mt * 16 * OS + kp * 32 + 16) * 2), a1[mt][0], a1[mt][1], a1[mt][2], a1[mt][3]);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint4 b = ldw<WS>(wb + (nt * 4 + kp) * 32);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma16816(yacc[mt][nt], a0[mt], b.x, b.y);
            mma16816(yacc[mt][nt], a1[mt], b.z, b.w);
          }
        }
      }
      // Y -> xn rows (all warps are past their last xn read: barrier above); rows >= n stay zero
      __half* xs = xn + s * kTaRows * XS + h * (C / 4);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int f = 16 * mt + g + 8 * r;
          if (f < n) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
              *reinterpret_cast<uint32_t*>(xs + f * XS + 8 * nt + 2 * q) = pack_h2(yacc[mt][nt][2 * r], yacc[mt][nt][2 * r + 1]);
          }
        }
    }
    __syncthreads();
    // ---- y = Y + x, coalesced: item = (f, s, 16-byte chunk)
    for (int idx = tid; idx < n * 2 * LP; idx += 256) {
      const int l = idx % LP, rr = idx / LP;
      const int f = rr >> 1, ss = rr & 1;
      if (ss < nseq) {
        uint4 v = *reinterpret_cast<const uint4*>(xn + (ss * kTaRows + f) * XS + l * 8);
        const size_t off = (static_cast<size_t>(f) * hw + ss) * C;
        const uint4 rv = __ldg(reinterpret_cast<const uint4*>(xb + off) + l);
        __half2* vh = reinterpret_cast<__half2*>(&v);
        const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a2 = __half22float2(vh[i]), b2 = __half22float2(rh[i]);
          vh[i] = wdno::h2_sat(a2.x + b2.x, a2.y + b2.y);
        }
        *(reinterpret_cast<uint4*>(yb + off) + l) = v;
      }
    }
  }
  (void)ROWS;
}

// ------------------------------------------------------------------ temporal attention, one WARP per pixel (C = 64)
// The block-per-pixel-pair kernel above synchronises its 8 warps four times per pair (LayerNorm tile shared by the head
// warps, per-head outputs gathered for the output projection) and runs at ~0.27 IPC per scheduler on the full-resolution
// level.  Here a warp owns a pixel end to end -- LayerNorm of its 24 tokens into a private tile, the four heads one after
// the other, and the output projection accumulated head by head straight from the P.V accumulator fragments (they ARE
// the A fragments of  Y += O_h Wout[:, 32h:32h+32]^T ) -- so there is no block-level barrier after the table setup and no
// O round trip through shared memory; 3 CTAs x 4 warps per SM.
template <int C>
__global__ void __launch_bounds__(128, 3) tattn_warp_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                                             const float* __restrict__ gamma, const uint4* __restrict__ wqk,
                                                             const uint4* __restrict__ wv, const uint4* __restrict__ wo,
                                                             const float* __restrict__ bias, const float* __restrict__ rot_cos,
                                                             const float* __restrict__ rot_sin, long long n_pix, long long hw,
                                                             int n, float scale, float eps) {
  static_assert(C == 64, "one-warp-per-pixel temporal attention is specialised for C = 64");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int XS = C + 8, KS = C / 16, LP = C / 8;
  constexpr int BS = 40, RS = 20;
  __half* xn_all = reinterpret_cast<__half*>(smem_raw);                  // [4 warps][32][XS]
  float* sbias = reinterpret_cast<float*>(xn_all + 4 * kTaRows * XS);      // [4][32][BS], -inf for keys >= n
  float2* scs = reinterpret_cast<float2*>(sbias + 4 * 32 * BS);            // [32][RS] (cos, sin)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  pdl_trigger();
  for (int i = tid; i < 4 * kTaRows * XS / 8; i += 128) reinterpret_cast<uint4*>(xn_all)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 4 * 32 * 32; i += 128) {
    const int hh = i >> 10, r = (i >> 5) & 31, c = i & 31;
    float b = 0.f;
    if (c >= n) b = -INFINITY;
    else if (r < n && bias != nullptr) b = __ldg(bias + (static_cast<size_t>(hh) * n + r) * n + c);
    sbias[(hh * 32 + r) * BS + c] = b;
  }
  for (int i = tid; i < 32 * 16; i += 128) {
    const int f = i >> 4;
    scs[f * RS + (i & 15)] = make_float2((rot_cos != nullptr && f < n) ? __ldg(rot_cos + f * 16 + (i & 15)) : 1.0f,
                                         (rot_sin != nullptr && f < n) ? __ldg(rot_sin + f * 16 + (i & 15)) : 0.0f);
  }
  __syncthreads();
  __half* xn = xn_all + warp * kTaRows * XS;
  const uint32_t xn_s = static_cast<uint32_t>(__cvta_generic_to_shared(xn));
  const uint32_t a_off = static_cast<uint32_t>((((lane & 15)) * XS + 8 * (lane >> 4)) * 2);   // A operand rows = tokens
  const uint32_t b_off = static_cast<uint32_t>((((lane & 7)) * XS + 8 * (lane >> 3)) * 2);    // B operand rows = tokens
  const int ln_l = lane & 7, ln_r = lane >> 3;   // LayerNorm: 8 lanes per token row, 4 rows per pass
  float gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gm[j] = __ldg(gamma + ln_l * 8 + j);
  pdl_wait();

  for (long long pix = static_cast<long long>(blockIdx.x) * 4 + warp; pix < n_pix; pix += static_cast<long long>(gridDim.x) * 4) {
    const long long bimg = pix / hw, pin = pix - bimg * hw;
    const __half* xb = x + (static_cast<size_t>(bimg) * n * hw + pin) * C;
    __half* yb = y + (static_cast<size_t>(bimg) * n * hw + pin) * C;
    // ---- LayerNorm of the pixel's n tokens -> xn rows 0..n-1 (rows >= n stay zero)
    {
      uint4 raw[8];
#pragma unroll
      for (int ps = 0; ps < 8; ++ps) {
        const int f = ln_r + 4 * ps;
        raw[ps] = make_uint4(0u, 0u, 0u, 0u);
        if (f < n) raw[ps] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(f) * hw * C) + ln_l);
      }
      __syncwarp();  // the previous pixel's staged output has been read
#pragma unroll
      for (int ps = 0; ps < 8; ++ps) {
        const int f = ln_r + 4 * ps;
        const __half2* hh = reinterpret_cast<const __half2*>(&raw[ps]);
        float fv[8];
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(hh[j]);
          fv[2 * j] = t.x;
          fv[2 * j + 1] = t.y;
          sum += t.x + t.y;
        }
#pragma unroll
        for (int sh = LP / 2; sh > 0; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
        const float mean = sum * (1.0f / C);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          fv[j] -= mean;
          sq = fmaf(fv[j], fv[j], sq);
        }
#pragma unroll
        for (int sh = LP / 2; sh > 0; sh >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, sh);
        const float rstd = rsqrtf(sq * (1.0f / C) + eps);
        if (f < n) {
          uint4 ov;
          uint32_t* o = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = pack_h2(fv[2 * j] * rstd * gm[2 * j], fv[2 * j + 1] * rstd * gm[2 * j + 1]);
          *reinterpret_cast<uint4*>(xn + f * XS + ln_l * 8) = ov;
        }
      }
    }
    __syncwarp();

    float yacc[2][C / 8][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < C / 8; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) yacc[i][j][c] = 0.f;

#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
      // ---- Q (scaled, rotated) -> A fragments ; K (rotated) -> B fragments
      uint32_t qa[2][2][4], kb[4][2][2];
#pragma unroll
      for (int sec = 0; sec < 2; ++sec) {
        float acc[2][4][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.f;
        const uint4* wb = wqk + static_cast<size_t>((sec * 16 + 4 * h) * (C / 32)) * 32 + lane;
#pragma unroll
        for (int kp = 0; kp < C / 32; ++kp) {
          uint32_t a0[2][4], a1[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            ldsm_x4(xn_s + a_off + static_cast<uint32_t>((mt * 16 * XS + kp * 32) * 2), a0[mt][0], a0[mt][1], a0[mt][2], a0[mt][3]);
            ldsm_x4(xn_s + a_off + static_cast<uint32_t>((mt * 16 * XS + kp * 32 + 16) * 2), a1[mt][0], a1[mt][1], a1[mt][2], a1[mt][3]);
          }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const uint4 b = __ldg(wb + (nt * (C / 32) + kp) * 32);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              mma16816(acc[mt][nt], a0[mt], b.x, b.y);
              mma16816(acc[mt][nt], a1[mt], b.z, b.w);
            }
          }
        }
        const float sc = (sec == 0) ? scale : 1.0f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int f = 16 * mt + g + 8 * r;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const float2 cs = scs[f * RS + 4 * nt + q];
              const float x0 = acc[mt][nt][2 * r] * sc, x1 = acc[mt][nt][2 * r + 1] * sc;
              acc[mt][nt][2 * r] = x0 * cs.x - x1 * cs.y;
              acc[mt][nt][2 * r + 1] = x1 * cs.x + x0 * cs.y;
            }
          }
        if (sec == 0) {
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              qa[mt][ks][0] = pack_h2(acc[mt][2 * ks][0], acc[mt][2 * ks][1]);
              qa[mt][ks][1] = pack_h2(acc[mt][2 * ks][2], acc[mt][2 * ks][3]);
              qa[mt][ks][2] = pack_h2(acc[mt][2 * ks + 1][0], acc[mt][2 * ks + 1][1]);
              qa[mt][ks][3] = pack_h2(acc[mt][2 * ks + 1][2], acc[mt][2 * ks + 1][3]);
            }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const int rr = 2 * (j & 1);
              kb[j][ks][0] = pack_h2(acc[j >> 1][2 * ks][rr], acc[j >> 1][2 * ks][rr + 1]);
              kb[j][ks][1] = pack_h2(acc[j >> 1][2 * ks + 1][rr], acc[j >> 1][2 * ks + 1][rr + 1]);
            }
        }
      }
      // ---- S = Q K^T + bias (keys >= n masked by the table), softmax over keys
      float sfr[2][4][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int c = 0; c < 4; ++c) sfr[mt][j][c] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) mma16816(sfr[mt][j], qa[mt][ks], kb[j][ks][0], kb[j][ks][1]);
        }
      uint32_t pa[2][2][4];
      float inv[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float* brow = sbias + (h * 32 + 16 * mt + g + 8 * r) * BS + 2 * q;
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 bv = *reinterpret_cast<const float2*>(brow + 8 * j);
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
      // ---- V^T[d][key] = Wv_h xn^T  (its accumulator fragments are the B fragments of P V)
      float vt[2][4][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) vt[i][j][c] = 0.f;
      {
        const uint4* wa = wv + static_cast<size_t>(h * 2) * KS * 32 + lane;
#pragma unroll
        for (int kp = 0; kp < C / 32; ++kp) {
          uint32_t bf[4][4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            ldsm_x4(xn_s + b_off + static_cast<uint32_t>((nt * 8 * XS + kp * 32) * 2), bf[nt][0], bf[nt][1], bf[nt][2], bf[nt][3]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint4 a0 = __ldg(wa + (mt * KS + 2 * kp) * 32), a1 = __ldg(wa + (mt * KS + 2 * kp + 1) * 32);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              mma16816(vt[mt][nt], a0.x, a0.y, a0.z, a0.w, bf[nt][0], bf[nt][1]);
              mma16816(vt[mt][nt], a1.x, a1.y, a1.z, a1.w, bf[nt][2], bf[nt][3]);
            }
          }
        }
      }
      // ---- O = P V ; normalised O fragments are the A fragments of the head's slice of the output projection
      uint32_t oa[2][2][4];
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) {
        const int me = nd >> 1, rr = 2 * (nd & 1);
        float ofr[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int c = 0; c < 4; ++c) ofr[mt][c] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint32_t b0 = pack_h2(vt[me][2 * kk][rr], vt[me][2 * kk][rr + 1]);
          const uint32_t b1 = pack_h2(vt[me][2 * kk + 1][rr], vt[me][2 * kk + 1][rr + 1]);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma16816(ofr[mt], pa[mt][kk], b0, b1);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          // n-tile nd covers d = 8 nd .. 8 nd + 7: k-step nd >> 1, low (a0, a1) or high (a2, a3) half
          oa[mt][nd >> 1][2 * (nd & 1)] = pack_h2(ofr[mt][0] * inv[mt][0], ofr[mt][1] * inv[mt][0]);
          oa[mt][nd >> 1][2 * (nd & 1) + 1] = pack_h2(ofr[mt][2] * inv[mt][1], ofr[mt][3] * inv[mt][1]);
        }
      }
      // ---- Y += O_h Wout[:, 32 h : 32 h + 32]^T
      {
        const uint4* wb = wo + static_cast<size_t>(h) * 32 + lane;
#pragma unroll
        for (int nt = 0; nt < C / 8; ++nt) {
          const uint4 b = __ldg(wb + nt * 4 * 32);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma16816(yacc[mt][nt], oa[mt][0], b.x, b.y);
            mma16816(yacc[mt][nt], oa[mt][1], b.z, b.w);
          }
        }
      }
    }
    // ---- Y -> the warp's tile (its LayerNorm rows are dead), then y = Y + x with coalesced 16-byte accesses
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int f = 16 * mt + g + 8 * r;
        if (f < n) {
#pragma unroll
          for (int nt = 0; nt < C / 8; ++nt)
            *reinterpret_cast<uint32_t*>(xn + f * XS + 8 * nt + 2 * q) = pack_h2(yacc[mt][nt][2 * r], yacc[mt][nt][2 * r + 1]);
        }
      }
    __syncwarp();
    for (int idx = lane; idx < n * LP; idx += 32) {
      const int f = idx / LP, l = idx - f * LP;
      uint4 v = *reinterpret_cast<const uint4*>(xn + f * XS + l * 8);
      const size_t off = static_cast<size_t>(f) * hw * C;
      const uint4 rv = __ldg(reinterpret_cast<const uint4*>(xb + off) + l);
      __half2* vh = reinterpret_cast<__half2*>(&v);
      const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 a2 = __half22float2(vh[i]), b2 = __half22float2(rh[i]);
        vh[i] = wdno::h2_sat(a2.x + b2.x, a2.y + b2.y);
      }
      *(reinterpret_cast<uint4*>(yb + off) + l) = v;
    }
  }
}

template <int C>
static int launch_tattn(const __half* x, __half* y, const float* gamma, const uint4* wqk, const uint4* wv, const uint4* wo,
                        const float* bias, const float* rot_cos, const float* rot_sin, long long n_pix, long long hw, int n,
                        float scale, float eps, cudaStream_t st) {
  if constexpr (C == 64) {
    // one warp per pixel (no block barriers, O never leaves registers); WDNO_TATTN_WARP=0 selects the pair kernel
    static const bool use_warp = [] { const char* e = getenv("WDNO_TATTN_WARP"); return !(e && e[0] == '0'); }();
    if (use_warp) {
      const int smem_w = 4 * kTaRows * (C + 8) * 2 + (4 * 32 * 40 + 2 * 32 * 20) * 4;
      static bool configured_w = false;
      if (!configured_w) {
        cudaError_t e = cudaFuncSetAttribute(tattn_warp_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_w);
        if (e != cudaSuccess) return set_cuda_error(e, "tattn_block: cudaFuncSetAttribute");
        configured_w = true;
      }
      const long long want = (n_pix + 3) / 4;
      const long long capw = static_cast<long long>(num_sms()) * 3;
      const unsigned gridw = static_cast<unsigned>(want < capw ? want : capw);
      launch_pdl(tattn_warp_kernel<C>, dim3(gridw), dim3(128), static_cast<size_t>(smem_w), st, x, y, gamma, wqk, wv, wo, bias,
                 rot_cos, rot_sin, n_pix, hw, n, scale, eps);
      return check_launch("tattn_block");
    }
  }
  constexpr bool WS = false;  // weight fragments come through L1 (staging them in shared memory measured the same: both share one data path)
  const int wbytes = WS ? (32 * (C / 32) + 4 * 2 * (C / 16) + (C / 8) * 4) * 32 * 16 : 0;
  const int smem = (2 * kTaRows * (C + 8) + 2 * kTaRows * (kLaHid + 8)) * 2 + (4 * 32 * 40 + 2 * 32 * 20) * 4 + wbytes;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tattn_kernel<C, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "tattn_block: cudaFuncSetAttribute");
    configured = true;
  }
  const long long n_pairs = (n_pix + 1) / 2;
  const long long cap = static_cast<long long>(num_sms()) * 2 * 4;
  const unsigned grid = static_cast<unsigned>(n_pairs < cap ? n_pairs : cap);
  launch_pdl(tattn_kernel<C, WS>, dim3(grid), dim3(256), static_cast<size_t>(smem), st, x, y, gamma, wqk, wv, wo, bias, rot_cos,
             rot_sin, n_pix, hw, n, scale, eps);
  return check_launch("tattn_block");
}

// merge + fold launch for the tcgen05 form (linattn_tc.cu): same kernel, M written as a UMMA operand when canon != 0
int launch_la_mid(const float* part, const float* wout, void* mpack, int C, int nparts, float scale, int canon, int n_img,
                  cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(la_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * kLaHid * 2);
    if (e != cudaSuccess) return set_cuda_error(e, "la_mid: cudaFuncSetAttribute");
    configured = true;
  }
  cudaError_t le = launch_pdl(la_mid_kernel, dim3(n_img), dim3(256), static_cast<size_t>(C * kLaHid * 2), st, part, wout, static_cast<__half*>(mpack), C,
                              nparts, scale, canon);
  if (le != cudaSuccess) return set_cuda_error(le, "la_mid: launch");
  return WDNO_OK;
}

static int la_split(int n_pos) { return ((n_pos + 63) / 64 >= 16) ? 2 : 1; }

template <int C>
static int launch_linattn(const __half* x, __half* y, const float* gamma, const uint4* wq, const uint4* wkv,
                          const float* wout, const float* bias, void* work, int n_img, int n_pos, float scale, float eps,
                          cudaStream_t st) {
  const int split = la_split(n_pos);
  const int nparts = split * 2;
  float* part = static_cast<float*>(work);
  __half* mpack = reinterpret_cast<__half*>(part + static_cast<size_t>(n_img) * nparts * kLaHeads * kLaPart);
  constexpr bool WS1 = (C <= 128);  // la1: K/V weight fragments staged in shared memory while the blocks still fit per SM
  static const bool la1_small = [] { const char* e = getenv("WDNO_LA1_NPH"); return !(e && e[0] == '2'); }();
  constexpr bool WS2 = (C == 64);   // la2: Wq and the image's M fragments in shared memory
  const int smem1 = 64 * (C + 8) * 2 + (WS1 ? 2 * 4 * 2 * (C / 16) * 32 * 16 : 0);
  const int smem1s = 32 * (C + 8) * 2 + (WS1 ? 2 * 4 * 2 * (C / 16) * 32 * 16 : 0);
  const int smem2 = 128 * (C + 8) * 2 + 2 * 128 * C * 2 + (WS2 ? (16 * (C / 32) + (C / 8) * 4) * 32 * 16 : 0);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(la2_kernel<C, WS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * kLaHid * 2);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la1_kernel<C, WS1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(la1_kernel<C, WS1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1s);
    if (e != cudaSuccess) return set_cuda_error(e, "linattn_block: cudaFuncSetAttribute");
    configured = true;
  }
  if (la1_small)   // same number of partials (split * 2): twice the blocks, one 32-position column each
    launch_pdl(la1_kernel<C, WS1, 1>, dim3(n_img * split * 2), dim3(128), static_cast<size_t>(smem1s), st, x, gamma, wkv, part, n_pos,
               split * 2, eps);
  else
    launch_pdl(la1_kernel<C, WS1, 2>, dim3(n_img * split), dim3(256), static_cast<size_t>(smem1), st, x, gamma, wkv, part, n_pos, split,
               eps);
  launch_pdl(la_mid_kernel, dim3(n_img), dim3(256), static_cast<size_t>(C * kLaHid * 2), st, static_cast<const float*>(part), wout,
             mpack, C, nparts, scale, 0);
  static const bool use_warp = [] { const char* e = getenv("WDNO_LA2_WARP"); return !(e && e[0] == '0'); }();
  if (use_warp) {
    const int smem_w = 4 * (16 * (C + 8) + 2 * 16 * C) * 2;
    static bool configured_w = false;
    if (!configured_w) {
      cudaError_t e = cudaFuncSetAttribute(la2_warp_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_w);
      if (e != cudaSuccess) return set_cuda_error(e, "linattn_block: cudaFuncSetAttribute");
      configured_w = true;
    }
    const int tiles16 = (n_pos + 15) / 16;
    const long long items16 = static_cast<long long>(n_img) * tiles16;
    if (items16 > 2147483647LL) return set_error(WDNO_E_INVALID, "linattn_block: too many tiles");
    const long long want = (items16 + 3) / 4, capw = static_cast<long long>(num_sms()) * 4;
    launch_pdl(la2_warp_kernel<C>, dim3(static_cast<unsigned>(want < capw ? want : capw)), dim3(128), static_cast<size_t>(smem_w), st, x,
               y, gamma, wq, reinterpret_cast<const uint4*>(mpack), bias, n_pos, tiles16, static_cast<int>(items16), eps);
    return check_launch("linattn_block");
  }
  const int tiles = (n_pos + 127) / 128;
  const int n_items = n_img * tiles;
  const int cap = num_sms() * 2;
  launch_pdl(la2_kernel<C, WS2>, dim3(n_items < cap ? n_items : cap), dim3(256), static_cast<size_t>(smem2), st, x, y, gamma, wq,
             reinterpret_cast<const uint4*>(mpack), bias, n_pos, tiles, n_items, eps);
  return check_launch("linattn_block");
}

}  // namespace wdno

using namespace wdno;

extern "C" int wdno_tattn_block(const void* x, void* y, const float* gamma, const void* wqk_pack, const void* wv_pack,
                                const void* wout_pack, const float* bias, const float* rot_cos, const float* rot_sin,
                                int64_t n_samples, int n_frames, int64_t hw, int C, float scale, float eps, void* stream) {
  if (!x || !y || !gamma || !wqk_pack || !wv_pack || !wout_pack || n_samples < 1 || hw < 1)
    return set_error(WDNO_E_INVALID, "tattn_block: bad arguments");
  if (n_frames < 1 || n_frames > 32) return set_error(WDNO_E_INVALID, "tattn_block: frames must be in [1,32]");
  if (hw & 1) return set_error(WDNO_E_INVALID, "tattn_block: H*W must be even");
  if ((rot_cos == nullptr) != (rot_sin == nullptr)) return set_error(WDNO_E_INVALID, "tattn_block: rotary tables must both be given");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* xi = static_cast<const __half*>(x);
  __half* yo = static_cast<__half*>(y);
  const uint4* a = static_cast<const uint4*>(wqk_pack);
  const uint4* b = static_cast<const uint4*>(wv_pack);
  const uint4* c = static_cast<const uint4*>(wout_pack);
  const long long npix = n_samples * hw;
  switch (C) {
    case 64: return launch_tattn<64>(xi, yo, gamma, a, b, c, bias, rot_cos, rot_sin, npix, hw, n_frames, scale, eps, st);
    case 128: return launch_tattn<128>(xi, yo, gamma, a, b, c, bias, rot_cos, rot_sin, npix, hw, n_frames, scale, eps, st);
    case 256: return launch_tattn<256>(xi, yo, gamma, a, b, c, bias, rot_cos, rot_sin, npix, hw, n_frames, scale, eps, st);
    default: return set_error(WDNO_E_INVALID, "tattn_block: C must be 64, 128 or 256");
  }
}

extern "C" int64_t wdno_linattn_work_bytes(int64_t n_img, int n_pos, int C) {
  if (n_img < 1 || n_pos < 1 || C < 1) return WDNO_E_INVALID;
  const int nparts = la_split(n_pos) * 2;
  return n_img * nparts * kLaHeads * kLaPart * 4 + n_img * static_cast<int64_t>(C) * kLaHid * 2;
}

extern "C" int wdno_linattn_block(const void* x, void* y, const float* gamma, const void* wq_pack, const void* wkv_pack,
                                  const float* wout, const float* bias, void* work, int64_t n_img, int n_pos, int C,
                                  float scale, float eps, void* stream) {
  if (!x || !y || !gamma || !wq_pack || !wkv_pack || !wout || !work || n_img < 1 || n_pos < 1)
    return set_error(WDNO_E_INVALID, "linattn_block: bad arguments");
  if (n_img > (1 << 20)) return set_error(WDNO_E_INVALID, "linattn_block: too many images");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* xi = static_cast<const __half*>(x);
  __half* yo = static_cast<__half*>(y);
  const uint4* wq = static_cast<const uint4*>(wq_pack);
  const uint4* wkv = static_cast<const uint4*>(wkv_pack);
  const int ni = static_cast<int>(n_img);
  switch (C) {
    case 64: return launch_linattn<64>(xi, yo, gamma, wq, wkv, wout, bias, work, ni, n_pos, scale, eps, st);
    case 128: return launch_linattn<128>(xi, yo, gamma, wq, wkv, wout, bias, work, ni, n_pos, scale, eps, st);
    case 256: return launch_linattn<256>(xi, yo, gamma, wq, wkv, wout, bias, work, ni, n_pos, scale, eps, st);
    default: return set_error(WDNO_E_INVALID, "linattn_block: C must be 64, 128 or 256");
  }
}
