// Attention cores of the two U-Nets (the qkv / out projections run on the tap-GEMM kernel).
//   softmax attention over short sequences: temporal attention (n = frames, rotary + T5 relative-position bias),
//     mid-block spatial attention (n = H*W tokens);   reference conv3d.py:277-353, 383, 450-453; unet.py:225-259
//   linear attention: softmax_d(q), softmax_n(k), ctx = k v^T, out = ctx^T q * scale;
//     reference conv3d.py:232-258; unet.py:183-223
// heads x dim_head is fixed to 4 x 32 (both reference models use the defaults attn_heads=4, attn_dim_head=32).
#include <cuda_fp16.h>
#include "cvt_sat.cuh"
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace wdno {

constexpr int kHeads = 4;
constexpr int kDh = 32;
constexpr int kHid = kHeads * kDh;   // 128
constexpr int kQkv = 3 * kHid;       // 384

// token index of (sequence s, token t):  (s / inner) * outerT + (s % inner) * innerT + t * tokT
struct SeqMap {
  long long inner, outerT, innerT, tokT;
};

__device__ __forceinline__ long long tok_index(const SeqMap& m, long long s, int t) {
  return (s / m.inner) * m.outerT + (s % m.inner) * m.innerT + static_cast<long long>(t) * m.tokT;
}

// ------------------------------------------------------------------ softmax attention, one thread per (head, query)
// Keys/values are staged chunk-wise (KCH tokens, all heads) in shared memory as fp32 with rotary applied to k;
// each thread keeps its (scaled, rotated) query and the output accumulator in registers and runs an online softmax.
constexpr int KCH = 32;
constexpr int kRow = kHid + 4;  // padded fp32 row (bank-conflict-free float4 reads of a broadcast row are trivially fine)

template <bool ROTARY>
__global__ void __launch_bounds__(512) softmax_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out,
                                                           const float* __restrict__ bias /*[H][n][n] or null*/,
                                                           const float* __restrict__ rot_cos /*[n][16]*/,
                                                           const float* __restrict__ rot_sin, SeqMap map, int n,
                                                           float scale) {
  __shared__ float ks[KCH * kRow];
  __shared__ float vs[KCH * kRow];
  const long long s = blockIdx.x;
  const int QT = blockDim.x / kHeads;  // query threads per head
  const int h = threadIdx.x / QT;
  const int iq0 = threadIdx.x - h * QT;
  for (int i0 = 0; i0 < n; i0 += QT) {
    const int i = i0 + iq0;
    const bool qok = i < n;
    float q[kDh], acc[kDh];
    float mrun = -INFINITY, lrun = 0.f;
#pragma unroll
    for (int d = 0; d < kDh; ++d) acc[d] = 0.f;
    if (qok) {
      const __half* qp = qkv + tok_index(map, s, i) * kQkv + h * kDh;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(qp) + c);
        const __half2* hh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(hh[j]);
          q[c * 8 + 2 * j] = t.x * scale;
          q[c * 8 + 2 * j + 1] = t.y * scale;
        }
      }
      if (ROTARY) {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const float cs = __ldg(rot_cos + i * 16 + m), sn = __ldg(rot_sin + i * 16 + m);
          const float x0 = q[2 * m], x1 = q[2 * m + 1];
          q[2 * m] = x0 * cs - x1 * sn;
          q[2 * m + 1] = x1 * cs + x0 * sn;
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < kDh; ++d) q[d] = 0.f;
    }
    for (int j0 = 0; j0 < n; j0 += KCH) {
      const int jn = min(KCH, n - j0);
      __syncthreads();
      // stage k (rotated) and v for tokens j0..j0+jn, all heads: one thread per half2 pair
      for (int e = threadIdx.x; e < jn * (kHid / 2); e += blockDim.x) {
        const int jt = e / (kHid / 2);
        const int c2 = e - jt * (kHid / 2);  // pair index within the 128 hidden dims
        const __half* base = qkv + tok_index(map, s, j0 + jt) * kQkv;
        const float2 kk = __half22float2(*reinterpret_cast<const __half2*>(base + kHid + 2 * c2));
        const float2 vv = __half22float2(*reinterpret_cast<const __half2*>(base + 2 * kHid + 2 * c2));
        float k0 = kk.x, k1 = kk.y;
        if (ROTARY) {
          const int m = c2 & 15;  // pair index inside the head
          const float cs = __ldg(rot_cos + (j0 + jt) * 16 + m), sn = __ldg(rot_sin + (j0 + jt) * 16 + m);
          const float x0 = k0, x1 = k1;
          k0 = x0 * cs - x1 * sn;
          k1 = x1 * cs + x0 * sn;
        }
        ks[jt * kRow + 2 * c2] = k0;
        ks[jt * kRow + 2 * c2 + 1] = k1;
        vs[jt * kRow + 2 * c2] = vv.x;
        vs[jt * kRow + 2 * c2 + 1] = vv.y;
      }
      __syncthreads();
      if (qok) {
        for (int jt = 0; jt < jn; ++jt) {
          const float4* kr = reinterpret_cast<const float4*>(ks + jt * kRow + h * kDh);
          float sdot = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 k4 = kr[c];
            sdot = fmaf(q[4 * c], k4.x, sdot);
            sdot = fmaf(q[4 * c + 1], k4.y, sdot);
            sdot = fmaf(q[4 * c + 2], k4.z, sdot);
            sdot = fmaf(q[4 * c + 3], k4.w, sdot);
          }
          if (bias != nullptr) sdot += __ldg(bias + (static_cast<size_t>(h) * n + i) * n + j0 + jt);
          const float mnew = fmaxf(mrun, sdot);
          const float corr = __expf(mrun - mnew);
          const float pj = __expf(sdot - mnew);
          lrun = lrun * corr + pj;
          mrun = mnew;
          const float4* vr = reinterpret_cast<const float4*>(vs + jt * kRow + h * kDh);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v4 = vr[c];
            acc[4 * c] = fmaf(acc[4 * c], corr, pj * v4.x);
            acc[4 * c + 1] = fmaf(acc[4 * c + 1], corr, pj * v4.y);
            acc[4 * c + 2] = fmaf(acc[4 * c + 2], corr, pj * v4.z);
            acc[4 * c + 3] = fmaf(acc[4 * c + 3], corr, pj * v4.w);
          }
        }
      }
    }
    if (qok) {
      const float inv = 1.0f / lrun;
      __half* op = out + tok_index(map, s, i) * kHid + h * kDh;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 ov;
        __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = wdno::h2_sat(acc[c * 8 + 2 * j] * inv, acc[c * 8 + 2 * j + 1] * inv);
        reinterpret_cast<uint4*>(op)[c] = ov;
      }
    }
  }
}

// ------------------------------------------------------------------ linear attention, one block per (image, head)
// qkv: [images][n][384] fp16 ; out: [images][n][128] fp16.
constexpr int LT = 64;  // positions per staged tile

__global__ void __launch_bounds__(256) linear_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int n,
                                                          float scale) {
  __shared__ float ek[LT][kDh + 1];
  __shared__ float vv[LT][kDh];
  __shared__ float red[8][kDh];
  __shared__ float cmax[kDh];
  __shared__ float zsum[kDh];
  __shared__ float ctx[kDh][kDh];  // [d][e], normalised
  const int img = blockIdx.x / kHeads;
  const int h = blockIdx.x - img * kHeads;
  const __half* base = qkv + static_cast<size_t>(img) * n * kQkv;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  // pass 1: column max of k over positions. thread (warp w, lane d) scans positions w, w+8, ...
  {
    float m = -INFINITY;
    for (int p = warp; p < n; p += 8) m = fmaxf(m, __half2float(base[static_cast<size_t>(p) * kQkv + kHid + h * kDh + lane]));
    red[warp][lane] = m;
    __syncthreads();
    if (warp == 0) {
      float mm = red[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) mm = fmaxf(mm, red[w][lane]);
      cmax[lane] = mm;
    }
    __syncthreads();
  }
  // pass 2: ctx[d][e] = sum_p exp(k[p][d]-cmax[d]) * v[p][e] ; thread (d = lane, eg = warp) owns e = 4*eg..4*eg+3
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, z = 0.f;
  for (int p0 = 0; p0 < n; p0 += LT) {
    const int pn = min(LT, n - p0);
    __syncthreads();
    for (int e = tid; e < pn * kDh; e += 256) {
      const int pp = e >> 5, d = e & 31;
      const __half* row = base + static_cast<size_t>(p0 + pp) * kQkv + h * kDh;
      ek[pp][d] = __expf(__half2float(row[kHid + d]) - cmax[d]);
      vv[pp][d] = __half2float(row[2 * kHid + d]);
    }
    __syncthreads();
    for (int pp = 0; pp < pn; ++pp) {
      const float e = ek[pp][lane];
      const float4 v4 = *reinterpret_cast<const float4*>(&vv[pp][warp * 4]);
      a0 = fmaf(e, v4.x, a0);
      a1 = fmaf(e, v4.y, a1);
      a2 = fmaf(e, v4.z, a2);
      a3 = fmaf(e, v4.w, a3);
      if (warp == 0) z += e;
    }
  }
  if (warp == 0) zsum[lane] = z;
  __syncthreads();
  {
    const float inv = 1.0f / zsum[lane];
    ctx[lane][warp * 4 + 0] = a0 * inv;
    ctx[lane][warp * 4 + 1] = a1 * inv;
    ctx[lane][warp * 4 + 2] = a2 * inv;
    ctx[lane][warp * 4 + 3] = a3 * inv;
  }
  __syncthreads();
  // pass 3: out[p][e] = scale * sum_d ctx[d][e] * softmax_d(q[p])[d] ; one thread per position
  for (int p = tid; p < n; p += 256) {
    const __half* qp = base + static_cast<size_t>(p) * kQkv + h * kDh;
    float q[kDh];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(qp) + c);
      const __half2* hh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(hh[j]);
        q[c * 8 + 2 * j] = t.x;
        q[c * 8 + 2 * j + 1] = t.y;
        m = fmaxf(m, fmaxf(t.x, t.y));
      }
    }
    float ssum = 0.f;
#pragma unroll
    for (int d = 0; d < kDh; ++d) {
      q[d] = __expf(q[d] - m);
      ssum += q[d];
    }
    const float qs = scale / ssum;
    float o[kDh];
#pragma unroll
    for (int e = 0; e < kDh; ++e) o[e] = 0.f;
#pragma unroll
    for (int d = 0; d < kDh; ++d) {
      const float qd = q[d] * qs;
#pragma unroll
      for (int e4 = 0; e4 < 8; ++e4) {
        const float4 c4 = *reinterpret_cast<const float4*>(&ctx[d][e4 * 4]);
        o[e4 * 4 + 0] = fmaf(c4.x, qd, o[e4 * 4 + 0]);
        o[e4 * 4 + 1] = fmaf(c4.y, qd, o[e4 * 4 + 1]);
        o[e4 * 4 + 2] = fmaf(c4.z, qd, o[e4 * 4 + 2]);
        o[e4 * 4 + 3] = fmaf(c4.w, qd, o[e4 * 4 + 3]);
      }
    }
    __half* op = out + (static_cast<size_t>(img) * n + p) * kHid + h * kDh;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 ov;
      __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = wdno::h2_sat(o[c * 8 + 2 * j], o[c * 8 + 2 * j + 1]);
      reinterpret_cast<uint4*>(op)[c] = ov;
    }
  }
}

}  // namespace wdno

using namespace wdno;

extern "C" int wdno_softmax_attn(const void* qkv, void* out, const float* bias, const float* rot_cos, const float* rot_sin,
                                 int64_t n_seq, int n_tok, int64_t inner, int64_t outerT, int64_t innerT, int64_t tokT,
                                 float scale, void* stream) {
  if (!qkv || !out || n_seq < 1 || n_tok < 1 || inner < 1) return set_error(WDNO_E_INVALID, "softmax_attn: bad arguments");
  if ((rot_cos == nullptr) != (rot_sin == nullptr)) return set_error(WDNO_E_INVALID, "softmax_attn: rotary tables must both be given");
  if (n_tok <= 32)  // short sequences (temporal attention): tensor-core kernel
    return launch_short_attn_mma(qkv, out, bias, rot_cos, rot_sin, n_seq, n_tok, inner, outerT, innerT, tokT, scale,
                                 static_cast<cudaStream_t>(stream));
  if (n_tok <= 512 && bias == nullptr && rot_cos == nullptr)  // mid-level spatial attention: tensor-core online softmax
    return launch_flash_attn_mma(qkv, out, n_seq, n_tok, inner, outerT, innerT, tokT, scale, static_cast<cudaStream_t>(stream));
  SeqMap m{inner, outerT, innerT, tokT};
  int qt = ((n_tok + 31) / 32) * 32;
  if (qt > 128) qt = 128;
  const int threads = qt * kHeads;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rot_cos)
    softmax_attn_kernel<true><<<static_cast<unsigned>(n_seq), threads, 0, st>>>(
        static_cast<const __half*>(qkv), static_cast<__half*>(out), bias, rot_cos, rot_sin, m, n_tok, scale);
  else
    softmax_attn_kernel<false><<<static_cast<unsigned>(n_seq), threads, 0, st>>>(
        static_cast<const __half*>(qkv), static_cast<__half*>(out), bias, rot_cos, rot_sin, m, n_tok, scale);
  return check_launch("softmax_attn");
}

extern "C" int wdno_linear_attn(const void* qkv, void* out, int64_t n_img, int n_pos, float scale, void* stream) {
  if (!qkv || !out || n_img < 1 || n_pos < 1) return set_error(WDNO_E_INVALID, "linear_attn: bad arguments");
  linear_attn_kernel<<<static_cast<unsigned>(n_img * kHeads), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(qkv), static_cast<__half*>(out), n_pos, scale);
  return check_launch("linear_attn");
}
