// Backward of the attention cores (training step).  The projections around them (to_qkv, to_out:
// conv3d.py:238-239,291-292) are 1x1 GEMMs handled by conv1x1 / wgrad; these kernels differentiate what lies between:
//   softmax attention (Attention.forward, conv3d.py:294-353): q' = rotary(scale q), k' = rotary(k), S = q' k'^T + bias,
//     P = softmax(S), O = P v   ->   dV = P^T dO, dP = dO V^T, dS = P o (dP - rowsum(P o dP)), dq' = dS k', dk' = dS^T q',
//     d bias += dS, inverse rotation of dq', dk';
//   linear attention (SpatialLinearAttention, conv3d.py:241-257; unet.py:203-222): q^ = softmax_d(q) * scale, k^ = softmax_n(k),
//     ctx = k^ v^T, out = ctx^T q^   ->   dctx = q^ dout^T, dq^ = ctx dout, dk^ = dctx v, dv = dctx^T k^, then the two softmax
//     backward passes (the sum over n that k's softmax needs is closed-form: sum_n k^ dk^ = sum_e dctx o ctx).
// Inputs / outputs are fp16 token rows ([.., 384] qkv / dqkv, [.., 128] dO) in the layouts of the forward kernels
// (attention.cu); all arithmetic is fp32.  heads x dim_head = 4 x 32.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/wdno_b200.h"
#include "common.cuh"
#include "cvt_sat.cuh"

namespace wdno {

namespace {

constexpr int kD = 32, kHid = 128, kQkv = 384;
constexpr int TS = 36;   // padded fp32 row: 144 B, 16-byte aligned (float4 broadcast reads, conflict-free float4 row writes)

struct SeqMap2 {
  long long inner, outerT, innerT, tokT;
};
__device__ __forceinline__ long long tok_idx(const SeqMap2& m, long long s, int t) {
  return (s / m.inner) * m.outerT + (s % m.inner) * m.innerT + static_cast<long long>(t) * m.tokT;
}
__device__ __forceinline__ void ld32h(const __half* p, float (&f)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + c);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 t = __half22float2(h[k]);
      f[c * 8 + 2 * k] = t.x;
      f[c * 8 + 2 * k + 1] = t.y;
    }
  }
}
__device__ __forceinline__ void st32h(__half* p, const float (&f)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 v;
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = h2_sat(f[c * 8 + 2 * k], f[c * 8 + 2 * k + 1]);
    reinterpret_cast<uint4*>(p)[c] = v;
  }
}

// ------------------------------------------------------------------ short sequences (n <= 32): one warp per (sequence, head)
// lane = token row.  Shared memory per warp: q', k', v, dO rows [32][33] fp32, P and dS [32][33], d-bias accumulator [32][33].
// per warp: 7 arrays of n rows

template <bool ROT>
__global__ void __launch_bounds__(128) short_attn_bwd_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dO,
                                                             const float* __restrict__ bias, const float* __restrict__ rc,
                                                             const float* __restrict__ rs, __half* __restrict__ dqkv,
                                                             float* __restrict__ dbias, SeqMap2 map, long long n_seq, int n,
                                                             float scale) {
  extern __shared__ float sm[];
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int RT = n * TS;    // floats per array (n rows)
  float* Q = sm + h * 7 * RT;
  float* K = Q + RT;
  float* V = K + RT;
  float* G = V + RT;        // dO
  float* P = G + RT;
  float* S = P + RT;        // dS (and scratch)
  float* Bacc = S + RT;
  const bool act = lane < n;
  if (act)
    for (int j = 0; j < n; ++j) Bacc[lane * TS + j] = 0.f;
  float cs[16], sn[16];
#pragma unroll
  for (int f = 0; f < 16; ++f) {
    cs[f] = (ROT && act) ? __ldg(rc + lane * 16 + f) : 1.f;
    sn[f] = (ROT && act) ? __ldg(rs + lane * 16 + f) : 0.f;
  }
  for (long long s = blockIdx.x; s < n_seq; s += gridDim.x) {
    const long long tk = act ? tok_idx(map, s, lane) : 0;
    float q[32], g[32];
    {
      float k[32], v[32];
#pragma unroll
      for (int d = 0; d < 32; ++d) { q[d] = 0.f; k[d] = 0.f; v[d] = 0.f; g[d] = 0.f; }
      if (act) {
        const __half* base = qkv + tk * kQkv + h * kD;
        ld32h(base, q);
        ld32h(base + kHid, k);
        ld32h(base + 2 * kHid, v);
        ld32h(dO + tk * kHid + h * kD, g);
#pragma unroll
        for (int f = 0; f < 16; ++f) {
          const float a = q[2 * f] * scale, b = q[2 * f + 1] * scale;
          q[2 * f] = a * cs[f] - b * sn[f];
          q[2 * f + 1] = b * cs[f] + a * sn[f];
          const float c = k[2 * f], e = k[2 * f + 1];
          k[2 * f] = c * cs[f] - e * sn[f];
          k[2 * f + 1] = e * cs[f] + c * sn[f];
        }
      }
      __syncwarp();   // the previous item's column pass is done with the tiles
      if (act) {
#pragma unroll
        for (int d = 0; d < 32; d += 4) {
          *reinterpret_cast<float4*>(Q + lane * TS + d) = make_float4(q[d], q[d + 1], q[d + 2], q[d + 3]);
          *reinterpret_cast<float4*>(K + lane * TS + d) = make_float4(k[d], k[d + 1], k[d + 2], k[d + 3]);
          *reinterpret_cast<float4*>(V + lane * TS + d) = make_float4(v[d], v[d + 1], v[d + 2], v[d + 3]);
          *reinterpret_cast<float4*>(G + lane * TS + d) = make_float4(g[d], g[d + 1], g[d + 2], g[d + 3]);
        }
      }
    }
    __syncwarp();
    float dq[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) dq[d] = 0.f;
    if (act) {
      float m = -INFINITY;
      for (int j = 0; j < n; ++j) {
        float sc = 0.f;
#pragma unroll
        for (int d = 0; d < 32; d += 4) {
          const float4 kv = *reinterpret_cast<const float4*>(K + j * TS + d);   // broadcast
          sc = fmaf(q[d], kv.x, fmaf(q[d + 1], kv.y, fmaf(q[d + 2], kv.z, fmaf(q[d + 3], kv.w, sc))));
        }
        if (bias != nullptr) sc += __ldg(bias + (static_cast<size_t>(h) * n + lane) * n + j);
        P[lane * TS + j] = sc;
        m = fmaxf(m, sc);
      }
      float l = 0.f;
      for (int j = 0; j < n; ++j) {
        const float p = __expf(P[lane * TS + j] - m);
        P[lane * TS + j] = p;
        l += p;
      }
      const float il = 1.f / l;
      float delta = 0.f;
      for (int j = 0; j < n; ++j) {
        const float p = P[lane * TS + j] * il;
        P[lane * TS + j] = p;
        float dp = 0.f;
#pragma unroll
        for (int d = 0; d < 32; d += 4) {
          const float4 vv = *reinterpret_cast<const float4*>(V + j * TS + d);
          dp = fmaf(g[d], vv.x, fmaf(g[d + 1], vv.y, fmaf(g[d + 2], vv.z, fmaf(g[d + 3], vv.w, dp))));
        }
        S[lane * TS + j] = dp;
        delta = fmaf(p, dp, delta);
      }
      for (int j = 0; j < n; ++j) {
        const float ds = P[lane * TS + j] * (S[lane * TS + j] - delta);
        S[lane * TS + j] = ds;
        Bacc[lane * TS + j] += ds;
#pragma unroll
        for (int d = 0; d < 32; d += 4) {
          const float4 kv = *reinterpret_cast<const float4*>(K + j * TS + d);
          dq[d] = fmaf(ds, kv.x, dq[d]);
          dq[d + 1] = fmaf(ds, kv.y, dq[d + 1]);
          dq[d + 2] = fmaf(ds, kv.z, dq[d + 2]);
          dq[d + 3] = fmaf(ds, kv.w, dq[d + 3]);
        }
      }
    }
    __syncwarp();
    if (act) {
      // column pass: this lane is key / value row j = lane
      float dk[32], dv[32];
#pragma unroll
      for (int d = 0; d < 32; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
      for (int i = 0; i < n; ++i) {
        const float ds = S[i * TS + lane], p = P[i * TS + lane];
#pragma unroll
        for (int d = 0; d < 32; d += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(Q + i * TS + d);
          const float4 gv = *reinterpret_cast<const float4*>(G + i * TS + d);
          dk[d] = fmaf(ds, qv.x, dk[d]);
          dk[d + 1] = fmaf(ds, qv.y, dk[d + 1]);
          dk[d + 2] = fmaf(ds, qv.z, dk[d + 2]);
          dk[d + 3] = fmaf(ds, qv.w, dk[d + 3]);
          dv[d] = fmaf(p, gv.x, dv[d]);
          dv[d + 1] = fmaf(p, gv.y, dv[d + 1]);
          dv[d + 2] = fmaf(p, gv.z, dv[d + 2]);
          dv[d + 3] = fmaf(p, gv.w, dv[d + 3]);
        }
      }
      // inverse rotation (the rotation matrix is orthogonal), then the q scale
#pragma unroll
      for (int f = 0; f < 16; ++f) {
        const float a = dq[2 * f], b = dq[2 * f + 1];
        dq[2 * f] = (a * cs[f] + b * sn[f]) * scale;
        dq[2 * f + 1] = (b * cs[f] - a * sn[f]) * scale;
        const float c = dk[2 * f], e = dk[2 * f + 1];
        dk[2 * f] = c * cs[f] + e * sn[f];
        dk[2 * f + 1] = e * cs[f] - c * sn[f];
      }
      __half* ob = dqkv + tk * kQkv + h * kD;
      st32h(ob, dq);
      st32h(ob + kHid, dk);
      st32h(ob + 2 * kHid, dv);
    }
  }
  __syncwarp();
  if (dbias != nullptr && act)
    for (int j = 0; j < n; ++j) atomicAdd(dbias + (static_cast<size_t>(h) * n + lane) * n + j, Bacc[lane * TS + j]);
}

// ------------------------------------------------------------------ longer sequences (32 < n <= 512): one block per (sequence, head)
// q, k, v, dO rows staged as fp32 [n][33] would not fit: fp16 rows of 40 halfs (80 B pitch); per-row m, l, delta in fp32.
constexpr int HP = 40;

__device__ __forceinline__ void row_from_smem(const __half* r, float (&f)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 v = *reinterpret_cast<const uint4*>(r + c * 8);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 t = __half22float2(h[k]);
      f[c * 8 + 2 * k] = t.x;
      f[c * 8 + 2 * k + 1] = t.y;
    }
  }
}

__global__ void __launch_bounds__(128) long_attn_bwd_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dO,
                                                            __half* __restrict__ dqkv, SeqMap2 map, int n, float scale) {
  extern __shared__ __align__(16) uint8_t smraw[];
  __half* Q = reinterpret_cast<__half*>(smraw);
  __half* K = Q + static_cast<size_t>(n) * HP;
  __half* V = K + static_cast<size_t>(n) * HP;
  __half* G = V + static_cast<size_t>(n) * HP;
  float* M = reinterpret_cast<float*>(G + static_cast<size_t>(n) * HP);
  float* Li = M + n;
  float* Dl = Li + n;
  const long long s = blockIdx.x >> 2;
  const int h = blockIdx.x & 3;
  for (int e = threadIdx.x; e < n * 4; e += blockDim.x) {
    const int t = e >> 2, c = e & 3;
    const long long tk = tok_idx(map, s, t);
    const __half* base = qkv + tk * kQkv + h * kD + c * 8;
    *reinterpret_cast<uint4*>(Q + t * HP + c * 8) = __ldg(reinterpret_cast<const uint4*>(base));
    *reinterpret_cast<uint4*>(K + t * HP + c * 8) = __ldg(reinterpret_cast<const uint4*>(base + kHid));
    *reinterpret_cast<uint4*>(V + t * HP + c * 8) = __ldg(reinterpret_cast<const uint4*>(base + 2 * kHid));
    *reinterpret_cast<uint4*>(G + t * HP + c * 8) = __ldg(reinterpret_cast<const uint4*>(dO + tk * kHid + h * kD + c * 8));
  }
  __syncthreads();
  // phase A: per query row: m, l (online), O -> delta = dO . O
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float q[32], g[32], o[32], kv[32];
    row_from_smem(Q + i * HP, q);
    row_from_smem(G + i * HP, g);
#pragma unroll
    for (int d = 0; d < 32; ++d) { q[d] *= scale; o[d] = 0.f; }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n; ++j) {
      row_from_smem(K + j * HP, kv);
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) sc = fmaf(q[d], kv[d], sc);
      const float mn = fmaxf(m, sc);
      const float corr = __expf(m - mn), p = __expf(sc - mn);
      l = l * corr + p;
      row_from_smem(V + j * HP, kv);
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] = fmaf(p, kv[d], o[d] * corr);
      m = mn;
    }
    float dl = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) dl = fmaf(g[d], o[d], dl);
    M[i] = m;
    Li[i] = 1.f / l;
    Dl[i] = dl / l;
  }
  __syncthreads();
  // phase B: dq_i = scale * sum_j dS_ij k_j
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float q[32], g[32], dq[32], kv[32];
    row_from_smem(Q + i * HP, q);
    row_from_smem(G + i * HP, g);
#pragma unroll
    for (int d = 0; d < 32; ++d) { q[d] *= scale; dq[d] = 0.f; }
    const float m = M[i], il = Li[i], dl = Dl[i];
    for (int j = 0; j < n; ++j) {
      row_from_smem(V + j * HP, kv);
      float dp = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) dp = fmaf(g[d], kv[d], dp);
      row_from_smem(K + j * HP, kv);
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) sc = fmaf(q[d], kv[d], sc);
      const float ds = __expf(sc - m) * il * (dp - dl);
#pragma unroll
      for (int d = 0; d < 32; ++d) dq[d] = fmaf(ds, kv[d], dq[d]);
    }
#pragma unroll
    for (int d = 0; d < 32; ++d) dq[d] *= scale;
    st32h(dqkv + tok_idx(map, s, i) * kQkv + h * kD, dq);
  }
  // phase C: per key row j: dk_j = sum_i dS_ij q'_i, dv_j = sum_i P_ij dO_i
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float kj[32], vj[32], dk[32], dv[32], r[32];
    row_from_smem(K + j * HP, kj);
    row_from_smem(V + j * HP, vj);
#pragma unroll
    for (int d = 0; d < 32; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
    for (int i = 0; i < n; ++i) {
      row_from_smem(Q + i * HP, r);
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) sc = fmaf(r[d], kj[d], sc);
      const float p = __expf(sc * scale - M[i]) * Li[i];
      float qs[32];
#pragma unroll
      for (int d = 0; d < 32; ++d) qs[d] = r[d] * scale;
      row_from_smem(G + i * HP, r);
      float dp = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) dp = fmaf(r[d], vj[d], dp);
      const float ds = p * (dp - Dl[i]);
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        dk[d] = fmaf(ds, qs[d], dk[d]);
        dv[d] = fmaf(p, r[d], dv[d]);
      }
    }
    __half* ob = dqkv + tok_idx(map, s, j) * kQkv + h * kD;
    st32h(ob + kHid, dk);
    st32h(ob + 2 * kHid, dv);
  }
}

// ------------------------------------------------------------------ linear attention
// work buffer per image: kstat [4][32][2] (max, 1/sum of exp over n), ctx [4][32][32], dctx [4][32][32]
constexpr int kLaWork = 4 * 32 * 2 + 2 * 4 * 32 * 32;

// (1) softmax statistics of k over the positions: one warp per (image, head), lane = d
__global__ void __launch_bounds__(128) la_bwd_kstat_kernel(const __half* __restrict__ qkv, float* __restrict__ work, int n_pos) {
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long img = blockIdx.x;
  const __half* kp = qkv + img * n_pos * static_cast<long long>(kQkv) + kHid + h * kD + lane;
  float mm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, ll[4] = {0.f, 0.f, 0.f, 0.f};
  int p = 0;
  for (; p + 4 <= n_pos; p += 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __half2float(__ldg(kp + static_cast<size_t>(p + u) * kQkv));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float mn = fmaxf(mm[u], v[u]);
      ll[u] = ll[u] * __expf(mm[u] - mn) + __expf(v[u] - mn);
      mm[u] = mn;
    }
  }
  for (; p < n_pos; ++p) {
    const float v = __half2float(__ldg(kp + static_cast<size_t>(p) * kQkv));
    const float mn = fmaxf(mm[0], v);
    ll[0] = ll[0] * __expf(mm[0] - mn) + __expf(v - mn);
    mm[0] = mn;
  }
  const float m = fmaxf(fmaxf(mm[0], mm[1]), fmaxf(mm[2], mm[3]));
  float l = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) l += (mm[u] == -INFINITY) ? 0.f : ll[u] * __expf(mm[u] - m);
  float* st = work + img * kLaWork + (h * 32 + lane) * 2;
  st[0] = m;
  st[1] = 1.f / l;
}

// (1) is latency bound when one warp walks all positions of an image serially; four independent (max, sum) chains per lane
// (2) ctx[d][e] += sum_p k^[d] v[e],  dctx[d][e] += sum_p q^[d] dout[e]  over a chunk of positions.
// block = 4 warps = 4 heads; lane = d; each lane accumulates its row d of both 32x32 matrices (64 registers); the per-position
// vectors v and dout are staged in the warp's shared-memory row and read back as float4 broadcasts (8 loads instead of 32
// shuffles each).
__global__ void __launch_bounds__(128) la_bwd_ctx_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dout,
                                                         float* __restrict__ work, int n_pos, int chunk, float scale) {
  __shared__ __align__(16) float stv[4][2][2][32];   // [warp][buffer][v | dout][e]
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long img = blockIdx.y;
  float* wk = work + img * kLaWork;
  const float km = wk[(h * 32 + lane) * 2], kil = wk[(h * 32 + lane) * 2 + 1];
  float cx[32], dc[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) { cx[e] = 0.f; dc[e] = 0.f; }
  const int p0 = blockIdx.x * chunk, p1 = min(n_pos, p0 + chunk);
  for (int pb = p0; pb < p1; pb += 4) {
    // the loop is latency bound on its four 64-byte loads per position: fetch four positions at once
    float qr4[4], kr4[4], vr4[4], gr4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = min(pb + u, p1 - 1);
      const __half* row = qkv + (img * n_pos + p) * static_cast<long long>(kQkv) + h * kD;
      qr4[u] = __half2float(__ldg(row + lane));
      kr4[u] = __half2float(__ldg(row + kHid + lane));
      vr4[u] = __half2float(__ldg(row + 2 * kHid + lane));
      gr4[u] = __half2float(__ldg(dout + (img * n_pos + p) * static_cast<long long>(kHid) + h * kD + lane));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (pb + u >= p1) break;
      const float qr = qr4[u], kr = kr4[u];
      const int bf = u & 1;
      stv[h][bf][0][lane] = vr4[u];
      stv[h][bf][1][lane] = gr4[u];
      // q^ = softmax over d (lanes) * scale
      float mx = qr;
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
      const float ex = __expf(qr - mx);
      float sum = ex;
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
      const float qh = ex / sum * scale;
      const float kh = __expf(kr - km) * kil;
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(&stv[h][bf][0][e]);
        const float4 g4 = *reinterpret_cast<const float4*>(&stv[h][bf][1][e]);
        cx[e] = fmaf(kh, v4.x, cx[e]);
        cx[e + 1] = fmaf(kh, v4.y, cx[e + 1]);
        cx[e + 2] = fmaf(kh, v4.z, cx[e + 2]);
        cx[e + 3] = fmaf(kh, v4.w, cx[e + 3]);
        dc[e] = fmaf(qh, g4.x, dc[e]);
        dc[e + 1] = fmaf(qh, g4.y, dc[e + 1]);
        dc[e + 2] = fmaf(qh, g4.z, dc[e + 2]);
        dc[e + 3] = fmaf(qh, g4.w, dc[e + 3]);
      }
    }
  }
  float* cxo = wk + 4 * 32 * 2 + (h * 32 + lane) * 32;
  float* dco = cxo + 4 * 32 * 32;
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    atomicAdd(cxo + e, cx[e]);
    atomicAdd(dco + e, dc[e]);
  }
}

// (3) per position: dq, dk, dv.  block = 4 warps = 4 heads, lane = d (and = e for dv); ctx / dctx of the image in shared memory,
// the position's v / dout / k^ vectors staged per warp (float4 broadcast reads).
__global__ void __launch_bounds__(128) la_bwd_pos_kernel(const __half* __restrict__ qkv, const __half* __restrict__ dout,
                                                         const float* __restrict__ work, __half* __restrict__ dqkv, int n_pos,
                                                         int chunk, float scale) {
  constexpr int TS2 = 36;
  extern __shared__ __align__(16) float la_sm[];
  float (*cxs)[32 * TS2] = reinterpret_cast<float (*)[32 * TS2]>(la_sm);                  // ctx   [4][32][36]
  float (*dcs)[32 * TS2] = reinterpret_cast<float (*)[32 * TS2]>(la_sm + 4 * 32 * TS2);   // dctx
  float (*dct)[32 * TS2] = reinterpret_cast<float (*)[32 * TS2]>(la_sm + 8 * 32 * TS2);   // dctx^T
  float (*stv)[2][3][32] = reinterpret_cast<float (*)[2][3][32]>(la_sm + 12 * 32 * TS2);  // [warp][buffer][v | dout | k^][.]
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long img = blockIdx.y;
  const float* wk = work + img * kLaWork;
  const float km = wk[(h * 32 + lane) * 2], kil = wk[(h * 32 + lane) * 2 + 1];
  const float* cxi = wk + 4 * 32 * 2 + h * 32 * 32;
  const float* dci = cxi + 4 * 32 * 32;
  for (int i = lane; i < 32 * 32; i += 32) {
    cxs[h][(i >> 5) * TS2 + (i & 31)] = cxi[i];
    dcs[h][(i >> 5) * TS2 + (i & 31)] = dci[i];
    dct[h][(i & 31) * TS2 + (i >> 5)] = dci[i];
  }
  __syncwarp();
  // t[d] = sum_e dctx[d][e] ctx[d][e]  (= sum over n of k^ dk^)
  float td = 0.f;
#pragma unroll
  for (int e = 0; e < 32; ++e) td = fmaf(dcs[h][lane * TS2 + e], cxs[h][lane * TS2 + e], td);
  const int p0 = blockIdx.x * chunk, p1 = min(n_pos, p0 + chunk);
  for (int pb = p0; pb < p1; pb += 4) {
    float qr4[4], kr4[4], vr4[4], gr4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long tk = img * n_pos + min(pb + u, p1 - 1);
      const __half* row = qkv + tk * kQkv + h * kD;
      qr4[u] = __half2float(__ldg(row + lane));
      kr4[u] = __half2float(__ldg(row + kHid + lane));
      vr4[u] = __half2float(__ldg(row + 2 * kHid + lane));
      gr4[u] = __half2float(__ldg(dout + tk * kHid + h * kD + lane));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (pb + u >= p1) break;
      const long long tk = img * n_pos + pb + u;
      const float qr = qr4[u];
      const float kh = __expf(kr4[u] - km) * kil;  // softmax over n
      const int bf = u & 1;
      stv[h][bf][0][lane] = vr4[u];
      stv[h][bf][1][lane] = gr4[u];
      stv[h][bf][2][lane] = kh;
      float mx = qr;
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));
      const float ex = __expf(qr - mx);
      float sum = ex;
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sh);
      const float qs = ex / sum;               // softmax over d (unscaled)
      __syncwarp();
      float dqh = 0.f, dkh = 0.f, dv = 0.f;
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(&stv[h][bf][0][e]);
        const float4 g4 = *reinterpret_cast<const float4*>(&stv[h][bf][1][e]);
        const float4 k4 = *reinterpret_cast<const float4*>(&stv[h][bf][2][e]);
        const float4 c4 = *reinterpret_cast<const float4*>(&cxs[h][lane * TS2 + e]);
        const float4 d4 = *reinterpret_cast<const float4*>(&dcs[h][lane * TS2 + e]);
        const float4 t4 = *reinterpret_cast<const float4*>(&dct[h][lane * TS2 + e]);   // dctx[d = e..e+3][e' = lane]
        dqh = fmaf(c4.x, g4.x, fmaf(c4.y, g4.y, fmaf(c4.z, g4.z, fmaf(c4.w, g4.w, dqh))));   // dq^[d] = sum_e ctx[d][e] dout[e]
        dkh = fmaf(d4.x, v4.x, fmaf(d4.y, v4.y, fmaf(d4.z, v4.z, fmaf(d4.w, v4.w, dkh))));   // dk^[d] = sum_e dctx[d][e] v[e]
        dv = fmaf(t4.x, k4.x, fmaf(t4.y, k4.y, fmaf(t4.z, k4.z, fmaf(t4.w, k4.w, dv))));     // dv[e'] = sum_d k^[d] dctx[d][e']
      }
      // softmax over d backward (q = scale * qs): dq_raw = scale * qs * (dq^ - sum_d qs dq^)
      float dot = qs * dqh;
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, sh);
      const float dq = scale * qs * (dqh - dot);
      const float dk = kh * (dkh - td);
      __half* ob = dqkv + tk * kQkv + h * kD + lane;
      ob[0] = h_sat(dq);
      ob[kHid] = h_sat(dk);
      ob[2 * kHid] = h_sat(dv);
    }
  }
}

}  // namespace

}  // namespace wdno

using namespace wdno;

extern "C" int wdno_softmax_attn_bwd(const void* qkv, const void* d_out, const float* bias, const float* rot_cos,
                                     const float* rot_sin, void* dqkv, float* dbias, int64_t n_seq, int n_tok, int64_t inner,
                                     int64_t outerT, int64_t innerT, int64_t tokT, float scale, void* stream) {
  if (!qkv || !d_out || !dqkv || n_seq < 1 || n_tok < 1 || inner < 1) return set_error(WDNO_E_INVALID, "softmax_attn_bwd: bad arguments");
  if ((rot_cos == nullptr) != (rot_sin == nullptr)) return set_error(WDNO_E_INVALID, "softmax_attn_bwd: rotary tables must both be given");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SeqMap2 m{inner, outerT, innerT, tokT};
  if (n_tok <= 32) {
    const size_t smem = static_cast<size_t>(4) * 7 * n_tok * TS * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
      cudaError_t e = cudaFuncSetAttribute(short_attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e == cudaSuccess) e = cudaFuncSetAttribute(short_attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return set_cuda_error(e, "softmax_attn_bwd: cudaFuncSetAttribute");
      configured = smem;
    }
    long long grid = static_cast<long long>(num_sms()) * 2;
    if (grid > n_seq) grid = n_seq;
    if (rot_cos)
      short_attn_bwd_kernel<true><<<static_cast<unsigned>(grid), 128, smem, st>>>(
          static_cast<const __half*>(qkv), static_cast<const __half*>(d_out), bias, rot_cos, rot_sin, static_cast<__half*>(dqkv),
          dbias, m, n_seq, n_tok, scale);
    else
      short_attn_bwd_kernel<false><<<static_cast<unsigned>(grid), 128, smem, st>>>(
          static_cast<const __half*>(qkv), static_cast<const __half*>(d_out), bias, rot_cos, rot_sin, static_cast<__half*>(dqkv),
          dbias, m, n_seq, n_tok, scale);
    return check_launch("softmax_attn_bwd");
  }
  if (n_tok > 512 || bias != nullptr || rot_cos != nullptr || dbias != nullptr)
    return set_error(WDNO_E_INVALID, "softmax_attn_bwd: sequences longer than 32 tokens support neither bias nor rotary (<= 512 tokens)");
  if (n_seq * 4 > 2147483647LL) return set_error(WDNO_E_INVALID, "softmax_attn_bwd: too many sequences");
  const size_t smem = static_cast<size_t>(n_tok) * (4 * HP * 2 + 3 * 4);
  static size_t configured_l = 0;
  if (smem > configured_l) {
    cudaError_t e = cudaFuncSetAttribute(long_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e, "softmax_attn_bwd: cudaFuncSetAttribute");
    configured_l = smem;
  }
  long_attn_bwd_kernel<<<static_cast<unsigned>(n_seq * 4), 128, smem, st>>>(
      static_cast<const __half*>(qkv), static_cast<const __half*>(d_out), static_cast<__half*>(dqkv), m, n_tok, scale);
  return check_launch("softmax_attn_bwd");
}

extern "C" int64_t wdno_linear_attn_bwd_work_bytes(int64_t n_img) {
  if (n_img < 1) return WDNO_E_INVALID;
  return n_img * static_cast<int64_t>(kLaWork) * 4;
}

extern "C" int wdno_linear_attn_bwd(const void* qkv, const void* d_out, void* dqkv, void* work, int64_t n_img, int n_pos,
                                    float scale, void* stream) {
  if (!qkv || !d_out || !dqkv || !work || n_img < 1 || n_pos < 1) return set_error(WDNO_E_INVALID, "linear_attn_bwd: bad arguments");
  if (n_img > 65535) return set_error(WDNO_E_INVALID, "linear_attn_bwd: more than 65535 images");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(work, 0, static_cast<size_t>(n_img) * kLaWork * 4, st);
  if (e != cudaSuccess) return set_cuda_error(e, "linear_attn_bwd: memset");
  la_bwd_kstat_kernel<<<static_cast<unsigned>(n_img), 128, 0, st>>>(static_cast<const __half*>(qkv), static_cast<float*>(work), n_pos);
  const int chunk = 64;
  const int nch = (n_pos + chunk - 1) / chunk;
  la_bwd_ctx_kernel<<<dim3(nch, static_cast<unsigned>(n_img)), 128, 0, st>>>(
      static_cast<const __half*>(qkv), static_cast<const __half*>(d_out), static_cast<float*>(work), n_pos, chunk, scale);
  constexpr int kPosSmem = (12 * 32 * 36 + 4 * 2 * 3 * 32) * 4;
  static bool configured = false;
  if (!configured) {
    cudaError_t e2 = cudaFuncSetAttribute(la_bwd_pos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPosSmem);
    if (e2 != cudaSuccess) return set_cuda_error(e2, "linear_attn_bwd: cudaFuncSetAttribute");
    configured = true;
  }
  la_bwd_pos_kernel<<<dim3(nch, static_cast<unsigned>(n_img)), 128, kPosSmem, st>>>(
      static_cast<const __half*>(qkv), static_cast<const __half*>(d_out), static_cast<const float*>(work),
      static_cast<__half*>(dqkv), n_pos, chunk, scale);
  return check_launch("linear_attn_bwd");
}
