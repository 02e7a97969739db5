/*
 * wdno_b200 -- C ABI of the B200-native WDNO sampling engine (libwdno_b200.so).
 *
 * Drop-in boundary for the one hot path of AI4Science-WestlakeU/wdno: the DDPM/DDIM denoising
 * loop over wavelet-coefficient tensors and the separable DWT/IDWT around it.  The reference has
 * no FFI of its own (its boundary is Python classes), so each entry point cites the reference
 * Python function whose arithmetic it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - no function allocates or frees caller memory; nothing is retained after return except
 *     work enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, <0 = WDNO_E_*; wdno_last_error() gives a thread-local message;
 *   - activations inside the U-Net are fp16, channels-last:  [B, D(frames), H, W, C];
 *     the public diffusion state is fp32 in the reference layout [B, F, C, H, W].
 */
#ifndef WDNO_B200_H
#define WDNO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WDNO_OK 0
#define WDNO_E_INVALID (-1)   /* bad argument / unsupported shape */
#define WDNO_E_CUDA (-2)      /* CUDA runtime error (message has cudaGetErrorString) */
#define WDNO_E_NO_DEVICE (-3) /* no sm_100 device */

const char* wdno_last_error(void);
int wdno_version(void);
/* compute capability major*10+minor of the current device, or <0 */
int wdno_device_cc(void);

/* ------------------------------------------------------------------------------------------
 * tap-GEMM: the tcgen05 implicit-GEMM engine behind every dense contraction of both U-Nets
 * (Conv3d 3x3x3 / 7x7x7 / 1x1x1, Conv2d, Linear, strided (1,4,4) conv, ConvTranspose3d (1,4,4)):
 *   reference: smoke/video_diffusion_pytorch/video_diffusion_pytorch_conv3d.py:159-163,189-230,
 *              238-239,291-292,393,469-472 ; burgers/ddpm_burgers/unet.py:35-45,129-181,190-194,
 *              233-234,317,336,361,369
 *   out[b,z,y,x,n] = sum_{tap,c} act(in[b, z+dz, y+dy, x+dx, c]) * w[n, tap, c]  (+bias, +resid)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t kz;    /* plane offset of this tap, 0..KD-1 */
  int32_t shift; /* ky * Wp + kx, in padded-row positions */
} wdno_tap;

typedef struct {
  int32_t src;       /* which source tensor (0/1): concat along channels */
  int32_t ch_off;    /* first channel inside that source */
  int32_t ph_y, ph_x; /* space-to-depth phase (src_mode 1) */
  int32_t tap_begin; /* index into taps[] */
  int32_t tap_count;
} wdno_kset;

typedef struct {
  int32_t out_ch_off;  /* first output channel of this N-chunk */
  int32_t n_valid;     /* valid columns (<= N) */
  int32_t ph_y, ph_x;  /* depth-to-space phase (out_mode 1) */
  int32_t set_begin;   /* index into sets[] */
  int32_t set_count;
  int32_t n_tiles;     /* total weight tiles of this chunk = sum(tap_count) */
  int32_t pad_;
  int64_t w_tile_off;  /* first weight tile of this chunk inside wpacked */
} wdno_nchunk;

typedef struct {
  /* sources: fp16 channels-last */
  const void* src[2];
  int32_t src_c[2];       /* total channels of each source tensor */
  const float* coef_a[2]; /* optional per-(b,channel) affine: act = silu(a*x + c); NULL = identity */
  const float* coef_c[2];
  int32_t src_mode;       /* 0 direct; 1 space-to-depth (src grid 2H x 2W); 2 nearest-up (src grid H/2 x W/2) */
  /* tap grid == output grid (before depth-to-space) */
  int32_t B, D, H, W;
  int32_t KD, pz, py, px; /* input plane for tap kz is z + kz - pz; padded position = (y+py, x+px) */
  int32_t Wp;             /* padded row width: W + KW/2 (the zero run between rows is shared by both neighbours) */
  int32_t maxshift;       /* largest tap shift */
  int32_t ZT, PT;         /* accumulators per unit: ZT planes x PT 128-position tiles (ZT*PT <= 4) */
  int32_t KC;             /* channels per K-set: 16, 32 or 64 */
  int32_t N;              /* MMA N: multiple of 16, <= 128 */
  int32_t n_chunks;
  const wdno_nchunk* chunks;
  const wdno_kset* sets;
  const wdno_tap* taps;
  const void* wpacked;    /* fp16 tiles [tile][KC/8][N][8] */
  /* output */
  int32_t out_mode;       /* 0 fp16 channels-last; 1 fp16 depth-to-space (out grid 2H x 2W); 2 fp32 [B,D,C,H,W] */
  int32_t out_c;          /* total channels of the output tensor */
  void* out;
  const float* bias;      /* indexed by global output channel; NULL = none */
  const void* resid;      /* fp16, same layout as out (modes 0/1); NULL = none */
  double* stats;          /* optional GroupNorm partial sums [B][G][2] (sum, sumsq) += ; NULL = none */
  int32_t G, cpg;         /* groups, channels per group (multiple of 8) */
  /* shared-memory plan */
  int32_t NSLOT, NBST, S_pad;
  int32_t grid;           /* persistent CTAs to launch (<= #SM) */
  int32_t TPS;            /* weight tiles per bulk-copy stage */
  int32_t reuse;          /* 1: (1x1 layers) slabs of all K-sets stay resident while every N-chunk is computed */
  int32_t n_taps;         /* entries of taps[] (<= 384; copied to shared memory) */
  int32_t bias_len;       /* floats readable at bias (padded output channels, <= 1024; staged in shared memory) */
  int32_t n_sets;         /* entries of sets[] (<= 128; copied to shared memory) */
  int32_t zstack;         /* 1: weight tiles hold the KD depth taps of one in-plane tap stacked along N ([kz][64] rows);
                             taps[] lists in-plane taps only; one MMA per input plane updates up to 4 output planes */
  int32_t strips;         /* >= 1.  > 1: every plane is cut into `strips` column strips of W tap-grid columns each; a strip is
                             a work-item dimension of its own, its slab carries real neighbour columns as halo
                             (Wp = W + 2*px) instead of the shared zero run.  Keeps the haloed slab small for wide planes
                             with large in-plane kernels (7x7 on 80x80).  Only src_mode 0, out_mode 0/2. */
  int32_t Wfull;          /* row width of the source / output tensors (== W when strips == 1) */
  int32_t fold;           /* 1: the D planes of a "sample" are D independent samples of a 2-D layer folded into the depth
                             axis (KD == 1 only) so that ZT of them share every weight tile: GroupNorm coefficients and
                             statistics are indexed by b*D + z instead of b */
  int32_t cluster;        /* 0 / 1: independent CTAs.  2: CTA pairs (thread-block clusters of 2) that fetch every weight tile ONCE with
                             a multicast bulk copy; needs an even grid <= wdno_tapgemm_max_cluster_ctas(), an even number of
                             (sample, plane group, tile) units and reuse == 0 */
} wdno_tapgemm_params;

/* bytes of dynamic shared memory the plan needs, or <0 */
int64_t wdno_tapgemm_smem_bytes(const wdno_tapgemm_params* p);
/* CTAs of 2-CTA clusters that can be resident at once with smem_bytes of dynamic shared memory each, or <0 */
int wdno_tapgemm_max_cluster_ctas(int64_t smem_bytes);
int wdno_tapgemm(const wdno_tapgemm_params* p, void* stream);

/* Plain 1x1 convolution / Linear (no fused GroupNorm prologue or statistics): the HBM-bound layers -- ResnetBlock.res_conv
 * (conv3d.py:216 ; unet.py:162), attention to_qkv / to_out (conv3d.py:291-292 ; unet.py:190-192,233-234) and the final
 * 1x1 convolution (conv3d.py:471 ; unet.py:369).
 *   out[m][n] = sum_k concat(src0[m][0..c0), src1[m][0..c1))[k] * w[n][k] + bias[n] (+ resid[m][n])
 * src*: fp16 channels-last [M][c*] (c0, c1 multiples of 32; src1 may be NULL with c1 = 0); w: fp16 [npad][c0+c1], rows
 * >= cout zero, npad a multiple of 64; bias: fp32 [npad] or NULL; resid: fp16 [M][cout] or NULL.
 * out_mode 0: fp16 [M][cout] (cout % 8 == 0);  out_mode 2: fp32 planar [M/hw][cout][hw] -- the reference's
 * [B,F,C,H,W] layout with hw = H*W (no residual). */
int wdno_conv1x1(const void* src0, int c0, const void* src1, int c1, const void* w, int npad, const float* bias,
                 const void* resid, void* out, int64_t M, int cout, int out_mode, int64_t hw, void* stream);


/* ------------------------------------------------------------------------------------------
 * U-Net helper kernels (HBM-bound).  reference: conv3d.py:139-151,165-174,189-230,405-410 ;
 * burgers/ddpm_burgers/unet.py:55-65,82-108,129-181
 * ------------------------------------------------------------------------------------------ */
/* x fp32 [B,F,C,H,W] -> out fp16 channels-last [B,F,H,W,Cp] (channels C..Cp-1 zero) */
int wdno_pack_bfchw_f16(const float* x, void* out, int B, int F, int C, int H, int W, int Cp, void* stream);
/* GroupNorm statistics (double [B][G][2] = sum,sumsq over `count` elements) -> per-(b,channel) affine
 * a,c with  GN(y)*(scale+1)+shift == a*y + c ; scale_shift = [B][ss_stride] rows (scale[C] | shift[C]) or NULL */
int wdno_gn_finalize(const double* stats, const float* gamma, const float* beta, const float* scale_shift,
                     int ss_stride, float* a, float* c, int B, int C, int G, double count, float eps, void* stream);
/* out = silu(a*y + c) (+ resid) ; y,resid,out fp16 [B, vox_per_sample, C] */
int wdno_gn_silu_add(const void* y, const float* a, const float* c, const void* resid, void* out, int B, int C,
                     int64_t vox_per_sample, void* stream);
/* channel LayerNorm without bias: (x-mean)*rsqrt(var+eps)*gamma (+ resid) over C for each of nvox voxels (fp16 in/out) */
int wdno_chan_layernorm(const void* x, const float* gamma, const void* resid, void* out, int64_t nvox, int C, float eps,
                        void* stream);
/* emb = W2*gelu(W1*sinusoid(time)+b1)+b2 ; also writes silu(emb).  time fp32 [B] */
int wdno_time_mlp(const float* time, const float* w1, const float* b1, const float* w2, const float* b2, float* emb,
                  float* emb_silu, int B, int dim, int tdim, float theta, void* stream);
/* out[b][j] = bias[j] + sum_k in[b][k]*w[j][k] (fp32; all ResnetBlock.mlp Linear layers concatenated along j) */
int wdno_small_linear(const float* in, const float* w, const float* bias, float* out, int B, int K, int J, void* stream);

/* ------------------------------------------------------------------------------------------
 * attention cores (4 heads x 32).  reference: conv3d.py:232-258,277-353 ; unet.py:183-259
 * token index of (sequence s, token t) = (s / inner)*outerT + (s % inner)*innerT + t*tokT ;
 * qkv fp16 [tokens][384] (q|k|v, each 4x32), out fp16 [tokens][128]
 * ------------------------------------------------------------------------------------------ */
int wdno_softmax_attn(const void* qkv, void* out, const float* bias /*[4][n][n] or NULL*/,
                      const float* rot_cos /*[n][16] or NULL*/, const float* rot_sin, int64_t n_seq, int n_tok,
                      int64_t inner, int64_t outerT, int64_t innerT, int64_t tokT, float scale, void* stream);
/* qkv fp16 [n_img][n_pos][384] -> out fp16 [n_img][n_pos][128] */
int wdno_linear_attn(const void* qkv, void* out, int64_t n_img, int n_pos, float scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * fused attention blocks: Residual(PreNorm(dim, attention)) in one pass over the residual stream
 * (LayerNorm, qkv projection, attention, output projection and the residual add never leave the SM).
 * x, y: fp16 channels-last [n_img][n_pos][C] (C = 64, 128 or 256; x != y).  Weight operands are pre-packed
 * in mma.sync fragment order by wdno_b200/attn_fused.py (pack_a_frags / pack_b_frags).
 *   linattn_block: SpatialLinearAttention, reference conv3d.py:165-184,232-258.
 *     wq_pack  fp16 B-fragments of W_q [128][C];  wkv_pack fp16 A-fragments [k|v][head][32][C];
 *     wout fp32 [C][128]; bias fp32 [C] or NULL; work: wdno_linattn_work_bytes() bytes of scratch.
 *   tattn_block: temporal Attention over the frame axis (n_frames <= 32 tokens per pixel; rotary on q,k; T5 relative
 *     position bias), reference conv3d.py:74-112,165-184,262-353,383.  x,y: [n_samples][n_frames][hw][C], hw even.
 *     wqk_pack fp16 B-fragments of W_qk [256][C]; wv_pack fp16 A-fragments [head][32][C]; wout_pack fp16 B-fragments of
 *     W_out [C][128]; bias fp32 [4][n_frames][n_frames] or NULL; rot_cos/rot_sin fp32 [n_frames][16] or both NULL.
 */
int64_t wdno_linattn_work_bytes(int64_t n_img, int n_pos, int C);
int wdno_tattn_block(const void* x, void* y, const float* gamma, const void* wqk_pack, const void* wv_pack,
                     const void* wout_pack, const float* bias, const float* rot_cos, const float* rot_sin,
                     int64_t n_samples, int n_frames, int64_t hw, int C, float scale, float eps, void* stream);
/* ------------------------------------------------------------------------------------------ */
int wdno_linattn_block(const void* x, void* y, const float* gamma, const void* wq_pack, const void* wkv_pack,
                       const float* wout, const float* bias, void* work, int64_t n_img, int n_pos, int C,
                       float scale, float eps, void* stream);

/* all-tcgen05 form of tattn_block for C = 64 (csrc/tattn_row.cu): every product a 128-row UMMA, softmax in the registers of the
 * thread that owns the accumulator row.  wqkv_canon fp16 [C/8][384][8] = W_qkv diag(gamma) and wout_canon fp16 [16][C][8] = W_out in
 * the UMMA K-major operand order.  The caller folds into W_qkv (a) the LayerNorm gain (all rows) and (b) scale * log2(e) into the
 * q rows 0..127 -- the kernel's softmax runs in base 2 and multiplies nothing at run time; the `scale` argument is therefore not
 * applied again (kept for signature symmetry with wdno_tattn_block).  bias / rot_* as for wdno_tattn_block (natural-log units). */
int wdno_tattn_block_row(const void* x, void* y, const void* wqkv_canon, const void* wout_canon, const float* bias,
                         const float* rot_cos, const float* rot_sin, int64_t n_samples, int n_frames, int64_t hw, int C,
                         float scale, float eps, void* stream);

/* tcgen05 form of linattn_block (csrc/linattn_tc.cu; C = 64 or 128): same workspace, same result up to fp16 rounding.
 * The LayerNorm gain is folded into the weight operands by the caller (W[:, c] * gamma[c]):
 *   wq_canon  fp16 [C/8][128][8] : W_q diag(gamma) [128][C] in the UMMA K-major operand order (8-channel chunk, row, channel in chunk);
 *   wkv_canon fp16 [C/8][256][8] : rows 0..127 = W_k diag(gamma) (head, d), rows 128..255 = W_v diag(gamma) (head, e).
 * wdno_linattn_tc_supported() -> 1 when (n_img, n_pos, C) is inside its envelope, else 0 (use wdno_linattn_block). */
int wdno_linattn_tc_supported(int64_t n_img, int n_pos, int C);
int wdno_linattn_block_tc(const void* x, void* y, const void* wq_canon, const void* wkv_canon,
                          const float* wout, const float* bias, void* work, int64_t n_img, int n_pos, int C,
                          float scale, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * diffusion step algebra on the fp32 state [B,F,C,H,W] (Burgers: F=1).
 * reference: smoke/ddpm/diffusion_2d.py:689-699,723-754,769-785,851-933,970-976,988-1050 ;
 *            burgers/ddpm_burgers/diffusion_1d.py:172-182,205-258,276-307,310-460,520-645
 * A condition program is the reference's sequence of in-place slice assignments; later ops override earlier.
 * ------------------------------------------------------------------------------------------ */
#define WDNO_MAX_COND_OPS 8
typedef struct {
  int32_t f0, f1, c0, c1, y0, y1, x0, x1; /* destination box (half-open) */
  const float* src;                       /* NULL = fill with zero */
  int64_t sb, sf, sc, sy, sx;             /* source strides (elements) */
  int32_t of, oc, oy, ox;                 /* source index = dst index - offset */
} wdno_cond_op;

/* coef_dev (device float[8]): DDIM {sqrt_recip_ac[t], sqrt_recipm1_ac[t], sqrt(ac[t_next]), c, sigma, last_flag, gscale, 0}
 *                             DDPM {sqrt_recip_ac[t], sqrt_recipm1_ac[t], post_mean_coef1, post_mean_coef2, exp(.5*logvar)|0, 0, gscale, 0}
 * cond_mode: 0 never, 1 all but the last step (smoke), 2 always (Burgers).  noise / guidance may be NULL. */
int wdno_ddim_step(float* x, const float* eps, const float* noise, const float* guidance, const float* coef_dev,
                   const wdno_cond_op* ops_host, int n_ops, int B, int F, int C, int H, int W, int cond_mode, void* stream);
int wdno_ddpm_step(float* x, const float* eps, const float* noise, const float* guidance, const float* coef_dev,
                   const wdno_cond_op* ops_host, int n_ops, int B, int F, int C, int H, int W, int cond_mode, void* stream);
int wdno_apply_conditions(float* x, const wdno_cond_op* ops_host, int n_ops, int B, int F, int C, int H, int W, void* stream);
int wdno_predict_x0(const float* x, const float* eps, const float* coef_dev, float* x0, int64_t total, int clip, void* stream);
int wdno_q_sample(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac, const int64_t* t,
                  float* out, int B, int64_t per_sample, void* stream);
/* acc[b] += sum (pred-target)^2 * w[c]   (w NULL, length 1 or length C) */
int wdno_mse_weighted(const float* pred, const float* target, const float* w, int w_len, int B, int F, int C, int H, int W,
                      double* acc, void* stream);
/* CUDA-graph support: copy step *step_dev's scalars (time_table[s], coef_table[s][8]) into time_out[B] / coef_out[8], ++step */
int wdno_step_begin(int* step_dev, const float* time_table, const float* coef_table, float* time_out, float* coef_out,
                    int B, int n_steps, void* stream);

/* Temporal attention block at C = 64 with both projections on tcgen05 (csrc/tattn_tc.cu): same operation and arguments as
 * wdno_tattn_block, weights in the UMMA canonical K-major layout instead of mma.sync fragment order:
 *   wqkv_canon fp16 [C/8][384][8]  (= to_qkv.weight[n][8c + j]),  wout_canon fp16 [16][C][8]  (= to_out.weight[n][8c + j]) */
int wdno_tattn_block_tc(const void* x, void* y, const float* gamma, const void* wqkv_canon, const void* wout_canon,
                        const float* bias, const float* rot_cos, const float* rot_sin, int64_t n_samples, int n_frames,
                        int64_t hw, int C, float scale, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training step (reference: Trainer.train, smoke/ddpm/diffusion_2d.py:1257-1307 and
 * burgers/ddpm_burgers/train_diffusion.py:187-237 -- there the backward is torch autograd over cuDNN/cuBLAS).
 * Gradients of activations are fp16 channels-last tensors in a SCALED domain (the host multiplies the loss gradient by a
 * power of two so that fp16 neither underflows nor saturates); every parameter-gradient kernel takes `scale` = 1 / that
 * factor and accumulates (+=) into fp32 buffers.
 * dgrad of a convolution is the forward tap-GEMM with transposed / flipped weight tiles (wdno_tapgemm). */
#define WDNO_WGRAD_MAX_GROUP_TAPS 3
typedef struct {
  int32_t dz;          /* source plane = z + dz */
  int32_t dy;          /* row shift of every tap of the group (in the X view's grid) */
  int32_t dx_min;      /* smallest column shift of the group */
  int32_t span;        /* largest - smallest column shift (<= 8) */
  int32_t n;           /* taps in the group (<= WDNO_WGRAD_MAX_GROUP_TAPS) */
  int32_t dxo[WDNO_WGRAD_MAX_GROUP_TAPS];  /* column shift of tap i minus dx_min */
  int64_t out[WDNO_WGRAD_MAX_GROUP_TAPS];  /* tap index of tap i inside dW's innermost axis */
} wdno_wgrad_group;

typedef struct {
  const void* x;       /* fp16 channels-last [B][D][Hs][Ws][Cx]: the layer input (or, for the transposed conv, dY) */
  const void* dy;      /* fp16 channels-last [B][D][H][W][Cy]: gradient of the layer output (transposed conv: the input) */
  float* dw;           /* fp32, += : dw[(m * n_total + n_off + n) * t_total + tap]  (m: dy channel, n: x channel) */
  float* dbias;        /* fp32 [Cy], += column sums of dy; NULL = none */
  int32_t B, D, H, W;  /* grid of dy (the tap grid) */
  int32_t Hs, Ws, Cx;  /* grid and channels of x */
  int32_t sy, sx, ph_y, ph_x; /* X view: Xv[y][x] = x[sy*y + ph_y][sx*x + ph_x] (stride-2 layers: one launch per phase) */
  int32_t Cy;          /* channels of dy (multiple of 8) */
  int32_t m_valid;     /* rows of dW (dy channels beyond it are zero padding and are not written) */
  int32_t cx_off, cx_n;/* channel window of x this launch covers (concatenated sources: one launch per source) */
  int32_t n_total, n_off, t_total; /* dW geometry: total input channels, offset of this window, taps per (m, n) pair */
  int32_t padw;        /* zero columns appended to a row when positions are linearised: max |column shift| */
  int32_t n_groups;
  const wdno_wgrad_group* groups; /* device array */
  int32_t split;       /* CTAs sharing the position axis (split-K; partial sums meet in fp32 atomics) */
  float scale;
} wdno_wgrad_params;
int wdno_wgrad(const wdno_wgrad_params* p, void* stream);

/* Weight gradient of the stride-1 'same' convolutions on tcgen05 (csrc/wgrad_tc.cu): MN-major UMMA operands straight from
 * the channels-last tensors, accumulators of up to 8 in-plane taps in TMEM, two dY planes stacked along M for 64-channel
 * layers (`stack`: rows 0..63 collect depth tap job.kz, rows 64..127 depth tap job.kz - 1).  Same dW indexing as wdno_wgrad. */
typedef struct {
  int32_t kz;          /* depth tap of the tile (rows 0..63 of a stacked tile) */
  int32_t n_taps;      /* in-plane taps of this job (<= 8) */
  int32_t qmin;        /* smallest linear in-plane shift (ky - py) * Wp + (kx - px) of the job: position of X slab row 0 */
  int32_t span;        /* largest - smallest linear shift */
  int32_t m0, n0;      /* first dY channel of the tile (unstacked) / first X channel inside the window */
  int32_t shift[8];    /* linear shift of tap i minus qmin */
  int64_t out_hi[8];   /* tap index inside dW (rows 0..63 of a stacked tile / the whole tile) */
  int64_t out_lo[8];   /* tap index for rows 64..127 of a stacked tile; -1 = discard */
} wdno_wgrad_tc_job;

typedef struct {
  const void* x;       /* fp16 channels-last [B][D][H][W][Cx] */
  const void* dy;      /* fp16 channels-last [B][D][H][W][Cy] */
  float* dw;           /* fp32, += : dw[(m * n_total + n_off + n) * t_total + tap] */
  int32_t B, D, H, W;
  int32_t Cx, Cy, m_valid;
  int32_t cx_off, cx_n, n_total, n_off, t_total;
  int32_t padw, pz;    /* zero columns per linearised row (max |column shift|); depth padding */
  int32_t stack;       /* 1: Cy == 64, planes (z, z + 1) stacked along M; 0: 128 dY channels per tile */
  int32_t nx, ncols;   /* MMA N (X channels per tile, multiple of 16, <= 64) and TMEM column stride per tap */
  int32_t stages;      /* cp.async ring depth (2 or 3) */
  int32_t n_jobs;
  const wdno_wgrad_tc_job* jobs;  /* device array */
  int32_t split;       /* CTAs per job along the position axis */
  float scale;
  int32_t swap_lbo_sbo;/* probe switch: exchange the two descriptor strides (tools/probe_wgrad_tc.py); product value: 0 */
  int32_t reserved0;
} wdno_wgrad_tc_params;
int64_t wdno_wgrad_tc_smem_bytes(const wdno_wgrad_tc_params* p, int span);
int wdno_wgrad_tc(const wdno_wgrad_tc_params* p, int max_span, void* stream);
/* out[c] += scale * sum over positions of dy[pos][c]  (bias gradient; fp16 channels-last in, fp32 out) */
int wdno_colsum_f16(const void* dy, int64_t n_pos, int C, float* out, float scale, void* stream);

/* GroupNorm + (scale+1, shift) + SiLU backward (conv3d.py:189-205; unet.py:129-147), z = a[b,c]*y + c[b,c], h = silu(z):
 *   reduce : sums[b][c] = (sum dz, sum dz*y) over the voxels, dz = dh * silu'(z)                 (double, +=)
 *   apply  : dy = a[b,c]*dz + k1[b,g]*y + k0[b,g]  -- k1, k0 from wdno_gn_bwd_finalize; fp16 out
 * finalize (one block per sample): from sums and the forward statistics (stats[b][g] = (sum y, sum y^2), count voxels*cpg):
 *   d_gamma[c] += A*(1+s), d_beta[c] += Bc*(1+s), d_scale[b,c] = A*gamma + Bc*beta, d_shift[b,c] = Bc   (times `scale`),
 *   with A = rstd*(Sy - mean*S1), Bc = S1; k1 = -rstd^2 * M2, k0 = -rstd*M1 + rstd^2*M2*mean, M1/M2 group means of
 *   gamma'(dz) and gamma'(dz*yhat).  ss = (scale|shift) row of the block's time-MLP output or NULL; d_ss likewise. */
int wdno_gn_bwd_reduce(const void* dh, const void* y, const float* a, const float* c, double* sums, int B, int C, int64_t vox,
                       void* stream);
int wdno_gn_bwd_finalize(const double* sums, const double* stats, const float* gamma, const float* beta, const float* ss,
                         int ss_stride, float* d_gamma, float* d_beta, float* d_ss, int dss_stride, float* k1, float* k0,
                         int B, int C, int G, double count, float eps, float scale, void* stream);
int wdno_gn_bwd_apply(const void* dh, const void* y, const float* a, const float* c, const float* k1, const float* k0,
                      const void* add, void* dy, int B, int C, int G, int64_t vox, void* stream);
/* d_eps (fp32 [B,F,C,H,W]) * mul -> fp16 channels-last [B,F,H,W,cp] (cp >= C, zero padded): head of the backward pass */
int wdno_pack_grad_f16(const float* g, void* out, int B, int F, int C, int H, int W, int cp, float mul, void* stream);
/* out[i] = a[i] + b[i] (fp16, saturating): gradient accumulation at skip connections */
int wdno_add_f16(const void* a, const void* b, void* out, int64_t n, void* stream);
/* nearest x2 up-sampling of H, W (nn.Upsample(scale_factor=2, 'nearest'), unet.py:35-39; x: fp16 [N][H][W][C] -> y [N][2H][2W][C]) --
 * the weight gradient of the convolution behind it needs the materialised input -- and its adjoint, the 2x2 sum (y -> x) */
int wdno_upsample2x_f16(const void* x, void* y, int64_t N, int H, int W, int C, void* stream);
int wdno_sumpool2x2_f16(const void* y, void* x, int64_t N, int H, int W, int C, void* stream);
/* channel LayerNorm backward (conv3d.py:165-174; unet.py:55-65): y = (x - mean) * rstd * gamma over C per voxel.
 * dx = rstd*(g*dy - mean_c(g*dy) - xhat*mean_c(g*dy*xhat)) (+ add), d_gamma[c] += scale * sum dy*xhat */
int wdno_chan_layernorm_bwd(const void* x, const void* dy, const float* gamma, const void* add, void* dx, float* d_gamma,
                            int64_t n_vox, int C, float eps, float scale, void* stream);
/* attention cores, backward (conv3d.py:241-257,294-353; unet.py:203-222,236-259).  qkv / dqkv: fp16 token rows [.., 384]
 * (q | k | v, 4 heads x 32), d_out: fp16 [.., 128] = gradient of the core's output (before to_out); token addressing as in
 * wdno_softmax_attn.  n_tok <= 32: rotary tables and the additive bias [4][n][n] are supported and dbias (fp32 [4][n][n]) += the
 * bias gradient; 32 < n_tok <= 512: plain softmax attention only. */
int wdno_softmax_attn_bwd(const void* qkv, const void* d_out, const float* bias, const float* rot_cos, const float* rot_sin,
                          void* dqkv, float* dbias, int64_t n_seq, int n_tok, int64_t inner, int64_t outerT, int64_t innerT,
                          int64_t tokT, float scale, void* stream);
/* linear attention over n_img images of n_pos positions; work: wdno_linear_attn_bwd_work_bytes(n_img) bytes of scratch */
int64_t wdno_linear_attn_bwd_work_bytes(int64_t n_img);
int wdno_linear_attn_bwd(const void* qkv, const void* d_out, void* dqkv, void* work, int64_t n_img, int n_pos, float scale,
                         void* stream);
/* fused gradient clipping + Adam + EMA over flat fp32 buffers (diffusion_2d.py:1286-1297: clip_grad_norm_(1.0), Adam,
 * ema.update()).  sumsq: device double (squared gradient norm, from wdno_sumsq); the clip factor is computed on the device. */
int wdno_sumsq(const float* g, int64_t n, double* out, void* stream);
int wdno_adam_clip_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n, const double* sumsq,
                       float max_norm, float lr, float beta1, float beta2, float eps, float bc1, float bc2,
                       float ema_decay, int ema_mode, void* stream);

/* Rows [elem_offset, elem_offset + n_local) of the tensor torch.randn(numel_full, device='cuda') would produce from
 * Philox state (seed, philox_offset), bit for bit, without drawing the rest: the noise rule of a batch-sharded run
 * (reference draws: diffusion_2d.py:866,907; diffusion_1d.py:389,431).  grid_full = the grid ATen launches for numel_full
 * values: min(SMs * (max threads per SM / 256), ceil(numel_full / 256)).  The caller advances the generator by
 * ((numel_full - 1) / (256 * grid_full * 4) + 1) * 4 afterwards, as ATen does. */
int wdno_randn_slice(float* out, int64_t n_local, int64_t elem_offset, int64_t numel_full, int grid_full, uint64_t seed,
                     uint64_t philox_offset, void* stream);

/* ------------------------------------------------------------------------------------------
 * separable DWT / IDWT along one axis of an fp32 tensor viewed as [outer][N][inner] (element strides).
 * Replaces the arithmetic of the third-party libraries the reference calls (not vendored in its tree):
 * pytorch_wavelets afb1d/sfb1d ('zero', 'periodization') and ptwt==0.1.6 wavedec3/waverec3 ('zero');
 * call sites: smoke/inference_2d.py:37-46,141-147,178-186,220-254 ; burgers/eval_ddpm_burgers.py:134-136,188-194 ;
 * burgers/ddpm_burgers/test_util.py:186-203.
 *   analysis : lo/hi[i] = sum_k X(2i + k - off) * taps[k]      (X zero-extended, or periodic with odd N edge-repeated)
 *   synthesis: y[m] = sum_{k:(m+off-k) even} lo((m+off-k)/2)*taps_lo[k] + hi(..)*taps_hi[k]
 * taps are HOST arrays (<= WDNO_MAX_TAPS).  The adjoints used by gradient guidance are the same two kernels
 * with the synthesis / analysis taps swapped in (see wdno_b200/wavelets.py).
 * ------------------------------------------------------------------------------------------ */
#define WDNO_MAX_TAPS 20
int wdno_dwt_analysis_axis(const float* x, float* lo, float* hi, int64_t outer, int N, int64_t inner, int nout,
                           int64_t x_ostride, int64_t lo_ostride, int64_t hi_ostride, const float* taps_lo_host,
                           const float* taps_hi_host, int L, int off, int periodic, void* stream);
int wdno_dwt_synthesis_axis(const float* lo, const float* hi, float* y, int64_t outer, int n, int64_t inner, int Nout,
                            int64_t lo_ostride, int64_t hi_ostride, int64_t y_ostride, const float* taps_lo_host,
                            const float* taps_hi_host, int L, int off, int periodic, void* stream);

/* Fused single-pass 3-D transforms ('zero' mode, level 1): ptwt.wavedec3 / waverec3 (call sites
 * smoke/inference_2d.py:41,141,184,220,250, wave_trans_2d.py:129-149) and, with the other filter pair, their adjoints (the
 * gradient of the guidance objective, every guided step).  One launch, all three passes in shared memory.
 * bands8: HOST array of 8 DEVICE pointers in ptwt key order aaa,aad,ada,add,daa,dad,dda,ddd (letters = D,H,W; a = low),
 * each [B][nd][nh][nw] with batch stride band_bstride (elements) and contiguous planes; x / y: [B][Nd][Nh][Nw] contiguous.
 *   analysis : band[i_d,i_h,i_w] = sum x(2i+k-off) t_d[k] t_h[k'] t_w[k'']      synthesis: the transposed-convolution form
 * wdno_dwt3d_supported: 1 if the tile plan fits shared memory for (L, nw, Nw); otherwise use the per-axis entry points. */
int wdno_dwt3d_supported(int L, int nw, int Nw);
int wdno_dwt3d_synthesis(const float* const* bands8, int64_t band_bstride, float* y, int64_t B, int nd, int nh, int nw, int Nd,
                         int Nh, int Nw, const float* taps_lo_host, const float* taps_hi_host, int L, int off, void* stream);
int wdno_dwt3d_analysis(const float* x, float* const* bands8, int64_t band_bstride, int64_t B, int Nd, int Nh, int Nw, int nd,
                        int nh, int nw, const float* taps_lo_host, const float* taps_hi_host, int L, int off, void* stream);

/* Fused one-level 2-D transforms of n_img images (modes 'zero' and 'periodization'): pytorch_wavelets DWTForward / DWTInverse,
 * one level per call (call sites burgers/eval_ddpm_burgers.py:134-136,188-194,
 * burgers/ddpm_burgers/test_util.py:200-203, data_burgers_1d.py:66-68, burgers/wave_trans.py:103-108,
 * smoke/inference_2d.py:178-180,244-246, smoke/wave_trans_2d.py:135-137).  One launch instead of three per-axis passes.
 * x / y: [n_img][H][W] contiguous.  bands4: HOST array of 4 DEVICE pointers (LL, LH, HL, HH) = (lo_w lo_h, lo_w hi_h,
 * hi_w lo_h, hi_w hi_h), each [n_img][nh][nw] with contiguous planes and its own image stride band_istride4[i] (elements), so
 * the detail bands may be the three slices of a [.., 3, nh, nw] tensor or slots of a packed [.., 4, nh, nw] tensor.
 * Per-axis formulas as wdno_dwt_analysis_axis / wdno_dwt_synthesis_axis with (offh, offw) and the same periodic flag.
 * wdno_dwt2d_supported: 1 if the geometry fits (else use the per-axis entry points). */
int wdno_dwt2d_supported(int L, int H, int W, int nh, int nw, int periodic);
int wdno_dwt2d_analysis(const float* x, float* const* bands4, const int64_t* band_istride4, int64_t n_img, int H, int W, int nh,
                        int nw, const float* taps_lo_host, const float* taps_hi_host, int L, int offh, int offw, int periodic,
                        void* stream);
int wdno_dwt2d_synthesis(const float* const* bands4, const int64_t* band_istride4, float* y, int64_t n_img, int nh, int nw, int H,
                         int W, const float* taps_lo_host, const float* taps_hi_host, int L, int offh, int offw, int periodic,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WDNO_B200_H */
