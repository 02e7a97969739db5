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
  int32_t Wp;             /* padded row width (W + KW - 1) */
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
} wdno_tapgemm_params;

/* bytes of dynamic shared memory the plan needs, or <0 */
int64_t wdno_tapgemm_smem_bytes(const wdno_tapgemm_params* p);
int wdno_tapgemm(const wdno_tapgemm_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WDNO_B200_H */
