// Error plumbing shared by all translation units of libwdno_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wdno_b200.h"

namespace wdno {

int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
// cudaPeekAtLastError after a launch -> WDNO_OK / WDNO_E_CUDA
int check_launch(const char* where);
int num_sms();
// attention_mma.cu: tensor-core softmax attention for sequences of <= 32 tokens
int launch_short_attn_mma(const void* qkv, void* out, const float* bias, const float* rot_cos, const float* rot_sin,
                          long long n_seq, int n_tok, long long inner, long long outerT, long long innerT, long long tokT,
                          float scale, cudaStream_t st);

// attention_mma.cu: tensor-core softmax attention for 32 < n <= 512 tokens without bias / rotary
int launch_flash_attn_mma(const void* qkv, void* out, long long n_seq, int n_tok, long long inner, long long outerT,
                          long long innerT, long long tokT, float scale, cudaStream_t st);

// attn_fused.cu: merge the context partials of the linear attention and fold to_out into them (canon: UMMA operand order)
int launch_la_mid(const float* part, const float* wout, void* mpack, int C, int nparts, float scale, int canon, int n_img,
                  cudaStream_t st);

// dwt3d_stream.cu: streaming (register sliding-window) form of the fused 3-D transforms; return 1 = launched, 0 = shape or
// alignment outside the envelope (use the tile kernels of dwt3d.cu), < 0 = error
int launch_ana3d_stream(const float* x, float* const* bands8, long long band_bstride, long long B, int Nd, int Nh, int Nw, int nd,
                        int nh, int nw, const float* t0, const float* t1, int L, int off, cudaStream_t st);
int launch_syn3d_stream(const float* const* bands8, long long band_bstride, float* y, long long B, int nd, int nh, int nw, int Nd,
                        int Nh, int Nw, const float* t0, const float* t1, int L, int off, cudaStream_t st);

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl may start (run its prologue) while
// the previous kernel of the stream drains; it must execute pdl_wait() before touching anything a predecessor wrote.
// Every kernel calls pdl_trigger() first so that its successor can be scheduled as early as resources allow.
bool pdl_enabled();  // WDNO_PDL=1 turns it on (default: plain stream order; measured neutral in round 1)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace wdno
