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

}  // namespace wdno
