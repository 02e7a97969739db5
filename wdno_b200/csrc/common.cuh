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

}  // namespace wdno
