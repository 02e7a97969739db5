// Library-level entry points: version, device probe, error string.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace wdno {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return WDNO_E_CUDA;
}

int check_launch(const char* where) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_cuda_error(e, where);
  }
  return WDNO_OK;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("WDNO_PDL");
    on = (e != nullptr && e[0] == '1') ? 1 : 0;  // measured neutral on the C3 step (98.9 vs 99.5 steps/s): opt-in
  }
  return on == 1;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return sms;
}

}  // namespace wdno

extern "C" const char* wdno_last_error(void) { return wdno::g_err; }

extern "C" int wdno_version(void) { return 100; }

extern "C" int wdno_device_cc(void) {
  int dev = 0, major = 0, minor = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return wdno::set_cuda_error(e, "wdno_device_cc");
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return wdno::set_cuda_error(e, "wdno_device_cc");
  e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) return wdno::set_cuda_error(e, "wdno_device_cc");
  return major * 10 + minor;
}
