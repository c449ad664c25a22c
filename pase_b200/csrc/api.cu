// Library-level entry points: error reporting, version, device query.
#include "common.cuh"
#include <cstdarg>
#include <cstring>

static thread_local char g_err[512] = "";

void pase_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

const char* pase_last_error(void) { return g_err; }

int pase_version(void) { return 100; }

int pase_device_info(int* out4) {
  PASE_CHECK_ARG(out4 != nullptr, "pase_device_info: null output");
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    pase_set_error("pase_device_info: %s", cudaGetErrorString(e));
    return (int)e;
  }
  cudaDeviceGetAttribute(&out4[0], cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&out4[1], cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&out4[2], cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&out4[3], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  return PASE_OK;
}

}  // extern "C"
