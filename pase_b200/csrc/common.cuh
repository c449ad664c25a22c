// Shared helpers for the pase_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include "../../include/pase_b200.h"

void pase_set_error(const char* fmt, ...);

#define PASE_CHECK_ARG(cond, ...)                     \
  do {                                                \
    if (!(cond)) {                                    \
      pase_set_error(__VA_ARGS__);                    \
      return PASE_ERR_ARG;                            \
    }                                                 \
  } while (0)

#define PASE_LAUNCH_CHECK(name)                                            \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) {                                              \
      pase_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return (int)e__;                                                     \
    }                                                                      \
  } while (0)

static inline int pase_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// reflect index for F.pad(mode='reflect'): t in [-padL, T+padR) -> [0,T)
__device__ __forceinline__ int reflect_idx(int t, int T) {
  if (t < 0) t = -t;
  if (t >= T) t = 2 * (T - 1) - t;
  return t;
}
