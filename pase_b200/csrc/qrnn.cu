// QRNN (window=2) gate activations + ForgetMult recurrence + output gate, forward and
// backward.  One thread per (sample, hidden unit); time is sequential (T' = T/160
// frames), loads are issued UNROLL steps ahead of the dependent recurrence.
// Replaces torchqrnn's runtime-compiled recurrent_forget_mult kernels
// (reference call site: pase/models/modules.py:52, frontend.py:256-259).
#include "common.cuh"

namespace {

constexpr int UNROLL = 8;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(128)
qrnn_scan_fwd_kernel(const float* __restrict__ Y, float* __restrict__ h, long ldh,
                     float* __restrict__ Cst, int T, int H) {
  pdl_wait();
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  const float* y = Y + (long)n * T * 3 * H + c;
  float* hp = h + (long)n * T * ldh + c;
  float* cp = Cst + (long)n * T * H + c;
  float cprev = 0.f;
  int t = 0;
  for (; t + UNROLL <= T; t += UNROLL) {
    float z[UNROLL], f[UNROLL], o[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      const float* yr = y + (long)(t + i) * 3 * H;
      z[i] = yr[0]; f[i] = yr[H]; o[i] = yr[2 * H];
    }
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      const float zz = tanhf(z[i]), ff = sigmoidf_(f[i]), oo = sigmoidf_(o[i]);
      cprev = ff * zz + (1.f - ff) * cprev;
      cp[(long)(t + i) * H] = cprev;
      hp[(long)(t + i) * ldh] = oo * cprev;
    }
  }
  for (; t < T; ++t) {
    const float* yr = y + (long)t * 3 * H;
    const float zz = tanhf(yr[0]), ff = sigmoidf_(yr[H]), oo = sigmoidf_(yr[2 * H]);
    cprev = ff * zz + (1.f - ff) * cprev;
    cp[(long)t * H] = cprev;
    hp[(long)t * ldh] = oo * cprev;
  }
}

__global__ void __launch_bounds__(128)
qrnn_scan_bwd_kernel(const float* __restrict__ Y, const float* __restrict__ Cst,
                     const float* __restrict__ dh, long lddh, float* __restrict__ dY, int T,
                     int H) {
  pdl_wait();
  const int n = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  const float* y = Y + (long)n * T * 3 * H + c;
  const float* cp = Cst + (long)n * T * H + c;
  const float* dhp = dh + (long)n * T * lddh + c;
  float* dy = dY + (long)n * T * 3 * H + c;
  float carry = 0.f;
  int t = T - 1;
  // the loads do not depend on the recurrence: issue UNROLL time steps of them up front
  for (; t - (UNROLL - 1) >= 0; t -= UNROLL) {
    float z[UNROLL], f[UNROLL], o[UNROLL], c[UNROLL], cm[UNROLL], g[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      const int tt = t - i;
      const float* yr = y + (long)tt * 3 * H;
      z[i] = yr[0]; f[i] = yr[H]; o[i] = yr[2 * H];
      c[i] = cp[(long)tt * H];
      cm[i] = tt > 0 ? cp[(long)(tt - 1) * H] : 0.f;
      g[i] = dhp[(long)tt * lddh];
    }
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) {
      const int tt = t - i;
      const float zz = tanhf(z[i]), ff = sigmoidf_(f[i]), oo = sigmoidf_(o[i]);
      const float dc = g[i] * oo + carry;
      carry = dc * (1.f - ff);
      float* dr = dy + (long)tt * 3 * H;
      dr[0] = dc * ff * (1.f - zz * zz);
      dr[H] = dc * (zz - cm[i]) * ff * (1.f - ff);
      dr[2 * H] = g[i] * c[i] * oo * (1.f - oo);
    }
  }
  for (; t >= 0; --t) {
    const float* yr = y + (long)t * 3 * H;
    const float zz = tanhf(yr[0]), ff = sigmoidf_(yr[H]), oo = sigmoidf_(yr[2 * H]);
    const float ct = cp[(long)t * H];
    const float cm1 = t > 0 ? cp[(long)(t - 1) * H] : 0.f;
    const float g = dhp[(long)t * lddh];
    const float dc = g * oo + carry;
    carry = dc * (1.f - ff);
    float* dr = dy + (long)t * 3 * H;
    dr[0] = dc * ff * (1.f - zz * zz);
    dr[H] = dc * (zz - cm1) * ff * (1.f - ff);
    dr[2 * H] = g * ct * oo * (1.f - oo);
  }
}

}  // namespace

extern "C" {

int pase_qrnn_scan_fwd(const float* Y, float* h, long ldh, float* Cst, int N, int T, int H,
                       void* stream) {
  PASE_CHECK_ARG(Y && h && Cst && N > 0 && T > 0 && H > 0 && ldh >= H,
                 "pase_qrnn_scan_fwd: bad args");
  dim3 grid((H + 127) / 128, N);
  PASE_LAUNCH((qrnn_scan_fwd_kernel), grid, 128, 0, (cudaStream_t)stream, Y, h, ldh, Cst, T, H);
  PASE_LAUNCH_CHECK("pase_qrnn_scan_fwd");
  return PASE_OK;
}

int pase_qrnn_scan_bwd(const float* Y, const float* Cst, const float* dh, long lddh, float* dY,
                       int N, int T, int H, void* stream) {
  PASE_CHECK_ARG(Y && Cst && dh && dY && N > 0 && T > 0 && H > 0 && lddh >= H,
                 "pase_qrnn_scan_bwd: bad args");
  dim3 grid((H + 127) / 128, N);
  PASE_LAUNCH((qrnn_scan_bwd_kernel), grid, 128, 0, (cudaStream_t)stream, Y, Cst, dh, lddh, dY, T, H);
  PASE_LAUNCH_CHECK("pase_qrnn_scan_bwd");
  return PASE_OK;
}

}  // extern "C"
