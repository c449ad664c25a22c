// QRNN (window=2) gate activations + ForgetMult recurrence + output gate, forward and
// backward.  One thread per (sample, hidden unit); time is sequential (T' = T/160
// frames), loads are issued UNROLL steps ahead of the dependent recurrence.
// Replaces torchqrnn's runtime-compiled recurrent_forget_mult kernels
// (reference call site: pase/models/modules.py:52, frontend.py:256-259).
#include "common.cuh"
#include <cstdlib>

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

// ---- time-segmented scans ----
// The recurrence c_t = f_t z_t + (1-f_t) c_{t-1} is linear in c: a segment of SEG steps maps
// its incoming state to  c_out = A c_in + B  with A = prod(1-f), B = the scan from zero.
// A block owns 32 hidden units x ceil(T/SEG) segments (threadIdx.y = segment): every thread
// scans its segment from zero (gate activations: the expensive part, now 13x more parallel
// at T'=200), the (A, B) pairs meet in shared memory, each thread folds the segments before
// its own (<= 31 FMAs) and corrects its SEG values.  The sequential kernels above keep one
// warp per scheduler busy for 200 dependent steps (45 / 64 us at N=32, H=512); these take
// the chain down to 2 x 16 short steps.
constexpr int SEG = 16;
constexpr int MAX_SEGS = 32;

template <int NSEG_MAX>
__global__ void __launch_bounds__(32 * NSEG_MAX, 1)
qrnn_scan_fwd_seg_kernel(const float* __restrict__ Y, float* __restrict__ h, long ldh,
                         float* __restrict__ Cst, int T, int H) {
  pdl_wait();
  __shared__ float sA[MAX_SEGS][33], sB[MAX_SEGS][33];
  const int lane = threadIdx.x, seg = threadIdx.y, n = blockIdx.y;
  const int c = blockIdx.x * 32 + lane;
  const bool live = c < H;
  const int t0 = seg * SEG;
  const int len = (T - t0) < SEG ? (T - t0) : SEG;
  const float* y = Y + ((long)n * T + t0) * 3 * H + c;
  float z[SEG], f[SEG], o[SEG];
#pragma unroll
  for (int i = 0; i < SEG; ++i) {
    const bool ok = live && i < len;
    const float* yr = y + (long)i * 3 * H;
    z[i] = ok ? yr[0] : 0.f;
    f[i] = ok ? yr[H] : 0.f;
    o[i] = ok ? yr[2 * H] : 0.f;
  }
  float loc[SEG], pr[SEG];
  float cl = 0.f, P = 1.f;
#pragma unroll
  for (int i = 0; i < SEG; ++i) {
    const float zz = tanhf(z[i]), ff = sigmoidf_(f[i]);
    const float keep = (i < len) ? 1.f - ff : 1.f;
    cl = (i < len) ? ff * zz + keep * cl : cl;
    P *= keep;
    loc[i] = cl;
    pr[i] = P;
    o[i] = sigmoidf_(o[i]);
  }
  sA[seg][lane] = P;
  sB[seg][lane] = cl;
  __syncthreads();
  float cin = 0.f;
  for (int s2 = 0; s2 < seg; ++s2) cin = sA[s2][lane] * cin + sB[s2][lane];
  if (!live) return;
  float* hp = h + ((long)n * T + t0) * ldh + c;
  float* cp = Cst + ((long)n * T + t0) * H + c;
#pragma unroll
  for (int i = 0; i < SEG; ++i) {
    if (i < len) {
      const float ct = fmaf(pr[i], cin, loc[i]);
      cp[(long)i * H] = ct;
      hp[(long)i * ldh] = o[i] * ct;
    }
  }
}

template <int NSEG_MAX>
__global__ void __launch_bounds__(32 * NSEG_MAX, 1)
qrnn_scan_bwd_seg_kernel(const float* __restrict__ Y, const float* __restrict__ Cst,
                         const float* __restrict__ dh, long lddh, float* __restrict__ dY, int T,
                         int H) {
  pdl_wait();
  __shared__ float sA[MAX_SEGS][33], sB[MAX_SEGS][33];
  const int lane = threadIdx.x, seg = threadIdx.y, n = blockIdx.y, nseg = blockDim.y;
  const int c = blockIdx.x * 32 + lane;
  const bool live = c < H;
  const int t0 = seg * SEG;
  const int len = (T - t0) < SEG ? (T - t0) : SEG;
  const float* y = Y + ((long)n * T + t0) * 3 * H + c;
  const float* cp = Cst + ((long)n * T + t0) * H + c;
  const float* dhp = dh + ((long)n * T + t0) * lddh + c;
  float* dy = dY + ((long)n * T + t0) * 3 * H + c;
  float z[SEG], f[SEG], o[SEG], cs[SEG], g[SEG];
#pragma unroll
  for (int i = 0; i < SEG; ++i) {
    const bool ok = live && i < len;
    const float* yr = y + (long)i * 3 * H;
    z[i] = ok ? yr[0] : 0.f;
    f[i] = ok ? yr[H] : 0.f;
    o[i] = ok ? yr[2 * H] : 0.f;
    cs[i] = ok ? cp[(long)i * H] : 0.f;
    g[i] = ok ? dhp[(long)i * lddh] : 0.f;
  }
  const float cm0 = (live && t0 > 0) ? cp[-(long)H] : 0.f;       // c_{t0-1}
  // right-to-left local scan with zero incoming carry; z[] / f[] are overwritten by the
  // coefficients of the correction pass (dz = dc kz, df = dc kf), cs[] / o[] by (local dc, Q)
  float carry = 0.f, Q = 1.f;
#pragma unroll
  for (int i = SEG - 1; i >= 0; --i) {
    const float zz = tanhf(z[i]), ff = sigmoidf_(f[i]), oo = sigmoidf_(o[i]);
    const float cm = i > 0 ? cs[i - 1] : cm0;
    const float ct = cs[i];
    const bool in = i < len;
    const float dc = in ? fmaf(g[i], oo, carry) : 0.f;
    if (live && in) dy[(long)i * 3 * H + 2 * H] = g[i] * ct * oo * (1.f - oo);
    z[i] = ff * (1.f - zz * zz);                 // kz
    f[i] = (zz - cm) * ff * (1.f - ff);          // kf
    cs[i] = dc;                                  // local dc
    o[i] = Q;                                    // prod_{s>t}(1-f_s) inside the segment
    if (in) {
      carry = dc * (1.f - ff);
      Q *= (1.f - ff);
    }
  }
  sA[seg][lane] = Q;
  sB[seg][lane] = carry;
  __syncthreads();
  float kin = 0.f;                               // carry entering from the right
  for (int s2 = nseg - 1; s2 > seg; --s2) kin = sA[s2][lane] * kin + sB[s2][lane];
  if (!live) return;
#pragma unroll
  for (int i = 0; i < SEG; ++i) {
    if (i < len) {
      const float dc = fmaf(o[i], kin, cs[i]);
      dy[(long)i * 3 * H] = dc * z[i];
      dy[(long)i * 3 * H + H] = dc * f[i];
    }
  }
}

inline bool seg_scan_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PASE_B200_QRNN_SEG");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

extern "C" {

int pase_qrnn_scan_fwd(const float* Y, float* h, long ldh, float* Cst, int N, int T, int H,
                       void* stream) {
  PASE_CHECK_ARG(Y && h && Cst && N > 0 && T > 0 && H > 0 && ldh >= H,
                 "pase_qrnn_scan_fwd: bad args");
  const int nseg = (T + SEG - 1) / SEG;
  if (nseg <= MAX_SEGS && seg_scan_enabled()) {
    if (nseg <= 16)
      PASE_LAUNCH((qrnn_scan_fwd_seg_kernel<16>), dim3((H + 31) / 32, N), dim3(32, nseg), 0,
                  (cudaStream_t)stream, Y, h, ldh, Cst, T, H);
    else if (nseg <= 24)
      PASE_LAUNCH((qrnn_scan_fwd_seg_kernel<24>), dim3((H + 31) / 32, N), dim3(32, nseg), 0,
                  (cudaStream_t)stream, Y, h, ldh, Cst, T, H);
    else
      PASE_LAUNCH((qrnn_scan_fwd_seg_kernel<MAX_SEGS>), dim3((H + 31) / 32, N), dim3(32, nseg), 0,
                  (cudaStream_t)stream, Y, h, ldh, Cst, T, H);
    PASE_LAUNCH_CHECK("pase_qrnn_scan_fwd");
    return PASE_OK;
  }
  dim3 grid((H + 127) / 128, N);
  PASE_LAUNCH((qrnn_scan_fwd_kernel), grid, 128, 0, (cudaStream_t)stream, Y, h, ldh, Cst, T, H);
  PASE_LAUNCH_CHECK("pase_qrnn_scan_fwd");
  return PASE_OK;
}

int pase_qrnn_scan_bwd(const float* Y, const float* Cst, const float* dh, long lddh, float* dY,
                       int N, int T, int H, void* stream) {
  PASE_CHECK_ARG(Y && Cst && dh && dY && N > 0 && T > 0 && H > 0 && lddh >= H,
                 "pase_qrnn_scan_bwd: bad args");
  const int nseg = (T + SEG - 1) / SEG;
  if (nseg <= MAX_SEGS && seg_scan_enabled()) {
    if (nseg <= 16)
      PASE_LAUNCH((qrnn_scan_bwd_seg_kernel<16>), dim3((H + 31) / 32, N), dim3(32, nseg), 0,
                  (cudaStream_t)stream, Y, Cst, dh, lddh, dY, T, H);
    else if (nseg <= 24)
      PASE_LAUNCH((qrnn_scan_bwd_seg_kernel<24>), dim3((H + 31) / 32, N), dim3(32, nseg), 0,
                  (cudaStream_t)stream, Y, Cst, dh, lddh, dY, T, H);
    else
      PASE_LAUNCH((qrnn_scan_bwd_seg_kernel<MAX_SEGS>), dim3((H + 31) / 32, N), dim3(32, nseg), 0,
                  (cudaStream_t)stream, Y, Cst, dh, lddh, dY, T, H);
    PASE_LAUNCH_CHECK("pase_qrnn_scan_bwd");
    return PASE_OK;
  }
  dim3 grid((H + 127) / 128, N);
  PASE_LAUNCH((qrnn_scan_bwd_kernel), grid, 128, 0, (cudaStream_t)stream, Y, Cst, dh, lddh, dY, T, H);
  PASE_LAUNCH_CHECK("pase_qrnn_scan_bwd");
  return PASE_OK;
}

}  // extern "C"
