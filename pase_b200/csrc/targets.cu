// On-device regression TARGETS: log-power spectrum + delta features +
// z-normalisation of a batch of waveform chunks (reference: pase/transforms.py:439-487 LPS --
// torch.stft(n_fft, hop, win, rectangular window, centre reflect padding), 10 log10(|X|^2 +
// 1e-19), librosa.feature.delta orders 1..der_order -- and transforms.py:183-202 ZNorm).
// The reference computes them per utterance on DataLoader workers and ships 79 MB per lps
// label and step to the GPU; here the chunk is already on the device:
//   pase_frame_wave   frames matrix (fp16 hi/lo pair operand): row (n, j) = the `win` samples
//                     the rectangular window keeps of STFT frame j (reflect indexing)
//   pase_tc_gemm_nt   frames x DFT basis (cos | -sin rows interleaved, only the `win`
//                     non-zero taps of the n_fft-point transform): 3xF16, fp32-equivalent
//   pase_lps_post     |X|^2 -> dB, Savitzky-Golay deltas along time (edge windows clamped =
//                     mode 'interp' for polyorder == deriv), z-norm, (N, F*(1+der), T') layout
#include "common.cuh"

namespace {

__global__ void frame_wave_kernel(const float* __restrict__ x, int T, int hop, int win, int start0,
                                  int frames, __half* __restrict__ hi, __half* __restrict__ lo,
                                  int lda) {
  pdl_wait();
  const int n = blockIdx.y;
  const long total = (long)frames * lda;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e / lda), m = (int)(e % lda);
    float v = 0.f;
    if (m < win) v = x[(long)n * T + reflect_idx(j * hop + start0 + m, T)];
    __half h, l;
    f16_split(v, h, l);
    const long o = ((long)n * frames + j) * lda + m;
    hi[o] = h;
    lo[o] = l;
  }
}

constexpr int LPS_BINS = 32;

// C: (N*frames, ldc) fp32, columns (2k, 2k+1) = (re, im) of bin k.  out: (N, (1+der)*nbins,
// frames).  fir: [der][width] correlation taps of the d-th derivative filter.
__global__ void __launch_bounds__(256)
lps_post_kernel(const float* __restrict__ C, long ldc, int frames, int nbins, int der, int width,
                const float* __restrict__ fir, const float* __restrict__ mean,
                const float* __restrict__ stdv, float* __restrict__ out) {
  pdl_wait();
  extern __shared__ float X[];                   // [LPS_BINS][frames + 1]
  const int pitch = frames + 1;
  const int n = blockIdx.y, k0 = blockIdx.x * LPS_BINS;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5, ngrp = blockDim.x >> 5;
  const int k = k0 + lane;
  for (int t = grp; t < frames; t += ngrp) {
    float v = 0.f;
    if (k < nbins) {
      const float2 z = *reinterpret_cast<const float2*>(C + ((long)n * frames + t) * ldc + 2 * k);
      const float mag = sqrtf(z.x * z.x + z.y * z.y);          // |X| first, like torch.norm
      v = 10.f * log10f(mag * mag + 1e-19f);
    }
    X[lane * pitch + t] = v;
  }
  __syncthreads();
  const int half = width >> 1;
  const int F = (1 + der) * nbins;
  for (int b = grp; b < LPS_BINS; b += ngrp) {
    const int kb = k0 + b;
    if (kb >= nbins) break;
    const float* row = X + b * pitch;
    for (int t = lane; t < frames; t += 32) {
      int c = t < half ? half : t;
      if (c > frames - 1 - half) c = frames - 1 - half;
      for (int d = 0; d <= der; ++d) {
        float v;
        if (d == 0) {
          v = row[t];
        } else {
          v = 0.f;
          const float* f = fir + (d - 1) * width;
          for (int i = 0; i < width; ++i) v = fmaf(f[i], row[c - half + i], v);
        }
        const int feat = d * nbins + kb;
        if (mean != nullptr) v = (v - mean[feat]) / stdv[feat];
        out[((long)n * F + feat) * frames + t] = v;
      }
    }
  }
}

}  // namespace

extern "C" {

int pase_frame_wave(const float* x, int N, int T, int hop, int win, int start0, int frames,
                    void* hi, void* lo, int lda, void* stream) {
  PASE_CHECK_ARG(x && hi && lo && N > 0 && T > 1 && hop > 0 && win > 0 && frames > 0,
                 "pase_frame_wave: bad args");
  PASE_CHECK_ARG(lda >= win && -start0 < T && (long)(frames - 1) * hop + start0 + win - 1 < 2L * T - 1,
                 "pase_frame_wave: window [%d, +%d) leaves the reflect range of T=%d", start0, win, T);
  long blocks = ((long)frames * lda + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  PASE_LAUNCH((frame_wave_kernel), dim3((unsigned)blocks, N), 256, 0, (cudaStream_t)stream, x, T,
              hop, win, start0, frames, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo),
              lda);
  PASE_LAUNCH_CHECK("pase_frame_wave");
  return PASE_OK;
}

int pase_lps_post(const float* C, long ldc, int N, int frames, int nbins, int der_order,
                  int width, const float* fir, const float* mean, const float* stdv, float* out,
                  void* stream) {
  PASE_CHECK_ARG(C && out && N > 0 && frames > 0 && nbins > 0 && ldc >= 2L * nbins &&
                     (ldc % 2) == 0 && aligned16(C),
                 "pase_lps_post: bad args");
  PASE_CHECK_ARG(der_order >= 0 && der_order <= 4 && (der_order == 0 || (fir && (width & 1) &&
                     width >= 3 && frames >= width)),
                 "pase_lps_post: der_order=%d width=%d frames=%d", der_order, width, frames);
  PASE_CHECK_ARG((mean == nullptr) == (stdv == nullptr), "pase_lps_post: mean and std go together");
  const size_t smem = (size_t)LPS_BINS * (frames + 1) * sizeof(float);
  PASE_CHECK_ARG(smem <= 200 * 1024, "pase_lps_post: %d frames exceed the shared-memory tile", frames);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(lps_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      pase_set_error("pase_lps_post: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  dim3 grid((nbins + LPS_BINS - 1) / LPS_BINS, N);
  PASE_LAUNCH((lps_post_kernel), grid, 256, smem, (cudaStream_t)stream, C, ldc, frames, nbins,
              der_order, width, fir, mean, stdv, out);
  PASE_LAUNCH_CHECK("pase_lps_post");
  return PASE_OK;
}

}  // extern "C"
