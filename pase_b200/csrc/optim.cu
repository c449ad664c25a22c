// Flat-buffer Adam: ONE launch updates every parameter of the model (encoder + all worker
// heads) with per-segment hyper-parameters.  Replaces the 13 torch.optim.Adam instances the
// reference steps one after the other (pase/models/WorkerScheduler/trainer.py:86-143,
// worker_scheduler.py:66-73), and folds the 1/world gradient scale of the data-parallel
// all-reduce into the same pass.  Semantics = torch.optim.Adam (amsgrad off, maximize off,
// L2 weight decay added to the gradient).
#include "common.cuh"

namespace {

constexpr int SEG_MAX = 64;

struct AdamSeg {           // one row of the device segment table (8 x 8 bytes)
  long start, end;         // element range [start, end) of the flat buffers
  float lr, beta1, beta2, eps, weight_decay, pad;
  long step_index;         // index into the device step-count vector
};

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                 float* __restrict__ v, long n, const AdamSeg* __restrict__ segs, int nseg,
                 const float* __restrict__ steps, float grad_scale) {
  pdl_wait();
  __shared__ AdamSeg ss[SEG_MAX];
  __shared__ float c1[SEG_MAX], c2[SEG_MAX];       // lr / bias1, 1 / sqrt(bias2)
  for (int i = threadIdx.x; i < nseg; i += blockDim.x) {
    ss[i] = segs[i];
    const float t = steps[ss[i].step_index];
    const float b1 = 1.f - powf(ss[i].beta1, t), b2 = 1.f - powf(ss[i].beta2, t);
    c1[i] = ss[i].lr / b1;
    c2[i] = rsqrtf(b2);
  }
  __syncthreads();
  const long stride = (long)gridDim.x * blockDim.x * 4;
  int s = 0;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    while (s + 1 < nseg && i >= ss[s].end) ++s;
    // a float4 never straddles two segments when every segment starts at a multiple of 4
    // (the host pads); the scalar tail handles a ragged end
    const AdamSeg& sg = ss[s];
    const float beta1 = sg.beta1, beta2 = sg.beta2, eps = sg.eps, wd = sg.weight_decay;
    const float step_size = c1[s], inv_sqrt_b2 = c2[s];
    const int cnt = (i + 3 < n) ? 4 : (int)(n - i);
    float pv[4], gv[4], mv[4], vv[4];
    if (cnt == 4) {
      const float4 P = *reinterpret_cast<const float4*>(p + i);
      const float4 G = *reinterpret_cast<const float4*>(g + i);
      const float4 M = *reinterpret_cast<const float4*>(m + i);
      const float4 V = *reinterpret_cast<const float4*>(v + i);
      pv[0] = P.x; pv[1] = P.y; pv[2] = P.z; pv[3] = P.w;
      gv[0] = G.x; gv[1] = G.y; gv[2] = G.z; gv[3] = G.w;
      mv[0] = M.x; mv[1] = M.y; mv[2] = M.z; mv[3] = M.w;
      vv[0] = V.x; vv[1] = V.y; vv[2] = V.z; vv[3] = V.w;
    } else {
      for (int k = 0; k < cnt; ++k) { pv[k] = p[i + k]; gv[k] = g[i + k]; mv[k] = m[i + k]; vv[k] = v[i + k]; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= cnt) break;
      if (i + k >= sg.end) continue;                 // padding between segments: untouched
      const float gg = gv[k] * grad_scale + wd * pv[k];
      mv[k] = beta1 * mv[k] + (1.f - beta1) * gg;
      vv[k] = beta2 * vv[k] + (1.f - beta2) * gg * gg;
      const float denom = sqrtf(vv[k]) * inv_sqrt_b2 + eps;
      pv[k] -= step_size * mv[k] / denom;
    }
    if (cnt == 4) {
      *reinterpret_cast<float4*>(p + i) = make_float4(pv[0], pv[1], pv[2], pv[3]);
      *reinterpret_cast<float4*>(m + i) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
      for (int k = 0; k < cnt; ++k) { p[i + k] = pv[k]; m[i + k] = mv[k]; v[i + k] = vv[k]; }
    }
  }
}

}  // namespace

extern "C" {

int pase_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n,
                   const long* seg_table, int nseg, const float* steps, float grad_scale,
                   void* stream) {
  PASE_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && seg_table && steps && n > 0,
                 "pase_adam_flat: null pointer / empty buffer");
  PASE_CHECK_ARG(nseg > 0 && nseg <= SEG_MAX, "pase_adam_flat: nseg=%d (1..%d)", nseg, SEG_MAX);
  PASE_CHECK_ARG(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq),
                 "pase_adam_flat: buffers must be 16-byte aligned");
  static_assert(sizeof(AdamSeg) == 48, "segment row = 6 x 8 bytes");
  long blocks = (n / 4 + 255) / 256;
  const long cap = (long)pase_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  PASE_LAUNCH((adam_flat_kernel), (unsigned)blocks, 256, 0, (cudaStream_t)stream, 
      param, grad, exp_avg, exp_avg_sq, n, reinterpret_cast<const AdamSeg*>(seg_table), nseg, steps,
      grad_scale);
  PASE_LAUNCH_CHECK("pase_adam_flat");
  return PASE_OK;
}

}  // extern "C"
