// Operand-format conversions for the 16-bit tensor-core GEMM modes (small tensors: weights,
// the [qrnn | pooled-skip] matrix, output-side gradients).  Activation-sized operands are
// written in their GEMM format directly by the producing kernels (act.cu).
#include "common.cuh"

namespace {

__global__ void cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ dst,
                                 long n) {
  const long stride = (long)gridDim.x * blockDim.x * 4;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      st4t(dst + i, *reinterpret_cast<const float4*>(x + i));
    } else {
      for (long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(x[j]);
    }
  }
}

__global__ void absmax_kernel(const float* __restrict__ x, long n, float* __restrict__ amax) {
  pdl_wait();
  float m = 0.f;
  const long stride = (long)gridDim.x * blockDim.x * 4;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      m = fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fabsf(v.y), fmaxf(fabsf(v.z), fabsf(v.w))));
    } else {
      for (long j = i; j < n; ++j) m = fmaxf(m, fabsf(x[j]));
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  // one global atomic per BLOCK (and few blocks, see absmax_blocks): every block's atomic lands
  // on the same address and they serialise in L2 -- with one per warp of a 2368-block grid the
  // kernel spent ~10 us on them for a 6 MB tensor
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = sm[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) b = fmaxf(b, sm[w]);
    if (b > 0.f) atomic_max_pos(amax, b);
  }
}

// hi = rn_f16(s x), lo' = rn_f16((s x - hi) 2^11); s = 1 (amax == NULL) or the power of two
// that places amax[0] below 2^14; scale_out = {1/s, s}.
__global__ void split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi,
                                 __half* __restrict__ lo, long n, const float* __restrict__ amax,
                                 float* __restrict__ scale_out) {
  pdl_wait();
  const float s = amax ? f16_grad_scale(amax[0] * 1.0001f) : 1.f;
  if (scale_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    scale_out[0] = 1.f / s;
    scale_out[1] = s;
  }
  const long stride = (long)gridDim.x * blockDim.x * 4;
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      st4_f16x2(hi + i, lo + i, make_float4(v.x * s, v.y * s, v.z * s, v.w * s));
    } else {
      for (long j = i; j < n; ++j) f16_split(x[j] * s, hi[j], lo[j]);
    }
  }
}

// amax[0] = max(amax[0], max_r ||X[r, :cols]||_2): one warp per row
__global__ void rownorm_max_kernel(const float* __restrict__ X, long ld, long rows, int cols,
                                   float* __restrict__ amax) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  float m = 0.f;
  for (long r = warp; r < rows; r += nwarps) {
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float v = X[r * ld + c];
      s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    m = fmaxf(m, sqrtf(s));
  }
  if (lane == 0 && m > 0.f) atomic_max_pos(amax, m);
}

// scale_out = {1/s, s} for values bounded by a[0]*a[1] + a[2] + a[3]
__global__ void bound_scale_kernel(const float* __restrict__ a, float* __restrict__ scale_out) {
  const float bound = fmaf(a[0], a[1], a[2] + a[3]);
  const float s = f16_grad_scale(bound * 1.0001f);
  scale_out[0] = 1.f / s;
  scale_out[1] = s;
}

inline unsigned cast_blocks(long n) {
  long b = (n / 4 + 255) / 256;
  long cap = (long)pase_num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

inline unsigned absmax_blocks(long n) {
  long b = (n / 4 + 255) / 256 / 4;              // >= 4 float4 per thread
  // 2 blocks per SM for small tensors (the same-address atomics of the tail dominate: 10 ->
  // 3.9 us at 6 MB), up to 8 per SM for large ones (streaming dominates)
  long per_sm = n * 4 / (8L << 20);
  per_sm = per_sm < 2 ? 2 : (per_sm > 8 ? 8 : per_sm);
  long cap = (long)pase_num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

extern "C" {

int pase_cast_bf16(const float* x, void* dst, long n, void* stream) {
  PASE_CHECK_ARG(x && dst && n > 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0,
                 "pase_cast_bf16: bad args / alignment");
  cast_bf16_kernel<<<cast_blocks(n), 256, 0, (cudaStream_t)stream>>>(
      x, reinterpret_cast<__nv_bfloat16*>(dst), n);
  PASE_LAUNCH_CHECK("pase_cast_bf16");
  return PASE_OK;
}

int pase_absmax(const float* x, long n, float* amax, void* stream) {
  PASE_CHECK_ARG(x && amax && n > 0 && aligned16(x), "pase_absmax: bad args / alignment");
  PASE_LAUNCH((absmax_kernel), absmax_blocks(n), 256, 0, (cudaStream_t)stream, x, n, amax);
  PASE_LAUNCH_CHECK("pase_absmax");
  return PASE_OK;
}

int pase_split_f16(const float* x, void* hi, void* lo, long n, const float* amax,
                   float* scale_out, void* stream) {
  PASE_CHECK_ARG(x && hi && lo && n > 0 && aligned16(x) &&
                     (reinterpret_cast<uintptr_t>(hi) & 7u) == 0 &&
                     (reinterpret_cast<uintptr_t>(lo) & 7u) == 0,
                 "pase_split_f16: bad args / alignment");
  PASE_CHECK_ARG(amax == nullptr || scale_out != nullptr,
                 "pase_split_f16: a scaled split needs scale_out");
  PASE_LAUNCH((split_f16_kernel), cast_blocks(n), 256, 0, (cudaStream_t)stream, 
      x, reinterpret_cast<__half*>(hi), reinterpret_cast<__half*>(lo), n, amax, scale_out);
  PASE_LAUNCH_CHECK("pase_split_f16");
  return PASE_OK;
}

int pase_rownorm_max(const float* X, long ld, long rows, int cols, float* amax, void* stream) {
  PASE_CHECK_ARG(X && amax && rows > 0 && cols > 0 && ld >= cols, "pase_rownorm_max: bad args");
  long blocks = (rows * 32 + 255) / 256;
  const long cap = (long)pase_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  rownorm_max_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(X, ld, rows, cols, amax);
  PASE_LAUNCH_CHECK("pase_rownorm_max");
  return PASE_OK;
}

int pase_bound_scale(const float* a4, float* scale_out, void* stream) {
  PASE_CHECK_ARG(a4 && scale_out, "pase_bound_scale: null pointer");
  bound_scale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(a4, scale_out);
  PASE_LAUNCH_CHECK("pase_bound_scale");
  return PASE_OK;
}

}  // extern "C"
