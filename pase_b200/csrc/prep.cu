// Operand preparation for the implicit-GEMM formulation: weight re-layouts and the
// SincConv band-pass generator (forward + gradient).  All tiny, launched once per
// step per layer.
#include "common.cuh"

namespace {

__global__ void conv_w_to_fwd_kernel(const float* __restrict__ W, float* __restrict__ Wt, int Cout,
                                     int Cin, int k) {
  const long total = (long)Cout * Cin * k;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    // i indexes Wt: [co][j][ci]
    const int ci = (int)(i % Cin);
    const long r = i / Cin;
    const int j = (int)(r % k);
    const int co = (int)(r / k);
    Wt[i] = W[((long)co * Cin + ci) * k + j];
  }
}

__global__ void conv_w_from_fwd_kernel(const float* __restrict__ dWt, float* __restrict__ dW,
                                       int Cout, int Cin, int k) {
  const long total = (long)Cout * Cin * k;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    // i indexes dW: [co][ci][j]
    const int j = (int)(i % k);
    const long r = i / k;
    const int ci = (int)(r % Cin);
    const int co = (int)(r / Cin);
    dW[i] = dWt[((long)co * k + j) * Cin + ci];
  }
}

// Wd[(p*Cin+ci), (v*Cout+co)] = W[co,ci, s*(taps-1-v)+p]  (0 when tap >= k)
__global__ void conv_w_to_dgrad_kernel(const float* __restrict__ W, float* __restrict__ Wd,
                                       int Cout, int Cin, int k, int s, int taps) {
  const long total = (long)s * Cin * taps * Cout;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    long r = i / Cout;
    const int v = (int)(r % taps);
    r /= taps;
    const int ci = (int)(r % Cin);
    const int p = (int)(r / Cin);
    const int j = s * (taps - 1 - v) + p;
    Wd[i] = (j < k) ? W[((long)co * Cin + ci) * k + j] : 0.f;
  }
}

// ConvTranspose1d weight W (Cin,Cout,k): forward operand
// Wu[(p*Cout+co), (v*Cin+ci)] = W[ci,co, s*(taps-1-v)+p]  (0 when tap >= k)
__global__ void deconv_w_to_fwd_kernel(const float* __restrict__ W, float* __restrict__ Wu, int Cin,
                                       int Cout, int k, int s, int taps) {
  const long total = (long)s * Cout * taps * Cin;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    long r = i / Cin;
    const int v = (int)(r % taps);
    r /= taps;
    const int co = (int)(r % Cout);
    const int p = (int)(r / Cout);
    const int j = s * (taps - 1 - v) + p;
    Wu[i] = (j < k) ? W[((long)ci * Cout + co) * k + j] : 0.f;
  }
}

__global__ void deconv_w_from_fwd_kernel(const float* __restrict__ dWu, float* __restrict__ dW,
                                         int Cin, int Cout, int k, int s, int taps) {
  const long total = (long)Cin * Cout * k;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int j = (int)(i % k);
    long r = i / k;
    const int co = (int)(r % Cout);
    const int ci = (int)(r / Cout);
    const int p = j % s;
    const int v = taps - 1 - j / s;
    dW[i] = dWu[(((long)p * Cout + co) * taps + v) * Cin + ci];
  }
}

// Wb[ci, j*Cout+co] = W[ci,co,j]
__global__ void deconv_w_to_bwd_kernel(const float* __restrict__ W, float* __restrict__ Wb, int Cin,
                                       int Cout, int k) {
  const long total = (long)Cin * Cout * k;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    long r = i / Cout;
    const int j = (int)(r % k);
    const int ci = (int)(r / k);
    Wb[i] = W[((long)ci * Cout + co) * k + j];
  }
}

// dst (cols x ldd) = src^T, src (rows x cols, ld lds); pads columns [rows, ldd) with 0
__global__ void transpose_pad_kernel(const float* __restrict__ src, long lds,
                                     float* __restrict__ dst, long ldd, int rows, int cols) {
  pdl_wait();
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(long)r * lds + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < ldd) dst[(long)c * ldd + r] = tile[threadIdx.x][i];
  }
}

// ---- sinc filters ----
__device__ __forceinline__ void sinc_edges(float l, float b, float min_low, float min_band,
                                           float sr, float& low, float& high, bool& clamped) {
  low = min_low + fabsf(l);
  const float h = low + min_band + fabsf(b);
  const float hi = sr * 0.5f;
  clamped = (h < min_low) || (h > hi);
  high = fminf(fmaxf(h, min_low), hi);
}

// one block per output channel
__global__ void sinc_make_kernel(const float* __restrict__ low_hz, const float* __restrict__ band_hz,
                                 const float* __restrict__ n_, const float* __restrict__ win,
                                 float* __restrict__ filt, float* __restrict__ Wp, int C, int k,
                                 int fold, int Kv, float min_low, float min_band, float sr) {
  pdl_wait();
  extern __shared__ float f[];       // [k]
  const int co = blockIdx.x;
  const int half = k / 2;
  float low, high;
  bool cl;
  sinc_edges(low_hz[co], band_hz[co], min_low, min_band, sr, low, high, cl);
  const float band = high - low;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float n = n_[i];
    const float left = ((sinf(high * n) - sinf(low * n)) / (n / 2.f)) * win[i];
    const float v = left / (2.f * band);
    f[i] = v;
    f[k - 1 - i] = v;
  }
  if (threadIdx.x == 0) f[half] = (2.f * band) / (2.f * band);
  __syncthreads();
  if (filt)
    for (int i = threadIdx.x; i < k; i += blockDim.x) filt[(long)co * k + i] = f[i];
  for (int p = 0; p < fold; ++p)
    for (int kk = threadIdx.x; kk < Kv; kk += blockDim.x) {
      const int j = kk - p;
      Wp[((long)p * C + co) * Kv + kk] = (j >= 0 && j < k) ? f[j] : 0.f;
    }
}

__global__ void sinc_grad_kernel(const float* __restrict__ dWp, const float* __restrict__ low_hz,
                                 const float* __restrict__ band_hz, const float* __restrict__ n_,
                                 const float* __restrict__ win, float* __restrict__ dlow,
                                 float* __restrict__ dband, int C, int k, int fold, int Kv,
                                 float min_low, float min_band, float sr) {
  pdl_wait();
  __shared__ double r_hi[32], r_lo[32];
  const int co = blockIdx.x;
  const int half = k / 2;
  float low, high;
  bool cl;
  const float l = low_hz[co], b = band_hz[co];
  sinc_edges(l, b, min_low, min_band, sr, low, high, cl);
  const float band = high - low;
  double ghi = 0.0, glo = 0.0;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    // gradient of the two symmetric taps i and k-1-i
    float g = 0.f;
    for (int p = 0; p < fold; ++p) {
      g += dWp[((long)p * C + co) * Kv + i + p];
      g += dWp[((long)p * C + co) * Kv + (k - 1 - i) + p];
    }
    const float n = n_[i];
    const double A = (double)win[i] / ((double)n / 2.0);
    const double S = (double)sinf(high * n) - (double)sinf(low * n);
    const double b2 = 2.0 * (double)band;
    const double dS = S / (b2 * (double)band);        // S / (2 band^2)
    ghi += (double)g * A * ((double)cosf(high * n) * (double)n / b2 - dS);
    glo += (double)g * A * (-(double)cosf(low * n) * (double)n / b2 + dS);
  }
  ghi = warp_sum_d(ghi);
  glo = warp_sum_d(glo);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { r_hi[w] = ghi; r_lo[w] = glo; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double H = 0.0, L = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { H += r_hi[i]; L += r_lo[i]; }
    if (cl) H = 0.0;                                   // clamp blocks the gradient
    const double sl = (l > 0.f) ? 1.0 : ((l < 0.f) ? -1.0 : 0.0);
    const double sb = (b > 0.f) ? 1.0 : ((b < 0.f) ? -1.0 : 0.0);
    dlow[co] = (float)((L + H) * sl);
    dband[co] = (float)(H * sb);
  }
}

inline unsigned nblk(long total) {
  long b = (total + 255) / 256;
  long cap = (long)pase_num_sms() * 8;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

// ---- batched weight re-layout: every conv block of a step in ONE launch ----
// table: njobs rows of PASE_WJOB longs {src, dst, hi, lo, Cout, Cin, k, s, taps, start, count, 0}
// (device memory, built once per buffer plan).  op 0: to_fwd, 1: to_dgrad, 2: from_fwd (dst is
// an element offset into dst_base, the per-call gradient buffer).  hi/lo != 0: also write the
// 3xTF32 weight split hi = rn_tf32(v), lo = rn_tf32(v - hi) (as pase_split_tf32 does).
constexpr int WJOB = 12;
constexpr int WJOB_MAX = 32;

__device__ __forceinline__ float w_tf32_rn(float v) {
  uint32_t u = __float_as_uint(v);
  u += 0xFFFu + ((u >> 13) & 1u);
  return __uint_as_float(u & 0xFFFFE000u);
}

// fmt: operand format of the hi/lo outputs -- 0: 3xTF32 split (fp32 hi/lo), 1: bf16 into
// `hi`, 2: fp16 pair (hi, lo' = (v-hi)*2^11).  dst (fp32, may be NULL) gets the plain value.
__device__ __forceinline__ void w_store(float* dst, void* hi, void* lo, long e, float v, int fmt) {
  if (dst != nullptr) dst[e] = v;
  if (hi == nullptr) return;
  if (fmt == 0) {
    const float h = w_tf32_rn(v);
    reinterpret_cast<float*>(hi)[e] = h;
    reinterpret_cast<float*>(lo)[e] = w_tf32_rn(v - h);
  } else if (fmt == 1) {
    reinterpret_cast<__nv_bfloat16*>(hi)[e] = __float2bfloat16_rn(v);
  } else {
    __half h, l;
    f16_split(v, h, l);
    reinterpret_cast<__half*>(hi)[e] = h;
    reinterpret_cast<__half*>(lo)[e] = l;
  }
}

__global__ void conv_w_batch_kernel(const long* __restrict__ table, int njobs, long total, int op,
                                    float* __restrict__ dst_base, int fmt) {
  pdl_wait();
  __shared__ long jt[WJOB_MAX * WJOB];
  for (int i = threadIdx.x; i < njobs * WJOB; i += blockDim.x) jt[i] = table[i];
  __syncthreads();
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    int j = 0;
    while (j + 1 < njobs && i >= jt[(j + 1) * WJOB + 9]) ++j;
    const long* J = jt + j * WJOB;
    const float* src = reinterpret_cast<const float*>(J[0]);
    float* dst = (op == 2) ? dst_base + J[1] : reinterpret_cast<float*>(J[1]);
    void* hi = reinterpret_cast<void*>(J[2]);
    void* lo = reinterpret_cast<void*>(J[3]);
    const int Cout = (int)J[4], Cin = (int)J[5], k = (int)J[6], sd = (int)J[7], taps = (int)J[8];
    const long e = i - J[9];
    float v;
    if (op == 0) {                     // e indexes Wt: [co][j][ci]
      const int ci = (int)(e % Cin);
      const long r = e / Cin;
      const int jj = (int)(r % k);
      const int co = (int)(r / k);
      v = src[((long)co * Cin + ci) * k + jj];
    } else if (op == 1) {              // e indexes Wd: [p][ci][v][co]
      const int co = (int)(e % Cout);
      long r = e / Cout;
      const int vv = (int)(r % taps);
      r /= taps;
      const int ci = (int)(r % Cin);
      const int pp = (int)(r / Cin);
      const int jj = sd * (taps - 1 - vv) + pp;
      v = (jj < k) ? src[((long)co * Cin + ci) * k + jj] : 0.f;
    } else {                           // e indexes dW: [co][ci][j]
      const int jj = (int)(e % k);
      const long r = e / k;
      const int ci = (int)(r % Cin);
      const int co = (int)(r / Cin);
      v = src[((long)co * k + jj) * Cin + ci];
    }
    w_store(dst, hi, lo, e, v, fmt);
  }
}

// Tiled variants (ops 3..5 = 0..2 through shared memory, coalesced on both sides; the gather
// forms above read W with a stride of k floats).  table[.., 11] = first thread block of the
// job; blocks per job: Cout (ops 3, 5: one output-channel slab of Cin*k floats each) or
// (Cout/32)*(Cin/WB_CI) (op 4: 32 output channels x WB_CI input channels).
constexpr int WB_CI = 8;
constexpr int WB_SMEM_FLOATS = 12000;          // 46.9 KB (+ the static job row < 48 KB)

__global__ void conv_w_batch_tiled_kernel(const long* __restrict__ table, int njobs, int op,
                                          float* __restrict__ dst_base, int fmt) {
  pdl_wait();
  extern __shared__ float tile[];
  __shared__ long J[WJOB];
  if (threadIdx.x < 32) {
    // job lookup by one warp: every lane tests one job's first block (one round of loads
    // instead of a serial walk of dependent global reads by thread 0)
    const int lane = threadIdx.x;
    int j = 0;
    for (int base = 0; base < njobs; base += 32) {
      const int jj = base + lane;
      const bool le = jj < njobs && table[jj * WJOB + 11] <= (long)blockIdx.x;
      const unsigned m = __ballot_sync(0xffffffffu, le);
      if (m) j = base + 31 - __clz(m);
    }
    if (lane < WJOB) J[lane] = table[j * WJOB + lane];
  }
  __syncthreads();
  const float* src = reinterpret_cast<const float*>(J[0]);
  float* dst = (op == 5) ? dst_base + J[1] : reinterpret_cast<float*>(J[1]);
  void* hi = reinterpret_cast<void*>(J[2]);
  void* lo = reinterpret_cast<void*>(J[3]);
  const int Cout = (int)J[4], Cin = (int)J[5], k = (int)J[6], sd = (int)J[7], taps = (int)J[8];
  const int b = (int)((long)blockIdx.x - J[11]);
  // (index pairs are advanced incrementally: the integer divisions by the runtime Cin / k
  // of a per-element e % Cin, e / Cin made these loops instruction-bound -- 35 us per launch
  // for 31 MB)
  const int nt = blockDim.x;
  if (op == 3) {                       // Wt[co][j][ci] = W[co][ci][j]
    const int n = Cin * k;
    const float* w = src + (long)b * n;
    for (int i = threadIdx.x; i < n; i += nt) tile[i] = w[i];
    __syncthreads();
    int ci = threadIdx.x % Cin, jj = threadIdx.x / Cin;
    const int dci = nt % Cin, djj = nt / Cin;
    for (int e = threadIdx.x; e < n; e += nt) {
      w_store(dst, hi, lo, (long)b * n + e, tile[ci * k + jj], fmt);
      ci += dci;
      jj += djj;
      if (ci >= Cin) {
        ci -= Cin;
        ++jj;
      }
    }
  } else if (op == 5) {                // dW[co][ci][j] = dWt[co][j][ci]
    const int n = Cin * k, pitch = Cin + 1;
    const float* w = src + (long)b * n;
    {
      int ci = threadIdx.x % Cin, jj = threadIdx.x / Cin;
      const int dci = nt % Cin, djj = nt / Cin;
      for (int i = threadIdx.x; i < n; i += nt) {
        tile[jj * pitch + ci] = w[i];
        ci += dci;
        jj += djj;
        if (ci >= Cin) {
          ci -= Cin;
          ++jj;
        }
      }
    }
    __syncthreads();
    int jj = threadIdx.x % k, ci = threadIdx.x / k;
    const int djj = nt % k, dci = nt / k;
    for (int e = threadIdx.x; e < n; e += nt) {
      dst[(long)b * n + e] = tile[jj * pitch + ci];
      jj += djj;
      ci += dci;
      if (jj >= k) {
        jj -= k;
        ++ci;
      }
    }
  } else {                             // Wd[p][ci][v][co] = W[co][ci][sd*(taps-1-v)+p] | 0
    const int nci = Cin / WB_CI;
    const int co0 = (b / nci) * 32, ci0 = (b % nci) * WB_CI;
    const int seg = WB_CI * k, pitch = seg + 1;
    {
      int col = threadIdx.x / seg, r = threadIdx.x % seg;
      const int dcol = nt / seg, dr = nt % seg;
      for (int i = threadIdx.x; i < 32 * seg; i += nt) {
        tile[col * pitch + r] = src[((long)(co0 + col) * Cin + ci0) * k + r];
        col += dcol;
        r += dr;
        if (r >= seg) {
          r -= seg;
          ++col;
        }
      }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int nrow = sd * WB_CI * taps;            // output rows of 32 consecutive co
    for (int r = threadIdx.x >> 5; r < nrow; r += nwarp) {
      const int v = r % taps;
      const int cl = (r / taps) % WB_CI;
      const int pp = r / (taps * WB_CI);
      const int jj = sd * (taps - 1 - v) + pp;
      const float val = (jj < k) ? tile[lane * pitch + cl * k + jj] : 0.f;
      const long e = (((long)pp * Cin + ci0 + cl) * taps + v) * Cout + co0 + lane;
      w_store(dst, hi, lo, e, val, fmt);
    }
  }
}

// ---- batched strided copy: every small gradient of a step -> its place in the flat buffer ----
// table: njobs rows of 6 longs {src, dst, rows, cols, src_ld, dst_ld}; blocks stride over the
// jobs' elements (job boundaries from an exclusive prefix sum held in shared memory).
constexpr int SC_MAX = 128;

__global__ void scatter_copy_kernel(const long* __restrict__ table, int njobs) {
  pdl_wait();
  __shared__ long jt[SC_MAX * 6];
  __shared__ long pre[SC_MAX + 1];
  for (int i = threadIdx.x; i < njobs * 6; i += blockDim.x) jt[i] = table[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    long acc = 0;
    for (int j = 0; j < njobs; ++j) {
      pre[j] = acc;
      acc += jt[j * 6 + 2] * jt[j * 6 + 3];
    }
    pre[njobs] = acc;
  }
  __syncthreads();
  const long total = pre[njobs];
  int j = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    while (i >= pre[j + 1]) ++j;
    const long* J = jt + j * 6;
    const long e = i - pre[j];
    const long cols = J[3];
    const long r = e / cols, c = e - r * cols;
    reinterpret_cast<float*>(J[1])[r * J[5] + c] = reinterpret_cast<const float*>(J[0])[r * J[4] + c];
  }
}

}  // namespace

extern "C" {

int pase_scatter_copy(const long* table, int njobs, long total, void* stream) {
  PASE_CHECK_ARG(table && njobs > 0 && njobs <= SC_MAX && total > 0,
                 "pase_scatter_copy: bad args (njobs=%d, at most %d)", njobs, SC_MAX);
  PASE_LAUNCH((scatter_copy_kernel), nblk(total), 256, 0, (cudaStream_t)stream, table, njobs);
  PASE_LAUNCH_CHECK("pase_scatter_copy");
  return PASE_OK;
}

int pase_conv_w_to_fwd(const float* W, float* Wt, int Cout, int Cin, int k, void* stream) {
  PASE_CHECK_ARG(W && Wt && Cout > 0 && Cin > 0 && k > 0, "pase_conv_w_to_fwd: bad args");
  conv_w_to_fwd_kernel<<<nblk((long)Cout * Cin * k), 256, 0, (cudaStream_t)stream>>>(W, Wt, Cout,
                                                                                    Cin, k);
  PASE_LAUNCH_CHECK("pase_conv_w_to_fwd");
  return PASE_OK;
}

int pase_conv_w_from_fwd(const float* dWt, float* dW, int Cout, int Cin, int k, void* stream) {
  PASE_CHECK_ARG(dWt && dW && Cout > 0 && Cin > 0 && k > 0, "pase_conv_w_from_fwd: bad args");
  conv_w_from_fwd_kernel<<<nblk((long)Cout * Cin * k), 256, 0, (cudaStream_t)stream>>>(
      dWt, dW, Cout, Cin, k);
  PASE_LAUNCH_CHECK("pase_conv_w_from_fwd");
  return PASE_OK;
}

int pase_conv_w_batch(const long* table, int njobs, long total, int op, float* dst_base, int fmt,
                      void* stream) {
  PASE_CHECK_ARG(table && njobs > 0 && njobs <= WJOB_MAX && total > 0,
                 "pase_conv_w_batch: bad args (njobs=%d, at most %d)", njobs, WJOB_MAX);
  PASE_CHECK_ARG(op >= 0 && op <= 5 && ((op % 3) != 2 || dst_base != nullptr),
                 "pase_conv_w_batch: op=%d (0 to_fwd, 1 to_dgrad, 2 from_fwd + dst_base; 3..5 "
                 "tiled)", op);
  PASE_CHECK_ARG(fmt >= 0 && fmt <= 2, "pase_conv_w_batch: fmt=%d (0 tf32 split, 1 bf16, 2 f16 "
                 "pair)", fmt);
  if (op >= 3) {
    // tiled: `total` thread blocks, 48 KB of shared memory (the caller guarantees that a
    // slab fits: (Cin+1)*k resp. 32*(8*k+1) floats <= 12000, Cout % 32 == 0, Cin % 8 == 0)
    PASE_CHECK_ARG(total < (1L << 31), "pase_conv_w_batch: too many blocks");
    PASE_LAUNCH((conv_w_batch_tiled_kernel), (unsigned)total, 256, WB_SMEM_FLOATS * sizeof(float), (cudaStream_t)stream, table, njobs, op, dst_base, fmt);
    PASE_LAUNCH_CHECK("pase_conv_w_batch");
    return PASE_OK;
  }
  PASE_LAUNCH((conv_w_batch_kernel), nblk(total), 256, 0, (cudaStream_t)stream, table, njobs, total, op,
                                                                    dst_base, fmt);
  PASE_LAUNCH_CHECK("pase_conv_w_batch");
  return PASE_OK;
}

int pase_conv_w_to_dgrad(const float* W, float* Wd, int Cout, int Cin, int k, int s, int taps,
                         void* stream) {
  PASE_CHECK_ARG(W && Wd && Cout > 0 && Cin > 0 && k > 0 && s > 0 && taps * s >= k,
                 "pase_conv_w_to_dgrad: bad args (k=%d s=%d taps=%d)", k, s, taps);
  conv_w_to_dgrad_kernel<<<nblk((long)s * Cin * taps * Cout), 256, 0, (cudaStream_t)stream>>>(
      W, Wd, Cout, Cin, k, s, taps);
  PASE_LAUNCH_CHECK("pase_conv_w_to_dgrad");
  return PASE_OK;
}

int pase_deconv_w_to_fwd(const float* W, float* Wu, int Cin, int Cout, int k, int s, int taps,
                         void* stream) {
  PASE_CHECK_ARG(W && Wu && Cout > 0 && Cin > 0 && k > 0 && s > 0 && taps * s >= k,
                 "pase_deconv_w_to_fwd: bad args");
  deconv_w_to_fwd_kernel<<<nblk((long)s * Cout * taps * Cin), 256, 0, (cudaStream_t)stream>>>(
      W, Wu, Cin, Cout, k, s, taps);
  PASE_LAUNCH_CHECK("pase_deconv_w_to_fwd");
  return PASE_OK;
}

int pase_deconv_w_from_fwd(const float* dWu, float* dW, int Cin, int Cout, int k, int s, int taps,
                           void* stream) {
  PASE_CHECK_ARG(dWu && dW && Cout > 0 && Cin > 0 && k > 0 && s > 0 && taps * s >= k,
                 "pase_deconv_w_from_fwd: bad args");
  deconv_w_from_fwd_kernel<<<nblk((long)Cin * Cout * k), 256, 0, (cudaStream_t)stream>>>(
      dWu, dW, Cin, Cout, k, s, taps);
  PASE_LAUNCH_CHECK("pase_deconv_w_from_fwd");
  return PASE_OK;
}

int pase_deconv_w_to_bwd(const float* W, float* Wb, int Cin, int Cout, int k, void* stream) {
  PASE_CHECK_ARG(W && Wb && Cout > 0 && Cin > 0 && k > 0, "pase_deconv_w_to_bwd: bad args");
  deconv_w_to_bwd_kernel<<<nblk((long)Cin * Cout * k), 256, 0, (cudaStream_t)stream>>>(W, Wb, Cin,
                                                                                      Cout, k);
  PASE_LAUNCH_CHECK("pase_deconv_w_to_bwd");
  return PASE_OK;
}

int pase_transpose_pad(const float* src, long lds, float* dst, long ldd, int rows, int cols,
                       void* stream) {
  PASE_CHECK_ARG(src && dst && rows > 0 && cols > 0 && ldd >= rows && lds >= cols,
                 "pase_transpose_pad: bad args");
  dim3 grid((unsigned)((ldd + 31) / 32), (cols + 31) / 32), block(32, 8);
  PASE_LAUNCH((transpose_pad_kernel), grid, block, 0, (cudaStream_t)stream, src, lds, dst, ldd, rows, cols);
  PASE_LAUNCH_CHECK("pase_transpose_pad");
  return PASE_OK;
}

int pase_sinc_make(const float* low_hz, const float* band_hz, const float* n_, const float* window_,
                   float* filt, float* Wp, int C, int k, int fold, int Kv, float min_low,
                   float min_band, float sr, void* stream) {
  PASE_CHECK_ARG(low_hz && band_hz && n_ && window_ && Wp, "pase_sinc_make: null pointer");
  PASE_CHECK_ARG(C > 0 && (k & 1) && fold >= 1 && Kv >= k + fold - 1,
                 "pase_sinc_make: need odd k and Kv >= k+fold-1 (k=%d fold=%d Kv=%d)", k, fold, Kv);
  PASE_LAUNCH((sinc_make_kernel), C, 128, k * sizeof(float), (cudaStream_t)stream, 
      low_hz, band_hz, n_, window_, filt, Wp, C, k, fold, Kv, min_low, min_band, sr);
  PASE_LAUNCH_CHECK("pase_sinc_make");
  return PASE_OK;
}

int pase_sinc_grad(const float* dWp, const float* low_hz, const float* band_hz, const float* n_,
                   const float* window_, float* dlow, float* dband, int C, int k, int fold, int Kv,
                   float min_low, float min_band, float sr, void* stream) {
  PASE_CHECK_ARG(dWp && low_hz && band_hz && n_ && window_ && dlow && dband,
                 "pase_sinc_grad: null pointer");
  PASE_CHECK_ARG(C > 0 && (k & 1) && fold >= 1 && Kv >= k + fold - 1, "pase_sinc_grad: bad shape");
  PASE_LAUNCH((sinc_grad_kernel), C, 128, 0, (cudaStream_t)stream, dWp, low_hz, band_hz, n_, window_, dlow,
                                                        dband, C, k, fold, Kv, min_low, min_band,
                                                        sr);
  PASE_LAUNCH_CHECK("pase_sinc_grad");
  return PASE_OK;
}

}  // extern "C"
