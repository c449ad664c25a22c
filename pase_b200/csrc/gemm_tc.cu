// tcgen05 / TMEM / TMA implicit-GEMM kernels for sm_100a.
//
//   pase_tc_gemm_nt : C[map(m), n] = alpha * sum_k A[m, k] * B[n, k] + bias[n]
//       A is addressed through a plain (non-overlapping) 2-D tensor [a_rows x R]:
//       element (m, k) lives at row m + k / R, column k % R, so the implicit im2col
//       matrix of a strided convolution (R = stride * Cin) needs no materialisation
//       and no overlapping TMA strides.  B is [N x K], K-major.
//   pase_tc_gemm_tn : C[i, j] += alpha * sum_r A[r, i] * B[r, j]   (weight gradients)
//       both operands MN-major in shared memory (transpose bits set in the UMMA
//       instruction descriptor); split over the reduction, fp32 red.add epilogue.
//
// Numerics: kind::tf32 MMAs with fp32 accumulation in TMEM.
//   mode 0: single TF32 pass (10-bit mantissa operands);
//   mode 1: 3xTF32 error-compensated: operands are pre-split into exactly
//           representable tf32 "hi" and "lo" parts (hi + lo == fp32 value to 2^-22),
//           D = Ahi*Bhi + Alo*Bhi + Ahi*Blo  -> fp32-equivalent products.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one
// elected lane), warps 2-17 = 16 epilogue warps (tcgen05.ld -> fp32 register sums -> bias /
// row map / BatchNorm column statistics -> staged coalesced stores).  K-major tiles are
// 128-byte rows with the 128B swizzle; two TMEM accumulator buffers per CTA.
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>

namespace {

constexpr int BM = 128;
constexpr int BKF = 32;                  // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;                // tf32
constexpr uint32_t SPIN_LIMIT = 200u * 1000u * 1000u;

// ------------------------------------------------------------------ PTX helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SPIN_LIMIT) {
      printf("pase tc gemm: mbarrier wait timed out (tag %d, block %d,%d,%d thread %d)\n", tag,
             blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
// Epilogue warps wait for a whole accumulator chunk (thousands of cycles): an optional
// back-off between polls leaves the issue slots of their scheduler to the MMA / TMA warps.
// PASE_B200_EPI_SLEEP_NS (read once on the host) sets it; 0 = plain polling.
__constant__ int c_epi_sleep_ns = 0;
__device__ __forceinline__ void mbar_wait_epi(uint64_t* bar, uint32_t parity, int tag) {
  uint32_t spins = 0;
  const int ns = c_epi_sleep_ns;
  while (!mbar_try_wait(bar, parity)) {
    if (ns > 0) __nanosleep(ns);
    if (++spins > SPIN_LIMIT) {
      printf("pase tc gemm: mbarrier wait timed out (tag %d, block %d,%d,%d thread %d)\n", tag,
             blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmap_prefetch(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(slot)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS)
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-convergent variants: every lane of the MMA warp executes the surrounding loop and
// one elected lane issues.  Keeping the warp converged lets ptxas hold descriptors and TMEM
// addresses in uniform registers; under `if (lane == 0)` every UTCHMMA is preceded by an
// ELECT / R2UR.BROADCAST / BRA.U.ANY sequence and the issue thread, not the tensor pipe,
// bounds the MMA rate (profiles/r01_history.md).
__device__ __forceinline__ void umma_tf32_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(
          smem_u32(bar))
      : "memory");
}
// ---- CTA-pair (cta_group::2) helpers --------------------------------------------------
// Two CTAs of a cluster (same TPC) execute ONE tcgen05.mma of M = 256: each holds its own 128
// rows of A and one half of B's N rows; the leader (cluster rank 0) issues, accumulators land
// in both CTAs' TMEM.  Operand bytes read from shared memory per flop drop by 25 % (N = 128)
// against cta_group::1 -- the limiter of the 1-CTA kernel (profiles/r01_history.md).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (an address in this CTA's window) in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA load: data lands in the executing CTA's shared memory, the transaction bytes are
// signalled on the mbarrier at `bar_cluster_addr` (the leader CTA's barrier)
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* map,
                                             uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
      "l"(0x1000000000000000ull)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(slot)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS)
               : "memory");
}
__device__ __forceinline__ void umma2_tf32_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of all prior MMAs of this thread; arrives on the barrier at the same CTA-relative
// offset in both CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc_w(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t.reg .b16 m;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b16 m, 3;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], m;\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), 128B swizzle.
//   K-major : rows of 128 B, 8-row atoms of 1024 B -> SBO = 1024 B, LBO unused.
//   MN-major: 128 B (32 fp32) along MN per row, 8 k-rows per 1024 B atom; SBO = 1024 B
//             (next 8 k), LBO = byte distance between consecutive 32-element MN blocks.
//   MN-major fp32/tf32 operands must use the 32-byte-base 128B swizzle (layout type 1,
//   TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of 128 B (MN) x 4 k-rows, SBO = 512 B.
template <int LAYOUT = 2>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  d |= (uint64_t)LAYOUT << 61;                 // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// Instruction descriptor: D=f32, A=B=tf32 (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct RowMap {
  int rows_in, t_valid, rows_out, fold, cols_per_fold;
};

// transposed butterfly: v[j] per lane (lane = row) -> lane j holds sum over lanes of v[j]
__device__ __forceinline__ float colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float keep = up ? v[i + off] : v[i];
      const float send = up ? v[i] : v[i + off];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Thread layout shared by both kernels: warp 0 = TMA producer, warp 1 = TMEM owner + MMA
// issuer, warps 2..17 = 16 epilogue warps.  Epilogue warp w reads TMEM lanes 32*(w%4)..
// (hardware restriction) and owns column chunk (w-2)/4 of the tile, so every warp folds /
// stores only 32 columns: short per-warp instruction streams, 4 warps per scheduler.
constexpr int N_EPI_WARPS = 16;
constexpr int NTHREADS_V3 = (2 + N_EPI_WARPS) * 32;
constexpr int OUT_STG_FLOATS = 32 * 8;                                // 32 rows x 8 cols per pass
constexpr int OUT_STAGE_BYTES = N_EPI_WARPS * OUT_STG_FLOATS * 4;     // 16 KB
constexpr int SMEM_LIMIT = 232448;                                    // 227 KB opt-in maximum
constexpr int BAR_BYTES = 256;
constexpr int MAX_STAGES = 8;

inline int pick_stages(int stage_bytes, int stat_bytes) {
  int st = (SMEM_LIMIT - 1024 - BAR_BYTES - OUT_STAGE_BYTES - stat_bytes) / stage_bytes;
  return st > MAX_STAGES ? MAX_STAGES : st;
}
inline int smem_bytes(int stages, int stage_bytes, int stat_bytes) {
  return stages * stage_bytes + 1024 + BAR_BYTES + OUT_STAGE_BYTES + stat_bytes;
}

// Coalesced store of one warp's 32 rows x 32 columns (v[j] = column j of this lane's row).
// Four passes of 8 columns through a 1 KB swizzled staging tile; after the transpose a
// lane owns 16 B of a row and 2 lanes cover a full 32 B sector.  `nlim` = first invalid
// column of this lane's row (0 for rows that must not be written).
template <bool RED>
__device__ __forceinline__ void store_chunk(float* stg, const float (&v)[32], int lane,
                                            float* __restrict__ C, long ldc, long orow, int nlim,
                                            int nb, bool vec_ok, int accumulate) {
  const int sw = (lane >> 2) & 1;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    *reinterpret_cast<float4*>(&stg[lane * 8 + 4 * (0 ^ sw)]) =
        make_float4(v[pass * 8 + 0], v[pass * 8 + 1], v[pass * 8 + 2], v[pass * 8 + 3]);
    *reinterpret_cast<float4*>(&stg[lane * 8 + 4 * (1 ^ sw)]) =
        make_float4(v[pass * 8 + 4], v[pass * 8 + 5], v[pass * 8 + 6], v[pass * 8 + 7]);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int r = it * 16 + (lane >> 1), h = lane & 1;
      const float4 val =
          *reinterpret_cast<const float4*>(&stg[r * 8 + 4 * (h ^ ((r >> 2) & 1))]);
      const long orow_r = __shfl_sync(0xffffffffu, orow, r);
      const int nlim_r = __shfl_sync(0xffffffffu, nlim, r);
      const int n = nb + pass * 8 + h * 4;
      float* cp = C + orow_r * ldc + n;
      float x[4] = {val.x, val.y, val.z, val.w};
      if (n + 3 < nlim_r && vec_ok) {
        if (RED) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(x[0]),
                       "f"(x[1]), "f"(x[2]), "f"(x[3])
                       : "memory");
        } else {
          if (accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(cp);
            x[0] += o.x; x[1] += o.y; x[2] += o.z; x[3] += o.w;
          }
          *reinterpret_cast<float4*>(cp) = make_float4(x[0], x[1], x[2], x[3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < nlim_r) {
            if (RED) atomicAdd(cp + j, x[j]);
            else cp[j] = accumulate ? cp[j] + x[j] : x[j];
          }
      }
    }
    __syncwarp();
  }
}

// one 64-bit shared-memory descriptor = constant bits | (address >> 4)
template <int LAYOUT>
__device__ __forceinline__ uint64_t desc_hi_bits(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_desc<LAYOUT>(0, lbo_bytes, sbo_bytes);
}

// BK = fp32 elements per k-block: 32 (128-byte rows, SWIZZLE_128B) or 16 (64-byte rows,
// SWIZZLE_64B; used with BN = 256 so that four pipeline stages still fit in shared memory).
// Why BN = 256: a 128xNx8 tf32 UMMA reads (128 + N) * 32 B of shared memory in N/2 clocks --
// 128 B/clk at N = 128, i.e. the whole shared-memory bandwidth of the SM before the TMA fills
// are even counted (measured tensor-pipe activity ~58 %); N = 256 needs 96 B/clk.
template <int BN, bool SPLIT, int BK>
struct NTCfg {
  static constexpr int ROW_BYTES = BK * 4;
  static constexpr int A_BYTES = BM * ROW_BYTES;
  static constexpr int B_BYTES = BN * ROW_BYTES;
  static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int TMEM_COLS = 2 * BN;      // two accumulator buffers
  static constexpr int CPW = (BN / 32 + 3) / 4; // 32-column chunks per epilogue warp
  static constexpr int EPI_ACTIVE = BN >= 128 ? 16 : 4 * (BN / 32);
  static constexpr int LAYOUT = BK == 32 ? 2 : 4;          // SWIZZLE_128B : SWIZZLE_64B
  static constexpr int SBO = 8 * ROW_BYTES;                // 8-row core-matrix group
};

// ------------------------------------------------------------------ NT kernel ----
// Persistent: CTA b processes tiles b, b+grid, ... (n-tile fastest so concurrently running
// CTAs share A rows in L2).  The MMA warp accumulates `flush_kb` k-blocks into one of two
// TMEM buffers; the epilogue warps fold each finished buffer into fp32 register sums with
// round-to-nearest adds (the tensor core's own accumulator rounds toward zero, which drifts
// over long K) while the MMAs of the next chunk / next tile run.
template <int BN, bool SPLIT, int BK>
__global__ void __launch_bounds__(NTHREADS_V3, 1)
tc_gemm_nt_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                  const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo,
                  int R, float* __restrict__ C, long ldc, int M, int N, int K, float alpha,
                  const float* __restrict__ bias, RowMap rm, double* __restrict__ colsum,
                  double* __restrict__ colsumsq, int accumulate, int flush_kb, int stages,
                  int stat_cols) {
  using Cfg = NTCfg<BN, SPLIT, BK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* acc_full = empty_bar + MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_out = reinterpret_cast<float*>(smem + stages * Cfg::STAGE_BYTES + BAR_BYTES);
  float* s_stats = s_out + OUT_STAGE_BYTES / 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (N + BN - 1) / BN, m_tiles = (M + BM - 1) / BM;
  const int total_tiles = n_tiles * m_tiles;
  const int nkb = (K + BK - 1) / BK;
  if (flush_kb <= 0 || flush_kb > nkb) flush_kb = nkb;
  const int nchunks = (nkb + flush_kb - 1) / flush_kb;
  const bool want_stats = colsum != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], Cfg::EPI_ACTIVE);       // one arrival per active epilogue warp
    }
    fence_barrier_init();
    tmap_prefetch(&mAhi);
    tmap_prefetch(&mBhi);
    if (SPLIT) {
      tmap_prefetch(&mAlo);
      tmap_prefetch(&mBlo);
    }
  }
  if (want_stats)
    for (int i = threadIdx.x; i < 2 * stat_cols; i += NTHREADS_V3) s_stats[i] = 0.f;
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        int arow = m0, acol = 0;                 // folded-row view: k = (arow - m0) * R + acol
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 1);
          uint8_t* st = tiles + s * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          const int kf = kb * BK;
          tma_load_2d(st, &mAhi, &full_bar[s], acol, arow);
          tma_load_2d(st + Cfg::A_BYTES, &mBhi, &full_bar[s], kf, n0);
          if (SPLIT) {
            tma_load_2d(st + Cfg::A_BYTES + Cfg::B_BYTES, &mAlo, &full_bar[s], acol, arow);
            tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mBlo, &full_bar[s], kf, n0);
          }
          acol += BK;
          while (acol >= R) {
            acol -= R;
            ++arow;
          }
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // all 32 lanes run the loop (converged); one elected lane issues each MMA / commit
    constexpr uint32_t idesc = make_idesc(BM, BN, 0, 0);
    const uint64_t dconst = desc_hi_bits<Cfg::LAYOUT>(16, Cfg::SBO);
    const uint32_t tiles_u32 = smem_u32(tiles);
    const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t c = 0;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int ch = 0; ch < nchunks; ++ch, ++c) {
        const uint32_t b = c & 1, aph = (c >> 1) & 1;
        mbar_wait(&acc_empty[b], aph ^ 1, 4);
        tc_fence_after();
        const uint32_t d_tmem = tm0 + b * BN;
        const int kb_lo = ch * flush_kb;
        const int kb_hi = (kb_lo + flush_kb < nkb) ? kb_lo + flush_kb : nkb;
        for (int kb = kb_lo; kb < kb_hi; ++kb) {
          mbar_wait(&full_bar[s], ph, 2);
          tc_fence_after();
          // descriptor = constant bits + (byte address >> 4); k-step of 8 tf32 = +32 B = +2
          const uint64_t dah = dconst + ((tiles_u32 + s * Cfg::STAGE_BYTES) >> 4);
          const uint64_t dbh = dah + (Cfg::A_BYTES >> 4);
          const uint64_t dal = dbh + (Cfg::B_BYTES >> 4);
          const uint64_t dbl = dal + (Cfg::A_BYTES >> 4);
          const uint32_t first = (kb == kb_lo) ? 0u : 1u;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t acc = (k == 0) ? first : 1u;
            if (SPLIT) {
              umma_tf32_w(d_tmem, dal + 2 * k, dbh + 2 * k, idesc, acc);
              umma_tf32_w(d_tmem, dah + 2 * k, dbl + 2 * k, idesc, 1);
              umma_tf32_w(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, 1);
            } else {
              umma_tf32_w(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, acc);
            }
          }
          umma_commit_w(&empty_bar[s]);          // frees the smem stage once the MMAs retire
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit_w(&acc_full[b]);             // this chunk's accumulator is complete
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: 16 warps = 4 TMEM lane quarters x 4 column chunks ------------
    const int e = warp - 2;
    const int q = warp & 3;                      // hardware: warp w may access lanes 32*(w%4)..
    const int cc0 = e >> 2;                      // first 32-column chunk of this warp
    if (cc0 < BN / 32) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
      float* stg = s_out + e * OUT_STG_FLOATS;
      uint32_t c = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        float sums[Cfg::CPW][32];
        for (int ch = 0; ch < nchunks; ++ch, ++c) {
          const uint32_t b = c & 1, aph = (c >> 1) & 1;
          mbar_wait_epi(&acc_full[b], aph, 3);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < Cfg::CPW; ++h) {
            uint32_t raw[32];
            tmem_ld32(lane_addr + b * BN + (cc0 + 4 * h) * 32, raw);
            if (ch == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) sums[h][j] = __uint_as_float(raw[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) sums[h][j] += __uint_as_float(raw[j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[b]);        // buffer may be overwritten now
        }
        // ---- output: bias, row map / validity, optional accumulate, BatchNorm statistics ----
        const int m = m0 + q * 32 + lane;
        long orow = 0;
        int nlim = 0;                            // columns [0, nlim) of this row are valid
        if (m < M) {
          const int g = m / rm.rows_in;
          const int u = m - g * rm.rows_in;
          orow = (long)g * rm.rows_out + u;
          const long lim = (long)(rm.t_valid - u * rm.fold) * rm.cols_per_fold;
          nlim = lim <= 0 ? 0 : (lim >= N ? N : (int)lim);
        }
#pragma unroll
        for (int h = 0; h < Cfg::CPW; ++h) {
          const int nb = n0 + (cc0 + 4 * h) * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = sums[h][j] * alpha;
            if (bias != nullptr && nb + j < N) x += __ldg(bias + nb + j);
            sums[h][j] = x;
          }
          store_chunk<false>(stg, sums[h], lane, C, ldc, orow, nlim, nb, vec_ok, accumulate);
          if (want_stats) {
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              sums[h][j] = (nb + j < nlim) ? sums[h][j] : 0.f;
              sq[j] = sums[h][j] * sums[h][j];
            }
            const float s1 = colsum32(sums[h], lane);
            const float s2 = colsum32(sq, lane);
            if (nb + lane < N) {
              atomicAdd(&s_stats[nb + lane], s1);
              atomicAdd(&s_stats[stat_cols + nb + lane], s2);
            }
          }
        }
      }
    }
    if (want_stats) {
      asm volatile("bar.sync 1, 512;" ::: "memory");     // the 16 epilogue warps only
      for (int col = threadIdx.x - 64; col < N; col += N_EPI_WARPS * 32) {
        const float a = s_stats[col], b2 = s_stats[stat_cols + col];
        if (a != 0.f || b2 != 0.f) {
          atomicAdd(colsum + col, (double)a);
          atomicAdd(colsumsq + col, (double)b2);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ NT kernel, CTA pair ----
// Same contract as tc_gemm_nt_kernel; a cluster of two CTAs computes a 256 x BN tile with
// tcgen05.mma.cta_group::2 (default for N % 128 == 0; PASE_B200_TC_2CTA=0 disables it).
//   * CTA rank r loads A rows [m0 + 128 r, +128) and B rows [n0 + r BN/2, + BN/2) of every
//     k-block into ITS shared memory at the same CTA-relative offsets; both signal the
//     transaction bytes on the LEADER's `full` barrier (rank 0), which expects 2 x STAGE_BYTES.
//   * the leader's MMA warp issues the M = 256 MMAs and commits with a cluster multicast, so
//     `empty` / `acc_full` flip in both CTAs; each CTA's producer waits on its own `empty`.
//   * each CTA's 16 epilogue warps drain the CTA's own TMEM (its 128 rows x BN columns) and
//     release the buffer on the leader's `acc_empty` (count = both CTAs' epilogue warps).
template <int BN, bool SPLIT>
struct NT2Cfg {
  static constexpr int BK = 32;
  static constexpr int ROW_BYTES = BK * 4;
  static constexpr int A_BYTES = BM * ROW_BYTES;
  static constexpr int BH = BN / 2;                           // B rows held by one CTA
  static constexpr int B_BYTES = BH * ROW_BYTES;
  static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);   // per CTA
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int CPW = (BN / 32 + 3) / 4;
  static constexpr int EPI_ACTIVE = BN >= 128 ? 16 : 4 * (BN / 32);
  static constexpr int SBO = 8 * ROW_BYTES;
};

template <int BN, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS_V3, 1)
tc_gemm_nt2_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                   const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo,
                   int R, float* __restrict__ C, long ldc, int M, int N, int K, float alpha,
                   const float* __restrict__ bias, RowMap rm, double* __restrict__ colsum,
                   double* __restrict__ colsumsq, int accumulate, int flush_kb, int stages,
                   int stat_cols) {
  using Cfg = NT2Cfg<BN, SPLIT>;
  constexpr int BK = Cfg::BK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* acc_full = empty_bar + MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_out = reinterpret_cast<float*>(smem + stages * Cfg::STAGE_BYTES + BAR_BYTES);
  float* s_stats = s_out + OUT_STAGE_BYTES / 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int n_tiles = (N + BN - 1) / BN, m_tiles = (M + 2 * BM - 1) / (2 * BM);
  const int total_tiles = n_tiles * m_tiles;
  const int nkb = (K + BK - 1) / BK;
  if (flush_kb <= 0 || flush_kb > nkb) flush_kb = nkb;
  const int nchunks = (nkb + flush_kb - 1) / flush_kb;
  const bool want_stats = colsum != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);                      // leader's: one arrive.expect_tx
      mbar_init(&empty_bar[s], 1);                     // one multicast commit per release
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);                      // one multicast commit per chunk
      mbar_init(&acc_empty[b], 2 * Cfg::EPI_ACTIVE);   // leader's: both CTAs' epilogue warps
    }
    fence_barrier_init();
    tmap_prefetch(&mAhi);
    tmap_prefetch(&mBhi);
    if (SPLIT) {
      tmap_prefetch(&mAlo);
      tmap_prefetch(&mBlo);
    }
  }
  if (want_stats)
    for (int i = threadIdx.x; i < 2 * stat_cols; i += NTHREADS_V3) s_stats[i] = 0.f;
  if (warp == 1) tmem_alloc2<Cfg::TMEM_COLS>(tmem_slot);      // same warp in both CTAs
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // barriers of both CTAs initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        const int m0 = (tile / n_tiles) * (2 * BM) + (int)rank * BM;
        const int n0 = (tile % n_tiles) * BN + (int)rank * Cfg::BH;
        int arow = m0, acol = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 21);
          uint8_t* st = tiles + s * Cfg::STAGE_BYTES;
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);      // the leader's barrier
          if (leader) mbar_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
          const int kf = kb * BK;
          tma2_load_2d(st, &mAhi, fb, acol, arow);
          tma2_load_2d(st + Cfg::A_BYTES, &mBhi, fb, kf, n0);
          if (SPLIT) {
            tma2_load_2d(st + Cfg::A_BYTES + Cfg::B_BYTES, &mAlo, fb, acol, arow);
            tma2_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mBlo, fb, kf, n0);
          }
          acol += BK;
          while (acol >= R) {
            acol -= R;
            ++arow;
          }
          if (++s == stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader) {
      // converged warp, elected issue (see umma_tf32_w); M = 256 across the pair
      constexpr uint32_t idesc = make_idesc(2 * BM, BN, 0, 0);
      const uint64_t dconst = desc_hi_bits<2>(16, Cfg::SBO);
      const uint32_t tiles_u32 = smem_u32(tiles);
      const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t c = 0;
      int s = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        for (int ch = 0; ch < nchunks; ++ch, ++c) {
          const uint32_t b = c & 1, aph = (c >> 1) & 1;
          mbar_wait(&acc_empty[b], aph ^ 1, 24);
          tc_fence_after();
          const uint32_t d_tmem = tm0 + b * BN;
          const int kb_lo = ch * flush_kb;
          const int kb_hi = (kb_lo + flush_kb < nkb) ? kb_lo + flush_kb : nkb;
          for (int kb = kb_lo; kb < kb_hi; ++kb) {
            mbar_wait(&full_bar[s], ph, 22);
            tc_fence_after();
            const uint64_t dah = dconst + ((tiles_u32 + s * Cfg::STAGE_BYTES) >> 4);
            const uint64_t dbh = dah + (Cfg::A_BYTES >> 4);
            const uint64_t dal = dbh + (Cfg::B_BYTES >> 4);
            const uint64_t dbl = dal + (Cfg::A_BYTES >> 4);
            const uint32_t first = (kb == kb_lo) ? 0u : 1u;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint32_t acc = (k == 0) ? first : 1u;
              if (SPLIT) {
                umma2_tf32_w(d_tmem, dal + 2 * k, dbh + 2 * k, idesc, acc);
                umma2_tf32_w(d_tmem, dah + 2 * k, dbl + 2 * k, idesc, 1);
                umma2_tf32_w(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, 1);
              } else {
                umma2_tf32_w(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, acc);
              }
            }
            umma2_commit_mc_w(&empty_bar[s]);        // frees the stage in both CTAs
            if (++s == stages) {
              s = 0;
              ph ^= 1;
            }
          }
          umma2_commit_mc_w(&acc_full[b]);           // accumulator chunk complete in both CTAs
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: this CTA's 128 rows x BN columns ----------------
    const int e = warp - 2;
    const int q = warp & 3;
    const int cc0 = e >> 2;
    if (cc0 < BN / 32) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
      float* stg = s_out + e * OUT_STG_FLOATS;
      uint32_t c = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        const int m0 = (tile / n_tiles) * (2 * BM) + (int)rank * BM;
        const int n0 = (tile % n_tiles) * BN;
        float sums[Cfg::CPW][32];
        for (int ch = 0; ch < nchunks; ++ch, ++c) {
          const uint32_t b = c & 1, aph = (c >> 1) & 1;
          mbar_wait_epi(&acc_full[b], aph, 23);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < Cfg::CPW; ++h) {
            uint32_t raw[32];
            tmem_ld32(lane_addr + b * BN + (cc0 + 4 * h) * 32, raw);
            if (ch == 0) {
#pragma unroll
              for (int j = 0; j < 32; ++j) sums[h][j] = __uint_as_float(raw[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) sums[h][j] += __uint_as_float(raw[j]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&acc_empty[b]), 0));
        }
        const int m = m0 + q * 32 + lane;
        long orow = 0;
        int nlim = 0;
        if (m < M) {
          const int g = m / rm.rows_in;
          const int u = m - g * rm.rows_in;
          orow = (long)g * rm.rows_out + u;
          const long lim = (long)(rm.t_valid - u * rm.fold) * rm.cols_per_fold;
          nlim = lim <= 0 ? 0 : (lim >= N ? N : (int)lim);
        }
#pragma unroll
        for (int h = 0; h < Cfg::CPW; ++h) {
          const int nb = n0 + (cc0 + 4 * h) * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = sums[h][j] * alpha;
            if (bias != nullptr && nb + j < N) x += __ldg(bias + nb + j);
            sums[h][j] = x;
          }
          store_chunk<false>(stg, sums[h], lane, C, ldc, orow, nlim, nb, vec_ok, accumulate);
          if (want_stats) {
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              sums[h][j] = (nb + j < nlim) ? sums[h][j] : 0.f;
              sq[j] = sums[h][j] * sums[h][j];
            }
            const float s1 = colsum32(sums[h], lane);
            const float s2 = colsum32(sq, lane);
            if (nb + lane < N) {
              atomicAdd(&s_stats[nb + lane], s1);
              atomicAdd(&s_stats[stat_cols + nb + lane], s2);
            }
          }
        }
      }
    }
    if (want_stats) {
      asm volatile("bar.sync 1, 512;" ::: "memory");     // the 16 epilogue warps only
      for (int col = threadIdx.x - 64; col < N; col += N_EPI_WARPS * 32) {
        const float a = s_stats[col], b2 = s_stats[stat_cols + col];
        if (a != 0.f || b2 != 0.f) {
          atomicAdd(colsum + col, (double)a);
          atomicAdd(colsumsq + col, (double)b2);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();            // neither CTA frees TMEM / exits while the pair still uses it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ NT window kernel ----
// Same contract as tc_gemm_nt_kernel, but the A operand is fetched once per 32-float column
// block instead of once per k-block.  With A(m, kk) = F[m + kk/R, kk%R] the k-blocks
// kk = q*R + cb*32 (q = 0..Q-1) of one column block cb read the SAME columns of F, shifted by
// q rows: one TMA box of (BM + Q - 1) rows serves all of them, the UMMA descriptor simply
// starts q rows (q*128 B) further down.  (Verified on B200, tools/rowshift_probe.py: K-major
// SWIZZLE_128B operands may start at any row with base_offset = 0 -- the swizzle is a function
// of absolute shared-memory address bits.)  For a k-tap stride-s convolution this divides the
// activation traffic by ~k/s; the weight tiles stream through a ring of B stages.
template <int BN, bool SPLIT>
struct NTWCfg {
  static constexpr int B_BYTES = BN * 128;
  static constexpr int B_STAGE_BYTES = (SPLIT ? 2 : 1) * B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int EPI_ACTIVE = 4 * (BN / 32);
};

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS_V3, 1)
tc_gemm_ntw_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                   const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo,
                   int R, float* __restrict__ C, long ldc, int M, int N, int K, float alpha,
                   const float* __restrict__ bias, RowMap rm, double* __restrict__ colsum,
                   double* __restrict__ colsumsq, int accumulate, int flush_kb, int stages,
                   int stat_cols, int win_rows) {
  using Cfg = NTWCfg<BN, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int win_bytes = win_rows * 128;                       // one hi (or lo) window
  const int win_pair = (SPLIT ? 2 : 1) * win_bytes;
  uint8_t* wins = smem;                                       // 2 window buffers (hi [, lo])
  uint8_t* btiles = smem + 2 * win_pair;
  uint8_t* after = btiles + stages * Cfg::B_STAGE_BYTES;
  uint64_t* b_full = reinterpret_cast<uint64_t*>(after);
  uint64_t* b_empty = b_full + MAX_STAGES;
  uint64_t* w_full = b_empty + MAX_STAGES;
  uint64_t* w_empty = w_full + 2;
  uint64_t* acc_full = w_empty + 2;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_out = reinterpret_cast<float*>(after + BAR_BYTES);
  float* s_stats = s_out + OUT_STAGE_BYTES / 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (N + BN - 1) / BN, m_tiles = (M + BM - 1) / BM;
  const int total_tiles = n_tiles * m_tiles;
  const int nkb = K / BKF;                                    // K % 32 == 0 (host-checked)
  const int ncb = (R < K ? R : K) / BKF;                      // column blocks of a folded row
  if (flush_kb <= 0 || flush_kb > nkb) flush_kb = nkb;
  const bool want_stats = colsum != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&w_full[b], 1);
      mbar_init(&w_empty[b], 1);
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], Cfg::EPI_ACTIVE);
    }
    fence_barrier_init();
    tmap_prefetch(&mAhi);
    tmap_prefetch(&mBhi);
    if (SPLIT) {
      tmap_prefetch(&mAlo);
      tmap_prefetch(&mBlo);
    }
  }
  if (want_stats)
    for (int i = threadIdx.x; i < 2 * stat_cols; i += NTHREADS_V3) s_stats[i] = 0.f;
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // q-steps of column block cb: kk = q*R + cb*32 < K
  auto nq_of = [&](int cb) { return (K - cb * BKF + R - 1) / R; };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t wi = 0, bi = 0;
      auto issue_window = [&](int tile, int cb) {
        const uint32_t w = wi & 1, wph = (wi >> 1) & 1;
        mbar_wait(&w_empty[w], wph ^ 1, 5);
        const int m0 = (tile / n_tiles) * BM;
        uint8_t* dst = wins + w * win_pair;
        mbar_expect_tx(&w_full[w], win_pair);
        tma_load_2d(dst, &mAhi, &w_full[w], cb * BKF, m0);
        if (SPLIT) tma_load_2d(dst + win_bytes, &mAlo, &w_full[w], cb * BKF, m0);
        ++wi;
      };
      int tile = blockIdx.x, cb = 0;
      if (tile < total_tiles) issue_window(tile, 0);
      while (tile < total_tiles) {
        const int n0 = (tile % n_tiles) * BN;
        const int nq = nq_of(cb);
        // successor (tile, cb) whose window is prefetched while this block's last B tiles load
        int ntile = tile, ncbn = cb + 1;
        if (ncbn == ncb) { ncbn = 0; ntile = tile + gridDim.x; }
        const int pre_at = nq > stages ? nq - stages : 0;
        for (int q = 0; q < nq; ++q, ++bi) {
          if (q == pre_at && ntile < total_tiles) issue_window(ntile, ncbn);
          const int s = bi % stages;
          const uint32_t ph = (bi / stages) & 1;
          mbar_wait(&b_empty[s], ph ^ 1, 1);
          uint8_t* st = btiles + s * Cfg::B_STAGE_BYTES;
          mbar_expect_tx(&b_full[s], Cfg::B_STAGE_BYTES);
          const int kf = q * R + cb * BKF;
          tma_load_2d(st, &mBhi, &b_full[s], kf, n0);
          if (SPLIT) tma_load_2d(st + Cfg::B_BYTES, &mBlo, &b_full[s], kf, n0);
        }
        tile = ntile;
        cb = ncbn;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, 0, 0);
      const uint64_t dconst = desc_hi_bits<2>(16, 1024);
      const uint32_t wins_u32 = smem_u32(wins), bt_u32 = smem_u32(btiles);
      uint32_t wi = 0, bi = 0, c = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int in_chunk = 0, done_kb = 0;
        uint32_t b = c & 1;
        mbar_wait(&acc_empty[b], ((c >> 1) & 1) ^ 1, 4);
        tc_fence_after();
        for (int cb = 0; cb < ncb; ++cb, ++wi) {
          const uint32_t w = wi & 1;
          mbar_wait(&w_full[w], (wi >> 1) & 1, 6);
          tc_fence_after();
          const uint32_t a_hi0 = wins_u32 + w * win_pair;
          const int nq = nq_of(cb);
          for (int q = 0; q < nq; ++q, ++bi) {
            const int s = bi % stages;
            mbar_wait(&b_full[s], (bi / stages) & 1, 2);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + b * BN;
            // A starts q rows into the window; k-step of 8 tf32 = +32 B
            const uint64_t dah = dconst + ((a_hi0 + q * 128) >> 4);
            const uint64_t dal = dconst + ((a_hi0 + win_bytes + q * 128) >> 4);
            const uint64_t dbh = dconst + ((bt_u32 + s * Cfg::B_STAGE_BYTES) >> 4);
            const uint64_t dbl = dbh + (Cfg::B_BYTES >> 4);
            const uint32_t first = in_chunk == 0 ? 0u : 1u;
#pragma unroll
            for (int k = 0; k < BKF / UMMA_K; ++k) {
              const uint32_t acc = (k == 0) ? first : 1u;
              if (SPLIT) {
                umma_tf32(d_tmem, dal + 2 * k, dbh + 2 * k, idesc, acc);
                umma_tf32(d_tmem, dah + 2 * k, dbl + 2 * k, idesc, 1);
                umma_tf32(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, 1);
              } else {
                umma_tf32(d_tmem, dah + 2 * k, dbh + 2 * k, idesc, acc);
              }
            }
            umma_commit(&b_empty[s]);
            ++in_chunk;
            ++done_kb;
            if (in_chunk == flush_kb || done_kb == nkb) {
              umma_commit(&acc_full[b]);           // chunk complete -> epilogue folds it
              ++c;
              in_chunk = 0;
              if (done_kb < nkb) {
                b = c & 1;
                mbar_wait(&acc_empty[b], ((c >> 1) & 1) ^ 1, 4);
                tc_fence_after();
              }
            }
          }
          umma_commit(&w_empty[w]);                // window free once its MMAs retire
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: identical to tc_gemm_nt_kernel ----------------
    const int e = warp - 2;
    const int q = warp & 3;
    const int cc = e >> 2;
    const int nchunks = (nkb + flush_kb - 1) / flush_kb;
    if (cc < BN / 32) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32);
      const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
      float* stg = s_out + e * OUT_STG_FLOATS;
      uint32_t c = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        float sums[32];
        for (int ch = 0; ch < nchunks; ++ch, ++c) {
          const uint32_t b = c & 1, aph = (c >> 1) & 1;
          mbar_wait_epi(&acc_full[b], aph, 3);
          tc_fence_after();
          uint32_t raw[32];
          tmem_ld32(taddr + b * BN, raw);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[b]);
          if (ch == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sums[j] = __uint_as_float(raw[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) sums[j] += __uint_as_float(raw[j]);
          }
        }
        const int m = m0 + q * 32 + lane;
        long orow = 0;
        int nlim = 0;
        if (m < M) {
          const int g = m / rm.rows_in;
          const int u = m - g * rm.rows_in;
          orow = (long)g * rm.rows_out + u;
          const long lim = (long)(rm.t_valid - u * rm.fold) * rm.cols_per_fold;
          nlim = lim <= 0 ? 0 : (lim >= N ? N : (int)lim);
        }
        const int nb = n0 + cc * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = sums[j] * alpha;
          if (bias != nullptr && nb + j < N) x += __ldg(bias + nb + j);
          sums[j] = x;
        }
        store_chunk<false>(stg, sums, lane, C, ldc, orow, nlim, nb, vec_ok, accumulate);
        if (want_stats) {
          float sq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            sums[j] = (nb + j < nlim) ? sums[j] : 0.f;
            sq[j] = sums[j] * sums[j];
          }
          const float s1 = colsum32(sums, lane);
          const float s2 = colsum32(sq, lane);
          if (nb + lane < N) {
            atomicAdd(&s_stats[nb + lane], s1);
            atomicAdd(&s_stats[stat_cols + nb + lane], s2);
          }
        }
      }
    }
    if (want_stats) {
      asm volatile("bar.sync 1, 512;" ::: "memory");
      for (int col = threadIdx.x - 64; col < N; col += N_EPI_WARPS * 32) {
        const float a = s_stats[col], b2 = s_stats[stat_cols + col];
        if (a != 0.f || b2 != 0.f) {
          atomicAdd(colsum + col, (double)a);
          atomicAdd(colsumsq + col, (double)b2);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ TN kernel ----
// C[i,j] += alpha * sum_r A[r,i] B[r,j].  A: 3-D tensor (I inner, rows-per-group, groups);
// B addressed through the folded-row trick: (r, j) -> row u + j/R of group g, column j % R.
constexpr int TN_KR = 32;                // reduction rows per stage (4 UMMA k-steps of 8)

template <int BN, bool SPLIT>
struct TNCfg {
  static constexpr int A_BYTES = TN_KR * BM * 4;          // 4 MN-blocks of [32 rows x 128 B]
  static constexpr int B_BYTES = TN_KR * BN * 4;
  static constexpr int STAGE_BYTES = (SPLIT ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int EPI_ACTIVE = 4 * (BN / 32);
};

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS_V3, 1)
tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                  const __grid_constant__ CUtensorMap mBhi, const __grid_constant__ CUtensorMap mBlo,
                  int R, float* __restrict__ C, long ldc, int I, int J, int groups,
                  int rows_per_group, float alpha, int chunks_per_split, int flush_ch,
                  int stages, int b_blocked) {
  using Cfg = TNCfg<BN, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* acc_full = empty_bar + MAX_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_out = reinterpret_cast<float*>(smem + stages * Cfg::STAGE_BYTES + BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * BN, i0 = blockIdx.y * BM;
  const int cpg = (rows_per_group + TN_KR - 1) / TN_KR;       // chunks per group
  const long total_chunks = (long)groups * cpg;
  const long c_begin = (long)blockIdx.z * chunks_per_split;
  long c_end = c_begin + chunks_per_split;
  if (c_end > total_chunks) c_end = total_chunks;
  const int nch = (int)(c_end - c_begin);
  if (flush_ch <= 0 || flush_ch > nch) flush_ch = nch > 0 ? nch : 1;
  const int nflush = (nch + flush_ch - 1) / flush_ch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], Cfg::EPI_ACTIVE);
    }
    fence_barrier_init();
    tmap_prefetch(&mAhi);
    tmap_prefetch(&mBhi);
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (nch <= 0) {                      // uniform across the CTA
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    return;
  }
  const int jq = j0 / R, jc = j0 % R;  // folded-row offset / column of this B tile

  if (warp == 0) {
    if (lane == 0) {
      int s = -1;
      uint32_t ph = 1;
      for (int it = 0; it < nch; ++it) {
        if (++s == stages || it == 0) {
          s = 0;
          ph ^= 1;
        }
        mbar_wait(&empty_bar[s], ph ^ 1, 11);
        const long ch = c_begin + it;
        const int g = (int)(ch / cpg);
        const int u0 = (int)(ch - (long)g * cpg) * TN_KR;
        uint8_t* st = tiles + s * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        uint8_t* a_hi = st;
        uint8_t* b_hi = st + Cfg::A_BYTES;
        uint8_t* a_lo = b_hi + Cfg::B_BYTES;
        uint8_t* b_lo = a_lo + Cfg::A_BYTES;
        // A: one 4-D box = (32 floats, 32 rows, BM/32 column blocks, 1 group) lands as
        // BM/32 consecutive [32 rows x 128 B] MN-blocks
        tma_load_4d(a_hi, &mAhi, &full_bar[s], 0, u0, i0 / 32, g);
        if (SPLIT) tma_load_4d(a_lo, &mAlo, &full_bar[s], 0, u0, i0 / 32, g);
        if (b_blocked) {
          // the BN columns of this tile lie inside one folded row (R % BN == 0)
          tma_load_4d(b_hi, &mBhi, &full_bar[s], 0, u0 + jq, jc / 32, g);
          if (SPLIT) tma_load_4d(b_lo, &mBlo, &full_bar[s], 0, u0 + jq, jc / 32, g);
        } else {
#pragma unroll
          for (int nb = 0; nb < BN / 32; ++nb) {
            // 32 consecutive columns never straddle a folded row (R % 32 == 0)
            const int col = jc + nb * 32;
            const int qq = jq + col / R, cc = col % R;
            tma_load_4d(b_hi + nb * TN_KR * 128, &mBhi, &full_bar[s], 0, u0 + qq, cc / 32, g);
            if (SPLIT)
              tma_load_4d(b_lo + nb * TN_KR * 128, &mBlo, &full_bar[s], 0, u0 + qq, cc / 32, g);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // converged warp, elected issue (see umma_tf32_w)
    constexpr uint32_t idesc = make_idesc(BM, BN, 1, 1);
    const uint64_t dconst = desc_hi_bits<1>(TN_KR * 128, 512);
    const uint32_t tiles_u32 = smem_u32(tiles);
    const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
    int s = 0;
    uint32_t ph = 0;
    for (int f = 0; f < nflush; ++f) {
      const uint32_t b = f & 1, aph = (f >> 1) & 1;
      mbar_wait(&acc_empty[b], aph ^ 1, 14);
      tc_fence_after();
      const uint32_t d_tmem = tm0 + b * BN;
      const int it_lo = f * flush_ch;
      const int it_hi = (it_lo + flush_ch < nch) ? it_lo + flush_ch : nch;
      for (int it = it_lo; it < it_hi; ++it) {
        mbar_wait(&full_bar[s], ph, 12);
        tc_fence_after();
        const uint64_t dah = dconst + ((tiles_u32 + s * Cfg::STAGE_BYTES) >> 4);
        const uint64_t dbh = dah + (Cfg::A_BYTES >> 4);
        const uint64_t dal = dbh + (Cfg::B_BYTES >> 4);
        const uint64_t dbl = dal + (Cfg::A_BYTES >> 4);
        const uint32_t first = (it == it_lo) ? 0u : 1u;
#pragma unroll
        for (int k = 0; k < TN_KR / UMMA_K; ++k) {
          const uint32_t acc = (k == 0) ? first : 1u;
          const uint32_t o = k * (1024 >> 4);          // 8 k-rows of 128 B
          if (SPLIT) {
            umma_tf32_w(d_tmem, dal + o, dbh + o, idesc, acc);
            umma_tf32_w(d_tmem, dah + o, dbl + o, idesc, 1);
            umma_tf32_w(d_tmem, dah + o, dbh + o, idesc, 1);
          } else {
            umma_tf32_w(d_tmem, dah + o, dbh + o, idesc, acc);
          }
        }
        umma_commit_w(&empty_bar[s]);
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
      }
      umma_commit_w(&acc_full[b]);
    }
    __syncwarp();
  } else {
    const int e = warp - 2;
    const int q = warp & 3;
    const int cc = e >> 2;
    if (cc < BN / 32) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32);
      float sums[32];
      for (int f = 0; f < nflush; ++f) {
        const uint32_t b = f & 1, aph = (f >> 1) & 1;
        mbar_wait_epi(&acc_full[b], aph, 13);
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld32(taddr + b * BN, raw);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[b]);
        if (f == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sums[j] = __uint_as_float(raw[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) sums[j] += __uint_as_float(raw[j]);
        }
      }
      const int i = i0 + q * 32 + lane;
      const bool vec_ok = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
#pragma unroll
      for (int j = 0; j < 32; ++j) sums[j] *= alpha;
      store_chunk<true>(s_out + e * OUT_STG_FLOATS, sums, lane, C, ldc, (long)i, i < I ? J : 0,
                        j0 + cc * 32, vec_ok, 0);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side ----
static bool pase_tc_use_window() {
  static int v = -1;
  if (v < 0) {
    // off by default: measured slower than the per-k-block kernel (the GEMMs are bound by
    // shared-memory bandwidth, not by operand traffic); PASE_B200_TC_WINDOW=1 enables it
    const char* e = getenv("PASE_B200_TC_WINDOW");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

// one-time upload of the epilogue poll back-off (default 0 = plain polling)
static void pase_tc_init_epi_sleep() {
  static bool done = false;
  if (done) return;
  done = true;
  const char* e = getenv("PASE_B200_EPI_SLEEP_NS");
  const int ns = e ? atoi(e) : 0;
  if (ns > 0) cudaMemcpyToSymbol(c_epi_sleep_ns, &ns, sizeof(int));
}

static bool pase_tc_use_2cta() {
  static int v = -1;
  if (v < 0) {
    // CTA-pair kernel (tcgen05 cta_group::2): on by default (NT 2.41 -> 2.16 ms per PASE+
    // step, profiles/r01_history.md); PASE_B200_TC_2CTA=0 selects the 1-CTA kernel
    const char* e = getenv("PASE_B200_TC_2CTA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || p == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// rank-2 or rank-3 fp32 tensor map, 128B swizzle, zero OOB fill.
int make_map(CUtensorMap* map, const float* base, int rank, const uint64_t* dims,
             const uint64_t* strides_bytes, const uint32_t* box, const char* what,
             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    pase_set_error("pase tc gemm: cuTensorMapEncodeTiled not available");
    return PASE_ERR_UNSUPPORTED;
  }
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base,
                   (const cuuint64_t*)dims, (const cuuint64_t*)strides_bytes,
                   (const cuuint32_t*)box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    pase_set_error("pase tc gemm: cuTensorMapEncodeTiled(%s) failed with %d (dims %llu,%llu,%llu "
                   "stride %llu,%llu box %u,%u,%u)",
                   what, (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                   (unsigned long long)(rank > 2 ? dims[2] : 0),
                   (unsigned long long)strides_bytes[0],
                   (unsigned long long)(rank > 2 ? strides_bytes[1] : 0), box[0], box[1],
                   rank > 2 ? box[2] : 0);
    return PASE_ERR_ARG;
  }
  return PASE_OK;
}

template <int BN, bool SPLIT, int BK>
int launch_nt(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh,
              const CUtensorMap& bl, int R, float* C, long ldc, int M, int N, int K, float alpha,
              const float* bias, RowMap rm, double* cs, double* cq, int accumulate, int flush_kb,
              cudaStream_t st) {
  using Cfg = NTCfg<BN, SPLIT, BK>;
  static bool attr = false;
  const int stat_cols = cs ? ((N + 31) / 32) * 32 : 0;
  const int stages = pick_stages(Cfg::STAGE_BYTES, 2 * stat_cols * 4);
  const int smem = smem_bytes(stages, Cfg::STAGE_BYTES, 2 * stat_cols * 4);
  if (stages < 2) {
    pase_set_error("pase_tc_gemm_nt: not enough shared memory for 2 pipeline stages");
    return PASE_ERR_UNSUPPORTED;
  }
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_nt_kernel<BN, SPLIT, BK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) {
      pase_set_error("pase_tc_gemm_nt: smem attribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr = true;
  }
  const long tiles = (long)((N + BN - 1) / BN) * ((M + BM - 1) / BM);
  const int grid = (int)(tiles < pase_num_sms() ? tiles : pase_num_sms());
  tc_gemm_nt_kernel<BN, SPLIT, BK><<<grid, NTHREADS_V3, smem, st>>>(
      ah, al, bh, bl, R, C, ldc, M, N, K, alpha, bias, rm, cs, cq, accumulate, flush_kb, stages,
      stat_cols);
  PASE_LAUNCH_CHECK("pase_tc_gemm_nt");
  return PASE_OK;
}

template <int BN, bool SPLIT>
int launch_nt2(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh,
               const CUtensorMap& bl, int R, float* C, long ldc, int M, int N, int K, float alpha,
               const float* bias, RowMap rm, double* cs, double* cq, int accumulate, int flush_kb,
               cudaStream_t st) {
  using Cfg = NT2Cfg<BN, SPLIT>;
  static bool attr = false;
  const int stat_cols = cs ? ((N + 31) / 32) * 32 : 0;
  const int stages = pick_stages(Cfg::STAGE_BYTES, 2 * stat_cols * 4);
  const int smem = smem_bytes(stages, Cfg::STAGE_BYTES, 2 * stat_cols * 4);
  if (stages < 2) {
    pase_set_error("pase_tc_gemm_nt (2-CTA): not enough shared memory for 2 pipeline stages");
    return PASE_ERR_UNSUPPORTED;
  }
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_nt2_kernel<BN, SPLIT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) {
      pase_set_error("pase_tc_gemm_nt (2-CTA): smem attribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr = true;
  }
  const long tiles = (long)((N + BN - 1) / BN) * ((M + 2 * BM - 1) / (2 * BM));
  const long max_pairs = pase_num_sms() / 2;
  const int grid = 2 * (int)(tiles < max_pairs ? tiles : max_pairs);   // cluster dims (2,1,1)
  tc_gemm_nt2_kernel<BN, SPLIT><<<grid, NTHREADS_V3, smem, st>>>(
      ah, al, bh, bl, R, C, ldc, M, N, K, alpha, bias, rm, cs, cq, accumulate, flush_kb, stages,
      stat_cols);
  PASE_LAUNCH_CHECK("pase_tc_gemm_nt(2cta)");
  return PASE_OK;
}

template <int BN, bool SPLIT>
int launch_ntw(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh,
               const CUtensorMap& bl, int R, float* C, long ldc, int M, int N, int K, float alpha,
               const float* bias, RowMap rm, double* cs, double* cq, int accumulate, int flush_kb,
               int win_rows, cudaStream_t st) {
  using Cfg = NTWCfg<BN, SPLIT>;
  static bool attr = false;
  const int stat_cols = cs ? ((N + 31) / 32) * 32 : 0;
  const int win_total = 2 * (SPLIT ? 2 : 1) * win_rows * 128;
  int stages = (SMEM_LIMIT - 1024 - BAR_BYTES - OUT_STAGE_BYTES - 2 * stat_cols * 4 - win_total) /
               Cfg::B_STAGE_BYTES;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) {
    pase_set_error("pase_tc_gemm_nt: not enough shared memory for the window pipeline");
    return PASE_ERR_UNSUPPORTED;
  }
  const int smem = win_total + stages * Cfg::B_STAGE_BYTES + 1024 + BAR_BYTES + OUT_STAGE_BYTES +
                   2 * stat_cols * 4;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_ntw_kernel<BN, SPLIT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) {
      pase_set_error("pase_tc_gemm_nt(window): smem attribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr = true;
  }
  const long tiles = (long)((N + BN - 1) / BN) * ((M + BM - 1) / BM);
  const int grid = (int)(tiles < pase_num_sms() ? tiles : pase_num_sms());
  tc_gemm_ntw_kernel<BN, SPLIT><<<grid, NTHREADS_V3, smem, st>>>(
      ah, al, bh, bl, R, C, ldc, M, N, K, alpha, bias, rm, cs, cq, accumulate, flush_kb, stages,
      stat_cols, win_rows);
  PASE_LAUNCH_CHECK("pase_tc_gemm_nt(window)");
  return PASE_OK;
}

template <int BN, bool SPLIT>
int launch_tn(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh,
              const CUtensorMap& bl, int R, float* C, long ldc, int I, int J, int groups,
              int rows_per_group, float alpha, int flush_ch, int b_blocked, cudaStream_t st) {
  using Cfg = TNCfg<BN, SPLIT>;
  static bool attr = false;
  const int stages = pick_stages(Cfg::STAGE_BYTES, 0);
  const int smem = smem_bytes(stages, Cfg::STAGE_BYTES, 0);
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_tn_kernel<BN, SPLIT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
    if (e != cudaSuccess) {
      pase_set_error("pase_tc_gemm_tn: smem attribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr = true;
  }
  const int ti = (I + BM - 1) / BM, tj = (J + BN - 1) / BN;
  const int cpg = (rows_per_group + TN_KR - 1) / TN_KR;
  const long total = (long)groups * cpg;
  // split the reduction so that the CTA count fills whole waves of the SMs (one CTA per SM):
  // among 1..4 waves pick the split with the best fill, keeping chains >= 8 chunks
  const long tiles = (long)ti * tj;
  const int sms = pase_num_sms();
  long best = 1;
  double best_eff = -1.0;
  for (long sp = 1; sp <= total && sp * tiles <= 4L * sms; ++sp) {
    const long cps_try = (total + sp - 1) / sp;
    if (cps_try < 8 && sp > 1) break;
    const long ctas = tiles * ((total + cps_try - 1) / cps_try);
    const long waves = (ctas + sms - 1) / sms;
    // time ~ waves * chunks-per-cta: lower is better
    const double cost = (double)waves * (double)(cps_try + 6);
    const double eff = 1.0 / cost;
    if (eff > best_eff) { best_eff = eff; best = sp; }
  }
  long splits = best;
  long cps = (total + splits - 1) / splits;
  splits = (total + cps - 1) / cps;
  dim3 grid(tj, ti, (unsigned)splits);
  tc_gemm_tn_kernel<BN, SPLIT><<<grid, NTHREADS_V3, smem, st>>>(ah, al, bh, bl, R, C, ldc, I, J, groups,
                                                            rows_per_group, alpha, (int)cps,
                                                            flush_ch, stages, b_blocked);
  PASE_LAUNCH_CHECK("pase_tc_gemm_tn");
  return PASE_OK;
}

// tf32 helpers: the tensor core reads the upper 19 bits of an fp32 operand (truncation).
__device__ __forceinline__ float tf32_trunc(float v) {
  return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
}
__device__ __forceinline__ float tf32_rn(float v) {          // round to nearest even
  uint32_t u = __float_as_uint(v);
  u += 0xFFFu + ((u >> 13) & 1u);
  return __uint_as_float(u & 0xFFFFE000u);
}
// split kernel.
//   hi == nullptr (activations): the operand itself is "hi" (hardware truncation), and
//        lo = rn(x - trunc(x));
//   hi != nullptr (weights): hi = rn(x), lo = rn(x - hi): explicit, exactly representable,
//        zero-mean residuals, so the dropped lo*lo / residual terms carry no coherent bias.
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                  float* __restrict__ lo, long n) {
  const long i4 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long stride = (long)gridDim.x * blockDim.x * 4;
  const bool explicit_hi = hi != nullptr;
  for (long i = i4; i < n; i += stride) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      const float vv[4] = {v.x, v.y, v.z, v.w};
      float h[4], l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        h[k] = explicit_hi ? tf32_rn(vv[k]) : tf32_trunc(vv[k]);
        l[k] = tf32_rn(vv[k] - h[k]);
      }
      if (explicit_hi) *reinterpret_cast<float4*>(hi + i) = make_float4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<float4*>(lo + i) = make_float4(l[0], l[1], l[2], l[3]);
    } else {
      for (long j = i; j < n; ++j) {
        const float v = x[j];
        const float h = explicit_hi ? tf32_rn(v) : tf32_trunc(v);
        if (explicit_hi) hi[j] = h;
        lo[j] = tf32_rn(v - h);
      }
    }
  }
}

}  // namespace

extern "C" {

int pase_split_tf32(const float* x, float* hi, float* lo, long n, void* stream) {
  // hi may be NULL: kind::tf32 ignores the low 13 mantissa bits of its fp32 operands
  // (verified on B200, tools/tf32_probe.py), so the original array serves as "hi".
  PASE_CHECK_ARG(x && lo && n > 0, "pase_split_tf32: bad args");
  PASE_CHECK_ARG(aligned16(x) && aligned16(lo) && (hi == nullptr || aligned16(hi)),
                 "pase_split_tf32: alignment");
  long blocks = (n / 4 + 255) / 256;
  long cap = (long)pase_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n);
  PASE_LAUNCH_CHECK("pase_split_tf32");
  return PASE_OK;
}

// A: [a_rows x R] fp32 (hi / lo parts, lo may be NULL for mode 0); B: [N x K] (ldb = K).
int pase_tc_gemm_nt(const float* Ahi, const float* Alo, long a_rows, int R, const float* Bhi,
                    const float* Blo, long ldb, float* C, long ldc, int M, int N, int K,
                    float alpha, const float* bias, int rows_in, int t_valid, int rows_out,
                    int fold, double* colsum, double* colsumsq, int accumulate, int mode,
                    void* stream) {
  PASE_CHECK_ARG(Ahi && Bhi && C && M > 0 && N > 0 && K > 0, "pase_tc_gemm_nt: bad args");
  pase_tc_init_epi_sleep();
  PASE_CHECK_ARG(mode == 0 || (Alo && Blo), "pase_tc_gemm_nt: mode 1 needs lo operands");
  PASE_CHECK_ARG(R >= 32 && (R % 32) == 0, "pase_tc_gemm_nt: R=%d must be a multiple of 32", R);
  PASE_CHECK_ARG((K % 4) == 0 && (ldb % 4) == 0 && ldb >= K, "pase_tc_gemm_nt: K/ldb alignment");
  PASE_CHECK_ARG(aligned16(Ahi) && aligned16(Bhi), "pase_tc_gemm_nt: operand alignment");
  PASE_CHECK_ARG(rows_in > 0 && rows_out > 0 && fold > 0 && (N % fold) == 0,
                 "pase_tc_gemm_nt: bad row map");
  PASE_CHECK_ARG((colsum == nullptr) == (colsumsq == nullptr), "pase_tc_gemm_nt: stats pair");
  PASE_CHECK_ARG(colsum == nullptr || N <= 2048, "pase_tc_gemm_nt: stats need N <= 2048");
  PASE_CHECK_ARG(colsum == nullptr || !accumulate, "pase_tc_gemm_nt: stats + accumulate");
  // 256-wide tiles (16-element k-blocks, see NTCfg) raise the MMA rate per flop by ~1.2x but
  // halve the tile count (wave quantisation) and double the epilogue per warp: measured on
  // the PASE+ shapes they only pay off for long reductions (profiles/r01_history.md)
  const bool wide = N >= 256 && (N % 256) == 0 && K >= 4096;
  const int BN = N <= 64 ? 64 : (wide ? 256 : 128);
  const int BKh = BN == 256 ? 16 : 32;
  // 3xTF32: fold the TMEM accumulator into fp32 register sums every K = 128
  const int flush_kb = (mode == 1) ? 128 / BKh : 0;
  CUtensorMap ah, al, bh, bl;
  // window kernel: A fetched once per 32-float column block (needs K % 32 == 0); the plain
  // per-k-block kernel remains for ragged K
  const bool pair2 = BN == 128 && (N % 128) == 0 && pase_tc_use_2cta();
  const bool window = !pair2 && BN != 256 && (K % 32) == 0 && pase_tc_use_window();
  const int qmax = (K + R - 1) / R - 1;
  const int win_rows = ((BM + qmax + 7) / 8) * 8;
  uint64_t adims[2] = {(uint64_t)R, (uint64_t)a_rows};
  uint64_t astr[1] = {(uint64_t)R * 4};
  uint32_t abox[2] = {(uint32_t)BKh, (uint32_t)(window ? win_rows : BM)};
  uint64_t bdims[2] = {(uint64_t)K, (uint64_t)N};
  uint64_t bstr[1] = {(uint64_t)ldb * 4};
  uint32_t bbox[2] = {(uint32_t)BKh, (uint32_t)(pair2 ? BN / 2 : BN)};   // pair: half per CTA
  const CUtensorMapSwizzle swz = BKh == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  PASE_CHECK_ARG(!window || win_rows <= 256, "pase_tc_gemm_nt: K/R=%d too large for a window",
                 qmax + 1);
  int rc;
  if ((rc = make_map(&ah, Ahi, 2, adims, astr, abox, "A.hi", swz)) != 0) return rc;
  if ((rc = make_map(&bh, Bhi, 2, bdims, bstr, bbox, "B.hi", swz)) != 0) return rc;
  if (mode == 1) {
    if ((rc = make_map(&al, Alo, 2, adims, astr, abox, "A.lo", swz)) != 0) return rc;
    if ((rc = make_map(&bl, Blo, 2, bdims, bstr, bbox, "B.lo", swz)) != 0) return rc;
  } else {
    al = ah;
    bl = bh;
  }
  RowMap rm{rows_in, t_valid, rows_out, fold, N / fold};
  cudaStream_t st = (cudaStream_t)stream;
  if (pair2) {
    return mode == 1 ? launch_nt2<128, true>(ah, al, bh, bl, R, C, ldc, M, N, K, alpha, bias, rm,
                                             colsum, colsumsq, accumulate, flush_kb, st)
                     : launch_nt2<128, false>(ah, al, bh, bl, R, C, ldc, M, N, K, alpha, bias, rm,
                                              colsum, colsumsq, accumulate, flush_kb, st);
  }
  if (window) {
#define PASE_NTW(BNV)                                                                          \
  (mode == 1 ? launch_ntw<BNV, true>(ah, al, bh, bl, R, C, ldc, M, N, K, alpha, bias, rm,       \
                                     colsum, colsumsq, accumulate, flush_kb, win_rows, st)      \
             : launch_ntw<BNV, false>(ah, al, bh, bl, R, C, ldc, M, N, K, alpha, bias, rm,      \
                                      colsum, colsumsq, accumulate, flush_kb, win_rows, st))
    if (BN == 64) return PASE_NTW(64);
    return PASE_NTW(128);
#undef PASE_NTW
  }
#define PASE_NT(BNV, BKV)                                                                       \
  (mode == 1 ? launch_nt<BNV, true, BKV>(ah, al, bh, bl, R, C, ldc, M, N, K, alpha, bias, rm,   \
                                         colsum, colsumsq, accumulate, flush_kb, st)            \
             : launch_nt<BNV, false, BKV>(ah, al, bh, bl, R, C, ldc, M, N, K, alpha, bias, rm,  \
                                          colsum, colsumsq, accumulate, flush_kb, st))
  if (BN == 64) return PASE_NT(64, 32);
  if (BN == 128) return PASE_NT(128, 32);
  return PASE_NT(256, 16);
#undef PASE_NT
}

// A: groups x [rows_per_group x lda] starting `offA` rows into each group of pitch `pitchA`
// rows (columns 0..I-1 used); B: folded rows of R floats, group pitch `pitchB` rows.
int pase_tc_gemm_tn(const float* Ahi, const float* Alo, long lda, int pitchA, int offA,
                    const float* Bhi, const float* Blo, int R, int pitchB, long b_rows_total,
                    float* C, long ldc, int I, int J, int groups, int rows_per_group, float alpha,
                    int accumulate, int mode, void* stream) {
  PASE_CHECK_ARG(Ahi && Bhi && C && I > 0 && J > 0 && groups > 0 && rows_per_group > 0,
                 "pase_tc_gemm_tn: bad args");
  pase_tc_init_epi_sleep();
  PASE_CHECK_ARG(mode == 0 || (Alo && Blo), "pase_tc_gemm_tn: mode 1 needs lo operands");
  PASE_CHECK_ARG(R >= 32 && (R % 32) == 0 && (lda % 4) == 0 && (I % 4) == 0 && (J % 32) == 0,
                 "pase_tc_gemm_tn: need R%%32==0, lda%%4==0, I%%4==0, J%%32==0 (R=%d lda=%ld I=%d "
                 "J=%d)", R, lda, I, J);
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    cudaError_t e =
        cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)J * sizeof(float), (size_t)I, st);
    if (e != cudaSuccess) {
      pase_set_error("pase_tc_gemm_tn: memset failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  const int BN = J <= 64 ? 64 : 128;
  const int flush_ch = (mode == 1) ? 4 : 0;      // fold TMEM into fp32 sums every 128 rows
  CUtensorMap ah, al, bh, bl;
  // A as (32 floats, rows-in-group, column blocks of 32, groups): the block dimension has the
  // smallest stride after the inner one, so one box brings BM/32 MN-blocks of [32 rows x 128 B]
  uint64_t adims[4] = {32, (uint64_t)rows_per_group, (uint64_t)((I + 31) / 32), (uint64_t)groups};
  uint64_t astr[3] = {(uint64_t)lda * 4, 128, (uint64_t)pitchA * lda * 4};
  uint32_t abox[4] = {32, TN_KR, (uint32_t)(BM / 32), 1};
  // B group g covers folded rows [g*pitchB, ...): rows beyond the allocation are zero-filled
  long rows_in_group = b_rows_total - (long)(groups - 1) * pitchB;
  if (rows_in_group > pitchB + (J + R - 1) / R + 1) rows_in_group = pitchB + (J + R - 1) / R + 1;
  const int b_blocked = (R % BN) == 0;
  uint64_t bdims[4] = {32, (uint64_t)rows_in_group, (uint64_t)(R / 32), (uint64_t)groups};
  uint64_t bstr[3] = {(uint64_t)R * 4, 128, (uint64_t)pitchB * R * 4};
  uint32_t bbox[4] = {32, TN_KR, (uint32_t)(b_blocked ? BN / 32 : 1), 1};
  // NOTE: when I % 32 != 0 the last A column block reads up to 31 floats past a row's I
  // columns (they only feed output rows >= I, which are never stored); the caller must
  // keep 32 floats of slack after the last row of A.
  int rc;
  const CUtensorMapSwizzle sw32 = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  if ((rc = make_map(&ah, Ahi + (long)offA * lda, 4, adims, astr, abox, "tn A.hi", sw32)) != 0)
    return rc;
  if ((rc = make_map(&bh, Bhi, 4, bdims, bstr, bbox, "tn B.hi", sw32)) != 0) return rc;
  if (mode == 1) {
    if ((rc = make_map(&al, Alo + (long)offA * lda, 4, adims, astr, abox, "tn A.lo", sw32)) != 0)
      return rc;
    if ((rc = make_map(&bl, Blo, 4, bdims, bstr, bbox, "tn B.lo", sw32)) != 0) return rc;
  } else {
    al = ah;
    bl = bh;
  }
#define PASE_TN(BNV)                                                                            \
  (mode == 1 ? launch_tn<BNV, true>(ah, al, bh, bl, R, C, ldc, I, J, groups, rows_per_group,    \
                                    alpha, flush_ch, b_blocked, st)                             \
             : launch_tn<BNV, false>(ah, al, bh, bl, R, C, ldc, I, J, groups, rows_per_group,   \
                                     alpha, flush_ch, b_blocked, st))
  if (BN == 64) return PASE_TN(64);
  return PASE_TN(128);
#undef PASE_TN
}

}  // extern "C"
