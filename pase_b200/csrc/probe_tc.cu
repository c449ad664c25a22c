// Hardware probe (tools/rowshift_probe.py): may a K-major SWIZZLE_128B UMMA operand start at
// a row that is not a multiple of 8 (start address not 1024-byte aligned)?  A window of
// `wrows` x 32 fp32 is loaded once by TMA; D = W[shift : shift+128, :] . B^T is computed with
// the descriptor start address advanced by shift*128 bytes and base_offset either 0 or
// (addr >> 7) & 7.  Not used by the product path.
#include "common.cuh"
#include <cuda.h>

namespace {

__device__ __forceinline__ uint32_t s32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__global__ void __launch_bounds__(128, 1)
rowshift_probe_kernel(const __grid_constant__ CUtensorMap mW, const __grid_constant__ CUtensorMap mB,
                      float* __restrict__ D, int shift, int use_base_offset, int wrows) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sW = smem;                       // wrows x 128 B (<= 256 rows = 32 KB)
  uint8_t* sB = smem + 32768;               // 128 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 16384);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(s32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)wrows * 128u + 128u * 128u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(sW)), "l"(reinterpret_cast<uint64_t>(&mW)), "r"(s32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s32(sB)), "l"(reinterpret_cast<uint64_t>(&mB)), "r"(s32(bar)), "r"(0), "r"(0) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = s32(sW) + (uint32_t)shift * 128u + k * 32u;
      const uint32_t b_addr = s32(sB) + k * 32u;
      uint64_t da = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) |
                    ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      if (use_base_offset) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
      const uint64_t db = (uint64_t)((b_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) |
                          ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t acc = k > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(done)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(done)) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cc * 32))
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + cc * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                          const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                          CUtensorMapFloatOOBfill);
}  // namespace

// W: (wrows x 32) fp32, B: (128 x 32) fp32, D: (128 x 128) out = W[shift:shift+128] . B^T
extern "C" int pase_tc_probe_rowshift(const float* W, const float* B, float* D, int wrows,
                                      int shift, int use_base_offset, void* stream) {
  PASE_CHECK_ARG(W && B && D && wrows >= 128 + shift && wrows <= 256 && shift >= 0,
                 "pase_tc_probe_rowshift: bad args");
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
    pase_set_error("probe: no cuTensorMapEncodeTiled");
    return PASE_ERR_UNSUPPORTED;
  }
  EncFn enc = reinterpret_cast<EncFn>(p);
  CUtensorMap mW, mB;
  cuuint64_t dW[2] = {32, (cuuint64_t)wrows}, dB[2] = {32, 128}, st[1] = {128};
  cuuint32_t bW[2] = {32, (cuuint32_t)wrows}, bB[2] = {32, 128}, es[2] = {1, 1};
  if (enc(&mW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)W, dW, st, bW, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      enc(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)B, dB, st, bB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    pase_set_error("probe: tensor map encode failed");
    return PASE_ERR_ARG;
  }
  const int smem = 32768 + 16384 + 64 + 1024;
  cudaFuncSetAttribute(rowshift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rowshift_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mW, mB, D, shift, use_base_offset, wrows);
  PASE_LAUNCH_CHECK("pase_tc_probe_rowshift");
  return PASE_OK;
}
