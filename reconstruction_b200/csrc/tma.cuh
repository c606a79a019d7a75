// tma.cuh — cp.async.bulk.tensor (TMA) tile loads completing on an mbarrier, and the host-side tensor-map encoder.
//
// A rectangular tile of a row-major map is fetched by ONE elected thread; the copy engine writes shared memory and
// signals the mbarrier with the byte count, elements outside the map arrive as zeros (no bounds tests in the kernel).
// Measured constraint on this B200 (tools/microbench/tma_probe.cu): the global address of the box's first element
// must be a multiple of 16 bytes (coordinates may be negative or past the end), the shared destination 128 B aligned,
// the row pitch a multiple of 16 bytes, box dimensions <= 256 elements.
#pragma once
#include <cuda.h>  // CUtensorMap (driver types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stddef.h>

#ifdef __CUDACC__
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  unsigned ok = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(unsigned smem_dst, const CUtensorMap* map, int x, int y, unsigned mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_dst),
               "l"(map), "r"(x), "r"(y), "r"(mbar)
               : "memory");
}
#endif

// Tensor maps: 2-D, row-major, no swizzle, zero fill outside the map.  false when the driver entry point is missing or
// the map violates a TMA constraint (the callers then take their plain-load path).
typedef CUresult (*SbEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline SbEncodeTiledFn sb_tma_encoder() {
  static SbEncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<SbEncodeTiledFn>(p);
    (void)cudaGetLastError();
    tried = true;
  }
  return fn;
}
// W, H, bw, bh in elements of `elem` bytes; pitch_bytes = distance between rows
static inline bool sb_tma_encode_2d(CUtensorMap* m, CUtensorMapDataType dt, int elem, const void* base, long W, long H, size_t pitch_bytes,
                                    int bw, int bh) {
  SbEncodeTiledFn enc = sb_tma_encoder();
  if (!enc || pitch_bytes % 16 != 0 || ((size_t)bw * elem) % 16 != 0 || bw > 256 || bh > 256 || ((size_t)base & 15) != 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
