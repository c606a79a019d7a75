// Probe of the TMA tile load used by k_refine_fused: 2-D f64 map, box 64x48, zero fill, tensor map as __grid_constant__.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

struct alignas(64) Maps { CUtensorMap m[2]; };

__global__ void k(const __grid_constant__ Maps tm, int z, int gx0, int gy0, double* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long bar;
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem), mb = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(64 * 48 * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s0),
                 "l"(&tm.m[z]), "r"(gx0), "r"(gy0), "r"(mb)
                 : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(mb), "r"(0) : "memory");
  const double* t = reinterpret_cast<const double*>(smem);
  for (int i = threadIdx.x; i < 64 * 48; i += blockDim.x) out[i] = t[i];
}

#define BODY(MAPPTR) \
  extern __shared__ __align__(128) unsigned char smem[]; \
  __shared__ unsigned long long bar; \
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem), mb = (unsigned)__cvta_generic_to_shared(&bar); \
  if (threadIdx.x == 0) { \
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1)); \
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); \
  } \
  __syncthreads(); \
  if (threadIdx.x == 0) { \
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(64 * 48 * 8) : "memory"); \
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s0), \
                 "l"(MAPPTR), "r"(gx0), "r"(gy0), "r"(mb) : "memory"); \
  } \
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(mb), "r"(0) : "memory"); \
  const double* t = reinterpret_cast<const double*>(smem); \
  for (int i = threadIdx.x; i < 64 * 48; i += blockDim.x) out[i] = t[i];

__global__ void kA(const __grid_constant__ CUtensorMap tm, int gx0, int gy0, double* out) { BODY(&tm) }
__global__ void kG(const CUtensorMap* tm, int gx0, int gy0, double* out) { BODY(tm) }

int main(int argc, char** argv) {
  const int W = 256, H = 192;
  std::vector<double> h(W * H);
  for (int i = 0; i < W * H; i++) h[i] = i;
  double *d, *o;
  cudaMalloc(&d, W * H * 8); cudaMalloc(&o, 64 * 48 * 8);
  cudaMemcpy(d, h.data(), W * H * 8, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: %d %d %p\n", (int)e, (int)q, p);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Maps tm; memset(&tm, 0, sizeof tm);
  cuuint64_t dims[2] = {W, H}, strides[1] = {W * 8};
  cuuint32_t box[2] = {64, 48}, es[2] = {1, 1};
  for (int z = 0; z < 2; z++) {
    CUresult r = ((Fn)p)(&tm.m[z], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d -> %d\n", z, (int)r);
  }
  const int gx0 = argc > 1 ? atoi(argv[1]) : 0, gy0 = argc > 2 ? atoi(argv[2]) : 0;
  kA<<<1, 256, 64 * 48 * 8>>>(tm.m[0], gx0, gy0, o);
  printf("kA (%d,%d): %s\n", gx0, gy0, cudaGetErrorString(cudaDeviceSynchronize()));
  std::vector<double> r(64 * 48);
  cudaMemcpy(r.data(), o, r.size() * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < 48; y++) for (int x = 0; x < 64; x++) {
    const int gx = x + gx0, gy = y + gy0;
    const double exp = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[gy * W + gx] : 0.0;
    bad += r[y * 64 + x] != exp;
  }
  printf("mismatches %d\n", bad);
  return 0;
}
