// Does an FP64 warp-instruction hold a sub-partition's dispatch port for both of its pipe cycles?  (the fused
// refinement kernel issues 61 FP64 + ~70 other instructions per pixel-sweep; if the answer is yes its sub-partitions are
// ~86 % occupied and the instruction count is the only lever left, if no they are ~58 % occupied.)
// Each thread runs NF independent DFMA chains and NI independent integer (IMAD) chains per iteration, 8 warps per sub-partition:
//   cycles per iteration ~ max(2 NF, NF + NI)   if the two streams only share the issue slot,
//   cycles per iteration ~ 2 NF + NI            if an FP64 instruction blocks dispatch for its second cycle.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue_mix fp64_issue_mix.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI>
__global__ void k(double* out, double a, double b, int m, int iters) {
  double x[NF > 0 ? NF : 1];
  int y[NI > 0 ? NI : 1];
#pragma unroll
  for (int i = 0; i < NF; i++) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < NI; i++) y[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < (NF > NI ? NF : NI); i++) {
      if (i < NF) x[i] = __fma_rn(x[i], a, b);
      if (i < NI) y[i] = y[i] * m + it;  // IMAD, data-dependent on the previous value: not foldable
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NF; i++) s += x[i];
#pragma unroll
  for (int i = 0; i < NI; i++) s += y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, int NI>
void run(double* out) {
  const int iters = 1 << 13, warps_per_smsp = 8, threads = warps_per_smsp * 4 * 32;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NF, NI><<<148, threads>>>(out, 1.0000001, 1e-9, 3, 64);
  cudaEventRecord(e0);
  k<NF, NI><<<148, threads>>>(out, 1.0000001, 1e-9, 3, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9 / iters / warps_per_smsp;  // cycles per warp-iteration per sub-partition
  printf("NF %d NI %2d: %.2f cycles per warp-iteration   [shared issue slot only: %d, dispatch blocked: %d]\n", NF, NI, cyc,
         2 * NF > NF + NI ? 2 * NF : NF + NI, 2 * NF + NI);
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  run<4, 0>(out); run<0, 8>(out);
  run<4, 2>(out); run<4, 4>(out); run<4, 6>(out); run<4, 8>(out); run<4, 12>(out);
  run<6, 7>(out);  // the refinement kernel's ratio (61 : 70)
  return 0;
}
