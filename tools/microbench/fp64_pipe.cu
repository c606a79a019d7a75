// FP64 pipe latency / throughput on the B200: dependent DFMA chains, ILP chains per thread, W warps per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipe fp64_pipe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, double a, double b, int iters) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = __fma_rn(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
void run(int warps_per_smsp, double* out) {
  const int iters = 1 << 14;
  const int threads = warps_per_smsp * 4 * 32;  // one block per SM
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<ILP><<<148, threads>>>(out, 1.0000001, 1e-9, 64);
  cudaEventRecord(e0);
  k<ILP><<<148, threads>>>(out, 1.0000001, 1e-9, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cycles = ms * 1e-3 * 1.965e9;
  const double inst_per_smsp = (double)iters * ILP * warps_per_smsp;
  printf("ILP %d warps/SMSP %2d: %.3f ms, %.2f cycles per DFMA warp-instr per SMSP, chain latency <= %.1f cycles\n", ILP, warps_per_smsp, ms,
         cycles / inst_per_smsp, cycles / iters / 1.0);
}

int main() {
  double* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  for (int w : {1, 2, 4, 8}) run<1>(w, out);
  for (int w : {1, 2, 4, 8}) run<2>(w, out);
  for (int w : {1, 2, 4, 8}) run<4>(w, out);
  for (int w : {1, 4, 8}) run<8>(w, out);
  return 0;
}
