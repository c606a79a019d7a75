// ncc_exact.cuh — the NCC primitive in the reference's exact arithmetic.
//
// CManageData::WindowToVec (CManageData.cpp:81-90) gathers a ws x ws x 3 window byte-column-major
// (k = j*ws + i, j = byte column, i = row), subtracts the mean and returns the L2 norm (0 -> 1).
// Armadillo 4.200 evaluates mean / norm / dot with two accumulators, even indices into the first,
// odd into the second, added at the end (arrayops_meat.hpp:902-921, fn_norm.hpp:108-127,
// op_dot_meat.hpp:36-55).  The sums of the raw bytes are exact integers, so the mean is one
// correctly rounded division; everything after it is rounded per operation and is reproduced
// here operation by operation (translation units are built with -fmad=false).
#pragma once
#include "common.cuh"

// (mean, norm) of the window whose first byte is p0 (row 0, byte column 3*x0); rows `pitch` apart.
template <int WS>
__device__ __forceinline__ double window_stats_exact(const uint8_t* __restrict__ p0, int pitch, double& mean) {
  constexpr int N = WS * WS * 3;
  int S = 0;
#pragma unroll
  for (int i = 0; i < WS; i++)
#pragma unroll
    for (int j = 0; j < 3 * WS; j++) S += p0[i * pitch + j];
  mean = (double)S / (double)N;
  double a1 = 0, a2 = 0;
#pragma unroll
  for (int j = 0; j < 3 * WS; j++)
#pragma unroll
    for (int i = 0; i < WS; i++) {
      const double u = (double)p0[i * pitch + j] - mean;
      const double uu = u * u;
      if ((j * WS + i) & 1) a2 += uu; else a1 += uu;
    }
  const double n = sqrt(a1 + a2);
  return n == 0 ? 1.0 : n;
}

// Same with every byte fetched through a bounds check against [0, size): bytes the reference's flat
// addressing would read outside the buffer are taken as 0 (the oracle's slack is zero-filled).
template <int WS>
__device__ __noinline__ double window_stats_checked(const uint8_t* __restrict__ base, long off0, long size, int pitch, double& mean) {
  constexpr int N = WS * WS * 3;
  int S = 0;
  for (int i = 0; i < WS; i++)
    for (int j = 0; j < 3 * WS; j++) {
      const long o = off0 + (long)i * pitch + j;
      S += (o >= 0 && o < size) ? base[o] : 0;
    }
  mean = (double)S / (double)N;
  double a1 = 0, a2 = 0;
  for (int j = 0; j < 3 * WS; j++)
    for (int i = 0; i < WS; i++) {
      const long o = off0 + (long)i * pitch + j;
      const double u = (double)((o >= 0 && o < size) ? base[o] : 0) - mean;
      const double uu = u * u;
      if ((j * WS + i) & 1) a2 += uu; else a1 += uu;
    }
  const double n = sqrt(a1 + a2);
  return n == 0 ? 1.0 : n;
}

// dot(vecL, vecR) with vecR[k] = byte[k] - meanR (not normalised), vecL read from shared memory
// (element k at vecL[k * strideL]).
template <int WS>
__device__ __forceinline__ double window_dot_exact(const double* __restrict__ vecL, int strideL, const uint8_t* __restrict__ p0, int pitch, double meanR) {
  double v1 = 0, v2 = 0;
#pragma unroll
  for (int j = 0; j < 3 * WS; j++)
#pragma unroll
    for (int i = 0; i < WS; i++) {
      const int k = j * WS + i;
      const double u = (double)p0[i * pitch + j] - meanR;
      const double pr = vecL[k * strideL] * u;
      if (k & 1) v2 += pr; else v1 += pr;
    }
  return v1 + v2;
}

template <int WS>
__device__ __noinline__ double window_dot_checked(const double* __restrict__ vecL, int strideL, const uint8_t* __restrict__ base, long off0, long size, int pitch, double meanR) {
  double v1 = 0, v2 = 0;
  for (int j = 0; j < 3 * WS; j++)
    for (int i = 0; i < WS; i++) {
      const int k = j * WS + i;
      const long o = off0 + (long)i * pitch + j;
      const double u = (double)((o >= 0 && o < size) ? base[o] : 0) - meanR;
      const double pr = vecL[k * strideL] * u;
      if (k & 1) v2 += pr; else v1 += pr;
    }
  return v1 + v2;
}

// Correctly rounded a / b for many a with one b: y = RN(1/b) once, then two FMA residual
// corrections (Markstein): q1 = RN(q0 + (a - q0 b) y) is already RN(a/b) for faithful q0; the second
// step is belt and braces.  Checked against hardware division on 1.8e9 operand pairs drawn from
// the window distribution.  Operands here are normal and far from overflow.
__device__ __forceinline__ double div_by_common(double a, double b, double y) {
  const double q0 = a * y;
  const double r0 = __fma_rn(-q0, b, a);
  const double q1 = __fma_rn(r0, y, q0);
  const double r1 = __fma_rn(-q1, b, a);
  return __fma_rn(r1, y, q1);
}
