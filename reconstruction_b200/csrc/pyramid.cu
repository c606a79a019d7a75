// pyramid.cu — K0 pyrDown, K1 FindMargin, and the per-pixel WindowToVec statistics.
#include "kernels.h"
#include "ncc_exact.cuh"

// ------------------------------------------------------------------------------------------------
// K0  cv::pyrDown as called by ConstructPyrm (CStereoMatching.cpp:1049-1050): 5x5 [1 4 6 4 1]^2,
// BORDER_REFLECT_101, dst = (src+1)/2, rounding (sum + 128) >> 8.  OpenCV is a third-party
// dependency of the reference; this follows its published definition (pinned against cv2 vectors
// in tests/golden/pyrdown_cv2.npz through the oracle, and against the oracle on the GPU).
// One thread per output byte; rows of the source are re-read through L1.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}

__global__ void __launch_bounds__(256) k_pyrdown(const uint8_t* __restrict__ src, int W, int H, int cn,
                                                 uint8_t* __restrict__ dst, int w, int h) {
  const int xb = blockIdx.x * blockDim.x + threadIdx.x;  // byte column in the destination row
  const int y = blockIdx.y;
  if (xb >= w * cn) return;
  const int x = xb / cn, c = xb - x * cn;
  int cx[5];
#pragma unroll
  for (int k = 0; k < 5; k++) cx[k] = reflect101(2 * x - 2 + k, W) * cn + c;
  const int wt[5] = {1, 4, 6, 4, 1};
  int acc = 0;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    const uint8_t* s = src + (size_t)reflect101(2 * y - 2 + k, H) * W * cn;
    const int hsum = s[cx[0]] + 4 * s[cx[1]] + 6 * s[cx[2]] + 4 * s[cx[3]] + s[cx[4]];
    acc += wt[k] * hsum;
  }
  dst[(size_t)y * w * cn + xb] = (uint8_t)((acc + 128) >> 8);
}

int launch_pyrdown(const uint8_t* src, int W, int H, int cn, uint8_t* dst, cudaStream_t st) {
  const int w = (W + 1) / 2, h = (H + 1) / 2;
  dim3 grid((w * cn + 255) / 256, h);
  k_pyrdown<<<grid, 256, 0, st>>>(src, W, H, cn, dst, w, h);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// K1  FindMargin (CStereoMatching.cpp:1011-1038): bounding box of mask == 255 inside the R-pixel
// border, initialised inverted (:1014-1017).  One block per row, warp-shuffle min/max, four atomics
// per non-empty row.
// ------------------------------------------------------------------------------------------------
__global__ void k_margin_init(int* out4, int W, int H, int R) {
  out4[0] = H - 1 - R; out4[1] = R; out4[2] = W - 1 - R; out4[3] = R;
}

__global__ void __launch_bounds__(128) k_find_margin(const uint8_t* __restrict__ mask, int W, int H, int R, int* out4) {
  const int y = R + blockIdx.x;
  if (y >= H - R) return;
  const uint8_t* p = mask + (size_t)y * W;
  int lo = 1 << 30, hi = -1;
  for (int x = R + threadIdx.x; x < W - R; x += blockDim.x)
    if (p[x] == 255) { lo = min(lo, x); hi = max(hi, x); }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ int slo[4], shi[4];
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 4; k++) { lo = min(lo, slo[k]); hi = max(hi, shi[k]); }
    if (hi >= 0) {
      atomicMin(out4 + 2, lo); atomicMax(out4 + 3, hi);
      atomicMin(out4 + 0, y);  atomicMax(out4 + 1, y);
    }
  }
}

int launch_find_margin(const uint8_t* mask, int W, int H, int R, int* out4, cudaStream_t st) {
  k_margin_init<<<1, 1, 0, st>>>(out4, W, H, R);
  if (H - 2 * R > 0) k_find_margin<<<H - 2 * R, 128, 0, st>>>(mask, W, H, R, out4);
  return 2;
}

// ------------------------------------------------------------------------------------------------
// WindowToVec statistics (CManageData.cpp:81-90): for every flat pixel index f whose (2R+1)^2
// window lies inside the payload, (mean, norm) of the zero-mean window vector, in the reference's
// summation order.  Both matching directions and the Rematch search reuse these per level, so
// the per-candidate work of the searches is the dot product only.
// ------------------------------------------------------------------------------------------------
template <int WS>
__global__ void __launch_bounds__(256) k_window_stats(const uint8_t* __restrict__ img, int W, long n_px, double2* __restrict__ stats,
                                                      int2* __restrict__ istats) {
  constexpr int R = WS / 2;
  const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_px) return;
  const long first = (long)R * W + R, last = n_px - first;  // f in [first, last): all bytes inside the payload
  if (f < first || f >= last) {
    stats[f] = make_double2(0.0, 1.0);
    if (istats) istats[f] = make_int2(0, 0);
    return;
  }
  const uint8_t* p0 = img + 3 * (f - first);
  double mean;
  const double nrm = window_stats_exact<WS>(p0, 3 * W, mean);
  stats[f] = make_double2(mean, nrm);
  if (istats) {  // exact integer sums of the same window, for the screening pass of the searches (match.cu)
    int S = 0, SS = 0;
#pragma unroll
    for (int i = 0; i < WS; i++)
#pragma unroll
      for (int j = 0; j < 3 * WS; j++) {
        const int b = p0[i * 3 * W + j];
        S += b;
        SS += b * b;
      }
    istats[f] = make_int2(S, SS);
  }
}

int launch_window_stats(const uint8_t* img, int W, int H, int R, double2* stats, int2* istats, cudaStream_t st) {
  const long n = (long)W * H;
  const int grid = (int)((n + 255) / 256);
  if (R == 2) k_window_stats<5><<<grid, 256, 0, st>>>(img, W, n, stats, istats);
  else if (R == 1) k_window_stats<3><<<grid, 256, 0, st>>>(img, W, n, stats, istats);
  else if (R == 3) k_window_stats<7><<<grid, 256, 0, st>>>(img, W, n, stats, istats);
  else return -1;
  return 1;
}

// integer-only statistics map (levels where the searches are screened: the double map is not needed)
__global__ void __launch_bounds__(256) k_window_istats5(const uint8_t* __restrict__ img, int W, long n_px, int2* __restrict__ istats) {
  const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_px) return;
  const long first = 2L * W + 2, last = n_px - first;
  if (f < first || f >= last) { istats[f] = make_int2(0, 0); return; }
  // 5 rows x 15 bytes as byte-packed words: sum = dp4a(w, 1111), sum of squares = dp4a(w, w)
  unsigned S = 0, SS = 0;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    unsigned w[4];
    load_row_words<4>(img, 3 * (f - first) + (long)i * 3 * W, w);
    w[3] &= 0x00ffffffu;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      S = __dp4a(w[q], 0x01010101u, S);
      SS = __dp4a(w[q], w[q], SS);
    }
  }
  istats[f] = make_int2((int)S, (int)SS);
}
int launch_window_istats(const uint8_t* img, int W, int H, int2* istats, cudaStream_t st) {
  const long n = (long)W * H;
  k_window_istats5<<<(int)((n + 255) / 256), 256, 0, st>>>(img, W, n, istats);
  return 1;
}

// 16-byte stores over the aligned body, scalar stores at the two ends (the maps come from cudaMalloc: the head is empty)
__global__ void __launch_bounds__(256) k_fill_s16(short* p, long n, short v) {
  const unsigned vv = (unsigned)(unsigned short)v * 0x00010001u;
  long head = (long)(((16 - ((size_t)p & 15)) & 15) / 2);
  if (head > n) head = n;
  const long body = (n - head) / 8;
  uint4* q = reinterpret_cast<uint4*>(p + head);
  const uint4 w = make_uint4(vv, vv, vv, vv);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < body; i += (long)gridDim.x * blockDim.x) q[i] = w;
  if (blockIdx.x == 0) {
    if ((long)threadIdx.x < head) p[threadIdx.x] = v;
    const long t = head + body * 8 + threadIdx.x;
    if (t < n) p[t] = v;  // at most 7 trailing elements
  }
}
int launch_fill_s16(short* p, long n, short v, cudaStream_t st) {
  if (n <= 0) return 0;
  const long body = n / 8;
  int blocks = (int)((body + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_fill_s16<<<blocks, 256, 0, st>>>(p, n, v);
  return 1;
}
