// cloud.cu — K10 DisparityToCloud<double> (CStereoMatching.cpp:682-761): ellipse-eroded mask,
// per-pixel reprojection with Q, rigid transform R_final*F + T_final, and emission of the points
// in row-major order over view 0's margin rectangle (quirk Q11).
//
// cv::erode with the MORPH_ELLIPSE element of size ceil(0.02*rows) (:703-705) only matters through
// the test `mask != 255` (:740), i.e. "every mask byte under the element is 255".  With
// run[y][x] = length of the run of 255s ending at x, row i of the element [j1,j2) is all-255 iff
// run[y+i-a][xb] >= xb-xa+1, so a pixel costs one lookup per element row instead of ks^2 bytes.
// Emission order is kept by a per-row count, an exclusive scan over rows, and an ordered ballot
// compaction inside each row.
#include <math.h>
#include "kernels.h"


// run length of mask == 255 ending at x (inclusive); one warp per row
__global__ void __launch_bounds__(128) k_mask_runs(const uint8_t* __restrict__ mask, int W, int H, unsigned short* __restrict__ run) {
  const int lane = threadIdx.x & 31;
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (y >= H) return;
  const uint8_t* p = mask + (size_t)y * W;
  int last_bad = -1;
  for (int base = 0; base < W; base += 32) {
    const int x = base + lane;
    int v = (x < W && p[x] != 255) ? x : -1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v = max(v, t);
    }
    v = max(v, last_bad);
    if (x < W) run[(size_t)y * W + x] = (unsigned short)min(x - v, 65535);
    last_bad = __shfl_sync(0xffffffffu, v, 31);
  }
}

// eroded-and-matched flag per pixel of the margin rectangle + per-row count; one block per row
__global__ void __launch_bounds__(256) k_cloud_flags(const double* __restrict__ disp, const unsigned short* __restrict__ run,
                                                     int W, int H, Bound m, int ks, const short* __restrict__ j12,
                                                     uint8_t* __restrict__ flag, int* __restrict__ row_count) {
  const int y = m.YL + blockIdx.x;
  const int a = ks / 2;
  int cnt = 0;
  for (int x = m.XL + threadIdx.x; x <= m.XR; x += blockDim.x) {
    bool ok = disp[(size_t)y * W + x] != (double)SB_NOMATCH;
    // element rows in groups of eight: the eight run-length lookups of a group are independent loads (a row-by-row early
    // exit chains up to ks dependent memory round trips per pixel); the AND is the same
    for (int i0 = 0; i0 < ks && ok; i0 += 8) {
      bool all = true;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int i = i0 + u;
        if (i >= ks) break;
        const int sy = y + i - a;
        const int j1 = j12[i], j2 = j12[ks + i];
        if (sy < 0 || sy >= H || j2 <= j1) continue;  // rows outside the image do not constrain
        const int xa = max(x + j1 - a, 0), xb = min(x + j2 - 1 - a, W - 1);
        if (xb < xa) continue;
        all &= (int)run[(size_t)sy * W + xb] >= xb - xa + 1;
      }
      ok = all;
    }
    flag[(size_t)y * W + x] = ok;
    cnt += ok;
  }
  __shared__ int s[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 8; k++) t += s[k];
    row_count[blockIdx.x] = t;
  }
}

// exclusive scan of the row counts (n <= a few thousand): one block
__global__ void __launch_bounds__(1024) k_row_scan(const int* __restrict__ row_count, int n, int* __restrict__ row_offset,
                                                   int* __restrict__ total) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int c = i < n ? row_count[i] : 0;
    int v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < warp; k++) woff += s_warp[k];
    const int carry = s_carry;
    if (i < n) row_offset[i] = carry + woff + v - c;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + woff + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) { row_offset[n] = s_carry; *total = s_carry; }
}

// reprojection (:745-749) + ordered emission; one block per row
__global__ void __launch_bounds__(256) k_cloud_emit(const double* __restrict__ disp, const uint8_t* __restrict__ img,
                                                    const uint8_t* __restrict__ flag, int W, Bound m, CloudParams p,
                                                    const int* __restrict__ row_offset, double* __restrict__ xyz,
                                                    uint8_t* __restrict__ bgr, int* __restrict__ pix) {
  const int y = m.YL + blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int s_cnt[8];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = row_offset[blockIdx.x];
  __syncthreads();
  const double qy = (double)y + p.q13;
  for (int base = m.XL; base <= m.XR; base += 256) {
    const int x = base + threadIdx.x;
    const size_t f = (size_t)y * W + x;
    const bool ok = x <= m.XR && flag[f];
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int k = 0; k < warp; k++) off += s_cnt[k];
    if (ok) {
      const size_t o = (size_t)off + __popc(bal & ((1u << lane) - 1));
      const double d = disp[f];
      const double iW = 1. / (p.q33 + p.q32 * d);
      const double F0 = (p.q03 + (double)x) * iW, F1 = qy * iW, F2 = p.q23 * iW;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        double s = 0;
        s += p.R[i * 3 + 0] * F0;
        s += p.R[i * 3 + 1] * F1;
        s += p.R[i * 3 + 2] * F2;
        xyz[o * 3 + i] = s + p.T[i];
      }
      bgr[o * 3 + 0] = img[f * 3 + 0];
      bgr[o * 3 + 1] = img[f * 3 + 1];
      bgr[o * 3 + 2] = img[f * 3 + 2];
      pix[o] = (int)f;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < 8; k++) t += s_cnt[k]; s_base += t; }
    __syncthreads();
  }
}

// getStructuringElement(MORPH_ELLIPSE, ks x ks) row extents [j1, j2) — OpenCV's published definition
void sb_ellipse_rows(int ks, short* j1, short* j2) {
  const int r = ks / 2, c = ks / 2;
  const double inv_r2 = r ? 1. / ((double)r * r) : 0;
  for (int i = 0; i < ks; i++) {
    j1[i] = 0; j2[i] = 0;
    if (ks == 1) { j2[i] = 1; continue; }
    const int dy = i - r;
    if (abs(dy) <= r) {
      const int dx = (int)nearbyint(c * sqrt((r * r - dy * dy) * inv_r2));
      j1[i] = (short)sb_imax(c - dx, 0);
      j2[i] = (short)sb_imin(c + dx + 1, ks);
    }
  }
}

int launch_cloud(const double* disp, const uint8_t* mask, const uint8_t* img, int W, int H, Bound m, int erode_ks,
                 const CloudParams& p, const CloudScratch& s, double* xyz, uint8_t* bgr, int* pix, int* n_points_dev,
                 cudaStream_t st) {
  if (m.width <= 0 || m.height <= 0) { cudaMemsetAsync(n_points_dev, 0, sizeof(int), st); return 0; }
  k_mask_runs<<<(H + 3) / 4, 128, 0, st>>>(mask, W, H, s.run);
  k_cloud_flags<<<m.height, 256, 0, st>>>(disp, s.run, W, H, m, erode_ks, s.ellipse, s.eroded, s.row_count);
  k_row_scan<<<1, 1024, 0, st>>>(s.row_count, m.height, s.row_offset, n_points_dev);
  k_cloud_emit<<<m.height, 256, 0, st>>>(disp, img, s.eroded, W, m, p, s.row_offset, xyz, bgr, pix);
  return 4;
}
