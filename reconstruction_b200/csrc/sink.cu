// sink.cu — the step after the path: what CCloudOptimization::filter does to the points of one camera pair
// before meshing (CloudOptimization/CCloudOptimization.cpp:64-121): pcl::StatisticalOutlierRemoval (meanK, stddev multiplier),
// pcl::NormalEstimationOMP (radius search) and the orientation of every normal towards the pair's first camera.
//
// PCL is a third-party dependency that is not under /root/reference (and not installed here), so this restates the published
// algorithms; parity is UNPINNED against PCL itself and is checked against oracle/sink_oracle.py (scipy cKDTree + numpy eigh).
//   * points are float32 (pcl::PointXYZ; InsertPoint narrows the f64 triple, CCloudOptimization.cpp:59-62);
//   * squared distances as FLANN's L2_Simple computes them: ((dx*dx + dy*dy) + dz*dz) in float, no FMA;
//   * SOR (pcl/filters/impl/statistical_outlier_removal.hpp): d_i = mean of the sqrt distances to the meanK nearest other
//     points; mean and sample standard deviation of d over all points (sequential double sums, done on the host in index
//     order); keep d_i <= mean + mul * stddev;
//   * normals (pcl/features/normal_3d.h): neighbours with d2 < r^2 among the KEPT points, covariance of the neighbourhood,
//     eigenvector of the smallest eigenvalue, curvature = lambda0 / trace, NaN when fewer than 3 neighbours; flipped towards
//     the default viewpoint (0,0,0) as ne.compute does (the setViewPoint call at :103 comes after compute and has no effect),
//     then towards CamCenter[idx] (:109-116).
//
// Device formulation: points are sorted by the id of their cell in a uniform grid (cell edge = the normal radius; x fastest), so
// the 27 cells around a query are 9 contiguous runs of the sorted array, found by binary search — no dense cell table, no limit
// on the extent.  One warp per query.  The k-th smallest distance is found by an 8-bit x 4 radix select over the candidates'
// float bit patterns (histogram in shared memory), so only sum-of-the-k-smallest is formed, never a neighbour list; ties at the
// k-th value are handled by count.  The search ring grows until the k-th distance is provably inside it (isolated outliers).
#include <cub/cub.cuh>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/stereo_b200.h"

namespace {

struct Grid {
  float minx, miny, minz, inv_c, c;
  int nx, ny, nz;
};

__device__ __forceinline__ int cell_coord(float p, float mn, float inv_c, int n) {
  int v = (int)floorf(__fmul_rn(__fsub_rn(p, mn), inv_c));
  return v < 0 ? 0 : (v >= n ? n - 1 : v);
}
__device__ __forceinline__ unsigned cell_id(const Grid& g, int x, int y, int z) { return ((unsigned)z * g.ny + y) * g.nx + x; }

__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));  // L2_Simple order
}

__device__ __forceinline__ unsigned lower_bound(const unsigned* __restrict__ keys, unsigned n, unsigned v) {
  unsigned lo = 0, hi = n;
  while (lo < hi) {
    const unsigned mid = (lo + hi) >> 1;
    if (keys[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void k_narrow_bbox(const double* __restrict__ xyz, long n, float4* __restrict__ p32, int* __restrict__ bbox /* 6 ordered ints */) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  float x = 0, y = 0, z = 0;
  bool ok = false;
  if (i < n) {
    x = (float)xyz[3 * i]; y = (float)xyz[3 * i + 1]; z = (float)xyz[3 * i + 2];
    ok = isfinite(x) && isfinite(y) && isfinite(z);
    p32[i] = make_float4(x, y, z, ok ? 1.0f : 0.0f);
  }
  auto enc = [](float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; };  // order-preserving
  int v[6] = {ok ? enc(x) : 0x7fffffff, ok ? enc(y) : 0x7fffffff, ok ? enc(z) : 0x7fffffff,
              ok ? enc(x) : (int)0x80000000, ok ? enc(y) : (int)0x80000000, ok ? enc(z) : (int)0x80000000};
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int t = __shfl_xor_sync(0xffffffffu, v[k], o);
      v[k] = k < 3 ? min(v[k], t) : max(v[k], t);
    }
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < 6; k++) { if (k < 3) atomicMin(bbox + k, v[k]); else atomicMax(bbox + k, v[k]); }
}

__global__ void k_cell_keys(const float4* __restrict__ p32, long n, Grid g, unsigned* __restrict__ keys, unsigned* __restrict__ idx) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = p32[i];
  idx[i] = (unsigned)i;
  keys[i] = p.w == 0.0f ? 0xffffffffu  // non-finite points sort last and are never candidates
                        : cell_id(g, cell_coord(p.x, g.minx, g.inv_c, g.nx), cell_coord(p.y, g.miny, g.inv_c, g.ny), cell_coord(p.z, g.minz, g.inv_c, g.nz));
}

__global__ void k_gather(const float4* __restrict__ p32, const unsigned* __restrict__ idx, long n, float4* __restrict__ sp) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float4 p = p32[idx[j]];
  p.w = __int_as_float((int)idx[j]);
  sp[j] = p;
}

// Runs of the sorted array covering the (2R+1)^3 block of cells around (cx,cy,cz): one run per (dy,dz) row, rows dealt to
// lanes for the binary searches.  row_ranges() yields the runs of rows [base, base+32) in (lo, hi) of lane = row - base.
__device__ __forceinline__ void row_ranges(const Grid& g, const unsigned* __restrict__ keys, unsigned n, int cx, int cy, int cz, int R, int base,
                                           int lane, unsigned& lo, unsigned& hi) {
  const int side = 2 * R + 1, r = base + lane;
  lo = hi = 0;
  if (r >= side * side) return;
  const int y = cy - R + r % side, z = cz - R + r / side;
  if (y < 0 || y >= g.ny || z < 0 || z >= g.nz) return;
  lo = lower_bound(keys, n, cell_id(g, max(cx - R, 0), y, z));
  hi = lower_bound(keys, n, cell_id(g, min(cx + R, g.nx - 1), y, z) + 1u);
}
// `f(valid, j, p)` is called by all 32 lanes together (warp-uniform loop), `valid` lanes holding one candidate each.  For R == 1
// (9 rows) the caller passes the runs it looked up once (clo, chi); wider rings look their rows up again in every sweep
// (rare: isolated points).
template <class F>
__device__ __forceinline__ void sweep_block(const Grid& g, const unsigned* __restrict__ keys, const float4* __restrict__ sp, unsigned n, int cx,
                                            int cy, int cz, int R, int lane, unsigned clo, unsigned chi, F& f) {
  const int rows = (2 * R + 1) * (2 * R + 1);
  for (int base = 0; base < rows; base += 32) {
    unsigned lo = clo, hi = chi;
    if (R != 1) row_ranges(g, keys, n, cx, cy, cz, R, base, lane, lo, hi);
    const int nr = min(32, rows - base);
    for (int k = 0; k < nr; k++) {
      const unsigned a = __shfl_sync(0xffffffffu, lo, k), b = __shfl_sync(0xffffffffu, hi, k);
      for (unsigned j0 = a; j0 < b; j0 += 32) {
        const unsigned j = j0 + lane;
        const bool valid = j < b;
        f(valid, valid ? j : a, sp[valid ? j : a]);
      }
    }
  }
}

#define SINK_KEY_CAP 1024  // candidate distances a warp keeps in shared memory between the passes of its radix select

// adds one digit of `key` to the per-warp histogram; lanes with the same digit elect one lane (neighbours share their leading
// digits: no same-address atomics).  Called by all 32 lanes.
__device__ __forceinline__ void hist_add(int* hist, bool take, unsigned key, int shift) {
  const unsigned bin = take ? ((key >> shift) & 255u) : 256u;
  const unsigned peers = __match_any_sync(0xffffffffu, bin);
  if (bin < 256u && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
}

struct HistPass {
  float qx, qy, qz;
  unsigned prefix;
  int shift;
  bool first;
  int* hist;
  unsigned* skeys;  // first pass: the candidates' distance bit patterns are kept here (up to SINK_KEY_CAP)
  int cnt;
  __device__ __forceinline__ void operator()(bool valid, unsigned, const float4& p) {
    const unsigned key = __float_as_uint(dist2(p.x, p.y, p.z, qx, qy, qz));
    if (first) {
      const int nv = __popc(__ballot_sync(0xffffffffu, valid));  // valid lanes are a prefix of the warp
      if (valid && cnt + nv <= SINK_KEY_CAP) skeys[cnt + (threadIdx.x & 31)] = key;
      cnt += nv;
    }
    hist_add(hist, valid && (first || (key >> (shift + 8)) == prefix), key, shift);
  }
};
struct SumPass {
  float qx, qy, qz;
  unsigned T;
  double sum;
  __device__ __forceinline__ void operator()(bool valid, unsigned, const float4& p) {
    const float d2 = dist2(p.x, p.y, p.z, qx, qy, qz);
    if (valid && __float_as_uint(d2) < T) sum += (double)__fsqrt_rn(d2);
  }
};

// mean distance to the mean_k nearest other points (the query itself is the (0-distance) first neighbour, as in PCL)
__global__ void __launch_bounds__(256) k_sor_mean_dist(Grid g, const unsigned* __restrict__ keys, const float4* __restrict__ sp, unsigned n_valid,
                                                       int mean_k, double* __restrict__ mean_dist, unsigned long long* __restrict__ counters) {
  __shared__ int s_hist[8][256];
  __shared__ unsigned s_keys[8][SINK_KEY_CAP];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* hist = s_hist[warp];
  unsigned* skeys = s_keys[warp];
  const unsigned nwarps = gridDim.x * 8;
  const int rmax = max(g.nx, max(g.ny, g.nz));
  for (unsigned j = blockIdx.x * 8 + warp; j < n_valid; j += nwarps) {
    const float4 q = sp[j];
    const int cx = cell_coord(q.x, g.minx, g.inv_c, g.nx), cy = cell_coord(q.y, g.miny, g.inv_c, g.ny), cz = cell_coord(q.z, g.minz, g.inv_c, g.nz);
    const int want = mean_k + 1;  // neighbours incl. the query
    double result = 0;
    unsigned clo, chi;
    row_ranges(g, keys, n_valid, cx, cy, cz, 1, 0, lane, clo, chi);
    for (int R = 1;;) {
      unsigned prefix = 0;
      int rank = want - 1, less = 0, total = 0, staged = -1;  // staged >= 0: that many keys sit in shared memory
      bool enough = true;
      for (int pass = 0; pass < 4 && enough; pass++) {
        for (int b = lane; b < 256; b += 32) hist[b] = 0;
        __syncwarp();
        if (staged >= 0) {  // later passes read the staged distances instead of the points
          for (int i0 = 0; i0 < staged; i0 += 32) {
            const int i = i0 + lane;
            const unsigned key = i < staged ? skeys[i] : 0u;
            hist_add(hist, i < staged && (key >> (32 - 8 * pass)) == prefix, key, 24 - 8 * pass);
          }
        } else {
          HistPass hp{q.x, q.y, q.z, prefix, 24 - 8 * pass, pass == 0, hist, skeys, 0};
          sweep_block(g, keys, sp, n_valid, cx, cy, cz, R, lane, clo, chi, hp);
          if (pass == 0 && hp.cnt <= SINK_KEY_CAP) staged = hp.cnt;
        }
        __syncwarp();
        int s = 0;
        int h[8];
#pragma unroll
        for (int b = 0; b < 8; b++) { h[b] = hist[lane * 8 + b]; s += h[b]; }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const int excl = incl - s;
        if (pass == 0) total = __shfl_sync(0xffffffffu, incl, 31);
        if (pass == 0 && total < want) { enough = false; __syncwarp(); break; }
        const bool mine = excl <= rank && rank < incl;
        const unsigned who = __ballot_sync(0xffffffffu, mine);
        const int src = __ffs(who) - 1;
        int bin = 0, below = excl;
        if (mine) {
#pragma unroll
          for (int b = 0; b < 8; b++) {
            if (rank >= below + h[b]) { below += h[b]; bin = b + 1; } else break;
          }
          bin += lane * 8;
        }
        bin = __shfl_sync(0xffffffffu, bin, src);
        below = __shfl_sync(0xffffffffu, below, src);
        less += below;
        rank -= below;
        prefix = (prefix << 8) | (unsigned)bin;
        __syncwarp();
      }
      const bool whole = R >= rmax;
      if (!enough && !whole) { R = min(2 * R, rmax); continue; }  // fewer than k+1 points in the block: widen
      const float lim = (float)R * g.c * 0.99999f;
      const unsigned T = enough ? prefix : 0x7f800000u;  // not enough points in the whole cloud: take them all
      if (enough && !whole && !(__uint_as_float(T) < lim * lim)) {
        // the k-th neighbour may lie outside the block, but it is no farther than sqrt(T): one ring of that size settles it
        R = min(max(R + 1, (int)ceilf(__fsqrt_ru(__uint_as_float(T)) / (g.c * 0.99999f))), rmax);
        continue;
      }
      double s = 0;
      if (staged >= 0) {
        for (int i = lane; i < staged; i += 32) {
          const unsigned key = skeys[i];
          if (key < T) s += (double)__fsqrt_rn(__uint_as_float(key));
        }
      } else {
        SumPass spass{q.x, q.y, q.z, T, 0.0};
        sweep_block(g, keys, sp, n_valid, cx, cy, cz, R, lane, clo, chi, spass);
        s = spass.sum;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (enough) s += (double)(want - less) * (double)__fsqrt_rn(__uint_as_float(T));
      result = s / (double)mean_k;
      if (lane == 0 && R > 1) atomicAdd(counters, 1ull);
      break;
    }
    __syncwarp();  // the next query reuses this warp's histogram and staged keys
    if (lane == 0) mean_dist[__float_as_int(q.w)] = result;
  }
}

__global__ void k_keep_flags(const double* __restrict__ mean_dist, const float4* __restrict__ p32, long n, double threshold, int* __restrict__ keep) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = (p32[i].w != 0.0f && !(mean_dist[i] > threshold)) ? 1 : 0;
}

struct CovPass {
  float qx, qy, qz, r2;
  const unsigned char* keep;  // keep flags in the order of the sorted array
  double s[9];
  int cnt;
  __device__ __forceinline__ void operator()(bool valid, unsigned j, const float4& p) {
    if (!valid || !keep[j]) return;
    if (!(dist2(p.x, p.y, p.z, qx, qy, qz) < r2)) return;
    const double dx = (double)p.x - (double)qx, dy = (double)p.y - (double)qy, dz = (double)p.z - (double)qz;  // exact
    s[0] += dx; s[1] += dy; s[2] += dz;
    s[3] += dx * dx; s[4] += dx * dy; s[5] += dx * dz; s[6] += dy * dy; s[7] += dy * dz; s[8] += dz * dz;
    cnt++;
  }
};

// smallest eigenpair of a symmetric 3x3 matrix by cyclic Jacobi rotations
__device__ void smallest_eigen(double a[3][3], double& lambda, double v[3], double& trace) {
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  trace = a[0][0] + a[1][1] + a[2][2];
  for (int sweep = 0; sweep < 12; sweep++) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off <= 1e-300 || off <= 1e-17 * fabs(trace)) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a[p][q] == 0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; k++) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int m = 0;
  if (a[1][1] < a[m][m]) m = 1;
  if (a[2][2] < a[m][m]) m = 2;
  lambda = a[m][m];
  const double nrm = sqrt(V[0][m] * V[0][m] + V[1][m] * V[1][m] + V[2][m] * V[2][m]);
  for (int k = 0; k < 3; k++) v[k] = V[k][m] / nrm;
}

__global__ void k_keep_sorted(const float4* __restrict__ sp, unsigned n_valid, const int* __restrict__ keep, unsigned char* __restrict__ keep_s) {
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_valid) keep_s[j] = (unsigned char)keep[__float_as_int(sp[j].w)];
}

// A warp takes 32 consecutive queries of the sorted array: the neighbourhood sums of query b are gathered by all lanes together
// and handed to lane b; then the 32 lanes solve their 32 eigenproblems side by side (a lane-0-only solve left 31 lanes idle
// for the longest part of the kernel: 46 -> see DESIGN.md 11).
__global__ void __launch_bounds__(256) k_normals(Grid g, const unsigned* __restrict__ keys, const float4* __restrict__ sp, unsigned n_valid,
                                                 const unsigned char* __restrict__ keep, const int* __restrict__ rank, float radius2, float cx_, float cy_, float cz_, float* __restrict__ out,
                                                 int* __restrict__ kept_index) {
  const int lane = threadIdx.x & 31;
  const unsigned nwarps = gridDim.x * 8;
  for (unsigned base = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 32u; base < n_valid; base += nwarps * 32u) {
    double ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // this lane's query: sums about the query point
    int mcnt = -1;                                // -1: no query (past the end or not kept)
    float4 mq = make_float4(0, 0, 0, 0);
    const int nq = (int)min(32u, n_valid - base);
    for (int b = 0; b < nq; b++) {
      const unsigned j = base + b;
      if (!keep[j]) continue;  // warp-uniform
      const float4 q = sp[j];
      const int cx = cell_coord(q.x, g.minx, g.inv_c, g.nx), cy = cell_coord(q.y, g.miny, g.inv_c, g.ny), cz = cell_coord(q.z, g.minz, g.inv_c, g.nz);
      CovPass cp{q.x, q.y, q.z, radius2, keep, {0, 0, 0, 0, 0, 0, 0, 0, 0}, 0};
      unsigned clo, chi;
      row_ranges(g, keys, n_valid, cx, cy, cz, 1, 0, lane, clo, chi);
      sweep_block(g, keys, sp, n_valid, cx, cy, cz, 1, lane, clo, chi, cp);
#pragma unroll
      for (int o = 16; o; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 9; k++) cp.s[k] += __shfl_xor_sync(0xffffffffu, cp.s[k], o);
        cp.cnt += __shfl_xor_sync(0xffffffffu, cp.cnt, o);
      }
      if (lane == b) {
#pragma unroll
        for (int k = 0; k < 9; k++) ms[k] = cp.s[k];
        mcnt = cp.cnt;
        mq = q;
      }
    }
    if (mcnt < 0) continue;
    const float4 q = mq;
    const int orig = __float_as_int(q.w);
    float* o = out + (size_t)rank[orig] * 7;
    if (kept_index) kept_index[rank[orig]] = orig;
    o[0] = q.x; o[1] = q.y; o[2] = q.z;
    if (mcnt < 3) {  // computePointNormal fails: NaN normal and curvature
      o[3] = o[4] = o[5] = o[6] = __int_as_float(0x7fc00000);
      continue;
    }
    const double n = mcnt, mx = ms[0] / n, my = ms[1] / n, mz = ms[2] / n;
    double a[3][3];
    a[0][0] = ms[3] / n - mx * mx; a[0][1] = a[1][0] = ms[4] / n - mx * my; a[0][2] = a[2][0] = ms[5] / n - mx * mz;
    a[1][1] = ms[6] / n - my * my; a[1][2] = a[2][1] = ms[7] / n - my * mz; a[2][2] = ms[8] / n - mz * mz;
    double lam, v[3], tr;
    smallest_eigen(a, lam, v, tr);
    float nx = (float)v[0], ny = (float)v[1], nz = (float)v[2];
    const float curv = tr != 0 ? (float)fabs(lam / tr) : 0.0f;
    // flipNormalTowardsViewpoint with the default viewpoint (0, 0, 0): vp - p = -p
    if (-q.x * nx - q.y * ny - q.z * nz < 0) { nx = -nx; ny = -ny; nz = -nz; }
    // CCloudOptimization.cpp:109-116: towards the camera centre of the pair's first view
    if (nx * (cx_ - q.x) + ny * (cy_ - q.y) + nz * (cz_ - q.z) < 0) { nx = -nx; ny = -ny; nz = -nz; }
    o[3] = nx; o[4] = ny; o[5] = nz; o[6] = curv;
  }
}

struct DevBuf {
  std::vector<void*> ptrs;
  template <class T> cudaError_t alloc(T** p, size_t n) {
    const cudaError_t e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
  ~DevBuf() { for (void* p : ptrs) cudaFree(p); }
};

thread_local std::string g_sink_error;

#define SK(x)                                                                                   \
  do {                                                                                          \
    const cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) { g_sink_error = std::string(#x) + ": " + cudaGetErrorString(e_); return SB200_ERR_CUDA; } \
  } while (0)

__global__ void k_count_cells(const unsigned* __restrict__ keys, unsigned n_valid, unsigned long long* __restrict__ out) {
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool first = j < n_valid && (j == 0 || keys[j] != keys[j - 1]);
  const unsigned b = __ballot_sync(0xffffffffu, first);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, (unsigned long long)__popc(b));
}

struct SortedGrid {
  Grid g;
  unsigned* keys = nullptr;  // sorted cell ids
  float4* sp = nullptr;      // points in that order, w = original index
  unsigned n_valid = 0;      // finite points (the others sort last)
};

// uniform grid of edge c over the bounding box, points sorted by cell id (stable: index order inside a cell)
int build_sorted_grid(DevBuf& mem, const float4* p32, int64_t n, float c, const float* lo, const float* ext, SortedGrid& out, cudaStream_t st) {
  for (;;) {
    const double cells = (floor(ext[0] / c) + 1) * (floor(ext[1] / c) + 1) * (floor(ext[2] / c) + 1);
    if (cells < 4.0e9) break;
    c *= 2;
  }
  Grid& g = out.g;
  g.minx = lo[0]; g.miny = lo[1]; g.minz = lo[2];
  g.c = c; g.inv_c = 1.0f / c;
  g.nx = (int)floorf(ext[0] * g.inv_c) + 1; g.ny = (int)floorf(ext[1] * g.inv_c) + 1; g.nz = (int)floorf(ext[2] * g.inv_c) + 1;
  unsigned *keys = nullptr, *idx = nullptr, *idx2 = nullptr;
  SK(mem.alloc(&keys, n)); SK(mem.alloc(&out.keys, n)); SK(mem.alloc(&idx, n)); SK(mem.alloc(&idx2, n));
  SK(mem.alloc(&out.sp, n));
  const unsigned nb = (unsigned)((n + 255) / 256);
  k_cell_keys<<<nb, 256, 0, st>>>(p32, n, g, keys, idx);
  size_t tmp_bytes = 0;
  SK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, out.keys, idx, idx2, (int)n, 0, 32, st));
  void* tmp = nullptr;
  SK(mem.alloc((char**)&tmp, tmp_bytes));
  SK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, out.keys, idx, idx2, (int)n, 0, 32, st));
  k_gather<<<nb, 256, 0, st>>>(p32, idx2, n, out.sp);
  // valid points = keys below 0xffffffff
  out.n_valid = (unsigned)n;
  unsigned last = 0;
  SK(cudaMemcpyAsync(&last, out.keys + (n - 1), 4, cudaMemcpyDeviceToHost, st));
  SK(cudaStreamSynchronize(st));
  if (last == 0xffffffffu) {  // binary search through single-element copies — only when a non-finite point exists
    unsigned l = 0, h = (unsigned)n;
    while (l < h) {
      const unsigned mid = (l + h) >> 1;
      unsigned v = 0;
      SK(cudaMemcpy(&v, out.keys + mid, 4, cudaMemcpyDeviceToHost));
      if (v < 0xffffffffu) l = mid + 1; else h = mid;
    }
    out.n_valid = l;
  }
  return SB200_OK;
}

int sink_filter_device(const double* d_xyz, int64_t n, int mean_k, double std_mul, double radius, const double* cam, float* out_host,
                       int32_t* kept_host, int64_t capacity, int64_t* n_kept, double* stats, cudaStream_t st) {
  if (mean_k < 1) { g_sink_error = "sor_meank must be >= 1"; return SB200_ERR_BAD_ARG; }
  DevBuf mem;
  float4* p32 = nullptr;
  int* bbox = nullptr;
  SK(mem.alloc(&p32, n));
  SK(mem.alloc(&bbox, 6));
  cudaEvent_t ev0, ev1;
  SK(cudaEventCreate(&ev0));
  SK(cudaEventCreate(&ev1));
  SK(cudaEventRecord(ev0, st));
  const int h_init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
  SK(cudaMemcpyAsync(bbox, h_init, sizeof h_init, cudaMemcpyHostToDevice, st));
  const unsigned nb = (unsigned)((n + 255) / 256);
  k_narrow_bbox<<<nb, 256, 0, st>>>(d_xyz, n, p32, bbox);
  int hb[6];
  SK(cudaMemcpyAsync(hb, bbox, sizeof hb, cudaMemcpyDeviceToHost, st));
  SK(cudaStreamSynchronize(st));
  auto dec = [](int b) { b = b >= 0 ? b : b ^ 0x7fffffff; float f; memcpy(&f, &b, 4); return f; };
  if (!(hb[0] <= hb[3])) { g_sink_error = "no finite point"; return SB200_ERR_BAD_ARG; }
  const float lo[3] = {dec(hb[0]), dec(hb[1]), dec(hb[2])};
  const float ext[3] = {dec(hb[3]) - lo[0], dec(hb[4]) - lo[1], dec(hb[5]) - lo[2]};

  // grid N: cell edge = the normal radius, so the radius search is exactly the 27 cells around the query
  SortedGrid gn;
  int rc = build_sorted_grid(mem, p32, n, (float)radius, lo, ext, gn, st);
  if (rc != SB200_OK) return rc;
  // grid S for the k-nearest search: sized from the mean occupancy m of grid N's non-empty cells so that, for points on a surface
  // (density m / c^2), the sphere holding k+1 points (radius^2 = (k+1) c^2 / (pi m)) stays inside one cell edge with margin
  unsigned long long* counters = nullptr;
  SK(mem.alloc(&counters, 2));
  SK(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned long long), st));
  k_count_cells<<<(gn.n_valid + 255) / 256, 256, 0, st>>>(gn.keys, gn.n_valid, counters + 1);
  unsigned long long h_cells = 0;
  SK(cudaMemcpyAsync(&h_cells, counters + 1, sizeof h_cells, cudaMemcpyDeviceToHost, st));
  SK(cudaStreamSynchronize(st));
  const double occupancy = (double)gn.n_valid / (double)(h_cells ? h_cells : 1);
  float c_sor = gn.g.c * (float)sqrt(0.8 * (mean_k + 1) / occupancy);
  SortedGrid gs_own;
  const SortedGrid* gs = &gn;
  if (c_sor < 0.8f * gn.g.c || c_sor > 1.1f * gn.g.c) {
    rc = build_sorted_grid(mem, p32, n, c_sor, lo, ext, gs_own, st);
    if (rc != SB200_OK) return rc;
    gs = &gs_own;
  }
  const unsigned n_valid = gn.n_valid;
  std::vector<double> h_mean((size_t)n);
  double* mean_dist = nullptr;
  SK(mem.alloc(&mean_dist, n));
  SK(cudaMemsetAsync(mean_dist, 0, sizeof(double) * n, st));
  const int grid = 148 * 8;
  k_sor_mean_dist<<<grid, 256, 0, st>>>(gs->g, gs->keys, gs->sp, n_valid, mean_k, mean_dist, counters);
  SK(cudaGetLastError());
  SK(cudaMemcpyAsync(h_mean.data(), mean_dist, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  unsigned long long h_widened = 0;
  SK(cudaMemcpyAsync(&h_widened, counters, sizeof h_widened, cudaMemcpyDeviceToHost, st));
  SK(cudaStreamSynchronize(st));
  // statistical_outlier_removal.hpp: sequential double sums over the valid points in index order
  std::vector<float4> h_p;  // validity flags only when needed
  double sum = 0, sq = 0;
  int64_t valid = 0;
  if (n_valid == (unsigned)n) {
    for (int64_t i = 0; i < n; i++) { sum += h_mean[i]; sq += h_mean[i] * h_mean[i]; }
    valid = n;
  } else {
    h_p.resize((size_t)n);
    SK(cudaMemcpy(h_p.data(), p32, sizeof(float4) * n, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; i++)
      if (h_p[i].w != 0.0f) { sum += h_mean[i]; sq += h_mean[i] * h_mean[i]; valid++; }
  }
  const double mean = sum / (double)valid;
  const double variance = valid > 1 ? (sq - sum * sum / (double)valid) / ((double)valid - 1) : 0.0;
  const double stddev = sqrt(variance > 0 ? variance : 0.0);
  const double threshold = mean + std_mul * stddev;

  int *keep = nullptr, *rank = nullptr;
  SK(mem.alloc(&keep, n + 1));
  SK(mem.alloc(&rank, n + 1));
  SK(cudaMemsetAsync(keep + n, 0, sizeof(int), st));
  k_keep_flags<<<nb, 256, 0, st>>>(mean_dist, p32, n, threshold, keep);
  size_t scan_bytes = 0;
  SK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, keep, rank, (int)(n + 1), st));
  void* tmp2 = nullptr;
  SK(mem.alloc((char**)&tmp2, scan_bytes));
  SK(cub::DeviceScan::ExclusiveSum(tmp2, scan_bytes, keep, rank, (int)(n + 1), st));
  int total = 0;
  SK(cudaMemcpyAsync(&total, rank + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  SK(cudaStreamSynchronize(st));
  *n_kept = total;
  if (stats) { stats[0] = mean; stats[1] = stddev; stats[2] = threshold; stats[4] = (double)h_widened; }
  if (total > capacity) { g_sink_error = "output capacity too small"; return SB200_ERR_BAD_ARG; }
  float* d_out = nullptr;
  int* d_kept = nullptr;
  SK(mem.alloc(&d_out, (size_t)total * 7));
  if (kept_host) SK(mem.alloc(&d_kept, (size_t)total));
  const float r2 = (float)(radius * radius);
  unsigned char* keep_s = nullptr;
  SK(mem.alloc(&keep_s, n));
  k_keep_sorted<<<(n_valid + 255) / 256, 256, 0, st>>>(gn.sp, n_valid, keep, keep_s);
  k_normals<<<grid, 256, 0, st>>>(gn.g, gn.keys, gn.sp, n_valid, keep_s, rank, r2, (float)cam[0], (float)cam[1], (float)cam[2], d_out, d_kept);
  SK(cudaGetLastError());
  SK(cudaEventRecord(ev1, st));
  SK(cudaMemcpyAsync(out_host, d_out, sizeof(float) * 7 * (size_t)total, cudaMemcpyDeviceToHost, st));
  if (kept_host) SK(cudaMemcpyAsync(kept_host, d_kept, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, st));
  SK(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, ev0, ev1);
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  if (stats) stats[3] = ms;
  return SB200_OK;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* sb200_sink_last_error(void) { return g_sink_error.c_str(); }

int sb200_sink_filter(int device, const double* xyz, int64_t n, int sor_meank, double sor_std_mul, double normal_radius,
                      const double* cam_center, float* out_xyz_normal_curv, int32_t* kept_index, int64_t capacity, int64_t* n_kept,
                      double* stats5) {
  g_sink_error.clear();
  if (!xyz || n <= 0 || n >= (1ll << 31) - 2 || !cam_center || !out_xyz_normal_curv || !n_kept || !(normal_radius > 0)) {
    g_sink_error = "bad argument";
    return SB200_ERR_BAD_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { g_sink_error = "no CUDA device (there is no CPU path)"; return SB200_ERR_NO_DEVICE; }
  SK(cudaSetDevice(device));
  cudaStream_t st;
  SK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  double* d_xyz = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_xyz, sizeof(double) * 3 * (size_t)n);
  if (e != cudaSuccess) { cudaStreamDestroy(st); g_sink_error = cudaGetErrorString(e); return SB200_ERR_CUDA; }
  int rc = SB200_OK;
  e = cudaMemcpyAsync(d_xyz, xyz, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { g_sink_error = cudaGetErrorString(e); rc = SB200_ERR_CUDA; }
  if (rc == SB200_OK) rc = sink_filter_device(d_xyz, n, sor_meank, sor_std_mul, normal_radius, cam_center, out_xyz_normal_curv, kept_index, capacity, n_kept, stats5, st);
  cudaFree(d_xyz);
  cudaStreamDestroy(st);
  return rc;
}

}  // extern "C"
#pragma GCC visibility pop
