// constraints.cu — K4 SmoothConstraint, K5 OrderConstraint, K6 UniquenessContraint, K8 MedianFilter.
// Integer stages: bit-exact by construction; each reformulates a sequential sweep of the reference
// as a gather / scan that produces the same bytes.
#include "kernels.h"

// ------------------------------------------------------------------------------------------------
// K4  SmoothConstraint (CStereoMatching.cpp:370-448).
// Pass 1 of the reference visits every valid pixel a of the margin rectangle and, for each valid
// neighbour b in {E, SW, S, SE}, bumps a [count, differ] byte pair of both ends (differ when
// |a-b| > 1, :3).  Increments commute, so a pixel's counters are a pure function of its 3x3
// neighbourhood — plus the stray increments of quirk Q4: the SE case bumps `qup[x]` / `qdown[x+2]`
// (byte offsets, not 2x), i.e. field (x&1) of pixel x>>1 in row y and of pixel (x+2)>>1 in row
// y+1, instead of the counts of a and b.  Pass 2 (:434-447) kills a pixel when count == 0 or
// 2*differ > count.  One thread per pixel, out of place.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_smooth(const short* __restrict__ in, short* __restrict__ out, int W, int H, Bound m) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t f = (size_t)y * W + x;
  const short c = in[f];
  if (x < m.XL || x > m.XR || y < m.YL || y > m.YR) { out[f] = c; return; }
  auto val = [&](int xx, int yy) -> int {  // disparity or NOMATCH outside the image
    return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? (int)in[(size_t)yy * W + xx] : SB_NOMATCH;
  };
  auto visited = [&](int xx, int yy) -> bool {  // pass 1 starts from valid pixels of the rectangle
    return xx >= m.XL && xx <= m.XR && yy >= m.YL && yy <= m.YR && in[(size_t)yy * W + xx] != SB_NOMATCH;
  };
  int count = 0, differ = 0;
  const int vc = c;
  if (vc != SB_NOMATCH) {
    // p as the visiting end: E, SW, S bump count; E, SW, S, SE bump differ
    const int e = val(x + 1, y), sw = val(x - 1, y + 1), s = val(x, y + 1), se = val(x + 1, y + 1);
    if (e != SB_NOMATCH) { count++; differ += abs(vc - e) > 1; }
    if (sw != SB_NOMATCH) { count++; differ += abs(vc - sw) > 1; }
    if (s != SB_NOMATCH) { count++; differ += abs(vc - s) > 1; }
    if (se != SB_NOMATCH) { differ += abs(vc - se) > 1; }
    // p as the visited neighbour of W (its E), NE (its SW), N (its S), NW (its SE: differ only)
    if (visited(x - 1, y)) { const int a = val(x - 1, y); count++; differ += abs(a - vc) > 1; }
    if (visited(x + 1, y - 1)) { const int a = val(x + 1, y - 1); count++; differ += abs(a - vc) > 1; }
    if (visited(x, y - 1)) { const int a = val(x, y - 1); count++; differ += abs(a - vc) > 1; }
    if (visited(x - 1, y - 1)) { const int a = val(x - 1, y - 1); differ += abs(a - vc) > 1; }
  }
  // stray SE increments (Q4): from a = (xa, y) via qup[xa]  and from a = (xa, y-1) via qdown[xa+2]
  {
    const int xa0 = 2 * x, xa1 = 2 * x + 1;
    if (xa0 < W && visited(xa0, y) && val(xa0 + 1, y + 1) != SB_NOMATCH) count++;
    if (xa1 < W && visited(xa1, y) && val(xa1 + 1, y + 1) != SB_NOMATCH) differ++;
    const int xb0 = 2 * x - 2, xb1 = 2 * x - 1;
    if (xb0 >= 0 && xb0 < W && visited(xb0, y - 1) && val(xb0 + 1, y) != SB_NOMATCH) count++;
    if (xb1 >= 0 && xb1 < W && visited(xb1, y - 1) && val(xb1 + 1, y) != SB_NOMATCH) differ++;
  }
  count &= 255; differ &= 255;  // the reference's counters are bytes
  out[f] = (count == 0 || (differ << 1) > count) ? (short)SB_NOMATCH : c;
}

int launch_smooth(const short* in, short* out, int W, int H, Bound m, cudaStream_t st) {
  dim3 grid((W + 255) / 256, H);
  k_smooth<<<grid, 256, 0, st>>>(in, out, W, H, m);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// K5  OrderConstraint (CStereoMatching.cpp:310-368).
// Per row: points (x, t = x + d) of the valid pixels; i and j (j < i) cross when t_j > t_i.  The
// reference builds the dense crossing matrix and repeatedly deletes the point with the most
// crossings (first maximum, op_max_meat.hpp) until none remain.  Here: one block per row, points
// compacted to shared memory, crossing counts by a bounded scan (a crossing needs
// x_i - x_j < d_j - d_i <= dmax - dmin), then the same greedy loop with a packed block arg-max.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_order(short* __restrict__ disp, int W, Bound m) {
  extern __shared__ unsigned char s_raw[];
  const int cap = m.width;
  short* line = (short*)s_raw;              // t = x + d
  short* xs = line + cap;                   // x
  unsigned short* cnt = (unsigned short*)(xs + cap);
  __shared__ int s_n, s_red[8], s_red2[8], s_flag;
  __shared__ unsigned s_key[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  short* p = disp + (size_t)(m.YL + blockIdx.x) * W;

  // ---- ordered compaction of the valid pixels --------------------------------------------------
  if (tid == 0) { s_n = 0; s_flag = 0; }
  __syncthreads();
  int dmin = 1 << 30, dmax = -(1 << 30);
  for (int base = m.XL; base <= m.XR; base += 256) {
    const int x = base + tid;
    const int d = x <= m.XR ? (int)p[x] : SB_NOMATCH;
    const bool v = d != SB_NOMATCH;
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    if (lane == 0) s_red[warp] = __popc(bal);
    __syncthreads();
    int off = s_n;
    for (int k = 0; k < warp; k++) off += s_red[k];
    if (v) {
      const int i = off + __popc(bal & ((1u << lane) - 1));
      line[i] = (short)(d + x);
      xs[i] = (short)x;
      cnt[i] = 0;
      dmin = min(dmin, d); dmax = max(dmax, d);
    }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int k = 0; k < 8; k++) t += s_red[k]; s_n += t; }
    __syncthreads();
  }
  const int n = s_n;
  if (n < 2) return;
  // ---- quick exit: a non-decreasing t sequence has no crossing ---------------------------------
  int local = 0;
  for (int i = tid + 1; i < n; i += 256) local |= line[i - 1] > line[i];
  if (local) s_flag = 1;
  // block range of d
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    dmin = min(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  }
  if (lane == 0) { s_red[warp] = dmin; s_red2[warp] = dmax; }
  __syncthreads();
  if (!s_flag) return;
  for (int k = 0; k < 8; k++) { dmin = min(dmin, s_red[k]); dmax = max(dmax, s_red2[k]); }
  const int range = dmax - dmin;  // a crossing pair is closer than this in x
  __syncthreads();
  // ---- crossing counts -------------------------------------------------------------------------
  int ones2 = 0;  // sum of counts = 2 * number of crossing pairs
  for (int i = tid; i < n; i += 256) {
    const int ti = line[i], xi = xs[i];
    int c = 0;
    for (int j = i - 1; j >= 0 && xi - xs[j] < range; j--) c += line[j] > ti;
    for (int j = i + 1; j < n && xs[j] - xi < range; j++) c += ti > line[j];
    cnt[i] = (unsigned short)c;
    ones2 += c;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ones2 += __shfl_xor_sync(0xffffffffu, ones2, o);
  if (lane == 0) s_red[warp] = ones2;
  __syncthreads();
  int ones = 0;
  for (int k = 0; k < 8; k++) ones += s_red[k];
  ones >>= 1;
  __syncthreads();
  // ---- greedy deletion (:354-364) ---------------------------------------------------------------
  // Deleted points keep count 0 and line = -32768 marks them dead for the crossing test below.
  while (ones > 0) {
    unsigned key = 0;  // (count << 16) | (65535 - i): max = largest count, then smallest index
    for (int i = tid; i < n; i += 256) key = max(key, ((unsigned)cnt[i] << 16) | (unsigned)(65535 - i));
#pragma unroll
    for (int o = 16; o; o >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, o));
    if (lane == 0) s_key[warp] = key;
    __syncthreads();
    for (int k = 0; k < 8; k++) key = max(key, s_key[k]);
    const int bi = 65535 - (int)(key & 0xffffu), bv = (int)(key >> 16);
    const int tb = line[bi], xb = xs[bi];
    __syncthreads();
    for (int k = tid; k < n; k += 256) {
      if (k == bi || xs[k] < 0) continue;  // xs < 0 marks a deleted point
      const bool cross = k < bi ? line[k] > tb : tb > line[k];
      if (cross) cnt[k]--;
    }
    if (tid == 0) { cnt[bi] = 0; xs[bi] = -1; p[xb] = (short)SB_NOMATCH; }
    ones -= bv;
    __syncthreads();
  }
}

int launch_order(short* disp, int W, int H, Bound m, cudaStream_t st) {
  (void)H;
  if (m.width <= 0 || m.height <= 0) return 0;
  const size_t smem = (size_t)m.width * 6;
  if (smem > 48 * 1024) {  // rows wider than 8192 px: opt in to the large carve-out (idempotent; per device)
    if (smem > 227 * 1024) return -1;
    cudaFuncSetAttribute(k_order, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  k_order<<<m.height, 256, smem, st>>>(disp, W, m);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// K6  one pass of UniquenessContraint_ (CStereoMatching.cpp:462-497): P filtered against Qm.
// For a valid p[x]: bl = max(int(p+0.5)+x-1, XL1), br = min(bl+2, XR1); keep if some q[im] in
// [bl,br] has |q+p| < 2; otherwise kill unless |q[bl+1]+p[x-1]| < 2 or |q[bl+1]+p[x+1]| < 2, where
// p[x-1] is the value AFTER this sweep touched it (quirk Q5).  A killed or NOMATCH p[x-1] always
// fails its test, so  kill(x) = g(x) | (pr(x) & kill(x-1))  with
//   pr = no partner & p[x+1] test fails,   g = pr & the ORIGINAL p[x-1] test fails
// — a carry chain, resolved per 32-pixel chunk with one 64-bit add (generate/propagate adder).
// One warp per row.
// ------------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ bool lt2(T a, T b);
template <> __device__ __forceinline__ bool lt2<short>(short a, short b) { return abs((int)a + (int)b) < 2; }
template <> __device__ __forceinline__ bool lt2<double>(double a, double b) { return fabs(a + b) < 2; }

// One CTA per row.  Phase 1: the warps compute the (generate, propagate) masks of all 32-pixel chunks of the row in
// parallel (every load of the row is in flight at once); phase 2: one warp resolves the carry chain chunk by chunk with
// 64-bit adds; phase 3: the kills are applied.  All tests read the values from BEFORE this pass, which is what the
// carry formulation needs (a killed p[x-1] and kill(x-1) = 1 give the same outcome).
#define SB_UNIQUE_MAX_CHUNKS 512  // rows up to 16384 pixels
template <class T>
__global__ void __launch_bounds__(256) k_unique(T* __restrict__ P, const T* __restrict__ Qm, int W, long n_px, Bound ms, Bound mt) {
  __shared__ unsigned s_g[SB_UNIQUE_MAX_CHUNKS], s_p[SB_UNIQUE_MAX_CHUNKS], s_kill[SB_UNIQUE_MAX_CHUNKS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int y = ms.YL + blockIdx.x;
  T* p = P + (size_t)y * W;
  const long row0 = (long)y * W;
  auto qa = [&](int im) -> T {  // flat, clamped to the buffer (memory safety only)
    long f = row0 + im;
    f = f < 0 ? 0 : (f >= n_px ? n_px - 1 : f);
    return Qm[f];
  };
  const int nch = (ms.XR - ms.XL + 32) / 32;
  for (int c = warp; c < nch; c += nw) {
    const int x = ms.XL + c * 32 + lane;
    bool g = false, pr = false;
    if (x <= ms.XR) {
      const T pv = p[x];
      if (pv != (T)SB_NOMATCH) {
        const int bl = sb_imax((int)((double)pv + 0.5) + x - 1, mt.XL);
        const int br = sb_imin(bl + 2, mt.XR);
        bool found = false;
        for (int im = bl; im <= br; im++) found = found || lt2<T>(qa(im), pv);
        if (!found) {
          const T qc = qa(bl + 1);
          pr = !lt2<T>(qc, p[x + 1]);
          g = pr && !lt2<T>(qc, p[x - 1]);
        }
      }
    }
    const unsigned G = __ballot_sync(0xffffffffu, g), Pm = __ballot_sync(0xffffffffu, pr);
    if (lane == 0) { s_g[c] = G; s_p[c] = Pm; }
  }
  __syncthreads();
  if (warp == 0 && lane == 0) {
    unsigned cin = 0;
    for (int c = 0; c < nch; c++) {
      // adder with generate G, propagate Pm: A = G, B = G | Pm; carry out of bit i = kill(i)
      const unsigned long long A = s_g[c], B = (unsigned long long)(s_g[c] | s_p[c]);
      const unsigned long long S = A + B + cin;
      const unsigned long long carries = (S ^ A ^ B) >> 1;
      s_kill[c] = (unsigned)carries;
      cin = (unsigned)((carries >> 31) & 1);
    }
  }
  __syncthreads();
  for (int c = warp; c < nch; c += nw) {
    const int x = ms.XL + c * 32 + lane;
    if (x <= ms.XR && ((s_kill[c] >> lane) & 1)) p[x] = (T)SB_NOMATCH;
  }
}

int launch_unique_s16(short* P, const short* Qm, int W, int H, Bound ms, Bound mt, cudaStream_t st) {
  if (ms.width <= 0 || ms.height <= 0) return 0;
  if (ms.width > 32 * SB_UNIQUE_MAX_CHUNKS) return -1;
  k_unique<short><<<ms.height, 256, 0, st>>>(P, Qm, W, (long)W * H, ms, mt);
  return 1;
}
int launch_unique_f64(double* P, const double* Qm, int W, int H, Bound ms, Bound mt, cudaStream_t st) {
  if (ms.width <= 0 || ms.height <= 0) return 0;
  if (ms.width > 32 * SB_UNIQUE_MAX_CHUNKS) return -1;
  k_unique<double><<<ms.height, 256, 0, st>>>(P, Qm, W, (long)W * H, ms, mt);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// K8  MedianFilter (CStereoMatching.cpp:763-815), one iteration.  Window = rows y-1..y+1 x columns
// x-1..x (two columns, quirk Q7); k = valid samples; centre NOMATCH: k >= 4 -> median else NOMATCH;
// centre valid: k <= 2 -> NOMATCH else median.  arma::median of an even count is
// lo + (hi - lo)/2 (op_median_meat.hpp:361-377).  Unmasked and out-of-rectangle pixels -> NOMATCH.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_median(const short* __restrict__ in, const uint8_t* __restrict__ mask,
                                                short* __restrict__ out, int W, int H, Bound m) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t f = (size_t)y * W + x;
  short res = (short)SB_NOMATCH;
  if (x >= m.XL && x <= m.XR && y >= m.YL && y <= m.YR && mask[f] == 255) {
    int v[6];
    int k = 0;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const int s = in[(size_t)(y - 1 + j) * W + (x - 1 + i)];
        v[i * 3 + j] = s;
        k += s != SB_NOMATCH;
      }
    const bool centre_missing = in[f] == SB_NOMATCH;
    if (centre_missing ? (k >= 4) : (k > 2)) {
      // rank of each valid sample among the valid ones (ties broken by position)
      const int half = k >> 1;
      int lo = 0, hi = 0;
#pragma unroll
      for (int a = 0; a < 6; a++) {
        if (v[a] == SB_NOMATCH) continue;
        int r = 0;
#pragma unroll
        for (int b = 0; b < 6; b++)
          if (v[b] != SB_NOMATCH) r += (v[b] < v[a]) || (v[b] == v[a] && b < a);
        if (r == half) hi = v[a];
        if (r == half - 1) lo = v[a];
      }
      res = (short)((k & 1) ? hi : lo + (hi - lo) / 2);
    }
  }
  out[f] = res;
}

int launch_median(const short* in, const uint8_t* mask, short* out, int W, int H, Bound m, cudaStream_t st) {
  dim3 grid((W + 255) / 256, H);
  k_median<<<grid, 256, 0, st>>>(in, mask, out, W, H, m);
  return 1;
}
