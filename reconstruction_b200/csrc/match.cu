// match.cu — K2 LowestLevelInitialMatch, K3 HighLevelInitialMatch, and the NCC search of K7 Rematch.
//
// All three are the same loop in the reference (CStereoMatching.cpp:202-218, :268-300, :535-562):
// for a masked source pixel, argmax over target columns im in [lo, hi] with mask1 == 255 of
//     dot(vecL / normL, vecR) / normR          (strict '>' from -1, first maximum wins)
// and differ only in where [lo, hi] comes from.  One kernel (k_ncc_search) serves them:
//   * G lanes share one source pixel: they build vecL/normL once into shared memory (element-wise
//     work, order-free), then each lane evaluates candidates lo+lane, lo+lane+G, ... in the
//     reference's summation order and the group merges (value, index) with shuffles;
//   * (mean, norm) of every target window come from the per-level statistics map, so a candidate
//     costs one 75-term dot product;
//   * a block owns a 256-pixel scanline segment, compacts its active pixels in shared memory and
//     deals them to its groups, so lanes are not parked on unmasked / already matched pixels.
// Target windows are addressed flat (y*W + im), as the reference's pointer arithmetic does, so a
// range that overruns the row (Rematch quirk at x == XL, :938-939) lands on the same bytes.
#include "kernels.h"
#include "ncc_exact.cuh"

enum { SEARCH_LOWEST = 0, SEARCH_RANGE_MAPS = 1, SEARCH_REMATCH = 2 };

template <int WS, int G, int MODE>
__global__ void __launch_bounds__(256) k_ncc_search(PairViews v, Bound ms, int lo_const, int hi_const,
                                                    const short* __restrict__ lo_map, const short* __restrict__ hi_map,
                                                    short* __restrict__ disp) {
  constexpr int R = WS / 2, N = WS * WS * 3, NG = 256 / G;
  __shared__ double s_vec[NG][N];
  __shared__ short s_list[256];
  __shared__ int s_n;

  const int y = ms.YL + blockIdx.y;
  const int x0 = ms.XL + blockIdx.x * 256;
  const int W = v.W, pitch = 3 * W;
  const long n_px = (long)W * v.H;
  const long stat_first = (long)R * W + R, stat_last = n_px - stat_first;

  // ---- compact the active pixels of this segment (ascending x) --------------------------------
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  {
    const int x = x0 + threadIdx.x;
    bool act = false;
    if (x <= ms.XR) {
      act = v.mask0[(size_t)y * W + x] == 255;
      if (MODE == SEARCH_REMATCH) act = act && disp[(size_t)y * W + x] == SB_NOMATCH;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, act);
    int base = 0;
    if ((threadIdx.x & 31) == 0 && bal) base = atomicAdd(&s_n, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (act) s_list[base + __popc(bal & ((1u << (threadIdx.x & 31)) - 1))] = (short)threadIdx.x;
  }
  __syncthreads();
  const int n_act = s_n;
  const int gid = threadIdx.x / G, gl = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1) << ((threadIdx.x & 31) / G * G));
  double* vec = s_vec[gid];

  for (int a = gid; a < n_act; a += NG) {
    const int x = x0 + s_list[a];
    const long f = (long)y * W + x;
    // ---- vecL = (window - mean) / normL --------------------------------------------------------
    const double2 sl = v.stat0[f];
    const double yl = 1.0 / sl.y;
    const uint8_t* pl = v.img0 + 3 * (f - stat_first);
    for (int k = gl; k < N; k += G) {
      const int j = k / WS, i = k - j * WS;
      vec[k] = div_by_common((double)pl[i * pitch + j] - sl.x, sl.y, yl);
    }
    __syncwarp(gmask);
    // ---- candidate range ----------------------------------------------------------------------
    int lo, hi;
    if (MODE == SEARCH_LOWEST) { lo = lo_const; hi = hi_const; }
    else { lo = lo_map[f]; hi = hi_map[f]; }
    double bv = -1.0;
    int bi = -1;
    for (int im = lo + gl; im <= hi; im += G) {
      const long ft = (long)y * W + im;
      if (ft < 0 || ft >= v.mask_bytes) continue;
      if (v.mask1[ft] != 255) continue;
      double val;
      if (ft >= stat_first && ft < stat_last) {
        const double2 sr = v.stat1[ft];
        val = window_dot_exact<WS>(vec, 1, v.img1 + 3 * (ft - stat_first), pitch, sr.x) / sr.y;
      } else {  // window leaves the payload: bounds-checked evaluation
        double mr;
        const long off0 = 3 * (ft - stat_first);
        const double nr = window_stats_checked<WS>(v.img1, off0, v.img_bytes, pitch, mr);
        val = window_dot_checked<WS>(vec, 1, v.img1, off0, v.img_bytes, pitch, mr) / nr;
      }
      if (val > bv) { bv = val; bi = im; }
    }
#pragma unroll
    for (int o = G / 2; o; o >>= 1) {
      const double ov = __shfl_xor_sync(gmask, bv, o, G);
      const int oi = __shfl_xor_sync(gmask, bi, o, G);
      if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) { bv = ov; bi = oi; }
    }
    if (gl == 0 && bi >= 0) disp[f] = (short)((unsigned short)bi - x);  // ushort temp_i; short(temp_i - x) (:271,:302)
    __syncwarp(gmask);
  }
}

// ------------------------------------------------------------------------------------------------
// The same exact search driven by a pixel list (what the screening pass below leaves unresolved): one group
// of G lanes per listed pixel, groups strided over the whole list, so a cluster of wide-range pixels (hole
// look-ahead) is spread over the GPU instead of serialising inside the block that owns their scanline segment.
// (mean, norm) of the windows are evaluated on the spot by the same routines that fill the statistics map.
// ------------------------------------------------------------------------------------------------
template <int WS, int G>
__global__ void __launch_bounds__(256) k_ncc_search_list(PairViews v, const unsigned* __restrict__ list, const unsigned* __restrict__ n_ptr,
                                                         unsigned cap, const short* __restrict__ lo_map,
                                                         const short* __restrict__ hi_map, int lo_const, int hi_const,
                                                         short* __restrict__ disp) {
  constexpr int R = WS / 2, N = WS * WS * 3, NG = 256 / G;
  __shared__ double s_vec[NG][N];
  const int W = v.W, pitch = 3 * W;
  const long n_px = (long)W * v.H;
  const long stat_first = (long)R * W + R, stat_last = n_px - stat_first;
  const unsigned n = min(*n_ptr, cap);
  const int gid = threadIdx.x / G, gl = threadIdx.x % G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1) << ((threadIdx.x & 31) / G * G));
  double* vec = s_vec[gid];
  for (unsigned e = blockIdx.x * NG + gid; e < n; e += gridDim.x * NG) {
    const long f = list[e];
    const int y = (int)(f / W), x = (int)(f - (long)y * W);
    const uint8_t* pl = v.img0 + 3 * (f - stat_first);  // source pixels lie inside the margin: window inside the payload
    double meanL;
    const double normL = window_stats_exact<WS>(pl, pitch, meanL);
    const double yl = 1.0 / normL;
    for (int k = gl; k < N; k += G) {
      const int j = k / WS, i = k - j * WS;
      vec[k] = div_by_common((double)pl[i * pitch + j] - meanL, normL, yl);
    }
    __syncwarp(gmask);
    const int lo = lo_map ? (int)lo_map[f] : lo_const, hi = lo_map ? (int)hi_map[f] : hi_const;
    double bv = -1.0;
    int bi = -1;
    for (int im = lo + gl; im <= hi; im += G) {
      const long ft = (long)y * W + im;
      if (ft < 0 || ft >= v.mask_bytes) continue;
      if (v.mask1[ft] != 255) continue;
      double val, mr;
      if (ft >= stat_first && ft < stat_last) {
        const uint8_t* pr = v.img1 + 3 * (ft - stat_first);
        const double nr = window_stats_exact<WS>(pr, pitch, mr);
        val = window_dot_exact<WS>(vec, 1, pr, pitch, mr) / nr;
      } else {  // window leaves the payload: bounds-checked evaluation
        const long off0 = 3 * (ft - stat_first);
        const double nr = window_stats_checked<WS>(v.img1, off0, v.img_bytes, pitch, mr);
        val = window_dot_checked<WS>(vec, 1, v.img1, off0, v.img_bytes, pitch, mr) / nr;
      }
      if (val > bv) { bv = val; bi = im; }
    }
#pragma unroll
    for (int o = G / 2; o; o >>= 1) {
      const double ov = __shfl_xor_sync(gmask, bv, o, G);
      const int oi = __shfl_xor_sync(gmask, bi, o, G);
      if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) { bv = ov; bi = oi; }
    }
    if (gl == 0 && bi >= 0) disp[f] = (short)((unsigned short)bi - x);  // ushort temp_i; short(temp_i - x) (:271,:302)
    __syncwarp(gmask);
  }
}

// ------------------------------------------------------------------------------------------------
// Screening pass of the range-map searches (5x5 windows, at most 5 candidates per pixel — the normal case of
// HighLevelInitialMatch, +-m_offset around twice the coarse disparity, :286-287).  The arg-max of
// NCC = cov / (nL nR) over the candidates of one pixel is the arg-max of  key = sign(num) num^2 / varR  with the
// EXACT integers  num = N Sum(LR) - Sum(L) Sum(R),  varR = N Sum(R^2) - Sum(R)^2  (nL is common and positive).
// Sum(LR) comes from dp4a over byte-packed window rows; Sum(R), Sum(R^2) from the per-level integer statistics map.
// A pixel is settled here only when the winner is unambiguous by a margin (relative 1e-4 on key, |rho| >= 1e-3)
// that is ~10 orders of magnitude above the rounding noise of the reference's double evaluation (~1e-14), so the
// reference's own strict-'>' scan must pick the same candidate.  Everything else — near ties, flat windows, wide
// ranges (hole look-ahead), windows touching the buffer edge — is appended to a pixel list and evaluated by
// k_ncc_search_list in the reference's exact arithmetic.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128) k_ncc_screen5(PairViews v, Bound ms, const short* __restrict__ lo_map,
                                                     const short* __restrict__ hi_map, short* __restrict__ disp,
                                                     unsigned* __restrict__ list, unsigned* __restrict__ n_list, unsigned cap) {
  const int x = ms.XL + blockIdx.x * 128 + threadIdx.x, y = ms.YL + blockIdx.y;
  if (x > ms.XR) return;
  const int W = v.W;
  const long f = (long)y * W + x;
  // everything that depends only on f is requested together (one memory round trip), tested afterwards
  const unsigned char m0 = v.mask0[f];
  const short dv = MODE == SEARCH_REMATCH ? disp[f] : (short)SB_NOMATCH;
  const int lo = lo_map[f], hi = hi_map[f];
  const int2 sl = v.istat0[f];
  if (m0 != 255) return;
  if (MODE == SEARCH_REMATCH && dv != SB_NOMATCH) return;
  if (hi < lo) return;  // no candidate: the pixel keeps its value
  const int varL = 75 * sl.y - sl.x * sl.x;
  if (hi - lo > 4 || lo < 2 || hi > W - 3 || x < 2 || x > W - 3 || y < 2 || y > v.H - 3 || varL == 0) {
    const unsigned e = atomicAdd(n_list, 1u);
    if (e < cap) list[e] = (unsigned)f;
    return;
  }
  unsigned L[5][4], Rw[5][7];
#pragma unroll
  for (int r = 0; r < 5; r++) {
    load_row_words<4>(v.img0, ((long)(y - 2 + r) * W + (x - 2)) * 3, L[r]);
    L[r][3] &= 0x00ffffffu;  // 15 bytes per row
    load_row_words<7>(v.img1, ((long)(y - 2 + r) * W + (lo - 2)) * 3, Rw[r]);
  }
  // masks and window statistics of all five candidates are requested up front, together with the window rows above, so
  // the kernel pays two memory round trips per pixel instead of one per candidate (lo + 4 may exceed hi: still inside the
  // row, see the range test above plus the buffer slack)
  unsigned char mk[5];
  int2 st1[5];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    mk[j] = v.mask1[(long)y * W + lo + j];
    st1[j] = v.istat1[(long)y * W + lo + j];
  }
  float best = -3.0e38f, second = -3.0e38f;
  int bj = 0, nvalid = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const int im = lo + j;
    const bool valid = im <= hi && mk[j] == 255;
    const int wi = (3 * j) / 4;
    const unsigned bs = ((3 * j) % 4) * 8;
    unsigned slr = 0;
#pragma unroll
    for (int r = 0; r < 5; r++)
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const unsigned q = bs ? __funnelshift_r(Rw[r][wi + i], Rw[r][wi + i + 1 < 7 ? wi + i + 1 : 6], bs) : Rw[r][wi + i];
        slr = __dp4a(L[r][i], q, slr);
      }
    if (valid) {
      const int2 sr = st1[j];
      const int num = 75 * (int)slr - sl.x * sr.x;
      const int varR = 75 * sr.y - sr.x * sr.x;
      const float fn = (float)num;
      const float key = varR == 0 ? 0.0f : fn * fabsf(fn) / (float)varR;
      nvalid++;
      if (key > best) { second = best; best = key; bj = j; }
      else if (key > second) second = key;
    }
  }
  if (nvalid == 0) return;
  const bool settled = best >= 1.0e-6f * (float)varL && (nvalid == 1 || best - second > 1.0e-4f * best);
  if (settled) {
    disp[f] = (short)((unsigned short)(lo + bj) - x);  // ushort temp_i; short(temp_i - x) (:271,:302)
  } else {
    const unsigned e = atomicAdd(n_list, 1u);
    if (e < cap) list[e] = (unsigned)f;
  }
}

// The same screening for ranges of any width (full-range search of the lowest level, hole look-ahead and Rematch ranges):
// G lanes per listed pixel (32 / G pixels per warp), the lanes of a group stride the candidates, (best, second best) merged
// across the group.  Settles a pixel under the same margin rule as k_ncc_screen5; the rest goes to `out_list` for the exact pass.
// ONFLY: the window sums (sum, sum of squares) are formed from the window words on the spot instead of being read from
// the per-level statistics map (K3's band path does not build that map).
template <int MODE, bool ONFLY, int G>
__global__ void __launch_bounds__(256) k_ncc_screen_wide(PairViews v, const unsigned* __restrict__ list, const unsigned* __restrict__ n_ptr,
                                                         unsigned cap, const short* __restrict__ lo_map, const short* __restrict__ hi_map,
                                                         int lo_const, int hi_const, short* __restrict__ disp,
                                                         unsigned* __restrict__ out_list, unsigned* __restrict__ n_out) {
  constexpr int PPW = 32 / G;  // pixels per warp
  const int W = v.W, lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
  const unsigned n = min(*n_ptr, cap);
  const unsigned nwarp = gridDim.x * (blockDim.x >> 5);
  for (unsigned e0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PPW; e0 < n; e0 += nwarp * PPW) {
    const unsigned e = e0 + grp;
    const bool have = e < n;
    const long f = have ? list[e] : 0;
    const int y = (int)(f / W), x = (int)(f - (long)y * W);
    const int lo = lo_map ? (int)lo_map[f] : lo_const, hi = lo_map ? (int)hi_map[f] : hi_const;
    bool bad = !have || lo < 2 || hi > W - 3 || x < 2 || x > W - 3 || y < 2 || y > v.H - 3 || hi < lo;  // uniform over the group
    int2 sl = make_int2(0, 0);
    int varL = 0;
    float best = -3.0e38f, second = -3.0e38f;
    int bi = -1, nvalid = 0;
    unsigned L[5][4];
    if (!bad) {
#pragma unroll
      for (int r = 0; r < 5; r++) {
        load_row_words<4>(v.img0, ((long)(y - 2 + r) * W + (x - 2)) * 3, L[r]);
        L[r][3] &= 0x00ffffffu;
      }
      if (ONFLY) {
#pragma unroll
        for (int r = 0; r < 5; r++)
#pragma unroll
          for (int i = 0; i < 4; i++) {
            sl.x = (int)__dp4a(L[r][i], 0x01010101u, (unsigned)sl.x);
            sl.y = (int)__dp4a(L[r][i], L[r][i], (unsigned)sl.y);
          }
      } else {
        sl = v.istat0[f];
      }
      varL = 75 * sl.y - sl.x * sl.x;
      bad = varL == 0;
    }
    if (!bad) {
      for (int im = lo + gl; im <= hi; im += G) {
        const long ft = (long)y * W + im;
        // narrow ranges (G < 32): the mask byte and the window rows are requested together (lo >= 2, hi <= W-3: the window is
        // inside the image), one memory round trip per candidate.  Wide ranges mostly run over unmasked columns beyond the
        // object: there the mask is tested first and the window of an unmasked column is never fetched.
        const unsigned char mk = v.mask1[ft];
        if (G == 32 && mk != 255) continue;
        int2 sr = make_int2(0, 0);
        if (!ONFLY) sr = v.istat1[ft];
        unsigned q[5][4];
#pragma unroll
        for (int r = 0; r < 5; r++) load_row_words<4>(v.img1, ((long)(y - 2 + r) * W + (im - 2)) * 3, q[r]);
        if (mk != 255) continue;
        unsigned slr = 0;
#pragma unroll
        for (int r = 0; r < 5; r++) {
#pragma unroll
          for (int i = 0; i < 4; i++) slr = __dp4a(L[r][i], q[r][i], slr);
          if (ONFLY) {
            q[r][3] &= 0x00ffffffu;
#pragma unroll
            for (int i = 0; i < 4; i++) {
              sr.x = (int)__dp4a(q[r][i], 0x01010101u, (unsigned)sr.x);
              sr.y = (int)__dp4a(q[r][i], q[r][i], (unsigned)sr.y);
            }
          }
        }
        const int num = 75 * (int)slr - sl.x * sr.x;
        const int varR = 75 * sr.y - sr.x * sr.x;
        const float fn = (float)num;
        const float key = varR == 0 ? 0.0f : fn * fabsf(fn) / (float)varR;
        nvalid++;
        if (key > best) { second = best; best = key; bi = im; }
        else if (key > second) second = key;
      }
    }
    __syncwarp();
#pragma unroll
    for (int o = G / 2; o; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o, G), os = __shfl_xor_sync(0xffffffffu, second, o, G);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o, G);
      nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o, G);
      const float lo2 = fminf(best, ob);
      if (ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
      second = fmaxf(fmaxf(second, os), lo2);
    }
    if (gl == 0 && have) {
      if (!bad && nvalid == 0) {
        // no masked candidate in range: the pixel keeps its value (NOMATCH)
      } else if (!bad && best >= 1.0e-6f * (float)varL && (nvalid == 1 || best - second > 1.0e-4f * best)) {
        disp[f] = (short)((unsigned short)bi - x);  // ushort temp_i; short(temp_i - x) (:271,:302)
      } else if (!(bad && hi < lo)) {
        const unsigned o = atomicAdd(n_out, 1u);
        if (o < cap) out_list[o] = (unsigned)f;
      }
    }
  }
}

// list of the masked pixels of the source margin rectangle (input of the full-range search of the lowest level)
__global__ void __launch_bounds__(256) k_list_masked(const uint8_t* __restrict__ mask0, int W, Bound ms, unsigned* __restrict__ list,
                                                     unsigned* __restrict__ n_list, unsigned cap) {
  const int x = ms.XL + blockIdx.x * 256 + threadIdx.x, y = ms.YL + blockIdx.y;
  if (x > ms.XR) return;
  const long f = (long)y * W + x;
  if (mask0[f] != 255) return;
  const unsigned e = atomicAdd(n_list, 1u);
  if (e < cap) list[e] = (unsigned)f;
}

// (atomics: the two matching directions of a level may run on two streams and share the totals)
__global__ void k_count_add(const unsigned* __restrict__ n, unsigned long long* __restrict__ total) { atomicAdd(total, (unsigned long long)*n); }
__global__ void k_count_add2(const unsigned* __restrict__ n3, unsigned long long* __restrict__ total2) {
  atomicAdd(total2, (unsigned long long)n3[0] + n3[2]);
  atomicAdd(total2 + 1, (unsigned long long)n3[1]);
}

template <int G, int MODE>
static int search_dispatch(const PairViews& v, Bound ms, int R, int lo, int hi, const short* lo_map, const short* hi_map,
                           short* disp, const SearchScratch* sc, cudaStream_t st) {
  if (ms.width <= 0 || ms.height <= 0) return 0;
  dim3 grid((ms.width + 255) / 256, ms.height);
  if (R == 2 && sc) {  // screen first (narrow ranges per thread, then any range per warp), exact arithmetic for what is left
    cudaMemsetAsync(sc->n_list, 0, 2 * sizeof(unsigned), st);
    if (MODE == SEARCH_LOWEST) {
      k_list_masked<<<dim3((ms.width + 255) / 256, ms.height), 256, 0, st>>>(v.mask0, v.W, ms, sc->list, sc->n_list, sc->cap);
    } else {
      dim3 gs((ms.width + 127) / 128, ms.height);
      k_ncc_screen5<MODE><<<gs, 128, 0, st>>>(v, ms, lo_map, hi_map, disp, sc->list, sc->n_list, sc->cap);
    }
    k_ncc_screen_wide<MODE, false, (MODE == SEARCH_LOWEST ? 32 : 8)><<<148 * 8, 256, 0, st>>>(v, sc->list, sc->n_list, sc->cap, lo_map, hi_map, lo, hi, disp, sc->list2, sc->n_list + 1);
    k_ncc_search_list<5, 32><<<148 * 4, 256, 0, st>>>(v, sc->list2, sc->n_list + 1, sc->cap, lo_map, hi_map, lo, hi, disp);
    k_count_add<<<1, 1, 0, st>>>(sc->n_list + 1, sc->counters + 1);
    return 4;
  }
  if (R == 2) k_ncc_search<5, G, MODE><<<grid, 256, 0, st>>>(v, ms, lo, hi, lo_map, hi_map, disp);
  else if (R == 1) k_ncc_search<3, G, MODE><<<grid, 256, 0, st>>>(v, ms, lo, hi, lo_map, hi_map, disp);
  else return -1;
  return 1;
}

// ------------------------------------------------------------------------------------------------
// K2  LowestLevelInitialMatch (:170-227): full-range search over [XL1, XR1], one warp per pixel.
// ------------------------------------------------------------------------------------------------
int launch_lowest_match(const PairViews& v, Bound ms, Bound mt, int R, short* out, const SearchScratch* sc, cudaStream_t st) {
  int n = launch_fill_s16(out, (long)v.W * v.H, (short)SB_NOMATCH, st);
  n += search_dispatch<32, SEARCH_LOWEST>(v, ms, R, mt.XL, mt.XR, nullptr, nullptr, out, sc, st);
  return n;
}

// ------------------------------------------------------------------------------------------------
// K3 ranges (:259-288).  Sequential-in-row state of the reference (quirk Q3): boundary_L/R start at
// [XL1, XR1] per row and are only updated at masked pixels:
//   coarse sample s valid : L = max(x + int(2s+.5) - off, XL1), R = min(x + int(2s+.5) + off, XR1)
//   coarse sample NOMATCH : L carried; R = min(i + int(2 s[i]) + off + 1, XR1) for the first valid
//                           coarse index i > t2 (coarse index, as written), else carried.
// Both are last-writer scans; one warp walks a row in 32-pixel chunks with ballot + shuffle.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_high_ranges(const uint8_t* __restrict__ mask0, int W, Bound ms, Bound mt, int off,
                                                     const double* __restrict__ prev, int pw,
                                                     short* __restrict__ lo_map, short* __restrict__ hi_map) {
  extern __shared__ short s_next[];  // [warps][pw]: first valid coarse index > i, or -1
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = ms.YL + blockIdx.x * (blockDim.x >> 5) + warp;
  if (y > ms.YR) return;
  short* nv = s_next + warp * pw;
  const double* s = prev + (size_t)((y + 1) >> 1) * pw;  // int((y+1)/2.0), y >= 0
  const int imax = ms.XR >> 1;                           // look-ahead stops at XR>>1 (:275)
  {
    int carry = -1;
    for (int base = ((pw - 1) >> 5) << 5; base >= 0; base -= 32) {
      const int i = base + lane;
      const bool valid = i < pw && i <= imax && s[i] != (double)SB_NOMATCH;
      const unsigned bal = __ballot_sync(0xffffffffu, valid);
      const unsigned above = lane == 31 ? 0u : (bal & (0xffffffffu << (lane + 1)));
      if (i < pw) nv[i] = (short)(above ? base + __ffs(above) - 1 : carry);
      if (bal) carry = base + __ffs(bal) - 1;
    }
  }
  __syncwarp();
  const int XL1 = mt.XL, XR1 = mt.XR;
  int carryL = XL1, carryR = XR1;
  const uint8_t* p = mask0 + (size_t)y * W;
  for (int base = ms.XL; base <= ms.XR; base += 32) {
    const int x = base + lane;
    const bool masked = x <= ms.XR && p[x] == 255;
    bool hasL = false, hasR = false;
    int valL = 0, valR = 0;
    if (masked) {
      const int t2 = (x + 1) >> 1;
      const double sv = s[t2];
      if (sv != (double)SB_NOMATCH) {
        const int c = x + (int)(sv * 2 + 0.5);
        hasL = hasR = true;
        valL = sb_imax(c - off, XL1);
        valR = sb_imin(c + off, XR1);
      } else {
        const int i = t2 < pw ? nv[t2] : -1;
        if (i >= 0) { hasR = true; valR = sb_imin(i + (int)(s[i] * 2) + off + 1, XR1); }
      }
    }
    const unsigned lower = (1u << lane) - 1;
    const unsigned balL = __ballot_sync(0xffffffffu, hasL), balR = __ballot_sync(0xffffffffu, hasR);
    const unsigned mL = balL & lower, mR = balR & lower;
    const int srcL = mL ? 31 - __clz(mL) : lane, srcR = mR ? 31 - __clz(mR) : lane;
    const int fromL = __shfl_sync(0xffffffffu, valL, srcL), fromR = __shfl_sync(0xffffffffu, valR, srcR);
    const int bL = hasL ? valL : (mL ? fromL : carryL);
    const int bR = hasR ? valR : (mR ? fromR : carryR);
    if (x <= ms.XR) {
      lo_map[(size_t)y * W + x] = (short)(masked ? bL : 1);
      hi_map[(size_t)y * W + x] = (short)(masked ? bR : 0);
    }
    if (balL) carryL = __shfl_sync(0xffffffffu, valL, 31 - __clz(balL));
    if (balR) carryR = __shfl_sync(0xffffffffu, valR, 31 - __clz(balR));
  }
}

int launch_high_match(const PairViews& v, Bound ms, Bound mt, int R, int offset, const double* prev, int pw, int ph,
                      short* lo_scratch, short* hi_scratch, short* out, const SearchScratch* sc, bool band, cudaStream_t st) {
  if (R == 2 && sc && band) {  // TMA / shared-memory band kernel; falls through where the level cannot take it
    const int nb = launch_high_match_band(v, ms, mt, offset, prev, pw, ph, lo_scratch, hi_scratch, out, sc, st);
    if (nb >= 0) return (ms.width <= 0 || ms.height <= 0) ? nb : nb + launch_range_lists(v, lo_scratch, hi_scratch, out, sc, st);
  }
  int n = launch_fill_s16(out, (long)v.W * v.H, (short)SB_NOMATCH, st);
  if (ms.width <= 0 || ms.height <= 0) return n;
  const int warps = 4;
  k_high_ranges<<<(ms.height + warps - 1) / warps, warps * 32, warps * pw * sizeof(short), st>>>(
      v.mask0, v.W, ms, mt, offset, prev, pw, lo_scratch, hi_scratch);
  n += 1;
  n += search_dispatch<8, SEARCH_RANGE_MAPS>(v, ms, R, 0, 0, lo_scratch, hi_scratch, out, sc, st);
  return n;
}

int launch_rematch_search(const PairViews& v, Bound ms, int R, const short* BL, const short* BR, short* disp,
                          const SearchScratch* sc, cudaStream_t st) {
  return search_dispatch<8, SEARCH_REMATCH>(v, ms, R, 0, 0, BL, BR, disp, sc, st);
}

// The two list kernels on a pixel list that is already filled (K3's band path, ncc_band.cu): any-width integer screening with the
// window sums formed on the spot, then the exact pass for what is left (eight lanes per pixel: the ranges are narrow).
int launch_range_lists(const PairViews& v, const short* lo_map, const short* hi_map, short* disp, const SearchScratch* sc, cudaStream_t st) {
  k_ncc_screen_wide<SEARCH_RANGE_MAPS, true, 8><<<148 * 8, 256, 0, st>>>(v, sc->list, sc->n_list, sc->cap, lo_map, hi_map, 0, 0, disp, sc->list2,
                                                                       sc->n_list + 1);
  k_ncc_screen_wide<SEARCH_RANGE_MAPS, true, 32><<<148 * 8, 256, 0, st>>>(v, sc->list_wide, sc->n_list + 2, sc->cap, lo_map, hi_map, 0, 0, disp,
                                                                        sc->list2, sc->n_list + 1);
  k_ncc_search_list<5, 8><<<148 * 2, 256, 0, st>>>(v, sc->list2, sc->n_list + 1, sc->cap, lo_map, hi_map, 0, 0, disp);
  k_count_add2<<<1, 1, 0, st>>>(sc->n_list, sc->counters);
  return 4;
}
