// ncc_band.cu — K3, the NCC cost-volume search of HighLevelInitialMatch (CStereoMatching.cpp:231-308), staged through
// shared memory by TMA.
//
// For a masked source pixel whose coarse sample s = prev[(y+1)/2][(x+1)/2] is valid the reference scans the five target
// columns x + int(2s+0.5) + {-2..2} (clamped to the target margin, :286-287) for the largest NCC of the 5x5x3 windows.
// The 2x2 block of pixels {2t-1, 2t} x {2u-1, 2u} shares its coarse sample, hence its five relative disparities, so one
// thread owns such a QUAD: the 6x6 union of its four source windows against a 6x10 strip of the target view gives all
// 20 window correlations from 36 per-pixel products per disparity (dp4a on B,G,R,0 words; the shared 4x4 core and the
// shared edge columns are accumulated once) — 45 integer MACs-of-4 per pixel instead of 125.
//
// A CTA owns a band of 16 scanlines x 128 source columns.  One elected thread fetches the source strip, the target strip
// (placed by the tile's smallest coarse disparity) and both mask strips with cp.async.bulk.tensor into shared memory; the
// 3-byte pixels are expanded once to 4-byte words, split into an even-column and an odd-column plane so that the
// stride-2 accesses of a warp's 32 quads (one scanline pair per warp) are bank-conflict free; the exact integer window
// sums (sum, sum of squares -> variance) of every staged column come from a column-wise running box filter on those
// planes.  Everything the reference compares is decided in exact integers: the arg-max of NCC over a pixel's candidates
// is the arg-max of  key = sign(num) num^2 / varR,  num = 75 SumLR - SumL SumR,  varR = 75 SumR^2 - (SumR)^2.  A pixel is
// settled here only if the winner is unambiguous by a margin (3e-5 relative on key — ten times the float evaluation error
// of the key, six orders of magnitude above the rounding noise of the reference's double evaluation — and |rho| >= 1e-3);
// near ties, flat windows and quads whose target strip left the staged box go to a pixel list and are decided by the
// list kernels of match.cu in the reference's exact arithmetic.  Pixels whose coarse sample is NOMATCH (holes: carried
// bounds, quirk Q3) are ranged by k_hole_ranges below and take the list path too.
#include "kernels.h"
#include "tma.cuh"

namespace {

constexpr int TW = 128;    // source columns per tile
constexpr int TR = 16;     // source rows per tile
constexpr int NT = 256;    // 8 warps: warp w owns the scanline pair (2w, 2w+1) of the tile
constexpr int ROWS = TR + 4;
constexpr int LBOXW = 100;  // u32 words per staged source row: 400 B = image columns xs .. xs+132 (and one byte)
constexpr int RPX = 192;    // staged target columns
constexpr int RBOXW = 144;  // 576 B
constexpr int M0W = 144;    // mask bytes per staged row
constexpr int M1W = 192;
constexpr int LPL = 80;     // words per parity plane of an expanded source row (67 used)
constexpr int RPL = 96;     // ... of an expanded target row.  A multiple of 32: when the coarse disparity steps by one inside a warp,
                            // the lanes on either side of the step read different planes at (nearly) the same word offset

// shared-memory layout (bytes).  The window-sum table of the target strip (SR) overlays the raw staging buffers, which are
// dead once the pixels have been expanded: 69 KB per CTA, three CTAs per SM.
constexpr int OFF_RAWL = 0;
constexpr int OFF_RAWR = 8064;                       // ROWS*400 = 8000, padded (the last 4-pixel group reads 8 bytes past a row)
constexpr int OFF_SR = 0;                            // TR*RPX*8 = 24576 >= OFF_RAWR + ROWS*RBOXW*4 = 19584
constexpr int OFF_M0 = 24576;
constexpr int OFF_M1 = OFF_M0 + TR * M0W;            // 26880
constexpr int OFF_LX = OFF_M1 + TR * M1W;            // 29952
constexpr int OFF_RX = OFF_LX + ROWS * 2 * LPL * 4;  // 42752
constexpr int OFF_SL = OFF_RX + ROWS * 2 * RPL * 4;  // 58112
constexpr int OFF_MISC = OFF_SL + TR * TW * 4;       // 66304
constexpr int SMEM_BYTES = OFF_MISC + 128;
static_assert(OFF_RAWR + ROWS * RBOXW * 4 <= OFF_M0 && TR * RPX * 8 <= OFF_M0, "raw buffers / SR overlay");
constexpr int FAKE_SUM = 110000;  // window sum stored for a target column that is not a candidate: its numerator is always negative

struct BandMaps {
  CUtensorMap img0, img1;    // u32 views of the BGR rows, boxes LBOXW x ROWS / RBOXW x ROWS
  CUtensorMap mask0, mask1;  // u8, boxes M0W x TR / M1W x TR
};

struct BandArgs {
  Bound ms, mt;
  int W, H, pw, ph;
  int xs0, ys0;  // image column of tile (0, .)'s box origin (multiple of 16, may be negative); first quad row (odd)
  const double* prev;
  short* disp;
  short *lo_map, *hi_map;
  unsigned *list, *n_list;
  unsigned cap;
};

// 3-byte pixels -> B,G,R,0 words, even and odd columns in separate planes (pixel p -> plane p & 1, word p >> 1).
// A warp takes the rows warp, warp + 8, ...; its lanes stride the 4-pixel groups of the row.
template <int RAW_WORDS, int GROUPS, int PL>
__device__ __forceinline__ void expand_rows(const unsigned* __restrict__ raw, unsigned* __restrict__ X, int warp, int lane) {
  for (int r = warp; r < ROWS; r += NT / 32) {
    const unsigned* s = raw + r * RAW_WORDS;
    unsigned* d = X + r * 2 * PL;
#pragma unroll
    for (int g = lane; g < GROUPS; g += 32) {
      const unsigned w0 = s[3 * g], w1 = s[3 * g + 1], w2 = s[3 * g + 2];
      const unsigned p0 = w0 & 0x00ffffffu, p1 = __funnelshift_r(w0, w1, 24) & 0x00ffffffu;
      const unsigned p2 = __funnelshift_r(w1, w2, 16) & 0x00ffffffu, p3 = w2 >> 8;
      *reinterpret_cast<uint2*>(d + 2 * g) = make_uint2(p0, p2);
      *reinterpret_cast<uint2*>(d + PL + 2 * g) = make_uint2(p1, p3);
    }
  }
}

// Exact sums over the 5x5 window centred on column c for the tile rows R0+2 .. R0+NR-3: a thread walks down its column with
// the row sums of the last five rows in registers.  emit(window row index - 2 = image row - ys, sum, sum of squares).
template <int PL, int R0, int NR, class Emit>
__device__ __forceinline__ void column_window_sums(const unsigned* __restrict__ X, int c, Emit emit) {
  int a[5];
#pragma unroll
  for (int k = 0; k < 5; k++) a[k] = (((c - 2 + k) & 1) ? PL : 0) + ((c - 2 + k) >> 1);
  int h1[5], h2[5];
  int v1 = 0, v2 = 0;
#pragma unroll
  for (int i = 0; i < NR; i++) {
    const unsigned* row = X + (R0 + i) * 2 * PL;
    int s1 = 0, s2 = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const unsigned px = row[a[k]];
      s1 = (int)__dp4a(px, 0x00010101u, (unsigned)s1);
      s2 = (int)__dp4a(px, px, (unsigned)s2);
    }
    if (i >= 5) { v1 -= h1[i % 5]; v2 -= h2[i % 5]; }
    h1[i % 5] = s1; h2[i % 5] = s2;
    v1 += s1; v2 += s2;
    if (i >= 4) emit(R0 + i - 4, v1, v2);
  }
}

__global__ void __launch_bounds__(NT, 3) k_ncc_band(const __grid_constant__ BandArgs a, const __grid_constant__ BandMaps tm) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned* rawL = reinterpret_cast<unsigned*>(smem + OFF_RAWL);
  unsigned* rawR = reinterpret_cast<unsigned*>(smem + OFF_RAWR);
  const uint8_t* m0 = smem + OFF_M0;
  const uint8_t* m1 = smem + OFF_M1;
  unsigned* LX = reinterpret_cast<unsigned*>(smem + OFF_LX);
  unsigned* RX = reinterpret_cast<unsigned*>(smem + OFF_RX);
  int2* SR = reinterpret_cast<int2*>(smem + OFF_SR);  // [TR][RPX] (sum, float bits of 1/var); overlays the raw buffers
  unsigned* SL = reinterpret_cast<unsigned*>(smem + OFF_SL);  // [TR][TW]  sum | ceil(var / 4096) << 15
  int* misc = reinterpret_cast<int*>(smem + OFF_MISC);  // [0..3] mbarriers (2 x 8 B), [4..11] warp minima, [12..19] warp maxima
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned barL = sbase + OFF_MISC, barR = barL + 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int xs = a.xs0 + (int)blockIdx.x * TW;  // box origin column; source columns of this tile: xs+3 .. xs+130
  const int ys = a.ys0 + (int)blockIdx.y * TR;  // first source row (odd)
  const int XL = a.ms.XL, XR = a.ms.XR, YL = a.ms.YL, YR = a.ms.YR, XL1 = a.mt.XL, XR1 = a.mt.XR;
  const int W = a.W;

  if (tid == 0) {
    mbar_init(barL, 1);
    mbar_init(barR, 1);
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(barL, (unsigned)(ROWS * LBOXW * 4 + TR * M0W));
    tma_load_2d(sbase + OFF_RAWL, &tm.img0, (xs / 16) * 12, ys - 2, barL);  // 3 xs / 4 words; xs is a multiple of 16
    tma_load_2d(sbase + OFF_M0, &tm.mask0, xs, ys, barL);
  }

  // ---- coarse samples of this thread's two quads (quad column q = lane, lane + 32; quad row = warp) ----
  const int y0 = ys + 2 * warp;
  const int u = (y0 + 1) >> 1;
  int D[2];
  bool sval[2];
  int dmin = 0x7fffffff, dmax = -0x7fffffff;
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int x0 = xs + 3 + 2 * (lane + 32 * j);
    const int t = (x0 + 1) >> 1;
    sval[j] = false;
    D[j] = 0;
    if (x0 <= XR && x0 + 1 >= XL && y0 <= YR && y0 + 1 >= YL && t < a.pw && u < a.ph) {
      const double sv = a.prev[(size_t)u * a.pw + t];
      if (sv != (double)SB_NOMATCH) {
        sval[j] = true;
        D[j] = sb_imax(-30000, sb_imin(30000, (int)(sv * 2 + 0.5)));  // :286 (clamped: no integer overflow below on absurd samples)
        dmin = min(dmin, D[j]);
        dmax = max(dmax, D[j]);
      }
    }
  }
  dmin = __reduce_min_sync(0xffffffffu, dmin);
  dmax = __reduce_max_sync(0xffffffffu, dmax);
  if (lane == 0) { misc[4 + warp] = dmin; misc[12 + warp] = dmax; }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; k++) { dmin = min(dmin, misc[4 + k]); dmax = max(dmax, misc[12 + k]); }
  if (dmin > dmax) {  // no quad of this tile has a coarse sample: nothing to match here (holes are k_hole_ranges' business)
    mbar_wait(barL, 0);  // the bulk copies must have landed before the CTA leaves
    return;
  }
  // target box origin: the leftmost strip starts at xs+3 + dmin - 4
  int xr = xs + dmin - 1;
  xr = (xr >= 0 ? xr / 16 : -((-xr + 15) / 16)) * 16;
  if (tid == 0) {
    mbar_expect_tx(barR, (unsigned)(ROWS * RBOXW * 4 + TR * M1W));
    tma_load_2d(sbase + OFF_RAWR, &tm.img1, (xr >= 0 ? xr / 16 : -((-xr) / 16)) * 12, ys - 2, barR);
    tma_load_2d(sbase + OFF_M1, &tm.mask1, xr, ys, barR);
  }
  mbar_wait(barL, 0);
  expand_rows<LBOXW, 34, LPL>(rawL, LX, warp, lane);
  mbar_wait(barR, 0);
  expand_rows<RBOXW, 48, RPL>(rawR, RX, warp, lane);
  __syncthreads();

  // ---- window sums: target columns 2..189 of the box (whole columns), source columns 3..130 (two half columns each) ----
  // (SR overlays the raw buffers: every warp is past its expansion reads here.)  Lanes take columns of one parity, so every
  // tap of a warp reads one plane at consecutive words.  188 + 256 tasks: the first 188 threads walk 20 rows of a target
  // column while the others take a 12-row half of a source column; the remaining halves follow on all threads.
  auto emitR = [&](int c) {
    const bool col_ok = xr + c >= XL1 && xr + c <= XR1;
    int2* dst = SR + ((c & 1) ? RPX / 2 : 0) + (c >> 1);
    column_window_sums<RPL, 0, ROWS>(RX, c, [&](int row, int s1, int s2) {
      const int var = 75 * s2 - s1 * s1;
      float rv;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rv) : "f"((float)var));
      if (var == 0) rv = 0.0f;
      const bool cand = col_ok && m1[row * M1W + c] == 255;
      // not a candidate: a sum that makes the numerator negative whatever the source window is (75 SumLR <= 19125 SumL)
      dst[row * RPX] = cand ? make_int2(s1, __float_as_int(rv)) : make_int2(FAKE_SUM, __float_as_int(1.0e-12f));
    });
  };
  auto emitL = [&](int h) {  // h in [0, 256): column parity-major, then half
    const int half = h >> 7, t = h & 127;
    const int c = 3 + ((t < 64) ? 2 * t + 1 : 2 * (t - 64));  // 4, 6, .., 130 | 3, 5, .., 129
    auto put = [&](int row, int s1, int s2) {
      const unsigned var = (unsigned)(75 * s2 - s1 * s1);
      SL[row * TW + c - 3] = (unsigned)s1 | (((var + 4095u) >> 12) << 15);
    };
    if (half == 0) column_window_sums<LPL, 0, 12>(LX, c, put);
    else column_window_sums<LPL, 8, 12>(LX, c, put);
  };
  if (tid < 188) emitR(tid < 94 ? 2 + 2 * tid : 3 + 2 * (tid - 94));
  else emitL(tid - 188);
  if (tid + 68 < 256) emitL(tid + 68);
  __syncthreads();

  // ---- the quads ----
#pragma unroll 1
  for (int j = 0; j < 2; j++) {
    const int q = lane + 32 * j;
    const int x0 = xs + 3 + 2 * q;
    const int Dq = j ? D[1] : D[0];
    const bool sv = j ? sval[1] : sval[0];
    unsigned act = 0;  // bit p = pixel (x0 + (p & 1), y0 + (p >> 1)) is a masked pixel of the source margin
    if (sv) {
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const int x = x0 + (p & 1), y = y0 + (p >> 1);
        if (x >= XL && x <= XR && y >= YL && y <= YR && m0[(2 * warp + (p >> 1)) * M0W + (x - xs)] == 255) act |= 1u << p;
      }
    }
    const int rc0 = x0 + Dq - 4 - xr;  // box column of the strip's first pixel
    unsigned tolist = 0;
    int bestj[4] = {0, 0, 0, 0};
    if (act && rc0 >= 0 && rc0 + 9 < RPX) {
      const unsigned* Lp = LX + (2 * warp) * 2 * LPL + q;
      const int b0 = ((rc0 & 1) ? RPL : 0) + (rc0 >> 1), b1 = (((rc0 + 1) & 1) ? RPL : 0) + ((rc0 + 1) >> 1);
      const unsigned* Rp = RX + (2 * warp) * 2 * RPL;
      int P00[5], P10[5], P01[5], P11[5], MM[5], A0[5], A5[5];
#pragma unroll
      for (int e = 0; e < 5; e++) { MM[e] = 0; A0[e] = 0; A5[e] = 0; }
#pragma unroll
      for (int r = 0; r < 6; r++) {
        unsigned L[6], Rr[10];
        const unsigned* lr = Lp + r * 2 * LPL;
        const unsigned* rr = Rp + r * 2 * RPL;
#pragma unroll
        for (int k = 0; k < 6; k++) L[k] = (k & 1) ? lr[(k + 1) / 2] : lr[LPL + k / 2];
#pragma unroll
        for (int k = 0; k < 10; k++) Rr[k] = (k & 1) ? rr[b1 + (k - 1) / 2] : rr[b0 + k / 2];
#pragma unroll
        for (int e = 0; e < 5; e++) {
          if (r == 0 || r == 5) {
            int m = 0;
#pragma unroll
            for (int k = 1; k <= 4; k++) m = (int)__dp4a(L[k], Rr[k + e], (unsigned)m);
            const int pa = (int)__dp4a(L[0], Rr[e], (unsigned)m), pb = (int)__dp4a(L[5], Rr[5 + e], (unsigned)m);
            if (r == 0) { P00[e] = pa; P10[e] = pb; } else { P01[e] = pa; P11[e] = pb; }
          } else {
#pragma unroll
            for (int k = 1; k <= 4; k++) MM[e] = (int)__dp4a(L[k], Rr[k + e], (unsigned)MM[e]);
            A0[e] = (int)__dp4a(L[0], Rr[e], (unsigned)A0[e]);
            A5[e] = (int)__dp4a(L[5], Rr[5 + e], (unsigned)A5[e]);
          }
        }
      }
      // ---- keys: exact integer numerators, float ratio with the candidate index in the three low mantissa bits ----
      int2 E[2][6];  // window sums of the six target columns rc0+2 .. rc0+7 on the quad's two rows (planes by column parity)
      {
        const int i0 = rc0 + 2;
        const int s0 = ((i0 & 1) ? RPX / 2 : 0) + (i0 >> 1), s1 = (((i0 + 1) & 1) ? RPX / 2 : 0) + ((i0 + 1) >> 1);
#pragma unroll
        for (int dy = 0; dy < 2; dy++)
#pragma unroll
          for (int m = 0; m < 6; m++) E[dy][m] = SR[(2 * warp + dy) * RPX + ((m & 1) ? s1 + (m - 1) / 2 : s0 + m / 2)];
      }
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const int row = 2 * warp + (p >> 1), dx = p & 1;
        const unsigned slp = SL[row * TW + 2 * q + dx];
        const int sumL = (int)(slp & 0x7fffu);
        const float varL = (float)(slp >> 15) * 4096.0f;  // rounded up to a multiple of 4096: the correlation floor below is not sharp
        float best = -3.0e38f, second = -3.0e38f;
#pragma unroll
        for (int e = 0; e < 5; e++) {
          const int S = (p == 0 ? P00[e] + A0[e] : p == 1 ? P10[e] + A5[e] : p == 2 ? P01[e] + A0[e] : P11[e] + A5[e]) + MM[e];
          const int2 st = E[p >> 1][dx + e];
          const int num = 75 * S - sumL * st.x;
          const float fn = (float)num;
          float key = fn * fabsf(fn) * __int_as_float(st.y);
          key = __int_as_float((__float_as_int(key) & ~7) | e);
          second = fmaxf(second, fminf(best, key));
          best = fmaxf(best, key);
        }
        if (act & (1u << p)) {
          // best <= 0 also covers "no candidate at all" (every key then comes from a FAKE_SUM entry): the list kernels sort it out
          const bool settled = best >= 1.0e-6f * varL && varL != 0.0f && best - second > 3.0e-5f * best;
          if (settled) bestj[p] = (__float_as_int(best) & 7) + 1;
          else tolist |= 1u << p;
        }
      }
    } else if (act) {
      tolist = act;  // the strip left the staged box: the list kernels take all four pixels
    }
    // ---- results: settled pixels straight to the map, the rest to the list with their candidate range ----
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const int x = x0 + (p & 1), y = y0 + (p >> 1);
      const size_t f = (size_t)y * W + x;
      if (bestj[p]) a.disp[f] = (short)(Dq - 2 + bestj[p] - 1);  // short(temp_i - x), temp_i = x + D - 2 + j (:302)
      const bool l = (tolist >> p) & 1u;
      const unsigned bal = __ballot_sync(0xffffffffu, l);
      if (bal) {
        unsigned base = 0;
        if (lane == (__ffs(bal) - 1)) base = atomicAdd(a.n_list, (unsigned)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
        if (l) {
          const unsigned e = base + __popc(bal & ((1u << lane) - 1));
          if (e < a.cap) {
            a.list[e] = (unsigned)f;
            a.lo_map[f] = (short)sb_imax(x + Dq - 2, XL1);
            a.hi_map[f] = (short)sb_imin(x + Dq + 2, XR1);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Holes (:259-288, quirk Q3).  boundary_L/R of a row start at [XL1, XR1] and are updated only at masked pixels:
//   coarse sample s valid : L = max(x + int(2s+.5) - off, XL1), R = min(x + int(2s+.5) + off, XR1)
//   coarse sample NOMATCH : L carried; R = min(i + int(2 s[i]) + off + 1, XR1) for the first valid coarse index
//                           i > t2 (coarse index, as written), else carried.
// Both are last-writer scans; one warp walks a row in 32-pixel chunks with ballot + shuffle, four chunks' loads in flight.
// Only the hole pixels need the result (the others are ranged by k_ncc_band itself): those with a non-empty range get it
// written to the range maps and are appended to the pixel list.
// ------------------------------------------------------------------------------------------------
constexpr int HU = 8;  // chunks whose loads are in flight together (the walk is a chain of dependent iterations: latency-bound)
__global__ void __launch_bounds__(128) k_hole_ranges(const uint8_t* __restrict__ mask0, int W, Bound ms, Bound mt, int off,
                                                     const double* __restrict__ prev, int pw, short* __restrict__ lo_map,
                                                     short* __restrict__ hi_map, unsigned* __restrict__ list,
                                                     unsigned* __restrict__ list_wide, unsigned* __restrict__ n_list, unsigned cap) {
  extern __shared__ int s_next[];  // [warps][pw]: (int(2 s[i']) << 14) | i'  for the first valid coarse index i' > i, low bits 0x3fff = none
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int y = ms.YL + blockIdx.x * (blockDim.x >> 5) + warp;
  if (y > ms.YR) return;
  int* nv = s_next + warp * pw;
  const double* s = prev + (size_t)((y + 1) >> 1) * pw;  // int((y+1)/2.0), y >= 0
  const int imax = ms.XR >> 1;                           // look-ahead stops at XR>>1 (:275)
  {
    int carry = 0x3fff;
    for (int base0 = ((pw - 1) >> 5) << 5; base0 >= 0; base0 -= 32 * HU) {
      double sv[HU];
#pragma unroll
      for (int k = 0; k < HU; k++) {
        const int i = base0 - 32 * k + lane;
        sv[k] = (i >= 0 && i < pw && i <= imax) ? s[i] : (double)SB_NOMATCH;
      }
#pragma unroll
      for (int k = 0; k < HU; k++) {
        const int base = base0 - 32 * k;
        if (base < 0) break;
        const int i = base + lane;
        const bool valid = sv[k] != (double)SB_NOMATCH;
        const int mine = valid ? (int)(((unsigned)(int)(sv[k] * 2) << 14) | (unsigned)i) : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        const unsigned above = lane == 31 ? 0u : (bal & (0xffffffffu << (lane + 1)));
        const int src = above ? __ffs(above) - 1 : 0;
        const int from = __shfl_sync(0xffffffffu, mine, src);
        if (i < pw) nv[i] = above ? from : carry;
        if (bal) carry = __shfl_sync(0xffffffffu, mine, __ffs(bal) - 1);
      }
    }
  }
  __syncwarp();
  const int XL1 = mt.XL, XR1 = mt.XR;
  int carryL = XL1, carryR = XR1;
  const uint8_t* p = mask0 + (size_t)y * W;
  for (int base0 = ms.XL; base0 <= ms.XR; base0 += 32 * HU) {
    unsigned char mk[HU];
    double svv[HU];
#pragma unroll
    for (int k = 0; k < HU; k++) {
      const int x = base0 + 32 * k + lane;
      mk[k] = x <= ms.XR ? p[x] : (unsigned char)0;
      svv[k] = x <= ms.XR ? s[(x + 1) >> 1] : (double)SB_NOMATCH;
    }
#pragma unroll
    for (int k = 0; k < HU; k++) {
      const int base = base0 + 32 * k;
      if (base > ms.XR) break;
      const int x = base + lane;
      const bool masked = mk[k] == 255;
      bool hasL = false, hasR = false, hole = false;
      int valL = 0, valR = 0;
      if (masked) {
        const int t2 = (x + 1) >> 1;
        const double sv = svv[k];
        if (sv != (double)SB_NOMATCH) {
          const int c = x + (int)(sv * 2 + 0.5);
          hasL = hasR = true;
          valL = sb_imax(c - off, XL1);
          valR = sb_imin(c + off, XR1);
        } else {
          hole = true;
          const int e = t2 < pw ? nv[t2] : 0x3fff;
          if ((e & 0x3fff) != 0x3fff) { hasR = true; valR = sb_imin((e & 0x3fff) + (e >> 14) + off + 1, XR1); }
        }
      }
      const unsigned lower = (1u << lane) - 1;
      const unsigned balL = __ballot_sync(0xffffffffu, hasL), balR = __ballot_sync(0xffffffffu, hasR);
      const unsigned mL = balL & lower, mR = balR & lower;
      const int srcL = mL ? 31 - __clz(mL) : lane, srcR = mR ? 31 - __clz(mR) : lane;
      const int fromL = __shfl_sync(0xffffffffu, valL, srcL), fromR = __shfl_sync(0xffffffffu, valR, srcR);
      const int bL = hasL ? valL : (mL ? fromL : carryL);
      const int bR = hasR ? valR : (mR ? fromR : carryR);
      // narrow ranges go to the list eight lanes screen, wide ones (carried bounds far apart) to the list a whole warp screens
      const bool put = hole && bR >= bL, wide = put && bR - bL >= 16;
      const unsigned balP = __ballot_sync(0xffffffffu, put);
      if (balP) {
        const unsigned balW = __ballot_sync(0xffffffffu, wide), balN = balP & ~balW;
        unsigned eN = 0, eW = 0;
        if (lane == 0) {
          if (balN) eN = atomicAdd(n_list, (unsigned)__popc(balN));
          if (balW) eW = atomicAdd(n_list + 2, (unsigned)__popc(balW));
        }
        eN = __shfl_sync(0xffffffffu, eN, 0);
        eW = __shfl_sync(0xffffffffu, eW, 0);
        if (put) {
          const unsigned e = wide ? eW + __popc(balW & lower) : eN + __popc(balN & lower);
          if (e < cap) {
            const size_t f = (size_t)y * W + x;
            (wide ? list_wide : list)[e] = (unsigned)f;
            lo_map[f] = (short)bL;
            hi_map[f] = (short)bR;
          }
        }
      }
      if (balL) carryL = __shfl_sync(0xffffffffu, valL, 31 - __clz(balL));
      if (balR) carryR = __shfl_sync(0xffffffffu, valR, 31 - __clz(balR));
    }
  }
}

}  // namespace

// >= 0: launches issued; -2: this level cannot take the band path (row pitch not a multiple of 16 bytes, no tensor-map encoder)
int launch_high_match_band(const PairViews& v, Bound ms, Bound mt, int offset, const double* prev, int pw, int ph, short* lo_map,
                           short* hi_map, short* out, const SearchScratch* sc, cudaStream_t st) {
  if (offset != 2 || v.W % 16 != 0 || !sc || pw > 8192) return -2;
  BandMaps tm;
  const long words = (long)v.W * 3 / 4;
  if (!sb_tma_encode_2d(&tm.img0, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, v.img0, words, v.H, (size_t)v.W * 3, LBOXW, ROWS) ||
      !sb_tma_encode_2d(&tm.img1, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, v.img1, words, v.H, (size_t)v.W * 3, RBOXW, ROWS) ||
      !sb_tma_encode_2d(&tm.mask0, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, v.mask0, v.W, v.H, (size_t)v.W, M0W, TR) ||
      !sb_tma_encode_2d(&tm.mask1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, v.mask1, v.W, v.H, (size_t)v.W, M1W, TR))
    return -2;
  int n = launch_fill_s16(out, (long)v.W * v.H, (short)SB_NOMATCH, st);
  if (ms.width <= 0 || ms.height <= 0) return n;
  cudaMemsetAsync(sc->n_list, 0, 4 * sizeof(unsigned), st);
  {
    // holes on the side stream, concurrently with the band kernel (both append to the same list; the range maps they write
    // are disjoint): the walk along a scanline is latency-bound and takes few SM resources
    cudaStream_t hs = sc->side ? sc->side : st;
    if (sc->side) {
      cudaEventRecord(sc->ev_fork, st);
      cudaStreamWaitEvent(sc->side, sc->ev_fork, 0);
    }
    int warps = 4;
    while (warps > 1 && (size_t)warps * pw * sizeof(int) > 48 * 1024) warps >>= 1;
    k_hole_ranges<<<(ms.height + warps - 1) / warps, warps * 32, warps * pw * sizeof(int), hs>>>(v.mask0, v.W, ms, mt, offset, prev, pw, lo_map,
                                                                                             hi_map, sc->list, sc->list_wide, sc->n_list, sc->cap);
    if (sc->side) cudaEventRecord(sc->ev_join, sc->side);
    n++;
  }
  {
    static unsigned long long attr_set = 0;  // per device, idempotent
    int dev = 0;
    cudaGetDevice(&dev);
    if (!((attr_set >> (dev & 63)) & 1ull)) {
      cudaFuncSetAttribute(k_ncc_band, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      __atomic_fetch_or(&attr_set, 1ull << (dev & 63), __ATOMIC_RELAXED);
    }
    BandArgs a;
    a.ms = ms; a.mt = mt; a.W = v.W; a.H = v.H; a.pw = pw; a.ph = ph;
    const int xlo = (ms.XL & 1) ? ms.XL : ms.XL - 1;  // first quad column (odd)
    a.xs0 = ((xlo - 3 + 1024) / 16) * 16 - 1024;       // floor to a multiple of 16 (xlo - 3 >= -2)
    a.ys0 = (ms.YL & 1) ? ms.YL : ms.YL - 1;
    a.prev = prev; a.disp = out; a.lo_map = lo_map; a.hi_map = hi_map;
    a.list = sc->list; a.n_list = sc->n_list; a.cap = sc->cap;
    const int gx = (ms.XR - (a.xs0 + 3) + TW) / TW, gy = (ms.YR - a.ys0 + TR) / TR;
    k_ncc_band<<<dim3(gx, gy), NT, SMEM_BYTES, st>>>(a, tm);
    n++;
  }
  if (sc->side) cudaStreamWaitEvent(st, sc->ev_join, 0);
  return n;
}
