// rematch.cu — K7a SetBoundary_smooth (CStereoMatching.cpp:817-942): per-pixel search bounds for
// Rematch, propagated from matched neighbours by four directional sweeps.  Constants and quirks are
// exactly as written (Q6):
//   down / up   : step MAX_DISPARITY; a matched pixel overwrites its own row's bounds (:863,:895)
//   right       : BL uses -1, BR uses +MAX_DISPARITY (:913-914)
//   left        : bounds become absolute (+= x) and are clamped to the target margin, then BL uses
//                 -MAX_DISPARITY, BR +1 (:921-928); at x == XL the BR clamp writes BL (:938-939)
//
// Every sweep carries one value along a column / row through per-pixel updates of the form
//     v' = max(v - a, c)        (BL;  BR is the mirror  v' = min(v + a, c))
// with a = step or "infinite" (the update ignores v: matched pixel, unmasked pixel).  These
// functions are closed under composition — (a1,c1) then (a2,c2) = (a1+a2, max(c1-a2, c2)) — so the
// sequential sweeps of the reference become scans:
//   vertical   : one block owns 32 columns x all rows; 32 warps take 32 row segments, summarise
//                their segment as one (a,c), exchange summaries through shared memory, then walk
//                the segment again with the right carry-in (down, then the same for up).
//   horizontal : one warp per row, 32-pixel chunks, warp-shuffle scan of (a,c), carry across chunks
//                (right, then left).
// All integer max/min/add: bit-exact with the sequential loops (checked against the oracle's BL/BR
// dumps at every level, tests/test_gpu_parity.py).
#include "kernels.h"

#define FN_INF (1 << 20)   // "infinite" step: the update ignores its input
#define FN_BIG (1 << 28)   // |c| bound meaning "no constant"

struct Fn { int a, c; };
// max-type: f(v) = max(v - a, c)
__device__ __forceinline__ Fn fmax_then(Fn f, Fn g) {  // apply f first, then g
  Fn r;
  r.a = min(f.a + g.a, FN_INF);
  r.c = max(max(f.c - g.a, g.c), -FN_BIG);
  return r;
}
__device__ __forceinline__ int fmax_eval(Fn f, int v) { return max(v - f.a, f.c); }
// min-type: f(v) = min(v + a, c)
__device__ __forceinline__ Fn fmin_then(Fn f, Fn g) {
  Fn r;
  r.a = min(f.a + g.a, FN_INF);
  r.c = min(min(f.c + g.a, g.c), FN_BIG);
  return r;
}
__device__ __forceinline__ int fmin_eval(Fn f, int v) { return min(v + f.a, f.c); }

// ------------------------------------------------------------------------------------------------
// vertical sweeps.  Block = 32 columns (lanes) x 32 row segments (warps).  Writes BL/BR for every
// row of [YL, YR] in its columns; columns outside [XL, XR] and rows outside [YL, YR] are filled
// with the initial (-10000, 10000) by the same kernel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_bounds_vertical(const short* __restrict__ disp, const uint8_t* __restrict__ mask,
                                                          int W, int H, Bound ms, short* __restrict__ BL, short* __restrict__ BR) {
  __shared__ Fn s_l[32][33], s_r[32][33];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + lane;
  const bool col_ok = x < W;
  const bool in_rect = col_ok && x >= ms.XL && x <= ms.XR;
  const int YL = ms.YL, YR = ms.YR;
  // rows outside the rectangle and columns outside it keep the initial value
  if (col_ok) {
    for (int y = seg; y < H; y += 32)
      if (!in_rect || y < YL || y > YR) { BL[(size_t)y * W + x] = (short)-10000; BR[(size_t)y * W + x] = (short)10000; }
  }
  const int rows = YR - YL + 1;
  const int SL = (rows + 31) / 32;
  const int y0 = YL + seg * SL, y1 = min(y0 + SL - 1, YR);  // this thread's segment (may be empty)

  // ---- down, phase 1: segment summary of "value pushed into the next row" -----------------------
  // row y (y <= YR-1 only: row YR is not processed by the down sweep, :842) :
  //   matched: const r-2 / r+2;  masked NOMATCH: max(v-2,-10000) / min(v+2,10000);  unmasked: const -10000 / 10000
  Fn fl = {0, -FN_BIG}, fr = {0, FN_BIG};
  if (in_rect)
    for (int y = y0; y <= y1 && y <= YR - 1; y++) {
      const size_t f = (size_t)y * W + x;
      Fn gl, gr;
      if (mask[f] == 255) {
        const int r = disp[f];
        if (r != SB_NOMATCH) { gl = {FN_INF, r - SB_MAX_DISPARITY}; gr = {FN_INF, r + SB_MAX_DISPARITY}; }
        else { gl = {SB_MAX_DISPARITY, -10000}; gr = {SB_MAX_DISPARITY, 10000}; }
      } else { gl = {FN_INF, -10000}; gr = {FN_INF, 10000}; }
      fl = fmax_then(fl, gl);
      fr = fmin_then(fr, gr);
    }
  s_l[seg][lane] = fl; s_r[seg][lane] = fr;
  __syncthreads();
  // ---- down, phase 2: carry-in = composition of the segments above, applied to the initial value
  int bl = -10000, br = 10000;
  for (int s = 0; s < seg; s++) { bl = fmax_eval(s_l[s][lane], bl); br = fmin_eval(s_r[s][lane], br); }
  __syncthreads();
  // ---- down, phase 3: walk the segment, write D; accumulate the UP summary of the segment ---------
  // up (:874-901), row y >= YL+1: cur = max(D[y], recv); matched: push const r-2; masked NOMATCH: push
  // cur-2 = max(recv-2, D[y]-2); unmasked: push nothing.  Composition order is high y first, so while
  // walking upward in y the new function is applied FIRST: total = g_y then total.
  Fn ul = {0, -FN_BIG}, ur = {0, FN_BIG};
  if (in_rect)
    for (int y = y0; y <= y1; y++) {
      const size_t f = (size_t)y * W + x;
      const bool m = mask[f] == 255;
      const int r = disp[f];
      int nbl = -10000, nbr = 10000;
      if (y <= YR - 1 && m) {
        if (r != SB_NOMATCH) { bl = r; br = r; }
        nbl = max(bl - SB_MAX_DISPARITY, -10000);
        nbr = min(br + SB_MAX_DISPARITY, 10000);
      }
      BL[f] = (short)bl; BR[f] = (short)br;  // D[y]
      if (y >= YL + 1) {
        Fn gl, gr;
        if (m) {
          if (r != SB_NOMATCH) { gl = {FN_INF, r - SB_MAX_DISPARITY}; gr = {FN_INF, r + SB_MAX_DISPARITY}; }
          else { gl = {SB_MAX_DISPARITY, bl - SB_MAX_DISPARITY}; gr = {SB_MAX_DISPARITY, br + SB_MAX_DISPARITY}; }
        } else { gl = {FN_INF, -FN_BIG}; gr = {FN_INF, FN_BIG}; }
        ul = fmax_then(gl, ul);
        ur = fmin_then(gr, ur);
      }
      bl = nbl; br = nbr;
    }
  s_l[seg][lane] = ul; s_r[seg][lane] = ur;
  __syncthreads();
  // ---- up, phase 2: carry-in from the segments below (higher y), nothing pushed into row YR -------
  int rl = -FN_BIG, rr = FN_BIG;
  for (int s = 31; s > seg; s--) { rl = fmax_eval(s_l[s][lane], rl); rr = fmin_eval(s_r[s][lane], rr); }
  // ---- up, phase 3: walk the segment downward in y, write the final vertical result ---------------
  if (in_rect)
    for (int y = y1; y >= y0; y--) {
      const size_t f = (size_t)y * W + x;
      bl = max((int)BL[f], rl);
      br = min((int)BR[f], rr);
      rl = -FN_BIG; rr = FN_BIG;
      if (y >= YL + 1 && mask[f] == 255) {
        const int r = disp[f];
        if (r != SB_NOMATCH) { bl = r; br = r; }
        rl = bl - SB_MAX_DISPARITY;
        rr = br + SB_MAX_DISPARITY;
      }
      BL[f] = (short)bl; BR[f] = (short)br;
    }
}

// ------------------------------------------------------------------------------------------------
// horizontal sweeps, one warp per row.
//   right (:907-917): R[t] = t > XL && masked(t-1) ? max(R[t-1] - 1, V[t]) : V[t]       (V = vertical result)
//                     BR:                              min(R[t-1] + 2, V[t])
//   left  (:918-940): cl(t) = value of bl[t] when the sweep reaches t;  cl(XR) = R[XR];
//                     cl(t) = masked(t+1) ? max(cl(t+1) - 2, XL1-(t+1)-2, R[t]) : R[t]
//                     BR: cr(t) = masked(t+1) ? min(cr(t+1) + 1, XR1-(t+1)+1, R[t]) : R[t]
//                     out[t] = masked(t) ? clamp(cl(t) + t) : cl(t)          (t == XL: the sic clamp)
// Both are scans of (a,c) updates whose first element ignores its input, so the scanned c is the value.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Fn shfl_up_fn(Fn f, int o) {
  Fn r;
  r.a = __shfl_up_sync(0xffffffffu, f.a, o);
  r.c = __shfl_up_sync(0xffffffffu, f.c, o);
  return r;
}
__device__ __forceinline__ Fn shfl_down_fn(Fn f, int o) {
  Fn r;
  r.a = __shfl_down_sync(0xffffffffu, f.a, o);
  r.c = __shfl_down_sync(0xffffffffu, f.c, o);
  return r;
}

__global__ void __launch_bounds__(128) k_bounds_horizontal(const uint8_t* __restrict__ mask, int W, Bound ms, Bound mt,
                                                           short* __restrict__ BL, short* __restrict__ BR) {
  const int lane = threadIdx.x & 31;
  const int y = ms.YL + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (y > ms.YR) return;
  const int XL = ms.XL, XR = ms.XR, XL1 = mt.XL, XR1 = mt.XR;
  short* bl = BL + (size_t)y * W;
  short* br = BR + (size_t)y * W;
  const uint8_t* mp = mask + (size_t)y * W;
  // ---- right: inclusive scan over t = XL..XR, low to high ---------------------------------------
  {
    Fn cl = {FN_INF, -FN_BIG}, cr = {FN_INF, FN_BIG};  // carry: composition of everything to the left
    for (int base = XL; base <= XR; base += 32) {
      const int t = base + lane;
      const bool ok = t <= XR;
      Fn fl = {0, -FN_BIG}, fr = {0, FN_BIG};  // identity for lanes past the end
      if (ok) {
        const bool prop = t > XL && mp[t - 1] == 255;
        fl = {prop ? 1 : FN_INF, (int)bl[t]};
        fr = {prop ? SB_MAX_DISPARITY : FN_INF, (int)br[t]};
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const Fn pl = shfl_up_fn(fl, o), pr = shfl_up_fn(fr, o);
        if (lane >= o) { fl = fmax_then(pl, fl); fr = fmin_then(pr, fr); }
      }
      fl = fmax_then(cl, fl);
      fr = fmin_then(cr, fr);
      if (ok) { bl[t] = (short)fl.c; br[t] = (short)fr.c; }
      cl.a = __shfl_sync(0xffffffffu, fl.a, 31); cl.c = __shfl_sync(0xffffffffu, fl.c, 31);
      cr.a = __shfl_sync(0xffffffffu, fr.a, 31); cr.c = __shfl_sync(0xffffffffu, fr.c, 31);
    }
  }
  __syncwarp();
  // ---- left: inclusive scan over t = XR..XL, high to low ----------------------------------------
  {
    Fn cl = {FN_INF, -FN_BIG}, cr = {FN_INF, FN_BIG};
    const int n = XR - XL + 1;
    for (int done = 0; done < n; done += 32) {
      const int t = XR - done - (31 - lane);  // lane 31 holds the highest t of the chunk
      const bool ok = t >= XL;
      Fn fl = {0, -FN_BIG}, fr = {0, FN_BIG};
      bool masked = false;
      if (ok) {
        masked = mp[t] == 255;
        const bool prop = t < XR && mp[t + 1] == 255;
        const int rl = bl[t], rr = br[t];
        fl = {prop ? SB_MAX_DISPARITY : FN_INF, prop ? max(XL1 - (t + 1) - SB_MAX_DISPARITY, rl) : rl};
        fr = {prop ? 1 : FN_INF, prop ? min(XR1 - (t + 1) + 1, rr) : rr};
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const Fn pl = shfl_down_fn(fl, o), pr = shfl_down_fn(fr, o);
        if (lane + o < 32) { fl = fmax_then(pl, fl); fr = fmin_then(pr, fr); }
      }
      fl = fmax_then(cl, fl);
      fr = fmin_then(cr, fr);
      if (ok) {
        int ol = fl.c, orr = fr.c;
        if (masked) {
          ol = (short)(ol + t); orr = (short)(orr + t);
          if (t > XL) {
            if (ol < XL1) ol = XL1;
            if (orr > XR1) orr = XR1;
          } else {  // :934-940
            if (ol < XL1) ol = XL1;
            if (orr > XR1) ol = XR1;  // sic: the BR clamp assigns BL
          }
        }
        bl[t] = (short)ol; br[t] = (short)orr;
      }
      cl.a = __shfl_sync(0xffffffffu, fl.a, 0); cl.c = __shfl_sync(0xffffffffu, fl.c, 0);
      cr.a = __shfl_sync(0xffffffffu, fr.a, 0); cr.c = __shfl_sync(0xffffffffu, fr.c, 0);
    }
  }
}

int launch_rematch_bounds(const short* disp, const uint8_t* mask, int W, int H, Bound ms, Bound mt, short* BL, short* BR,
                          cudaStream_t st) {
  if (ms.YL >= ms.YR || ms.XL >= ms.XR) return 0;  // the reference exit(0)s here; the caller reports it
  k_bounds_vertical<<<(W + 31) / 32, 1024, 0, st>>>(disp, mask, W, H, ms, BL, BR);
  k_bounds_horizontal<<<(ms.height + 3) / 4, 128, 0, st>>>(mask, W, ms, mt, BL, BR);
  return 2;
}
