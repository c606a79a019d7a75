// rematch.cu — K7a SetBoundary_smooth (CStereoMatching.cpp:817-942): per-pixel search bounds for
// Rematch, propagated from matched neighbours by four directional sweeps.  The sweeps carry state
// along a column / row, with the constants and the quirks exactly as written (Q6):
//   down / up   : step MAX_DISPARITY; a matched pixel overwrites its own row's bounds (:863,:895)
//   right       : BL uses -1, BR uses +MAX_DISPARITY (:913-914)
//   left        : bounds become absolute (+= x) and are clamped to the target margin, then BL uses
//                 -MAX_DISPARITY, BR +1 (:921-928); at x == XL the BR clamp writes BL (:938-939)
// The carried state lives in registers; the loads that feed it (mask, disparity, the other sweep's
// output) do not depend on it, so they pipeline.  Vertical sweeps: one thread per column
// (coalesced across x).  Horizontal sweeps: one thread per row.
#include "kernels.h"

__global__ void k_fill_bounds(short* __restrict__ BL, short* __restrict__ BR, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { BL[i] = (short)-10000; BR[i] = (short)10000; }
}

__global__ void __launch_bounds__(128) k_bounds_vertical(const short* __restrict__ disp, const uint8_t* __restrict__ mask, int W,
                                                         Bound ms, short* __restrict__ BL, short* __restrict__ BR) {
  const int x = ms.XL + blockIdx.x * blockDim.x + threadIdx.x;
  if (x > ms.XR) return;
  const int YL = ms.YL, YR = ms.YR;
  // ---- down (:842-869): row y pushes into row y+1, which still holds its initial value ---------
  int bl = -10000, br = 10000;  // bounds of the current row as left by the row above
  for (int y = YL; y <= YR - 1; y++) {
    const size_t f = (size_t)y * W + x;
    int nbl = -10000, nbr = 10000;
    if (mask[f] == 255) {
      const int r = disp[f];
      if (r != SB_NOMATCH) { bl = r; br = r; }
      nbl = sb_imax(bl - SB_MAX_DISPARITY, -10000);
      nbr = sb_imin(br + SB_MAX_DISPARITY, 10000);
    }
    BL[f] = (short)bl; BR[f] = (short)br;
    bl = nbl; br = nbr;
  }
  BL[(size_t)YR * W + x] = (short)bl; BR[(size_t)YR * W + x] = (short)br;
  // ---- up (:874-901): row y pushes into row y-1 (max / min with what the down sweep left) ------
  int rbl = -10000, rbr = 10000;  // what the row below pushed
  for (int y = YR; y >= YL + 1; y--) {
    const size_t f = (size_t)y * W + x;
    bl = sb_imax((int)BL[f], rbl);
    br = sb_imin((int)BR[f], rbr);
    rbl = -10000; rbr = 10000;
    if (mask[f] == 255) {
      const int r = disp[f];
      if (r != SB_NOMATCH) { bl = r; br = r; }
      rbl = bl - SB_MAX_DISPARITY;
      rbr = br + SB_MAX_DISPARITY;
    }
    BL[f] = (short)bl; BR[f] = (short)br;
  }
  {
    const size_t f = (size_t)YL * W + x;
    BL[f] = (short)sb_imax((int)BL[f], rbl);
    BR[f] = (short)sb_imin((int)BR[f], rbr);
  }
}

__global__ void __launch_bounds__(64) k_bounds_horizontal(const uint8_t* __restrict__ mask, int W, Bound ms, Bound mt,
                                                          short* __restrict__ BL, short* __restrict__ BR) {
  const int y = ms.YL + blockIdx.x * blockDim.x + threadIdx.x;
  if (y > ms.YR) return;
  const int XL = ms.XL, XR = ms.XR, XL1 = mt.XL, XR1 = mt.XR;
  short* bl = BL + (size_t)y * W;
  short* br = BR + (size_t)y * W;
  const uint8_t* mp = mask + (size_t)y * W;
  // ---- right (:907-917) -------------------------------------------------------------------------
  int cl = bl[XL], cr = br[XL];
  for (int x = XL; x <= XR - 1; x++) {
    int nl = bl[x + 1], nr = br[x + 1];
    if (mp[x] == 255) {
      nl = (short)sb_imax(cl - 1, nl);
      nr = (short)sb_imin(cr + SB_MAX_DISPARITY, nr);
      bl[x + 1] = (short)nl; br[x + 1] = (short)nr;
    }
    cl = nl; cr = nr;
  }
  // ---- left (:918-933) --------------------------------------------------------------------------
  cl = bl[XR]; cr = br[XR];
  for (int x = XR; x >= XL + 1; x--) {
    int nl = bl[x - 1], nr = br[x - 1];
    if (mp[x] == 255) {
      cl = (short)(cl + x); cr = (short)(cr + x);
      if (cl < XL1) cl = XL1;
      if (cr > XR1) cr = XR1;
      bl[x] = (short)cl; br[x] = (short)cr;
      nl = (short)sb_imax(cl - x - SB_MAX_DISPARITY, nl);
      nr = (short)sb_imin(cr - x + 1, nr);
      bl[x - 1] = (short)nl; br[x - 1] = (short)nr;
    }
    cl = nl; cr = nr;
  }
  if (mp[XL] == 255) {  // :934-940
    cl = (short)(cl + XL); cr = (short)(cr + XL);
    if (cl < XL1) cl = XL1;
    if (cr > XR1) cl = XR1;  // sic: the BR clamp assigns BL
    bl[XL] = (short)cl; br[XL] = (short)cr;
  }
}

int launch_rematch_bounds(const short* disp, const uint8_t* mask, int W, int H, Bound ms, Bound mt, short* BL, short* BR,
                          cudaStream_t st) {
  const long n = (long)W * H;
  k_fill_bounds<<<(int)((n + 255) / 256), 256, 0, st>>>(BL, BR, n);
  if (ms.YL >= ms.YR || ms.XL >= ms.XR) return 1;  // the reference exit(0)s here; the caller reports it
  k_bounds_vertical<<<(ms.width + 127) / 128, 128, 0, st>>>(disp, mask, W, ms, BL, BR);
  k_bounds_horizontal<<<(ms.height + 63) / 64, 64, 0, st>>>(mask, W, ms, mt, BL, BR);
  return 3;
}
