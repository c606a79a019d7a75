// TEST INFRASTRUCTURE — CPU oracle ("port") for the stereo-matching hot path.
//
// A restatement, on plain row-major buffers, of the reference algorithm in
//   reconstruction/CStereoMatching.cpp:36-113 (stage order), :170-308 (initial match),
//   :310-497 (constraints), :499-570 + :817-942 (rematch), :572-680 (refine),
//   :682-761 (triangulation), :763-815 (median), :1011-1053 (margin, pyramid)
//   reconstruction/CManageData.cpp:81-90 (WindowToVec)
// and of the Armadillo 4.200 arithmetic those call (two-accumulator even/odd sums:
// arrayops_meat.hpp:902-921, fn_norm.hpp:108-127, op_dot_meat.hpp:36-55; median:
// op_median_meat.hpp:361-377; first-max: op_max_meat.hpp:108-152).
// OpenCV pieces (pyrDown, ellipse erode) are restated from OpenCV's published
// behaviour (third-party, OpenCV 2.4.5, not under /root/reference) and pinned
// against cv2 4.13 golden vectors in tests/golden/.
//
// PINNING: tests/test_oracle_cpu.py checks this file bit-for-bit (s16 and f64)
// against oracle/_ref (the reference's own sources compiled here) and against
// the committed fixtures generated from it (tests/golden/make_golden.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
// the library built from this file.  Build: -O2 -fopenmp -ffp-contract=off.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NOMATCH (-10000)     // CStereoMatching.h:9
#define MAX_DISPARITY 2      // CStereoMatching.cpp:4
#define IMIN(a, b) ((a) > (b) ? (b) : (a))
#define IMAX(a, b) ((a) < (b) ? (b) : (a))

namespace {

struct Boundary { int YL, YR, XL, XR, width, height; };  // CManageData.h:10-14

// ---------------------------------------------------------------- a-W
// arrayops::accumulate (arrayops_meat.hpp:902-921)
inline double accumulate2(const double* s, int n) {
  double a1 = 0, a2 = 0;
  int i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { a1 += s[i]; a2 += s[j]; }
  if (i < n) a1 += s[i];
  return a1 + a2;
}
// arma_vec_norm_2 (fn_norm.hpp:108-127,170)
inline double norm2(const double* s, int n) {
  double a1 = 0, a2 = 0;
  int i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { a1 += s[i] * s[i]; a2 += s[j] * s[j]; }
  if (i < n) a1 += s[i] * s[i];
  return sqrt(a1 + a2);
}
// op_dot::direct_dot_arma (op_dot_meat.hpp:36-55)
inline double dot2(const double* a, const double* b, int n) {
  double v1 = 0, v2 = 0;
  int i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { v1 += a[i] * b[i]; v2 += a[j] * b[j]; }
  if (i < n) v1 += a[i] * b[i];
  return v1 + v2;
}
// CManageData::WindowToVec (CManageData.cpp:81-90): byte-column outer, row inner
inline double window_to_vec(const uint8_t* const* rows, int x, int ws, double* u) {
  int k = 0;
  for (int j = x * 3; j < (ws + x) * 3; j++)
    for (int i = 0; i < ws; i++) u[k++] = rows[i][j];
  const int n = ws * ws * 3;
  const double mean = accumulate2(u, n) / double(n);  // op_mean::direct_mean (op_mean_meat.hpp:77-85)
  for (int i = 0; i < n; i++) u[i] -= mean;
  const double nu = norm2(u, n);
  return nu == 0 ? 1 : nu;
}

struct Level {
  int w, h;
  std::vector<uint8_t> img[2], mask[2];
};

struct Ctx {
  int L, W0, H0, OW, OH, R, offset, refine_iters;
  double ws;
  std::vector<Level> lv;
  Boundary margin[2];
  int dw, dh, elem;  // current disparity size / element size (0,2,8)
  std::vector<short> ds[2];
  std::vector<double> dd[2];
  std::vector<short> BL[2], BR[2];
  double Q[16], Rf[9], Tf[3];
  std::vector<double> pts;
  std::vector<uint8_t> pts_bgr;
  std::vector<int32_t> pts_pix;
};

// ---------------------------------------------------------------- a-11 (OpenCV pyrDown, restated)
inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}
void pyr_down(const uint8_t* src, int W, int H, int cn, uint8_t* dst) {
  const int w = (W + 1) / 2, h = (H + 1) / 2;
#pragma omp parallel for
  for (int y = 0; y < h; y++) {
    std::vector<int> hrow((size_t)5 * w * cn);
    for (int k = 0; k < 5; k++) {
      const uint8_t* s = src + (size_t)reflect101(2 * y - 2 + k, H) * W * cn;
      int* hr = &hrow[(size_t)k * w * cn];
      for (int x = 0; x < w; x++) {
        const int x0 = reflect101(2 * x - 2, W) * cn, x1 = reflect101(2 * x - 1, W) * cn, x2 = 2 * x * cn,
                  x3 = reflect101(2 * x + 1, W) * cn, x4 = reflect101(2 * x + 2, W) * cn;
        for (int c = 0; c < cn; c++) hr[x * cn + c] = s[x0 + c] + 4 * s[x1 + c] + 6 * s[x2 + c] + 4 * s[x3 + c] + s[x4 + c];
      }
    }
    uint8_t* d = dst + (size_t)y * w * cn;
    const size_t st = (size_t)w * cn;
    for (size_t i = 0; i < st; i++)
      d[i] = (uint8_t)((hrow[i] + 4 * hrow[st + i] + 6 * hrow[2 * st + i] + 4 * hrow[3 * st + i] + hrow[4 * st + i] + 128) >> 8);
  }
}

// getStructuringElement(MORPH_ELLIPSE) row extents [j1, j2) (OpenCV, restated)
void ellipse_rows(int ks, std::vector<int>& j1, std::vector<int>& j2) {
  const int r = ks / 2, c = ks / 2;
  const double inv_r2 = r ? 1. / ((double)r * r) : 0;
  j1.assign(ks, 0);
  j2.assign(ks, 0);
  for (int i = 0; i < ks; i++) {
    if (ks == 1) { j2[i] = 1; continue; }
    const int dy = i - r;
    if (abs(dy) <= r) {
      const int dx = (int)nearbyint(c * sqrt((r * r - dy * dy) * inv_r2));
      j1[i] = IMAX(c - dx, 0);
      j2[i] = IMIN(c + dx + 1, ks);
    }
  }
}
// erode with that element; anchor = centre; outside-image pixels do not constrain
void erode_ellipse(const uint8_t* src, int W, int H, int ks, uint8_t* dst) {
  std::vector<int> j1, j2;
  ellipse_rows(ks, j1, j2);
  const int a = ks / 2;
#pragma omp parallel for
  for (int y = 0; y < H; y++) {
    for (int x = 0; x < W; x++) {
      int m = 255;
      for (int i = 0; i < ks && m; i++) {
        const int sy = y + i - a;
        if (sy < 0 || sy >= H || j2[i] <= j1[i]) continue;
        const int xa = IMAX(x + j1[i] - a, 0), xb = IMIN(x + j2[i] - 1 - a, W - 1);
        const uint8_t* s = src + (size_t)sy * W;
        for (int xx = xa; xx <= xb; xx++) m = IMIN(m, (int)s[xx]);
      }
      dst[(size_t)y * W + x] = (uint8_t)m;
    }
  }
}

// ---------------------------------------------------------------- a-1 FindMargin (:1011-1038)
void find_margin(Boundary& m, const uint8_t* mask, int W, int H, int R) {
  m.YL = H - 1 - R; m.YR = R; m.XL = W - 1 - R; m.XR = R;
  for (int y = R; y < H - R; y++) {
    const uint8_t* p = mask + (size_t)y * W;
    bool flag = false;
    for (int x = R; x < W - R; x++) {
      if (p[x] != 255) continue;
      m.XL = IMIN(m.XL, x); m.XR = IMAX(m.XR, x); flag = true;
    }
    if (flag) { m.YL = IMIN(m.YL, y); m.YR = IMAX(m.YR, y); }
  }
  m.width = m.XR - m.XL + 1;
  m.height = m.YR - m.YL + 1;
}

// shared argmax loop of a-2 / a-3 / a-7 (:207-218, :289-300, :551-562); returns -1 if nothing beat -1
inline int ncc_argmax(const uint8_t* const* rowsL, const uint8_t* const* rowsR, const uint8_t* q, int x, int lo, int hi, int R) {
  const int wsz = 2 * R + 1, n = wsz * wsz * 3;
  double vecL[147], vecR[147];
  const double normL = window_to_vec(rowsL, x - R, wsz, vecL);
  for (int i = 0; i < n; i++) vecL[i] /= normL;
  int best = -1;
  double bestv = -1;
  for (int im = lo; im <= hi; im++) {
    if (q[im] != 255) continue;
    const double normR = window_to_vec(rowsR, im - R, wsz, vecR);
    const double v = dot2(vecL, vecR, n) / normR;
    if (v > bestv) { best = im; bestv = v; }
  }
  return best;
}

struct Views {  // source view = index 0, target view = index 1 (image / image_inv of MatchOneLayer :43-50)
  const uint8_t *img0, *img1, *mask0, *mask1;
  int W, H;
};
Views views(const Ctx* c, int level, bool zeroOne) {
  const Level& l = c->lv[level];
  Views v;
  v.W = l.w; v.H = l.h;
  v.img0 = l.img[zeroOne ? 0 : 1].data(); v.img1 = l.img[zeroOne ? 1 : 0].data();
  v.mask0 = l.mask[zeroOne ? 0 : 1].data(); v.mask1 = l.mask[zeroOne ? 1 : 0].data();
  return v;
}

// ---------------------------------------------------------------- a-2 (:170-227)
void lowest_level_match(Ctx* c, int level, std::vector<short>& disp, bool zeroOne) {
  const Views v = views(c, level, zeroOne);
  const int W = v.W, H = v.H, R = c->R;
  disp.assign((size_t)W * H, (short)NOMATCH);
  const Boundary &ms = c->margin[!zeroOne], &mt = c->margin[zeroOne];
#pragma omp parallel for
  for (int y = ms.YL; y <= ms.YR; y++) {
    const uint8_t *rl[7], *rr[7];
    for (int i = -R; i <= R; i++) { rl[i + R] = v.img0 + (size_t)(y + i) * W * 3; rr[i + R] = v.img1 + (size_t)(y + i) * W * 3; }
    const uint8_t *p = v.mask0 + (size_t)y * W, *q = v.mask1 + (size_t)y * W;
    short* s = &disp[(size_t)y * W];
    for (int x = ms.XL; x <= ms.XR; x++) {
      if (p[x] != 255) continue;
      const int b = ncc_argmax(rl, rr, q, x, mt.XL, mt.XR, R);
      if (b != -1) s[x] = (short)(b - x);
    }
  }
}

// ---------------------------------------------------------------- a-3 (:231-308)
void high_level_match(Ctx* c, int level, const std::vector<double>& prev, int pw, std::vector<short>& out, bool zeroOne) {
  const Views v = views(c, level, zeroOne);
  const int W = v.W, H = v.H, R = c->R, off = c->offset;
  out.assign((size_t)W * H, (short)NOMATCH);
  const Boundary &ms = c->margin[!zeroOne], &mt = c->margin[zeroOne];
  const int XL1 = mt.XL, XR1 = mt.XR;
#pragma omp parallel for
  for (int y = ms.YL; y <= ms.YR; y++) {
    const uint8_t *rl[7], *rr[7];
    for (int i = -R; i <= R; i++) { rl[i + R] = v.img0 + (size_t)(y + i) * W * 3; rr[i + R] = v.img1 + (size_t)(y + i) * W * 3; }
    const uint8_t *p = v.mask0 + (size_t)y * W, *q = v.mask1 + (size_t)y * W;
    short* d = &out[(size_t)y * W];
    const double* s = &prev[(size_t)int((y + 1) / 2.0) * pw];
    int bL = XL1, bR = XR1;  // carried along the row (Q3)
    for (int x = ms.XL; x <= ms.XR; x++) {
      if (p[x] != 255) continue;
      const int t2 = int((x + 1) / 2.0);
      if (s[t2] == NOMATCH) {
        for (int i = t2 + 1; i <= ms.XR >> 1; i++)
          if (s[i] != NOMATCH) { bR = IMIN(i + int(s[i] * 2) + off + 1, XR1); break; }
      } else {
        bL = IMAX(x + int(s[t2] * 2 + 0.5) - off, XL1);
        bR = IMIN(x + int(s[t2] * 2 + 0.5) + off, XR1);
      }
      const int b = ncc_argmax(rl, rr, q, x, bL, bR, R);
      if (b != -1) d[x] = (short)((unsigned short)b - x);  // ushort temp_i; short(temp_i - x) (Q9)
    }
  }
}

// ---------------------------------------------------------------- a-4 (:370-448)
void smooth_constraint(Ctx* c, std::vector<short>& disp, int W, int H, bool zeroOne) {
  const Boundary& m = c->margin[!zeroOne];
  std::vector<uint8_t> tmp((size_t)W * H * 2 + 64, 0);
  for (int y = m.YL; y <= m.YR; y++) {
    const short *pup = &disp[(size_t)y * W], *pdown = &disp[(size_t)(y + 1) * W];
    uint8_t *qup = &tmp[(size_t)y * W * 2], *qdown = &tmp[(size_t)(y + 1) * W * 2];
    for (int x = m.XL; x <= m.XR; x++) {
      if (pup[x] == NOMATCH) continue;
      const int dx = x << 1;
      if (pup[x + 1] != NOMATCH) {  // east
        qup[dx]++; qup[dx + 2]++;
        if (abs(pup[x] - pup[x + 1]) > 1) { qup[dx + 1]++; qup[dx + 3]++; }
      }
      if (pdown[x - 1] != NOMATCH) {  // south-west
        qup[dx]++; qdown[dx - 2]++;
        if (abs(pup[x] - pdown[x - 1]) > 1) { qup[dx + 1]++; qdown[dx - 1]++; }
      }
      if (pdown[x] != NOMATCH) {  // south
        qup[dx]++; qdown[dx]++;
        if (abs(pup[x] - pdown[x]) > 1) { qup[dx + 1]++; qdown[dx + 1]++; }
      }
      if (pdown[x + 1] != NOMATCH) {  // south-east: byte offsets x and x+2, as written (Q4)
        qup[x]++; qdown[x + 2]++;
        if (abs(pup[x] - pdown[x + 1]) > 1) { qup[dx + 1]++; qdown[dx + 3]++; }
      }
    }
  }
#pragma omp parallel for
  for (int y = m.YL; y <= m.YR; y++) {
    const uint8_t* pc = &tmp[(size_t)y * W * 2 + (m.XL << 1)];
    short* pd = &disp[(size_t)y * W];
    for (int x = m.XL; x <= m.XR; x++) {
      if (pc[0] == 0 || (pc[1] << 1) > pc[0]) pd[x] = NOMATCH;
      pc += 2;
    }
  }
}

// ---------------------------------------------------------------- a-5 (:310-368)
void order_constraint(Ctx* c, std::vector<short>& disp, int W, int H, bool zeroOne) {
  const Boundary& m = c->margin[!zeroOne];
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = m.YL; y <= m.YR; y++) {
    short* p = &disp[(size_t)y * W];
    std::vector<short> line, idx;
    for (int x = m.XL; x <= m.XR; x++) {
      if (p[x] == NOMATCH) continue;
      line.push_back((short)(p[x] + x));
      idx.push_back((short)x);
    }
    const int n = (int)line.size();
    // cross(i,j) for j<i  <=>  line[j] > line[i]; count = row sum of the symmetrised matrix
    std::vector<int> cnt(n, 0);
    long ones = 0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < i; j++)
        if (line[j] > line[i]) { cnt[i]++; cnt[j]++; ones++; }
    std::vector<char> dead(n, 0);
    while (ones) {
      int bi = 0, bv = -32768;  // first max, strict '>' in ascending index (op_max_meat.hpp:108-152)
      for (int i = 0; i < n; i++)
        if (cnt[i] > bv) { bv = cnt[i]; bi = i; }
      for (int k = 0; k < n; k++) {
        if (dead[k] || k == bi) continue;
        const bool cross = k < bi ? line[k] > line[bi] : line[bi] > line[k];
        if (cross) cnt[k]--;
      }
      cnt[bi] = 0;
      dead[bi] = 1;
      ones -= bv;
      p[idx[bi]] = NOMATCH;
    }
  }
}

// ---------------------------------------------------------------- a-6 (:450-497)
template <class T> inline double tabs(T v) { return v < 0 ? -v : v; }
template <class T>
void uniqueness_pass(Ctx* c, T* P, const T* Qm, int W, bool zeroOne) {
  const Boundary &ms = c->margin[!zeroOne], &mt = c->margin[zeroOne];
#pragma omp parallel for
  for (int y = ms.YL; y <= ms.YR; y++) {
    T* p = P + (size_t)y * W;
    const T* q = Qm + (size_t)y * W;
    for (int x = ms.XL; x <= ms.XR; x++) {
      if (p[x] == NOMATCH) continue;
      const int bl = IMAX(int(p[x] + 0.5) + x - 1, mt.XL);
      const int br = IMIN(bl + 2, mt.XR);
      int im;
      for (im = bl; im <= br; im++)
        if (tabs(q[im] + p[x]) < 2) break;
      if (im > br) {
        if (tabs(q[bl + 1] + p[x - 1]) >= 2 && tabs(q[bl + 1] + p[x + 1]) >= 2) p[x] = NOMATCH;  // sees in-sweep kills (Q5)
      }
    }
  }
}
template <class T>
void uniqueness(Ctx* c, T* d0, T* d1, int W) {  // :456-460
  uniqueness_pass<T>(c, d0, d1, W, true);
  uniqueness_pass<T>(c, d1, d0, W, false);
  uniqueness_pass<T>(c, d0, d1, W, true);
}

// ---------------------------------------------------------------- a-7 (:817-942, :499-570)
void set_boundary_smooth(Ctx* c, const short* disp, const uint8_t* mask, int W, int H, std::vector<short>& BL, std::vector<short>& BR, bool zeroOne) {
  const Boundary &ms = c->margin[!zeroOne], &mt = c->margin[zeroOne];
  const int YL = ms.YL, YR = ms.YR, XL = ms.XL, XR = ms.XR, XL1 = mt.XL, XR1 = mt.XR;
  BL.assign((size_t)W * H, (short)-10000);
  BR.assign((size_t)W * H, (short)10000);
  if (YL >= YR || XL >= XR) return;  // the reference exit(0)s here (:827-830); callers check margins first
  for (int y = YL; y <= YR - 1; y++) {  // down
    const short* src = disp + (size_t)y * W;
    const uint8_t* mp = mask + (size_t)y * W;
    short *bl0 = &BL[(size_t)y * W], *br0 = &BR[(size_t)y * W], *bl1 = &BL[(size_t)(y + 1) * W], *br1 = &BR[(size_t)(y + 1) * W];
    for (int x = XL; x <= XR; x++) {
      if (mp[x] != 255) continue;
      const short r = src[x];
      if (r == NOMATCH) {
        bl1[x] = (short)IMAX(bl0[x] - MAX_DISPARITY, bl1[x]);
        br1[x] = (short)IMIN(br0[x] + MAX_DISPARITY, br1[x]);
      } else {
        bl0[x] = r; br0[x] = r;
        bl1[x] = (short)IMAX(r - MAX_DISPARITY, bl1[x]);
        br1[x] = (short)IMIN(r + MAX_DISPARITY, br1[x]);
      }
    }
  }
  for (int y = YR; y >= YL + 1; y--) {  // up
    const short* src = disp + (size_t)y * W;
    const uint8_t* mp = mask + (size_t)y * W;
    short *bl0 = &BL[(size_t)y * W], *br0 = &BR[(size_t)y * W], *bl1 = &BL[(size_t)(y - 1) * W], *br1 = &BR[(size_t)(y - 1) * W];
    for (int x = XL; x <= XR; x++) {
      if (mp[x] != 255) continue;
      const short r = src[x];
      if (r == NOMATCH) {
        bl1[x] = (short)IMAX(bl0[x] - MAX_DISPARITY, bl1[x]);
        br1[x] = (short)IMIN(br0[x] + MAX_DISPARITY, br1[x]);
      } else {
        bl0[x] = r; br0[x] = r;
        bl1[x] = (short)IMAX(r - MAX_DISPARITY, bl1[x]);
        br1[x] = (short)IMIN(r + MAX_DISPARITY, br1[x]);
      }
    }
  }
#pragma omp parallel for
  for (int y = YL; y <= YR; y++) {  // left / right, constants as written (Q6)
    short *bl = &BL[(size_t)y * W], *br = &BR[(size_t)y * W];
    const uint8_t* mp = mask + (size_t)y * W;
    for (int x = XL; x <= XR - 1; x++)
      if (mp[x] == 255) {
        bl[x + 1] = (short)IMAX(bl[x] - 1, bl[x + 1]);
        br[x + 1] = (short)IMIN(br[x] + MAX_DISPARITY, br[x + 1]);
      }
    for (int x = XR; x >= XL + 1; x--)
      if (mp[x] == 255) {
        bl[x] += x; br[x] += x;
        if (bl[x] < XL1) bl[x] = XL1;
        if (br[x] > XR1) br[x] = XR1;
        bl[x - 1] = (short)IMAX(bl[x] - x - MAX_DISPARITY, bl[x - 1]);
        br[x - 1] = (short)IMIN(br[x] - x + 1, br[x - 1]);
      }
    if (mp[XL] == 255) {
      bl[XL] += XL; br[XL] += XL;
      if (bl[XL] < XL1) bl[XL] = XL1;
      if (br[XL] > XR1) bl[XL] = XR1;  // sic (:938-939)
    }
  }
}

void rematch(Ctx* c, int level, std::vector<short>& disp, int dir, bool zeroOne) {
  const Views v = views(c, level, zeroOne);
  const int W = v.W, H = v.H, R = c->R;
  const Boundary& ms = c->margin[!zeroOne];
  set_boundary_smooth(c, disp.data(), v.mask0, W, H, c->BL[dir], c->BR[dir], zeroOne);
  const std::vector<short>&BL = c->BL[dir], &BR = c->BR[dir];
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = ms.YL; y <= ms.YR; y++) {
    const uint8_t *rl[7], *rr[7];
    for (int i = -R; i <= R; i++) { rl[i + R] = v.img0 + (size_t)(y + i) * W * 3; rr[i + R] = v.img1 + (size_t)(y + i) * W * 3; }
    const uint8_t *p = v.mask0 + (size_t)y * W, *q = v.mask1 + (size_t)y * W;
    short* s = &disp[(size_t)y * W];
    for (int x = ms.XL; x <= ms.XR; x++) {
      if (p[x] != 255 || s[x] != NOMATCH) continue;
      const int b = ncc_argmax(rl, rr, q, x, BL[(size_t)y * W + x], BR[(size_t)y * W + x], R);
      if (b != -1) s[x] = (short)(b - x);
    }
  }
}

// ---------------------------------------------------------------- a-8 (:763-815)
void median_filter(Ctx* c, std::vector<short>& disp, const uint8_t* mask, int W, int H, bool zeroOne) {
  const Boundary& m = c->margin[!zeroOne];
  std::vector<short> out((size_t)W * H, (short)NOMATCH);
#pragma omp parallel for
  for (int y = m.YL; y <= m.YR; y++) {
    const short* wp[3] = {&disp[(size_t)(y - 1) * W], &disp[(size_t)y * W], &disp[(size_t)(y + 1) * W]};
    const uint8_t* mp = mask + (size_t)y * W;
    short* p = &out[(size_t)y * W];
    for (int x = m.XL; x <= m.XR; x++) {
      if (mp[x] != 255) continue;
      long long u[9];
      int k = 0;
      for (int i = x - 1; i < x + 1; i++)  // two columns only (Q7)
        for (int j = 0; j < 3; j++)
          if (wp[j][i] != NOMATCH) u[k++] = wp[j][i];
      const bool centre_missing = wp[1][x] == NOMATCH;
      if (centre_missing ? (k >= 4) : (k > 2)) {
        std::sort(u, u + k);
        const int half = k / 2;  // op_median::direct_median + robust_mean (lo + (hi-lo)/2)
        p[x] = (short)((k % 2) == 0 ? u[half - 1] + (u[half] - u[half - 1]) / 2 : u[half]);
      } else {
        p[x] = NOMATCH;
      }
    }
  }
  disp.swap(out);
}

// ---------------------------------------------------------------- a-9 (:572-680)
void disparity_refine(Ctx* c, int level, const std::vector<short>& in, std::vector<double>& outd, int iteration, bool zeroOne) {
  const Views v = views(c, level, zeroOne);
  const int W = v.W, H = v.H;
  const Boundary& ms = c->margin[!zeroOne];
  const double ws = c->ws;
  std::vector<double> A((size_t)W * H), B;
  for (size_t i = 0; i < A.size(); i++) A[i] = in[i];
  B = A;
  double *dout = A.data(), *cur = B.data();  // disparity_out / Current_Disparity
  for (int iter = 0; iter < iteration; iter++) {
#pragma omp parallel for
    for (int y = ms.YL + 1; y <= ms.YR - 1; y++) {
      const double *p0 = dout + (size_t)(y - 1) * W, *p1 = dout + (size_t)y * W, *p2 = dout + (size_t)(y + 1) * W;
      double* pc = cur + (size_t)y * W;
      const uint8_t *rl[3], *rr[3];
      for (int i = -1; i <= 1; i++) { rl[i + 1] = v.img0 + (size_t)(y + i) * W * 3; rr[i + 1] = v.img1 + (size_t)(y + i) * W * 3; }
      double vecL[27], vecR[27], xi[3], pdp = 0, pwp = 0;
      for (int x = ms.XL + 1; x <= ms.XR - 1; x++) {
        if (p1[x] == NOMATCH) continue;
        const double dC = p1[x], dE = p1[x + 1], dW = p1[x - 1], dN = p0[x], dS = p2[x];
        const int mode = (dE != NOMATCH && dW != NOMATCH) + (dS != NOMATCH && dN != NOMATCH) * 2;
        if (mode != 0) {
          const double normL = window_to_vec(rl, x - 1, 3, vecL);
          const int im = int(dC - 1.5) + x;
          for (int i = 0; i < 3; i++) {
            const double normR = window_to_vec(rr, im + i, 3, vecR);  // no bounds / mask test (Q8)
            xi[i] = (1 - dot2(vecL, vecR, 27) / (normL * normR)) / 2;
          }
          int index = xi[0] >= xi[1];
          if (xi[index] > xi[2]) index = 2;
          switch (index) {
            case 0: pwp = xi[1] - xi[0]; pdp = dC - 0.5; break;
            case 1:
              pwp = 0.5 * (xi[0] + xi[2]) - xi[1];
              pdp = dC + 0.5 * (xi[0] - xi[2]) / (xi[0] + xi[2] - 2 * xi[1]);
              if (pwp == 0) pdp = 0;
              break;
            case 2: pwp = xi[1] - xi[2]; pdp = dC + 0.5; break;
          }
        }
        switch (mode) {
          case 0: pc[x] = dC; break;
          case 1: pc[x] = (pdp * pwp + ws * (dE + dW) / 2) / (pwp + ws); break;
          case 2: pc[x] = (pdp * pwp + ws * (dN + dS) / 2) / (pwp + ws); break;
          case 3: {
            const double ex = fabs(dE - dC) - fabs(dW - dC), ey = fabs(dS - dC) - fabs(dN - dC);
            const double wx = exp(-(ex * ex)), wy = exp(-(ey * ey));
            double ds;
            if (wx + wy == 0) ds = (dE + dW + dS + dN) / 4;
            else ds = (wx * (dE + dW) + wy * (dN + dS)) / (2 * (wx + wy));
            pc[x] = (pdp * pwp + ws * ds) / (pwp + ws);
          }
        }
      }
    }
    std::swap(dout, cur);  // :675-677
  }
  outd.assign(dout, dout + (size_t)W * H);  // :679
}

// ---------------------------------------------------------------- a-10 (:682-761)
long disparity_to_cloud(Ctx* c) {
  const int level = c->L - 1;
  const Level& l = c->lv[level];
  const int W = l.w, H = l.h;
  const Boundary& m = c->margin[0];  // IsZeroOne = true -> margin[!true]
  const double scale = double(c->W0) / c->OW * (1 << level);
  double q[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) q[i][j] = c->Q[i * 4 + j];
  for (int i = 0; i < 4; i++) q[i][3] *= scale;
  const double qz = q[2][3], qw = q[3][3];
  std::vector<uint8_t> mask((size_t)W * H);
  erode_ellipse(l.mask[0].data(), W, H, (int)ceil(0.02 * H), mask.data());
  c->pts.clear(); c->pts_bgr.clear(); c->pts_pix.clear();
  const double* D = c->dd[0].data();
  const uint8_t* img = l.img[0].data();
  for (int y = m.YL; y <= m.YR; y++) {
    const double qy = y + q[1][3];
    for (int x = m.XL; x <= m.XR; x++) {
      if (mask[(size_t)y * W + x] != 255) continue;
      const double d = D[(size_t)y * W + x];
      if (d == NOMATCH) continue;
      const double iW = 1. / (qw + q[3][2] * d);
      const double F[3] = {(q[0][3] + double(x)) * iW, qy * iW, qz * iW};
      for (int i = 0; i < 3; i++) {  // R_final*Fout + T_final, plain left-to-right sum
        double s = 0;
        for (int k = 0; k < 3; k++) s += c->Rf[i * 3 + k] * F[k];
        c->pts.push_back(s + c->Tf[i]);
      }
      for (int k = 0; k < 3; k++) c->pts_bgr.push_back(img[((size_t)y * W + x) * 3 + k]);
      c->pts_pix.push_back(y * W + x);
    }
  }
  return (long)c->pts_pix.size();
}

void construct_pyramid(Ctx* c) {  // :1040-1053
  for (int id = 0; id < 2; id++)
    for (int i = c->L - 1; i > 0; i--) {
      Level &s = c->lv[i], &d = c->lv[i - 1];
      pyr_down(s.img[id].data(), s.w, s.h, 3, d.img[id].data());
      pyr_down(s.mask[id].data(), s.w, s.h, 1, d.mask[id].data());
    }
}

int run_stage(Ctx* c, int level, int stage) {  // MatchOneLayer :51-109
  const Level& l = c->lv[level];
  const int W = l.w, H = l.h;
  switch (stage) {
    case 1:
      find_margin(c->margin[0], l.mask[0].data(), W, H, c->R);
      find_margin(c->margin[1], l.mask[1].data(), W, H, c->R);
      return 0;
    case 2:
      if (level == 0) {
        lowest_level_match(c, level, c->ds[0], true);
        lowest_level_match(c, level, c->ds[1], false);
      } else {
        std::vector<short> o0, o1;
        high_level_match(c, level, c->dd[0], c->dw, o0, true);
        high_level_match(c, level, c->dd[1], c->dw, o1, false);
        c->ds[0].swap(o0); c->ds[1].swap(o1);
      }
      c->dw = W; c->dh = H; c->elem = 2;
      return 0;
    case 3: smooth_constraint(c, c->ds[0], W, H, true); smooth_constraint(c, c->ds[1], W, H, false); return 0;
    case 4: order_constraint(c, c->ds[0], W, H, true); order_constraint(c, c->ds[1], W, H, false); return 0;
    case 5: case 7: uniqueness<short>(c, c->ds[0].data(), c->ds[1].data(), W); return 0;
    case 6: rematch(c, level, c->ds[0], 0, true); rematch(c, level, c->ds[1], 1, false); return 0;
    case 8:
      median_filter(c, c->ds[0], l.mask[0].data(), W, H, true);
      median_filter(c, c->ds[1], l.mask[1].data(), W, H, false);
      return 0;
    case 9: {
      const int it = c->refine_iters >= 0 ? c->refine_iters : 30 + level * 30;
      disparity_refine(c, level, c->ds[0], c->dd[0], it, true);
      disparity_refine(c, level, c->ds[1], c->dd[1], it, false);
      c->elem = 8;
      return 0;
    }
    case 10: uniqueness<double>(c, c->dd[0].data(), c->dd[1].data(), W); return 0;
  }
  return -1;
}

}  // namespace

extern "C" {

void* orc_create(int pyrm_num, int lowest_w, int lowest_h, int origin_w, int origin_h, int radius, double ws, int offset) {
  Ctx* c = new Ctx();
  c->L = pyrm_num; c->W0 = lowest_w; c->H0 = lowest_h; c->OW = origin_w; c->OH = origin_h;
  c->R = radius; c->ws = ws; c->offset = offset; c->refine_iters = -1;
  c->dw = c->dh = c->elem = 0;
  c->lv.resize(pyrm_num);
  for (int i = 0; i < pyrm_num; i++) {
    Level& l = c->lv[i];
    l.w = lowest_w << i; l.h = lowest_h << i;
    for (int k = 0; k < 2; k++) {
      // slack after the payload: the reference's flat addressing may run a few bytes past a row (Q8)
      l.img[k].assign((size_t)l.w * l.h * 3 + 4096, 0);
      l.mask[k].assign((size_t)l.w * l.h + 4096, 0);
    }
  }
  memset(c->margin, 0, sizeof(c->margin));
  return c;
}
void orc_destroy(void* h) { delete (Ctx*)h; }
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_pair(void* h, const uint8_t* i0, const uint8_t* i1, const uint8_t* m0, const uint8_t* m1) {
  Ctx* c = (Ctx*)h;
  Level& t = c->lv[c->L - 1];
  memcpy(t.img[0].data(), i0, (size_t)t.w * t.h * 3);
  memcpy(t.img[1].data(), i1, (size_t)t.w * t.h * 3);
  memcpy(t.mask[0].data(), m0, (size_t)t.w * t.h);
  memcpy(t.mask[1].data(), m1, (size_t)t.w * t.h);
  construct_pyramid(c);
  c->elem = 0;
}
void orc_set_calib(void* h, const double* Q, const double* R, const double* T) {
  Ctx* c = (Ctx*)h;
  memcpy(c->Q, Q, sizeof(c->Q)); memcpy(c->Rf, R, sizeof(c->Rf)); memcpy(c->Tf, T, sizeof(c->Tf));
}
void orc_get_level(void* h, int level, int view, uint8_t* img, uint8_t* mask) {
  Ctx* c = (Ctx*)h;
  const Level& l = c->lv[level];
  if (img) memcpy(img, l.img[view].data(), (size_t)l.w * l.h * 3);
  if (mask) memcpy(mask, l.mask[view].data(), (size_t)l.w * l.h);
}
void orc_get_margins(void* h, int* o) {
  Ctx* c = (Ctx*)h;
  for (int k = 0; k < 2; k++) {
    const Boundary& b = c->margin[k];
    int* p = o + 6 * k;
    p[0] = b.YL; p[1] = b.YR; p[2] = b.XL; p[3] = b.XR; p[4] = b.width; p[5] = b.height;
  }
}
void orc_set_refine_iters(void* h, int n) { ((Ctx*)h)->refine_iters = n; }
int orc_run_stage(void* h, int level, int stage) { return run_stage((Ctx*)h, level, stage); }
void orc_match_one_layer(void* h, int level) {
  for (int s = 1; s <= 10; s++) run_stage((Ctx*)h, level, s);
}
int orc_disp_elem_size(void* h, int) { return ((Ctx*)h)->elem; }
void orc_get_disparity(void* h, int dir, void* out) {
  Ctx* c = (Ctx*)h;
  if (c->elem == 2) memcpy(out, c->ds[dir].data(), c->ds[dir].size() * 2);
  else memcpy(out, c->dd[dir].data(), c->dd[dir].size() * 8);
}
void orc_set_disparity(void* h, int dir, const void* in, int rows, int cols, int es) {
  Ctx* c = (Ctx*)h;
  const size_t n = (size_t)rows * cols;
  if (es == 2) c->ds[dir].assign((const short*)in, (const short*)in + n);
  else c->dd[dir].assign((const double*)in, (const double*)in + n);
  c->dw = cols; c->dh = rows; c->elem = es;
}
void orc_get_rematch_bounds(void* h, int dir, int16_t* bl, int16_t* br) {
  Ctx* c = (Ctx*)h;
  memcpy(bl, c->BL[dir].data(), c->BL[dir].size() * 2);
  memcpy(br, c->BR[dir].data(), c->BR[dir].size() * 2);
}
long orc_to_cloud(void* h) { return disparity_to_cloud((Ctx*)h); }
long orc_num_points(void* h) { return (long)((Ctx*)h)->pts_pix.size(); }
void orc_get_points(void* h, double* xyz) {
  Ctx* c = (Ctx*)h;
  memcpy(xyz, c->pts.data(), c->pts.size() * 8);
}
void orc_get_point_attrs(void* h, uint8_t* bgr, int32_t* pix) {
  Ctx* c = (Ctx*)h;
  memcpy(bgr, c->pts_bgr.data(), c->pts_bgr.size());
  memcpy(pix, c->pts_pix.data(), c->pts_pix.size() * 4);
}
long orc_match_pair(void* h) {
  Ctx* c = (Ctx*)h;
  construct_pyramid(c);
  c->elem = 0;
  for (int i = 0; i < c->L; i++)
    for (int s = 1; s <= 10; s++) run_stage(c, i, s);
  return disparity_to_cloud(c);
}

double orc_window_to_vec(const uint8_t* img, int pitch, int y0, int x, int ws, double* out) {
  const uint8_t* rows[7];
  for (int i = 0; i < ws; i++) rows[i] = img + (size_t)(y0 + i) * pitch;
  return window_to_vec(rows, x, ws, out);
}
double orc_ncc_match_value(const uint8_t* a, const uint8_t* b, int pitch, int y0, int xl, int xr, int ws) {
  const uint8_t *rl[7], *rr[7];
  for (int i = 0; i < ws; i++) { rl[i] = a + (size_t)(y0 + i) * pitch; rr[i] = b + (size_t)(y0 + i) * pitch; }
  double vl[147], vr[147];
  const int n = ws * ws * 3;
  const double nl = window_to_vec(rl, xl, ws, vl);
  for (int i = 0; i < n; i++) vl[i] /= nl;
  const double nr = window_to_vec(rr, xr, ws, vr);
  return dot2(vl, vr, n) / nr;
}
void orc_pyrdown(const uint8_t* src, int w, int h, int cn, uint8_t* dst) { pyr_down(src, w, h, cn, dst); }
void orc_erode_ellipse(const uint8_t* src, int w, int h, int ks, uint8_t* dst) { erode_ellipse(src, w, h, ks, dst); }
void orc_structuring_ellipse(int ks, uint8_t* dst) {
  std::vector<int> j1, j2;
  ellipse_rows(ks, j1, j2);
  for (int i = 0; i < ks; i++)
    for (int j = 0; j < ks; j++) dst[i * ks + j] = (uint8_t)(j >= j1[i] && j < j2[i]);
}

}  // extern "C"
