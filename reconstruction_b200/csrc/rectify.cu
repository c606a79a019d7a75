// rectify.cu — CStereoMatching::Rectify (CStereoMatching.cpp:117-168), the step right before the hot path
// (SURVEY.md 8f-1).  The reference does it with four OpenCV calls; OpenCV is a third-party dependency whose sources
// are not under /root/reference, so this follows the published algorithms (OpenCV calib3d / imgproc) and is pinned
// against cv2 4.13 vectors (tests/golden/rectify_cv2.npz):
//   sb_stereo_rectify      cv::stereoRectify (Bouguet), zero distortion, flags = 0, alpha = -1, newImageSize = imageSize
//                          as called at :128-131 — host, double precision
//   k_rectify_maps         cv::initUndistortRectifyMap(..., CV_16SC2) (:144): fixed-point source coordinates, 5 fractional bits
//   k_remap_linear<CN>     cv::remap(INTER_LINEAR, BORDER_CONSTANT 0) on those maps (:154,:156): exact integer arithmetic
//   k_hmin_level/k_erode   cv::erode with getStructuringElement(MORPH_ELLIPSE, 3*2^(L-1)) (:157-158): grey-level minimum
#include <math.h>
#include <string.h>

#include "kernels.h"

// ------------------------------------------------------------------------------------------------ host: calibration
namespace {

void mat3_mul(const double* a, const double* b, double* c, bool bt = false) {  // c = a * (bt ? b^T : b)
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += a[i * 3 + k] * (bt ? b[j * 3 + k] : b[k * 3 + j]);
      c[i * 3 + j] = s;
    }
}
void mat3_vec(const double* a, const double* v, double* o) {
  for (int i = 0; i < 3; i++) o[i] = a[i * 3] * v[0] + a[i * 3 + 1] * v[1] + a[i * 3 + 2] * v[2];
}

// cv::Rodrigues, vector -> matrix
void rodrigues_v2m(const double* r, double* R) {
  const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (theta < 2.2204460492503131e-16) {
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
  const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
  const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
  const double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  for (int i = 0; i < 9; i++) R[i] = c * ((i % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[i] + s * rx[i];
}

// cv::Rodrigues, matrix -> vector.  OpenCV first re-orthogonalises R through an SVD; the inputs here are products of
// rotation matrices, orthogonal to rounding, so that step changes the result only at the 1e-16 level and is omitted.
void rodrigues_m2v(const double* R, double* r) {
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  const double theta = acos(c);
  if (s < 1e-5) {
    if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
    double t = (R[0] + 1) * 0.5;
    rx = sqrt(t > 0. ? t : 0.);
    t = (R[4] + 1) * 0.5;
    ry = sqrt(t > 0. ? t : 0.) * (R[1] < 0 ? -1. : 1.);
    t = (R[8] + 1) * 0.5;
    rz = sqrt(t > 0. ? t : 0.) * (R[2] < 0 ? -1. : 1.);
    if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
    const double k = theta / sqrt(rx * rx + ry * ry + rz * rz);
    r[0] = rx * k; r[1] = ry * k; r[2] = rz * k;
    return;
  }
  const double vth = 1 / (2 * s) * theta;
  r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

// cv::invert(3x3, DECOMP_LU): OpenCV uses the cofactor formula for 3x3 matrices
bool inv3(const double* m, double* t) {
  double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  if (d == 0.) return false;
  d = 1. / d;
  t[0] = (m[4] * m[8] - m[5] * m[7]) * d;
  t[1] = (m[2] * m[7] - m[1] * m[8]) * d;
  t[2] = (m[1] * m[5] - m[2] * m[4]) * d;
  t[3] = (m[5] * m[6] - m[3] * m[8]) * d;
  t[4] = (m[0] * m[8] - m[2] * m[6]) * d;
  t[5] = (m[2] * m[3] - m[0] * m[5]) * d;
  t[6] = (m[3] * m[7] - m[4] * m[6]) * d;
  t[7] = (m[1] * m[6] - m[0] * m[7]) * d;
  t[8] = (m[0] * m[4] - m[1] * m[3]) * d;
  return true;
}

}  // namespace

// stereoRectify for the reference's call (:128-131).  K: 3x3, R: 3x3, T: 3.  Outputs R1, R2 (3x3), P1, P2 (3x4), Q (4x4).
// compat: which OpenCV's stereoRectify is followed where the versions differ.  245 = OpenCV 2.4.5, the version the reference
// links (include/opencv/version.hpp:50-53 in the reference tree): the SMALLER of the two focal lengths, image corners at
// (nx, ny); 413 = OpenCV 4.13, the version the golden vectors were produced with: the MEAN focal length, corners at
// (nx-1, ny-1).  2.4.5 itself cannot be run here (Windows .lib only), so that branch follows its published source and is not
// pinned by a vector; the 4.13 branch is (tests/golden/rectify_cv2.npz).
void sb_stereo_rectify(const double* K1, const double* K2, int nx, int ny, const double* R, const double* T, double* R1, double* R2,
                       double* P1, double* P2, double* Q, int compat) {
  double om[3], r_r[9], t[3], uu[3] = {0, 0, 0}, ww[3], wR[9];
  rodrigues_m2v(R, om);
  for (int i = 0; i < 3; i++) om[i] *= -0.5;  // each camera rotates half way
  rodrigues_v2m(om, r_r);
  mat3_vec(r_r, T, t);
  const int idx = fabs(t[0]) > fabs(t[1]) ? 0 : 1;
  const double c = t[idx], nt = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  uu[idx] = c > 0 ? 1 : -1;
  ww[0] = t[1] * uu[2] - t[2] * uu[1];  // global rotation that puts the baseline on the x (or y) axis
  ww[1] = t[2] * uu[0] - t[0] * uu[2];
  ww[2] = t[0] * uu[1] - t[1] * uu[0];
  const double nw = sqrt(ww[0] * ww[0] + ww[1] * ww[1] + ww[2] * ww[2]);
  if (nw > 0.0) {
    const double k = acos(fabs(c) / nt) / nw;
    for (int i = 0; i < 3; i++) ww[i] *= k;
  }
  rodrigues_v2m(ww, wR);
  mat3_mul(wR, r_r, R1, true);  // R1 = wR * r_r^T
  mat3_mul(wR, r_r, R2, false);
  mat3_vec(R2, T, t);
  // new focal length from the two f_y (horizontal pair) / f_x (vertical pair): their mean (4.13) or the smaller one (2.4.5).
  // Zero distortion: no k1 correction.
  const double f1 = K1[(idx ^ 1) * 4], f2 = K2[(idx ^ 1) * 4];
  const double fc_new = compat >= 300 ? (f1 + f2) * 0.5 : (f1 < f2 ? f1 : f2);
  const int cx = compat >= 300 ? nx - 1 : nx, cy = compat >= 300 ? ny - 1 : ny;  // where the far image corners are taken
  // new principal points: the image corners go through undistortPoints / projectPoints in SINGLE precision (CvPoint2D32f)
  double cc[2][2];
  for (int k = 0; k < 2; k++) {
    const double* A = k == 0 ? K1 : K2;
    const double* Rk = k == 0 ? R1 : R2;
    const double ifx = 1. / A[0], ify = 1. / A[4];
    double sx = 0, sy = 0;
    for (int i = 0; i < 4; i++) {
      const float px = (float)((i % 2) * cx), py = (float)((i < 2 ? 0 : 1) * cy);
      const float xn = (float)(((double)px - A[2]) * ifx), yn = (float)(((double)py - A[5]) * ify);  // undistortPoints, k = 0
      const double X = Rk[0] * xn + Rk[1] * yn + Rk[2], Y = Rk[3] * xn + Rk[4] * yn + Rk[5], Z = Rk[6] * xn + Rk[7] * yn + Rk[8];
      const double iz = Z ? 1. / Z : 1.;
      sx += (double)(float)(X * iz * fc_new);  // projectPoints with fc_new, cc = 0, stored as float
      sy += (double)(float)(Y * iz * fc_new);
    }
    cc[k][0] = (nx - 1) / 2. - sx * 0.25;
    cc[k][1] = (ny - 1) / 2. - sy * 0.25;
  }
  if (idx == 0) cc[0][1] = cc[1][1] = (cc[0][1] + cc[1][1]) * 0.5;  // flags = 0: only the shared axis is averaged
  else cc[0][0] = cc[1][0] = (cc[0][0] + cc[1][0]) * 0.5;
  memset(P1, 0, 12 * sizeof(double));
  memset(P2, 0, 12 * sizeof(double));
  P1[0] = P1[5] = P2[0] = P2[5] = fc_new;
  P1[2] = cc[0][0]; P1[6] = cc[0][1]; P1[10] = 1;
  P2[2] = cc[1][0]; P2[6] = cc[1][1]; P2[10] = 1;
  P2[idx * 4 + 3] = t[idx] * fc_new;
  memset(Q, 0, 16 * sizeof(double));
  Q[0] = Q[5] = 1;
  Q[3] = -cc[0][0];
  Q[7] = -cc[0][1];
  Q[11] = fc_new;
  Q[14] = -1. / t[idx];
  Q[15] = (idx == 0 ? cc[0][0] - cc[1][0] : cc[0][1] - cc[1][1]) / t[idx];
}

// The calibration half of Rectify (:121-145): everything the image half and DisparityToCloud need.
void sb_rectify_calib(const double* K0, const double* Rt0, const double* K1, const double* Rt1, int origin_w, int origin_h, int lowest_w,
                      int pyrm_num, double* R_new, double* P_scaled, double* P_final, double* Q, double* R_final, double* T_final, int compat) {
  double R0[9], R1m[9], t0[3], t1[3], R[9], T[3], tmp[3];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) { R0[i * 3 + j] = Rt0[i * 4 + j]; R1m[i * 3 + j] = Rt1[i * 4 + j]; }
    t0[i] = Rt0[i * 4 + 3];
    t1[i] = Rt1[i * 4 + 3];
  }
  mat3_mul(R1m, R0, R, true);  // R = R1 * R0^T (:125)
  mat3_vec(R, t0, tmp);
  for (int i = 0; i < 3; i++) T[i] = -tmp[i] + t1[i];  // T = -R*t0 + t1 (:126)
  double P[2][12];
  sb_stereo_rectify(K0, K1, origin_w, origin_h, R, T, R_new, R_new + 9, P[0], P[1], Q, compat);
  // R_final = R0^T * R_new0^T (:132), T_final = -R0^T * t0 (:133)
  double R0t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R0t[i * 3 + j] = R0[j * 3 + i];
  mat3_mul(R0t, R_new, R_final, true);
  mat3_vec(R0t, t0, tmp);
  for (int i = 0; i < 3; i++) T_final[i] = -tmp[i];
  // Extrinsic_final = [R_final^T | -R_final^T T_final] (:134-137)
  double E[16] = {0};
  E[15] = 1;
  double Rft[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Rft[i * 3 + j] = R_final[j * 3 + i];
  mat3_vec(Rft, T_final, tmp);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) E[i * 4 + j] = Rft[i * 3 + j];
    E[i * 4 + 3] = -tmp[i];
  }
  Q[14] = -Q[14];  // :138
  const double scale = double(lowest_w) / origin_w * (1 << (pyrm_num - 1));  // :140
  for (int v = 0; v < 2; v++) {
    for (int i = 0; i < 8; i++) P[v][i] *= scale;  // P.rowRange(0,2) *= scale (:143)
    memcpy(P_scaled + 12 * v, P[v], sizeof P[v]);
    for (int i = 0; i < 3; i++)  // P = P * Extrinsic_final (:145)
      for (int j = 0; j < 4; j++) {
        double s = 0;
        for (int k = 0; k < 4; k++) s += P[v][i * 4 + k] * E[k * 4 + j];
        P_final[12 * v + i * 4 + j] = s;
      }
  }
}

// iR = (P[:, :3] * R)^-1 of initUndistortRectifyMap
bool sb_rectify_inverse(const double* P_scaled, const double* R_new, double* iR) {
  double A[9], M[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A[i * 3 + j] = P_scaled[i * 4 + j];
  mat3_mul(A, R_new, M, false);
  return inv3(M, iR);
}

// ------------------------------------------------------------------------------------------------ device
// initUndistortRectifyMap, zero distortion, m1type CV_16SC2: per destination pixel (j, i) the source position
//   [x y w]' = iR [j i 1]',  u = fx x/w + u0,  v = fy y/w + v0   in double, then 5-bit fixed point.
// OpenCV does not evaluate iR [j i 1]' per pixel: it starts a row at (i iR[1] + iR[2], ...) and ADDS iR[0], iR[3], iR[6]
// once per column, rounding at every step.  A closed form differs from that in the last bits and flips about one map
// entry in a thousand by one fixed-point step, so the accumulation is reproduced: k_rectify_row_starts walks every row
// once (one thread per row, three independent chains of additions) and stores the running values at every 32nd column;
// k_rectify_maps restarts from those and replays 32 additions per thread.  Pinned bit for bit against cv2 4.13's maps
// (its scalar and its dispatched loop give the same values; tests/golden/rectify_cv2.npz).
#define SB_RECT_SEG 32
__global__ void __launch_bounds__(64) k_rectify_row_starts(int W, int H, RectifyView rv, double* __restrict__ starts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H) return;
  const double* ir = rv.iR;
  double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
  const int nseg = (W + SB_RECT_SEG - 1) / SB_RECT_SEG;
  double* o = starts + (size_t)i * nseg * 3;
  for (int sgm = 0; sgm < nseg; sgm++) {
    o[sgm * 3] = _x; o[sgm * 3 + 1] = _y; o[sgm * 3 + 2] = _w;
#pragma unroll 8
    for (int k = 0; k < SB_RECT_SEG; k++) { _x += ir[0]; _y += ir[3]; _w += ir[6]; }
  }
}

// one block per row; thread t owns columns [32 t, 32 t + 32); results are staged in shared memory and stored coalesced
__global__ void __launch_bounds__(128) k_rectify_maps(int W, int H, RectifyView rv, const double* __restrict__ starts,
                                                      short2* __restrict__ map1, unsigned short* __restrict__ map2) {
  __shared__ short2 s1[128 * SB_RECT_SEG];
  __shared__ unsigned short s2[128 * SB_RECT_SEG];
  const int i = blockIdx.y;
  const int nseg = (W + SB_RECT_SEG - 1) / SB_RECT_SEG;
  const double* ir = rv.iR;
  for (int seg0 = 0; seg0 < nseg; seg0 += 128) {
    const int sgm = seg0 + threadIdx.x;
    if (sgm < nseg) {
      const double* st = starts + ((size_t)i * nseg + sgm) * 3;
      double _x = st[0], _y = st[1], _w = st[2];
#pragma unroll 4
      for (int k = 0; k < SB_RECT_SEG; k++, _x += ir[0], _y += ir[3], _w += ir[6]) {
        const double w = 1. / _w, x = _x * w, y = _y * w;
        const double u = rv.fx * x + rv.u0, v = rv.fy * y + rv.v0;
        const int iu = __double2int_rn(u * 32.0), iv = __double2int_rn(v * 32.0);  // saturate_cast<int> = round half to even
        // rotated by the thread index: the 32 threads of a warp write 32 different banks
        const int slot = threadIdx.x * SB_RECT_SEG + ((k + threadIdx.x) & (SB_RECT_SEG - 1));
        s1[slot] = make_short2((short)(iu >> 5), (short)(iv >> 5));
        s2[slot] = (unsigned short)((iv & 31) * 32 + (iu & 31));
      }
    }
    __syncthreads();
    const int j0 = seg0 * SB_RECT_SEG;
    for (int q = threadIdx.x; q < 128 * SB_RECT_SEG; q += 128) {
      const int j = j0 + q;
      if (j < W) {
        const int t = q / SB_RECT_SEG, k = q % SB_RECT_SEG;
        const int slot = t * SB_RECT_SEG + ((k + t) & (SB_RECT_SEG - 1));
        map1[(size_t)i * W + j] = s1[slot];
        map2[(size_t)i * W + j] = s2[slot];
      }
    }
    __syncthreads();
  }
}

// remap, INTER_LINEAR on fixed-point maps: weights (32-fx)(32-fy)*32 ... sum to 2^15; result (sum + 2^14) >> 15.
// BORDER_CONSTANT with value 0: taps outside the source read 0.
template <int CN>
__global__ void __launch_bounds__(256) k_remap_linear(const uint8_t* __restrict__ src, int sw, int sh, const short2* __restrict__ map1,
                                                      const unsigned short* __restrict__ map2, int W, int H, uint8_t* __restrict__ dst) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= W) return;
  const short2 m = map1[(size_t)i * W + j];
  const int f = map2[(size_t)i * W + j] & 1023;
  const int fx = f & 31, fy = f >> 5;
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  const int sx = m.x, sy = m.y;
  const bool x0 = sx >= 0 && sx < sw, x1 = sx + 1 >= 0 && sx + 1 < sw, y0 = sy >= 0 && sy < sh, y1 = sy + 1 >= 0 && sy + 1 < sh;
#pragma unroll
  for (int c = 0; c < CN; c++) {
    const int p00 = (x0 && y0) ? src[((size_t)sy * sw + sx) * CN + c] : 0;
    const int p01 = (x1 && y0) ? src[((size_t)sy * sw + sx + 1) * CN + c] : 0;
    const int p10 = (x0 && y1) ? src[((size_t)(sy + 1) * sw + sx) * CN + c] : 0;
    const int p11 = (x1 && y1) ? src[((size_t)(sy + 1) * sw + sx + 1) * CN + c] : 0;
    const int s = p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11;
    dst[((size_t)i * W + j) * CN + c] = (uint8_t)min(255, (s + (1 << 14)) >> 15);
  }
}

// Grey-level erosion with the ellipse element: out(y, x) = min over element rows i of the minimum of row y+i-a over
// [x+j1-a, x+j2-1-a] (positions outside the image do not constrain: cv::erode's default border is +inf).  Row-range minima
// come from a sparse table M_k(y, x) = min src(y, x .. x+2^k-1) built by doubling.
__global__ void __launch_bounds__(256) k_hmin_level(const uint8_t* __restrict__ prev, int W, long n, int half, uint8_t* __restrict__ out) {
  const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const int x = (int)(f % W);
  const uint8_t a = prev[f];
  out[f] = (x + half < W) ? (uint8_t)min((int)a, (int)prev[f + half]) : a;
}

__global__ void __launch_bounds__(256) k_erode_ellipse(const uint8_t* __restrict__ tab, long plane, int W, int H, int ks,
                                                       const short* __restrict__ j12, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int a = ks / 2;
  int v = 255;
  for (int i = 0; i < ks; i++) {
    const int sy = y + i - a;
    const int j1 = j12[i], j2 = j12[ks + i];
    if (sy < 0 || sy >= H || j2 <= j1) continue;
    const int xa = max(x + j1 - a, 0), xb = min(x + j2 - 1 - a, W - 1);
    if (xb < xa) continue;
    const int k = 31 - __clz(xb - xa + 1);
    const uint8_t* row = tab + (long)k * plane + (long)sy * W;
    v = min(v, min((int)row[xa], (int)row[xb - (1 << k) + 1]));
  }
  out[(size_t)y * W + x] = (uint8_t)v;
}

// starts: scratch of H * ceil(W / 32) * 3 doubles
int launch_rectify_maps(int W, int H, const RectifyView& rv, double* starts, short2* map1, unsigned short* map2, cudaStream_t st) {
  k_rectify_row_starts<<<(H + 63) / 64, 64, 0, st>>>(W, H, rv, starts);
  k_rectify_maps<<<dim3(1, H), 128, 0, st>>>(W, H, rv, starts, map1, map2);
  return 2;
}

int launch_remap(const uint8_t* src, int sw, int sh, int cn, const short2* map1, const unsigned short* map2, int W, int H, uint8_t* dst,
                 cudaStream_t st) {
  dim3 g((W + 255) / 256, H);
  if (cn == 3) k_remap_linear<3><<<g, 256, 0, st>>>(src, sw, sh, map1, map2, W, H, dst);
  else k_remap_linear<1><<<g, 256, 0, st>>>(src, sw, sh, map1, map2, W, H, dst);
  return 1;
}

// tab: (levels) planes of W*H bytes, plane 0 = the image to erode (filled by the caller)
int launch_erode_ellipse(uint8_t* tab, int levels, int W, int H, int ks, const short* j12_dev, uint8_t* out, cudaStream_t st) {
  const long n = (long)W * H;
  int nl = 0;
  for (int k = 1; k < levels; k++) {
    k_hmin_level<<<(int)((n + 255) / 256), 256, 0, st>>>(tab + (k - 1) * n, W, n, 1 << (k - 1), tab + k * n);
    nl++;
  }
  k_erode_ellipse<<<dim3((W + 255) / 256, H), 256, 0, st>>>(tab, n, W, H, ks, j12_dev, out);
  return nl + 1;
}
