// TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load what this builds.
//
// Compiles the reference's OWN reconstruction/CStereoMatching.cpp and
// reconstruction/CManageData.cpp, unmodified and where they lie under
// $(REF_ROOT) (default /root/reference), against oracle/ref_build/shim (our
// minimal cv::) and the reference's vendored Armadillo 4.200, and exposes the
// private stage functions through a small C ABI so the CUDA path (and the
// oracle restatement in oracle/stereo_oracle.cpp) can be diffed stage by stage
// at the reference's own dump points (CStereoMatching.cpp:63-111).
//
// No reference source is copied: the two .cpp files are #included by name and
// resolved through -I$(REF_ROOT)/reconstruction at build time.  Output goes to
// oracle/_ref/ only (git-ignored, travels to the GPU box as a built .so).
//
// Boundary: Rectify (OpenCV stereoRectify/remap, parity unpinned)
// is bypassed; callers hand in the staged rectified top-level images, masks
// and Q / R_final / T_final — exactly what Rectify leaves behind
// (CStereoMatching.cpp:117-168).

#include <opencv/core.hpp>
#include <armadillo/armadillo>
#include <stdio.h>
#include <iostream>
#include <vector>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

// open the class so the harness can call the per-stage members
#define private public
#include "CStereoMatching.h"
#undef private

#include "CManageData.cpp"
#include "CStereoMatching.cpp"

// ---- sink stub: CCloudOptimization is a DLL import in the reference
// (CStereoMatching.h:17-32); record InsertPoint calls in arrival order. ----
static std::vector<double> g_points;
void CCloudOptimization::Init(int, double, int, double, double, CManageData* d, bool) { m_ImageData = d; }
void CCloudOptimization::InsertPoint(cv::Mat p) {
  g_points.push_back(p.at<double>(0, 0));
  g_points.push_back(p.at<double>(1, 0));
  g_points.push_back(p.at<double>(2, 0));
}
void CCloudOptimization::filter(int) {}
void CCloudOptimization::run() {}

namespace {
struct RefCtx {
  CManageData data;
  CStereoMatching m;
  CCloudOptimization cloud;
  cv::Mat disparity[2];
  cv::Mat BL[2], BR[2];
  int refine_iters;
};

cv::Mat wrap_copy(const void* src, int rows, int cols, int type) {
  cv::Mat tmp(rows, cols, type, const_cast<void*>(src));
  return tmp.clone();
}
}  // namespace

extern "C" {

void* ref_create(int pyrm_num, int lowest_w, int lowest_h, int origin_w, int origin_h, int radius, double ws,
                 int offset) {
  RefCtx* c = new RefCtx();
  CManageData& d = c->data;
  d.m_PyrmNum = pyrm_num;
  d.m_LowestLevelSize = cv::Size(lowest_w, lowest_h);
  d.m_OriginSize = cv::Size(origin_w, origin_h);
  d.m_CampairNum = 1;
  d.m_CameraNum = 2;
  d.isoutput = 0;
  d.cam.resize(1);
  d.cam[0].resize(2);
  d.cam[0][0].camID = 0;
  d.cam[0][1].camID = 1;
  // same allocation pattern as CManageData::Init (CManageData.cpp:70-76)
  d.imagePyrm = new cv::Mat*[pyrm_num];
  d.maskPyrm = new cv::Mat*[pyrm_num];
  for (int i = 0; i < pyrm_num; i++) {
    d.imagePyrm[i] = new cv::Mat[2];
    d.maskPyrm[i] = new cv::Mat[2];
  }
  c->m.Init(&c->data, &c->cloud, radius, ws, offset);
  c->m.Verbose = 0;
  c->refine_iters = -1;
  return c;
}

void ref_destroy(void* h) { delete (RefCtx*)h; }

void ref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}
int ref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// staged rectified top-level inputs (tight pitch); then the reference's own ConstructPyrm
void ref_set_pair(void* h, const uint8_t* img0, const uint8_t* img1, const uint8_t* mask0, const uint8_t* mask1) {
  RefCtx* c = (RefCtx*)h;
  const int L = c->data.m_PyrmNum;
  const int W = c->data.m_LowestLevelSize.width << (L - 1), H = c->data.m_LowestLevelSize.height << (L - 1);
  c->data.cam[0][0].image = wrap_copy(img0, H, W, CV_8UC3);
  c->data.cam[0][1].image = wrap_copy(img1, H, W, CV_8UC3);
  c->data.cam[0][0].mask = wrap_copy(mask0, H, W, CV_8UC1);
  c->data.cam[0][1].mask = wrap_copy(mask1, H, W, CV_8UC1);
  c->m.ConstructPyrm(0);
  c->disparity[0].release();
  c->disparity[1].release();
}

void ref_set_calib(void* h, const double* Q16, const double* R9, const double* T3) {
  RefCtx* c = (RefCtx*)h;
  c->m.Q = wrap_copy(Q16, 4, 4, CV_64FC1);
  c->m.R_final = wrap_copy(R9, 3, 3, CV_64FC1);
  c->m.T_final = wrap_copy(T3, 3, 1, CV_64FC1);
}

void ref_level_size(void* h, int level, int* w, int* hgt) {
  RefCtx* c = (RefCtx*)h;
  *w = c->data.imagePyrm[level][0].cols;
  *hgt = c->data.imagePyrm[level][0].rows;
}

void ref_get_level(void* h, int level, int view, uint8_t* img, uint8_t* mask) {
  RefCtx* c = (RefCtx*)h;
  const cv::Mat& I = c->data.imagePyrm[level][view];
  const cv::Mat& M = c->data.maskPyrm[level][view];
  for (int y = 0; y < I.rows; y++) {
    if (img) memcpy(img + (size_t)y * I.cols * 3, I.ptr(y), (size_t)I.cols * 3);
    if (mask) memcpy(mask + (size_t)y * M.cols, M.ptr(y), (size_t)M.cols);
  }
}

static void margins_out(RefCtx* c, int* out12) {
  for (int k = 0; k < 2; k++) {
    const Boundary& b = c->m.margin[k];
    int* o = out12 + 6 * k;
    o[0] = b.YL; o[1] = b.YR; o[2] = b.XL; o[3] = b.XR; o[4] = b.width; o[5] = b.height;
  }
}

void ref_get_margins(void* h, int* out12) { margins_out((RefCtx*)h, out12); }

void ref_set_refine_iters(void* h, int n) { ((RefCtx*)h)->refine_iters = n; }

// One step of MatchOneLayer (CStereoMatching.cpp:51-109), numbered as in
// SURVEY.md §3.3.  Returns 0, or -1 for an unknown stage.
int ref_run_stage(void* h, int level, int stage) {
  RefCtx* c = (RefCtx*)h;
  CStereoMatching& m = c->m;
  CManageData& d = c->data;
  cv::Mat image[2], mask[2], image_inv[2], mask_inv[2];
  image[0] = d.imagePyrm[level][0];
  image[1] = d.imagePyrm[level][1];
  mask[0] = d.maskPyrm[level][0];
  mask[1] = d.maskPyrm[level][1];
  image_inv[0] = image[1];
  image_inv[1] = image[0];
  mask_inv[0] = mask[1];
  mask_inv[1] = mask[0];
  cv::Mat* disp = c->disparity;
  switch (stage) {
    case 1:
      m.FindMargin(m.margin[0], d.maskPyrm[level][0]);
      m.FindMargin(m.margin[1], d.maskPyrm[level][1]);
      return 0;
    case 2:
      if (level == 0) {
        m.LowestLevelInitialMatch(image, mask, disp[0], true);
        m.LowestLevelInitialMatch(image_inv, mask_inv, disp[1], false);
      } else {
        m.HighLevelInitialMatch(image, mask, disp[0], level, true);
        m.HighLevelInitialMatch(image_inv, mask_inv, disp[1], level, false);
      }
      return 0;
    case 3:
      m.SmoothConstraint(disp[0], true);
      m.SmoothConstraint(disp[1], false);
      return 0;
    case 4:
      m.OrderConstraint(disp[0], true);
      m.OrderConstraint(disp[1], false);
      return 0;
    case 5:
    case 7:
      m.UniquenessContraint<short>(disp);
      return 0;
    case 6:
      // dump point of the commented SaveMat(BL,"bl.dat") (CStereoMatching.cpp:515)
      m.SetBoundary_smooth<short>(disp[0], mask[0], c->BL[0], c->BR[0], true);
      m.SetBoundary_smooth<short>(disp[1], mask_inv[0], c->BL[1], c->BR[1], false);
      m.Rematch(image, mask, disp[0], true);
      m.Rematch(image_inv, mask_inv, disp[1], false);
      return 0;
    case 8:
      m.MedianFilter(disp[0], mask[0], 1, true);
      m.MedianFilter(disp[1], mask[1], 1, false);
      return 0;
    case 9: {
      const int it = c->refine_iters >= 0 ? c->refine_iters : 30 + level * 30;
      m.DisparityRefine(disp[0], image, it, true);
      m.DisparityRefine(disp[1], image_inv, it, false);
      return 0;
    }
    case 10:
      m.UniquenessContraint<double>(disp);
      return 0;
  }
  return -1;
}

// The reference's own MatchOneLayer, untouched stage order (for cross-checking
// the stage runner above and for timing).
void ref_match_one_layer(void* h, int level) {
  RefCtx* c = (RefCtx*)h;
  c->m.MatchOneLayer(c->disparity, level);
}

int ref_disp_elem_size(void* h, int dir) {
  RefCtx* c = (RefCtx*)h;
  return c->disparity[dir].empty() ? 0 : (int)c->disparity[dir].elemSize();
}

void ref_get_disparity(void* h, int dir, void* out) {
  RefCtx* c = (RefCtx*)h;
  const cv::Mat& D = c->disparity[dir];
  const size_t rb = (size_t)D.cols * D.elemSize();
  for (int y = 0; y < D.rows; y++) memcpy((char*)out + y * rb, D.ptr(y), rb);
}

void ref_set_disparity(void* h, int dir, const void* in, int rows, int cols, int elem_size) {
  RefCtx* c = (RefCtx*)h;
  c->disparity[dir] = wrap_copy(in, rows, cols, elem_size == 2 ? CV_16SC1 : CV_64FC1);
}

void ref_get_rematch_bounds(void* h, int dir, int16_t* bl, int16_t* br) {
  RefCtx* c = (RefCtx*)h;
  const cv::Mat& A = c->BL[dir];
  const cv::Mat& B = c->BR[dir];
  for (int y = 0; y < A.rows; y++) {
    memcpy(bl + (size_t)y * A.cols, A.ptr(y), (size_t)A.cols * 2);
    memcpy(br + (size_t)y * B.cols, B.ptr(y), (size_t)B.cols * 2);
  }
}

// DisparityToCloud<double> exactly as MatchAllLayer calls it (CStereoMatching.cpp:27-29)
long ref_to_cloud(void* h) {
  RefCtx* c = (RefCtx*)h;
  const int L = c->data.m_PyrmNum;
  g_points.clear();
  c->data.cam[0][0].bound = c->m.margin[0];
  c->data.cam[0][1].bound = c->m.margin[1];
  c->m.DisparityToCloud<double>(c->disparity[0], c->data.maskPyrm[L - 1][0], c->m.Q, L - 1, true, 0);
  return (long)(g_points.size() / 3);
}

void ref_get_points(void* h, double* xyz) {
  (void)h;
  memcpy(xyz, g_points.data(), g_points.size() * sizeof(double));
}

// whole per-pair span timed by the CPU baseline: ConstructPyrm + MatchOneLayer x L + DisparityToCloud
long ref_match_pair(void* h) {
  RefCtx* c = (RefCtx*)h;
  const int L = c->data.m_PyrmNum;
  c->m.ConstructPyrm(0);
  c->disparity[0].release();
  c->disparity[1].release();
  for (int i = 0; i < L; i++) c->m.MatchOneLayer(c->disparity, i);
  return ref_to_cloud(h);
}

// ---- primitives, for known-answer tests ----
double ref_window_to_vec(const uint8_t* img, int pitch, int y0, int x, int ws, double* out) {
  CManageData d;
  std::vector<uchar*> rows(ws);
  for (int i = 0; i < ws; i++) rows[i] = const_cast<uchar*>(img) + (size_t)(y0 + i) * pitch;
  arma::vec u(ws * ws * 3);
  const double n = d.WindowToVec(rows.data(), x, ws, u);
  for (int k = 0; k < ws * ws * 3; k++) out[k] = u(k);
  return n;
}

double ref_ncc_match_value(const uint8_t* imgL, const uint8_t* imgR, int pitch, int y0, int xl, int xr, int ws) {
  // value compared in the argmax loops (CStereoMatching.cpp:203-212)
  CManageData d;
  std::vector<uchar*> rl(ws), rr(ws);
  for (int i = 0; i < ws; i++) {
    rl[i] = const_cast<uchar*>(imgL) + (size_t)(y0 + i) * pitch;
    rr[i] = const_cast<uchar*>(imgR) + (size_t)(y0 + i) * pitch;
  }
  arma::vec vl(ws * ws * 3), vr(ws * ws * 3);
  const double nl = d.WindowToVec(rl.data(), xl, ws, vl);
  vl /= nl;
  const double nr = d.WindowToVec(rr.data(), xr, ws, vr);
  return arma::dot(vl, vr) / nr;
}

void ref_pyrdown(const uint8_t* src, int w, int h, int cn, uint8_t* dst) {
  cv::Mat s(h, w, cn == 3 ? CV_8UC3 : CV_8UC1, const_cast<uint8_t*>(src));
  cv::Mat d;
  cv::pyrDown(s, d);
  for (int y = 0; y < d.rows; y++) memcpy(dst + (size_t)y * d.cols * cn, d.ptr(y), (size_t)d.cols * cn);
}

void ref_erode_ellipse(const uint8_t* src, int w, int h, int ksize, uint8_t* dst) {
  cv::Mat s(h, w, CV_8UC1, const_cast<uint8_t*>(src));
  cv::Mat d;
  cv::erode(s, d, cv::getStructuringElement(cv::MORPH_ELLIPSE, cv::Size(ksize, ksize)));
  for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * w, d.ptr(y), (size_t)w);
}

void ref_structuring_ellipse(int ksize, uint8_t* dst) {
  cv::Mat e = cv::getStructuringElement(cv::MORPH_ELLIPSE, cv::Size(ksize, ksize));
  for (int y = 0; y < ksize; y++) memcpy(dst + (size_t)y * ksize, e.ptr(y), (size_t)ksize);
}

}  // extern "C"
