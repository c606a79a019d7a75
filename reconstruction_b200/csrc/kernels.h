// kernels.h — host-callable launchers of the sm_100a kernels (one per reference sweep).
// Every launcher enqueues on `st` and returns the number of kernel launches it issued.
#pragma once
#include "common.cuh"

struct PairViews {  // source / target view of one matching direction (image / image_inv, :43-50)
  const uint8_t *img0, *img1, *mask0, *mask1;
  const double2 *stat0, *stat1;  // per-pixel (mean, norm) of the 5x5x3 window (WindowToVec)
  const int2 *istat0, *istat1;   // per-pixel exact (sum, sum of squares) of the same window
  int W, H;
  long img_bytes;   // readable bytes of an image buffer (payload + slack)
  long mask_bytes;  // readable bytes of a mask buffer
};

// K0  ConstructPyrm (CStereoMatching.cpp:1040-1053): pyrDown of an interleaved u8 image, cn = 1 or 3.
int launch_pyrdown(const uint8_t* src, int W, int H, int cn, uint8_t* dst, cudaStream_t st);
// K1  FindMargin (:1011-1038): out[4] = {YL, YR, XL, XR} (device ints, initialised by the kernel).
int launch_find_margin(const uint8_t* mask, int W, int H, int R, int* out4, cudaStream_t st);
// per-pixel WindowToVec statistics of the (2R+1)^2*3 window centred on each pixel (CManageData.cpp:81-90)
int launch_window_stats(const uint8_t* img, int W, int H, int R, double2* stats, int2* istats, cudaStream_t st);

// Scratch of the screened searches (K3, K7): the pixels the integer screening pass could not settle.
struct SearchScratch {
  unsigned* list;                // flat pixel indices left by the per-thread screening pass (or all masked pixels, lowest level)
  unsigned* list2;               // ... left by the per-warp screening pass: input of the exact pass
  unsigned* list_wide;           // K3 band path: listed pixels whose candidate range is wide (a whole warp screens each)
  unsigned* n_list;              // device counters [4] of the current search: [0] list, [1] list2, [2] list_wide
  unsigned cap;
  unsigned long long* counters;  // [2]: [0] += pixels listed, [1] += pixels left to the exact pass, after every search (instrumentation)
  cudaStream_t side;             // optional second stream (+ fork / join events) for work that runs beside the main kernel
  cudaEvent_t ev_fork, ev_join;
};
// integer-only (sum, sum of squares) map of the 5x5x3 windows
int launch_window_istats(const uint8_t* img, int W, int H, int2* istats, cudaStream_t st);

// K2  LowestLevelInitialMatch (:170-227)
int launch_lowest_match(const PairViews& v, Bound ms, Bound mt, int R, short* out, const SearchScratch* sc, cudaStream_t st);
// K3  HighLevelInitialMatch (:231-308): prev = refined f64 map of the coarser level (pw x ph)
// band: take the TMA / shared-memory band kernel (ncc_band.cu) where the level allows it (needs sc, R == 2, offset == 2)
int launch_high_match(const PairViews& v, Bound ms, Bound mt, int R, int offset, const double* prev, int pw, int ph,
                      short* lo_scratch, short* hi_scratch, short* out, const SearchScratch* sc, bool band, cudaStream_t st);
// K3 band path (ncc_band.cu): NOMATCH fill, hole ranges, the band kernel; leaves the undecided pixels in sc->list with their
// ranges in lo_map / hi_map.  -2: the level cannot take this path.
int launch_high_match_band(const PairViews& v, Bound ms, Bound mt, int offset, const double* prev, int pw, int ph, short* lo_map,
                           short* hi_map, short* out, const SearchScratch* sc, cudaStream_t st);
int launch_range_lists(const PairViews& v, const short* lo_map, const short* hi_map, short* disp, const SearchScratch* sc, cudaStream_t st);
// K4  SmoothConstraint (:370-448): in -> out (out-of-place gather formulation)
int launch_smooth(const short* in, short* out, int W, int H, Bound m, cudaStream_t st);
// K5  OrderConstraint (:310-368): in place
int launch_order(short* disp, int W, int H, Bound m, cudaStream_t st);
// K6  one pass of UniquenessContraint_ (:462-497): P is filtered against Qm, in place
int launch_unique_s16(short* P, const short* Qm, int W, int H, Bound ms, Bound mt, cudaStream_t st);
int launch_unique_f64(double* P, const double* Qm, int W, int H, Bound ms, Bound mt, cudaStream_t st);
// K7  SetBoundary_smooth (:817-942) then the NCC search of Rematch (:499-570)
int launch_rematch_bounds(const short* disp, const uint8_t* mask, int W, int H, Bound ms, Bound mt, short* BL, short* BR,
                          cudaStream_t st);
// sc == nullptr: no screening pass (every candidate evaluated in exact arithmetic, statistics map required)
int launch_rematch_search(const PairViews& v, Bound ms, int R, const short* BL, const short* BR, short* disp,
                          const SearchScratch* sc, cudaStream_t st);
// K8  MedianFilter (:763-815), one iteration
int launch_median(const short* in, const uint8_t* mask, short* out, int W, int H, Bound m, cudaStream_t st);

// K9  DisparityRefine (:572-680)
struct RefineScratch {
  double* A;           // W*H  ping
  double* B;           // W*H  pong
  double2* table;      // K_REFINE * W*H entries (pwp, c)
  unsigned short* code;  // W*H  packed (table base, mode)
  unsigned long long* counters;  // [2]: [1] = out-of-table evaluations
  unsigned* miss_count;  // [SB_REFINE_MAX_ITERS] one counter per sweep
  unsigned* miss_list;   // [miss_cap] flat pixel indices of the current sweep's out-of-table pixels
  unsigned miss_cap;
  cudaEvent_t ev_begin, ev_end;  // optional: bracket the sweeps (profiling)
};
#define SB_REFINE_MAX_ITERS 1024
#define SB_REFINE_K 4      // table entries per pixel: im - im0 in [-2, 1]
#define SB_REFINE_KLO (-2)
// Fused form: both matching directions advance T sweeps per launch (temporal blocking in shared memory).
struct RefineFusedDir {
  const uint8_t *img0, *img1;  // source / target image of this direction
  long img_bytes;
  Bound ms;                    // source margin
  const double* src;
  double* dst;
  const double2* table;
  const unsigned short* code;
  double2* table_rw;           // same buffers, writable (k_refine_rebase only)
  unsigned short* code_rw;
  unsigned* miss_count;        // counter of this launch
  unsigned* miss_list;
  unsigned miss_cap;
};
struct RefineFusedArgs {
  RefineFusedDir d[2];
  int W, H, T;
  int use_tma;  // tile load phase through cp.async.bulk.tensor (needs W % 8 == 0), else plain loads
  long n_px;
  double ws;
  unsigned long long* counters;
};
// variant: tile shape / table access, see k_refine_dims in refine.cu; < 0 = pick per level.
// s[d].A / s[d].B / table / code / miss_* are per direction; s[0].ev_begin/ev_end bracket the sweeps.
// allow_tma: tile load phase through cp.async.bulk.tensor where the level's pitch permits it (SB200_REFINE_TMA=0 turns it off)
int launch_refine_fused(const PairViews v[2], const Bound ms[2], short* const in[2], int iterations, double ws, int T,
                        int variant, int allow_tma, const RefineScratch s[2], double* result[2], cudaStream_t st);

// K10 DisparityToCloud<double> (:682-761)
struct CloudScratch {
  unsigned short* run;   // W*H  horizontal run length of mask == 255
  uint8_t* eroded;       // W*H  1 where the eroded mask is 255
  int* row_count;        // H + 1
  int* row_offset;       // H + 1
  const short* ellipse;  // [2*ks]: j1[ks] then j2[ks], row extents of the MORPH_ELLIPSE element
};
// getStructuringElement(MORPH_ELLIPSE, ks x ks) row extents [j1, j2) (host)
void sb_ellipse_rows(int ks, short* j1, short* j2);
struct CloudParams {
  double q03, q13, q23, q32, q33;  // Q with column 3 already scaled (:697-698)
  double R[9], T[3];
};
int launch_cloud(const double* disp, const uint8_t* mask, const uint8_t* img, int W, int H, Bound m, int erode_ks,
                 const CloudParams& p, const CloudScratch& s, double* xyz, uint8_t* bgr, int* pix, int* n_points_dev,
                 cudaStream_t st);

// Rectify (:117-168), see rectify.cu
struct RectifyView {
  double iR[9];            // (P[:, :3] * R_new)^-1
  double fx, fy, u0, v0;   // of the ORIGINAL camera matrix
};
// compat: 245 = OpenCV 2.4.5's stereoRectify (what the reference links), 413 = OpenCV 4.13's (what the golden vectors pin)
void sb_stereo_rectify(const double* K1, const double* K2, int nx, int ny, const double* R, const double* T, double* R1, double* R2,
                       double* P1, double* P2, double* Q, int compat);
void sb_rectify_calib(const double* K0, const double* Rt0, const double* K1, const double* Rt1, int origin_w, int origin_h, int lowest_w,
                      int pyrm_num, double* R_new, double* P_scaled, double* P_final, double* Q, double* R_final, double* T_final, int compat);
bool sb_rectify_inverse(const double* P_scaled, const double* R_new, double* iR);
int launch_rectify_maps(int W, int H, const RectifyView& rv, double* starts, short2* map1, unsigned short* map2, cudaStream_t st);
int launch_remap(const uint8_t* src, int sw, int sh, int cn, const short2* map1, const unsigned short* map2, int W, int H, uint8_t* dst,
                 cudaStream_t st);
int launch_erode_ellipse(uint8_t* tab, int levels, int W, int H, int ks, const short* j12_dev, uint8_t* out, cudaStream_t st);

// s16 -> f64 conversion (Mat::convertTo, :585,587) and fills
int launch_fill_s16(short* p, long n, short v, cudaStream_t st);
