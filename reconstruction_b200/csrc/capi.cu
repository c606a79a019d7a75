// capi.cu — the context behind include/stereo_b200.h: device buffers for one camera pair, the fixed
// stage order of MatchOneLayer (CStereoMatching.cpp:36-113) and the C entry points.
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <string>
#include <vector>

#include "../../include/stereo_b200.h"
#include "kernels.h"

static const unsigned long long h_exp_tab[256] = {
#include "exp_table.inc"
};
extern "C" SB200_API double sb200_exp_host(double x) { return sb_exp_twin(x, h_exp_tab); }

namespace {

struct Level {
  int w = 0, h = 0;
  uint8_t* img[2] = {nullptr, nullptr};
  uint8_t* mask[2] = {nullptr, nullptr};
  long img_bytes = 0, mask_bytes = 0;
  Bound margin[2];
};

}  // namespace

struct sb200_ctx {
  int device = 0, L = 0, W0 = 0, H0 = 0, OW = 0, OH = 0, R = 2, offset = 2, refine_override = -1;
  double ws = 0.5;
  cudaStream_t st = nullptr;
  // K3 (HighLevelInitialMatch): direction 1 runs on `side` beside direction 0 on `st`, each direction's hole ranges on its own
  // stream beside its band kernel; fork / join through events (capturable into a CUDA graph)
  cudaStream_t side = nullptr, hole_s[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_hf[2] = {nullptr, nullptr}, ev_hj[2] = {nullptr, nullptr};
  cudaEvent_t ev_done = nullptr;  // blocking-sync event: host waits sleep instead of spinning (sync_ctx)
  std::vector<Level> lv;
  int* d_margins = nullptr;  // [L][2][4]
  bool uploaded = false, calib_set = false;
  // disparity state (cv::Mat disparity[2] of MatchAllLayer)
  short* ds[2] = {nullptr, nullptr};
  short* ds_tmp = nullptr;
  short *range_lo[2] = {nullptr, nullptr}, *range_hi[2] = {nullptr, nullptr};  // per direction (the two directions of K3 run concurrently)
  short *BL[2] = {nullptr, nullptr}, *BR[2] = {nullptr, nullptr};
  double* f64buf[4] = {nullptr, nullptr, nullptr, nullptr};
  double* dd[2] = {nullptr, nullptr};
  int dw = 0, dh = 0, elem = 0;
  // per-level window statistics
  double2* stats[2] = {nullptr, nullptr};
  int2* istats[2] = {nullptr, nullptr};
  unsigned long long* search_counters = nullptr;  // [2]: [1] = pixels the screening pass left to the exact search
  unsigned* search_list[2] = {nullptr, nullptr};
  unsigned* search_list2[2] = {nullptr, nullptr};
  unsigned* search_listw[2] = {nullptr, nullptr};
  unsigned* search_n = nullptr;  // [2][4]
  SearchScratch ss{}, ss1{};     // per direction; everything sequential uses ss
  bool screen = true;                             // SB200_SCREEN=0 disables the integer screening pass
  bool band = true;                               // SB200_BAND=0: K3 through the register-resident screening kernel instead of ncc_band.cu
  int stats_level = -1;
  // refinement: per-direction table / code / miss list (rs[0] also owns the counters)
  RefineScratch rs[2]{};
  bool refine_tma = true;                 // SB200_REFINE_TMA=0: plain loads in the tile load phase
  int refine_T = 5, refine_variant = -1;  // sweeps fused per launch, tile shape (< 0: per level); SB200_REFINE_T / _TILE
  // triangulation
  CloudScratch cs{};
  short* d_ellipse = nullptr;
  int erode_ks = 0;
  double Q[16], Rf[9], Tf[3];
  double* xyz = nullptr;
  uint8_t* bgr = nullptr;
  int* pix = nullptr;
  int* d_npoints = nullptr;
  int* h_npoints = nullptr;  // pinned
  int64_t n_points = 0;
  // Rectify scratch (allocated on first use)
  uint8_t *rc_src_img = nullptr, *rc_src_mask = nullptr, *rc_tab = nullptr;
  size_t rc_src_cap = 0;
  double* rc_starts = nullptr;  // running row values of initUndistortRectifyMap at every 32nd column
  short2* rc_map1 = nullptr;
  unsigned short* rc_map2 = nullptr;
  short* rc_ellipse = nullptr;
  int rc_ks = 0, rc_levels = 0;
  bool rc_maps_given = false;
  int rc_views_done = 0;
  // CUDA graph of a whole pair (MatchOneLayer x L + DisparityToCloud): the fixed stage order is captured from the stream
  // and replayed as one graph launch; the next pair re-captures and updates the executable graph in place (same topology,
  // new grid sizes / margins / tensor maps).  SB200_GRAPH=0, or profiling on, enqueues kernel by kernel instead.
  bool use_graph = true;
  cudaGraphExec_t gexec = nullptr;
  int64_t graph_launches = 0, graph_rebuilds = 0;
  // instrumentation
  int64_t launches = 0;
  bool profiling = false;
  struct Ev { int stage; cudaEvent_t a, b; int level; };
  std::vector<Ev> events;
  double stage_ms[16] = {0};
  double stage_level_ms[16][SB_MAX_LEVELS] = {{0}};
  // DisparityRefine sweep kernel per pyramid level, since the last sb200_get_refine_profile reset
  double sweep_ms[SB_MAX_LEVELS] = {0};
  int64_t sweep_launches[SB_MAX_LEVELS] = {0}, sweep_px_iters[SB_MAX_LEVELS] = {0};
  std::string err;
  Bound cur_margin[2];
  int cur_level = -1;
};

namespace {

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      char b_[512];                                                                       \
      snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      c->err = b_;                                                                        \
      return SB200_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

template <class T> cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }

// Wait for the context's stream WITHOUT spinning: an event created with cudaEventBlockingSync puts the host thread to sleep.
// With several pairs in flight per GPU and one process per GPU there are more waiting host threads than cores on the box
// (8 ranks x (3 matchers + 1 exchange thread) on 16 hardware threads); spinning waiters then delay the threads that are
// trying to enqueue work.
cudaError_t sync_ctx(sb200_ctx* c) {
  if (!c->ev_done) return cudaStreamSynchronize(c->st);
  cudaError_t e = cudaEventRecord(c->ev_done, c->st);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(c->ev_done);
}

struct StageTimer {
  sb200_ctx* c;
  sb200_ctx::Ev ev{};
  bool on;
  StageTimer(sb200_ctx* c_, int stage, int level = -1) : c(c_), on(c_->profiling) {
    if (!on) return;
    ev.stage = stage;
    ev.level = level;
    cudaEventCreate(&ev.a);
    cudaEventCreate(&ev.b);
    cudaEventRecord(ev.a, c->st);
  }
  ~StageTimer() {
    if (!on) return;
    cudaEventRecord(ev.b, c->st);
    c->events.push_back(ev);
  }
};

PairViews make_views(const sb200_ctx* c, int level, bool zeroOne) {
  const Level& l = c->lv[level];
  PairViews v;
  const int s = zeroOne ? 0 : 1, t = zeroOne ? 1 : 0;
  v.img0 = l.img[s]; v.img1 = l.img[t];
  v.mask0 = l.mask[s]; v.mask1 = l.mask[t];
  v.stat0 = c->stats[s]; v.stat1 = c->stats[t];
  v.istat0 = c->istats[s]; v.istat1 = c->istats[t];
  v.W = l.w; v.H = l.h;
  v.img_bytes = l.img_bytes; v.mask_bytes = l.mask_bytes;
  return v;
}

int ensure_stats(sb200_ctx* c, int level) {
  if (c->stats_level == level) return SB200_OK;
  const Level& l = c->lv[level];
  // screened searches only need the integer map; the exact pass evaluates the few windows it touches on the spot.
  // Unscreened configurations (SB200_SCREEN=0, MatchBlockRadius != 2) keep the double map.
  const bool int_only = c->screen && c->R == 2;
  for (int k = 0; k < 2; k++) {
    const int n = int_only ? launch_window_istats(l.img[k], l.w, l.h, c->istats[k], c->st)
                           : launch_window_stats(l.img[k], l.w, l.h, c->R, c->stats[k], c->R == 2 ? c->istats[k] : nullptr, c->st);
    if (n < 0) { c->err = "unsupported MatchBlockRadius (1..3)"; return SB200_ERR_BAD_ARG; }
    c->launches += n;
  }
  c->stats_level = level;
  return SB200_OK;
}

bool degenerate(const Bound& m) { return m.YL >= m.YR || m.XL >= m.XR; }

int run_stage_impl(sb200_ctx* c, int level, int stage) {
  if (!c->uploaded) { c->err = "no pair uploaded"; return SB200_ERR_STATE; }
  if (level < 0 || level >= c->L) { c->err = "level out of range"; return SB200_ERR_BAD_ARG; }
  const Level& l = c->lv[level];
  const int W = l.w, H = l.h;
  StageTimer timer(c, stage, level);
  if (stage == SB200_STAGE_FIND_MARGIN) {  // margins were reduced on the device at upload
    c->cur_margin[0] = l.margin[0];
    c->cur_margin[1] = l.margin[1];
    c->cur_level = level;
    return SB200_OK;
  }
  if (c->cur_level != level) { c->err = "run FindMargin (stage 1) of this level first"; return SB200_ERR_STATE; }
  const Bound m0 = c->cur_margin[0], m1 = c->cur_margin[1];
  // direction d: source margin = margin[d], target margin = margin[1-d]  (margin[!IsZeroOne], margin[IsZeroOne])
  const Bound msrc[2] = {m0, m1}, mtgt[2] = {m1, m0};
  switch (stage) {
    case SB200_STAGE_INITIAL_MATCH: {
      const bool band = c->band && c->screen && c->R == 2 && c->offset == 2 && level > 0 && l.w % 16 == 0;
      if (!band) {  // the band kernel forms its window sums itself; every other search reads the per-level statistics map
        int rc = ensure_stats(c, level);
        if (rc) return rc;
      }
      if (level == 0) {
        for (int d = 0; d < 2; d++) {
          const int n = launch_lowest_match(make_views(c, level, d == 0), msrc[d], mtgt[d], c->R, c->ds[d], (c->screen && c->R == 2) ? &c->ss : nullptr, c->st);
          if (n < 0) { c->err = "unsupported MatchBlockRadius"; return SB200_ERR_BAD_ARG; }
          c->launches += n;
        }
      } else {
        if (c->elem != 8 || c->dw != c->lv[level - 1].w || c->dh != c->lv[level - 1].h) {
          c->err = "HighLevelInitialMatch needs the refined f64 maps of the previous level";
          return SB200_ERR_STATE;
        }
        if (band) {  // direction 1 on the side stream
          CK(cudaEventRecord(c->ev_fork, c->st));
          CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
        }
        for (int d = 0; d < 2; d++) {
          cudaStream_t sd = (band && d == 1) ? c->side : c->st;
          const int n = launch_high_match(make_views(c, level, d == 0), msrc[d], mtgt[d], c->R, c->offset, c->dd[d], c->dw,
                                          c->dh, c->range_lo[d], c->range_hi[d], c->ds[d], (c->screen && c->R == 2) ? (d ? &c->ss1 : &c->ss) : nullptr, band, sd);
          if (n < 0) {
            if (band) {  // join the side stream before leaving (it may be part of a stream capture)
              cudaEventRecord(c->ev_join, c->side);
              cudaStreamWaitEvent(c->st, c->ev_join, 0);
            }
            c->err = "unsupported MatchBlockRadius";
            return SB200_ERR_BAD_ARG;
          }
          c->launches += n;
        }
        if (band) {
          CK(cudaEventRecord(c->ev_join, c->side));
          CK(cudaStreamWaitEvent(c->st, c->ev_join, 0));
        }
      }
      c->dw = W; c->dh = H; c->elem = 2;
      break;
    }
    case SB200_STAGE_SMOOTH:
    case SB200_STAGE_ORDER:
    case SB200_STAGE_UNIQUE_1:
    case SB200_STAGE_REMATCH:
    case SB200_STAGE_UNIQUE_2:
    case SB200_STAGE_MEDIAN:
    case SB200_STAGE_REFINE: {
      if (c->elem != 2 || c->dw != W || c->dh != H) { c->err = "stage needs s16 disparity maps of this level"; return SB200_ERR_STATE; }
      if (stage == SB200_STAGE_SMOOTH) {
        for (int d = 0; d < 2; d++) {
          c->launches += launch_smooth(c->ds[d], c->ds_tmp, W, H, msrc[d], c->st);
          std::swap(c->ds[d], c->ds_tmp);
        }
      } else if (stage == SB200_STAGE_ORDER) {
        for (int d = 0; d < 2; d++) {
          const int n = launch_order(c->ds[d], W, H, msrc[d], c->st);
          if (n < 0) { c->err = "OrderConstraint: margin wider than the kernel supports"; return SB200_ERR_BAD_ARG; }
          c->launches += n;
        }
      } else if (stage == SB200_STAGE_UNIQUE_1 || stage == SB200_STAGE_UNIQUE_2) {  // :456-460
        for (int pass = 0; pass < 3; pass++) {
          const int a = pass & 1;
          const int n = launch_unique_s16(c->ds[a], c->ds[1 - a], W, H, msrc[a], mtgt[a], c->st);
          if (n < 0) { c->err = "UniquenessContraint: margin wider than the kernel supports (16384 px)"; return SB200_ERR_BAD_ARG; }
          c->launches += n;
        }
      } else if (stage == SB200_STAGE_REMATCH) {
        int rc = ensure_stats(c, level);
        if (rc) return rc;
        for (int d = 0; d < 2; d++) {
          if (degenerate(msrc[d])) { c->err = "degenerate margin in SetBoundary_smooth (reference exits)"; return SB200_ERR_DEGENERATE_MARGIN; }
          const PairViews v = make_views(c, level, d == 0);
          c->launches += launch_rematch_bounds(c->ds[d], v.mask0, W, H, msrc[d], mtgt[d], c->BL[d], c->BR[d], c->st);
          const int n = launch_rematch_search(v, msrc[d], c->R, c->BL[d], c->BR[d], c->ds[d], (c->screen && c->R == 2) ? &c->ss : nullptr, c->st);
          if (n < 0) { c->err = "unsupported MatchBlockRadius"; return SB200_ERR_BAD_ARG; }
          c->launches += n;
        }
      } else if (stage == SB200_STAGE_MEDIAN) {
        for (int d = 0; d < 2; d++) {
          c->launches += launch_median(c->ds[d], l.mask[d], c->ds_tmp, W, H, msrc[d], c->st);
          std::swap(c->ds[d], c->ds_tmp);
        }
      } else {  // refine: both directions advance together, refine_T sweeps per launch
        const int it = c->refine_override >= 0 ? c->refine_override : 30 + level * 30;  // :95
        RefineScratch s[2];
        PairViews pv[2];
        short* in[2];
        for (int d = 0; d < 2; d++) {
          s[d] = c->rs[d];
          s[d].A = c->f64buf[2 * d];
          s[d].B = c->f64buf[2 * d + 1];
          s[d].counters = c->rs[0].counters;
          s[d].ev_begin = s[d].ev_end = nullptr;
          pv[d] = make_views(c, level, d == 0);
          in[d] = c->ds[d];
        }
        if (c->profiling) {
          sb200_ctx::Ev ev{};
          ev.stage = 12;
          ev.level = level;
          cudaEventCreate(&ev.a);
          cudaEventCreate(&ev.b);
          s[0].ev_begin = ev.a; s[0].ev_end = ev.b;
          c->events.push_back(ev);
        }
        double* res[2] = {nullptr, nullptr};
        const int n = launch_refine_fused(pv, msrc, in, it, c->ws, c->refine_T, c->refine_variant, c->refine_tma ? 1 : 0, s, res, c->st);
        if (n < 0) { c->err = "too many refinement sweeps"; return SB200_ERR_BAD_ARG; }
        c->launches += n;
        for (int d = 0; d < 2; d++) {
          c->dd[d] = res[d];
          if (c->profiling && msrc[d].width > 2 && msrc[d].height > 2)
            c->sweep_px_iters[level] += (int64_t)it * msrc[d].width * msrc[d].height;
        }
        if (c->profiling) c->sweep_launches[level] += (it + c->refine_T - 1) / c->refine_T;
        c->elem = 8;
      }
      break;
    }
    case SB200_STAGE_UNIQUE_3: {
      if (c->elem != 8 || c->dw != W || c->dh != H) { c->err = "stage needs f64 disparity maps of this level"; return SB200_ERR_STATE; }
      for (int pass = 0; pass < 3; pass++) {
        const int a = pass & 1;
        const int n = launch_unique_f64(c->dd[a], c->dd[1 - a], W, H, msrc[a], mtgt[a], c->st);
        if (n < 0) { c->err = "UniquenessContraint: margin wider than the kernel supports (16384 px)"; return SB200_ERR_BAD_ARG; }
        c->launches += n;
      }
      break;
    }
    default:
      c->err = "unknown stage";
      return SB200_ERR_BAD_ARG;
  }
  CK(cudaGetLastError());
  return SB200_OK;
}

int triangulate_impl(sb200_ctx* c, bool sync) {
  if (!c->calib_set) { c->err = "calibration not set"; return SB200_ERR_STATE; }
  const int level = c->L - 1;
  const Level& l = c->lv[level];
  if (c->elem != 8 || c->dw != l.w || c->dh != l.h) { c->err = "triangulation needs the refined top-level map"; return SB200_ERR_STATE; }
  StageTimer timer(c, 11);
  const double scale = double(c->W0) / c->OW * (1 << level);  // :692
  CloudParams p;
  p.q03 = c->Q[3] * scale; p.q13 = c->Q[7] * scale; p.q23 = c->Q[11] * scale; p.q33 = c->Q[15] * scale;  // :697-698
  p.q32 = c->Q[14];
  memcpy(p.R, c->Rf, sizeof p.R);
  memcpy(p.T, c->Tf, sizeof p.T);
  c->launches += launch_cloud(c->dd[0], l.mask[0], l.img[0], l.w, l.h, l.margin[0], c->erode_ks, p, c->cs, c->xyz, c->bgr,
                              c->pix, c->d_npoints, c->st);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(c->h_npoints, c->d_npoints, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  if (sync) {
    CK(sync_ctx(c));
    c->n_points = *c->h_npoints;
  }
  return SB200_OK;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* sb200_status_string(int s) {
  switch (s) {
    case SB200_OK: return "ok";
    case SB200_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
    case SB200_ERR_BAD_ARG: return "bad argument";
    case SB200_ERR_CUDA: return "CUDA error";
    case SB200_ERR_STATE: return "call order violated";
    case SB200_ERR_DEGENERATE_MARGIN: return "degenerate mask margin";
  }
  return "unknown status";
}

const char* sb200_last_error(const sb200_ctx* c) { return c ? c->err.c_str() : ""; }

int sb200_ctx_create(sb200_ctx** out, int device, int pyrm_num, int lowest_w, int lowest_h, int origin_w, int origin_h,
                     int radius, double ws, int offset) {
  if (!out) return SB200_ERR_BAD_ARG;
  *out = nullptr;
  if (pyrm_num < 1 || pyrm_num > SB_MAX_LEVELS || lowest_w < 8 || lowest_h < 8 || radius < 1 || radius > 2) return SB200_ERR_BAD_ARG;
  // widest row the per-row kernels take (k_unique: 512 chunks of 32 px; s16 column indices everywhere); checked here so
  // that no stage can silently skip its work later
  if (((long)lowest_w << (pyrm_num - 1)) > 16384 || ((long)lowest_h << (pyrm_num - 1)) > 16384) return SB200_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return SB200_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return SB200_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return SB200_ERR_NO_DEVICE;
  sb200_ctx* c = new sb200_ctx();
  c->device = device; c->L = pyrm_num; c->W0 = lowest_w; c->H0 = lowest_h;
  c->OW = origin_w > 0 ? origin_w : lowest_w << (pyrm_num - 1);
  c->OH = origin_h > 0 ? origin_h : lowest_h << (pyrm_num - 1);
  c->R = radius; c->ws = ws; c->offset = offset;
  *out = c;  // returned even on failure so the caller can read sb200_last_error, then destroy
  CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_done, cudaEventBlockingSync | cudaEventDisableTiming));
  c->lv.resize(pyrm_num);
  for (int i = 0; i < pyrm_num; i++) {
    Level& l = c->lv[i];
    l.w = lowest_w << i; l.h = lowest_h << i;
    l.img_bytes = (long)l.w * l.h * 3 + SB_IMG_SLACK;
    l.mask_bytes = (long)l.w * l.h + SB_IMG_SLACK;
    for (int k = 0; k < 2; k++) {
      CK(dalloc(&l.img[k], l.img_bytes));
      CK(dalloc(&l.mask[k], l.mask_bytes));
      CK(cudaMemsetAsync(l.img[k], 0, l.img_bytes, c->st));
      CK(cudaMemsetAsync(l.mask[k], 0, l.mask_bytes, c->st));
    }
  }
  const size_t n = (size_t)c->lv[pyrm_num - 1].w * c->lv[pyrm_num - 1].h;
  const size_t pad = 256;
  CK(dalloc(&c->d_margins, pyrm_num * 8));
  for (int d = 0; d < 2; d++) {
    CK(dalloc(&c->ds[d], n + pad));
    CK(dalloc(&c->BL[d], n + pad));
    CK(dalloc(&c->BR[d], n + pad));
    CK(dalloc(&c->stats[d], n + pad));
    CK(dalloc(&c->istats[d], n + pad));
  }
  CK(dalloc(&c->search_counters, 2));
  CK(dalloc(&c->search_n, 8));
  CK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  for (int d = 0; d < 2; d++) {
    CK(dalloc(&c->search_list[d], n + pad));
    CK(dalloc(&c->search_list2[d], n + pad));
    CK(dalloc(&c->search_listw[d], n + pad));
    CK(dalloc(&c->range_lo[d], n + pad));
    CK(dalloc(&c->range_hi[d], n + pad));
    CK(cudaStreamCreateWithFlags(&c->hole_s[d], cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_hf[d], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_hj[d], cudaEventDisableTiming));
    SearchScratch& q = d ? c->ss1 : c->ss;
    q.list = c->search_list[d]; q.list2 = c->search_list2[d]; q.list_wide = c->search_listw[d]; q.n_list = c->search_n + 4 * d;
    q.cap = (unsigned)n; q.counters = c->search_counters;
    q.side = c->hole_s[d]; q.ev_fork = c->ev_hf[d]; q.ev_join = c->ev_hj[d];
  }
  CK(cudaMemsetAsync(c->search_counters, 0, 2 * sizeof(unsigned long long), c->st));
  if (const char* e = getenv("SB200_SCREEN")) c->screen = atoi(e) != 0;
  if (const char* e = getenv("SB200_BAND")) c->band = atoi(e) != 0;
  if (const char* e = getenv("SB200_GRAPH")) c->use_graph = atoi(e) != 0;
  CK(dalloc(&c->ds_tmp, n + pad));
  for (int k = 0; k < 4; k++) CK(dalloc(&c->f64buf[k], n + pad));
  for (int d = 0; d < 2; d++) {
    CK(dalloc(&c->rs[d].table, (size_t)SB_REFINE_K * n + pad));
    CK(dalloc(&c->rs[d].code, n + pad));
    CK(dalloc(&c->rs[d].miss_count, SB_REFINE_MAX_ITERS));
    CK(dalloc(&c->rs[d].miss_list, n + pad));
    c->rs[d].miss_cap = (unsigned)n;
    c->rs[d].ev_begin = c->rs[d].ev_end = nullptr;
  }
  CK(dalloc(&c->rs[0].counters, 2));
  c->rs[1].counters = c->rs[0].counters;
  CK(cudaMemsetAsync(c->rs[0].counters, 0, 2 * sizeof(unsigned long long), c->st));
  if (const char* e = getenv("SB200_REFINE_T")) c->refine_T = atoi(e) > 0 ? atoi(e) : c->refine_T;
  if (const char* e = getenv("SB200_REFINE_TILE")) c->refine_variant = atoi(e);
  if (c->refine_variant > 7) c->refine_variant = -1;
  if (const char* e = getenv("SB200_REFINE_TMA")) c->refine_tma = atoi(e) != 0;
  CK(dalloc(&c->cs.run, n + pad));
  CK(dalloc(&c->cs.eroded, n + pad));
  CK(dalloc(&c->cs.row_count, (size_t)c->lv[pyrm_num - 1].h + 2));
  CK(dalloc(&c->cs.row_offset, (size_t)c->lv[pyrm_num - 1].h + 2));
  c->erode_ks = (int)ceil(0.02 * c->lv[pyrm_num - 1].h);  // :703
  {
    std::vector<short> j12(2 * (size_t)c->erode_ks);
    sb_ellipse_rows(c->erode_ks, j12.data(), j12.data() + c->erode_ks);
    CK(dalloc(&c->d_ellipse, j12.size()));
    CK(cudaMemcpy(c->d_ellipse, j12.data(), j12.size() * sizeof(short), cudaMemcpyHostToDevice));
    c->cs.ellipse = c->d_ellipse;
  }
  CK(dalloc(&c->xyz, 3 * n));
  CK(dalloc(&c->bgr, 3 * n));
  CK(dalloc(&c->pix, n));
  CK(dalloc(&c->d_npoints, 1));
  CK(cudaMallocHost((void**)&c->h_npoints, sizeof(int)));
  *c->h_npoints = 0;
  CK(sync_ctx(c));
  return SB200_OK;
}

void sb200_ctx_destroy(sb200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->st) cudaStreamSynchronize(c->st);
  for (auto& e : c->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  if (c->gexec) cudaGraphExecDestroy(c->gexec);
  for (auto& l : c->lv)
    for (int k = 0; k < 2; k++) { cudaFree(l.img[k]); cudaFree(l.mask[k]); }
  cudaFree(c->d_margins);
  for (int d = 0; d < 2; d++) { cudaFree(c->ds[d]); cudaFree(c->BL[d]); cudaFree(c->BR[d]); cudaFree(c->stats[d]); cudaFree(c->istats[d]); }
  cudaFree(c->ds_tmp); cudaFree(c->search_counters); cudaFree(c->search_n);
  for (int d = 0; d < 2; d++) {
    cudaFree(c->range_lo[d]); cudaFree(c->range_hi[d]); cudaFree(c->search_list[d]); cudaFree(c->search_list2[d]); cudaFree(c->search_listw[d]);
    if (c->hole_s[d]) { cudaStreamSynchronize(c->hole_s[d]); cudaStreamDestroy(c->hole_s[d]); }
    if (c->ev_hf[d]) cudaEventDestroy(c->ev_hf[d]);
    if (c->ev_hj[d]) cudaEventDestroy(c->ev_hj[d]);
  }
  for (int k = 0; k < 4; k++) cudaFree(c->f64buf[k]);
  for (int d = 0; d < 2; d++) { cudaFree(c->rs[d].table); cudaFree(c->rs[d].code); cudaFree(c->rs[d].miss_count); cudaFree(c->rs[d].miss_list); }
  cudaFree(c->rs[0].counters);
  cudaFree(c->cs.run); cudaFree(c->cs.eroded); cudaFree(c->cs.row_count); cudaFree(c->cs.row_offset);
  cudaFree(c->d_ellipse);
  cudaFree(c->rc_src_img); cudaFree(c->rc_src_mask); cudaFree(c->rc_tab); cudaFree(c->rc_map1); cudaFree(c->rc_map2); cudaFree(c->rc_starts); cudaFree(c->rc_ellipse);
  cudaFree(c->xyz); cudaFree(c->bgr); cudaFree(c->pix); cudaFree(c->d_npoints);
  if (c->h_npoints) cudaFreeHost(c->h_npoints);
  if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->st) cudaStreamDestroy(c->st);
  delete c;
}

// ConstructPyrm (:1040-1053) + FindMargin (:1011-1038) for every level, from the top-level buffers
static int build_pyramid(sb200_ctx* c) {
  for (int k = 0; k < 2; k++)
    for (int i = c->L - 1; i > 0; i--) {
      c->launches += launch_pyrdown(c->lv[i].img[k], c->lv[i].w, c->lv[i].h, 3, c->lv[i - 1].img[k], c->st);
      c->launches += launch_pyrdown(c->lv[i].mask[k], c->lv[i].w, c->lv[i].h, 1, c->lv[i - 1].mask[k], c->st);
    }
  for (int i = 0; i < c->L; i++)
    for (int k = 0; k < 2; k++)
      c->launches += launch_find_margin(c->lv[i].mask[k], c->lv[i].w, c->lv[i].h, c->R, c->d_margins + (i * 2 + k) * 4, c->st);
  CK(cudaGetLastError());
  std::vector<int> hm((size_t)c->L * 8);
  CK(cudaMemcpyAsync(hm.data(), c->d_margins, hm.size() * sizeof(int), cudaMemcpyDeviceToHost, c->st));
  CK(sync_ctx(c));
  for (int i = 0; i < c->L; i++)
    for (int k = 0; k < 2; k++) {
      const int* m = &hm[(size_t)(i * 2 + k) * 4];
      Bound& b = c->lv[i].margin[k];
      b.YL = m[0]; b.YR = m[1]; b.XL = m[2]; b.XR = m[3];
      b.width = b.XR - b.XL + 1;   // :1036-1037
      b.height = b.YR - b.YL + 1;
    }
  c->uploaded = true;
  c->elem = 0; c->dw = c->dh = 0;
  c->stats_level = -1;
  c->cur_level = -1;
  return SB200_OK;
}

static int pair_stage(sb200_ctx* c, const uint8_t* bgr0, const uint8_t* bgr1, const uint8_t* mask0, const uint8_t* mask1,
                      cudaMemcpyKind kind) {
  if (!c || !bgr0 || !bgr1 || !mask0 || !mask1) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  StageTimer timer(c, 0);
  Level& top = c->lv[c->L - 1];
  const size_t npx = (size_t)top.w * top.h;
  const uint8_t* im[2] = {bgr0, bgr1};
  const uint8_t* mk[2] = {mask0, mask1};
  for (int k = 0; k < 2; k++) {
    CK(cudaMemcpyAsync(top.img[k], im[k], npx * 3, kind, c->st));
    CK(cudaMemcpyAsync(top.mask[k], mk[k], npx, kind, c->st));
  }
  return build_pyramid(c);
}

static int rectify_prepare(sb200_ctx* c, size_t src_px) {
  Level& top = c->lv[c->L - 1];
  const size_t n = (size_t)top.w * top.h;
  if (!c->rc_map1) {
    c->rc_ks = 3 * (1 << (c->L - 1));  // :157
    c->rc_levels = 1;
    while ((2 << (c->rc_levels - 1)) <= c->rc_ks) c->rc_levels++;
    CK(dalloc(&c->rc_map1, n));
    CK(dalloc(&c->rc_map2, n));
    CK(dalloc(&c->rc_starts, (size_t)top.h * ((top.w + 31) / 32) * 3));
    CK(dalloc(&c->rc_tab, (size_t)c->rc_levels * n));
    std::vector<short> j12(2 * (size_t)c->rc_ks);
    sb_ellipse_rows(c->rc_ks, j12.data(), j12.data() + c->rc_ks);
    CK(dalloc(&c->rc_ellipse, j12.size()));
    CK(cudaMemcpy(c->rc_ellipse, j12.data(), j12.size() * sizeof(short), cudaMemcpyHostToDevice));
  }
  if (src_px > c->rc_src_cap) {
    cudaFree(c->rc_src_img); cudaFree(c->rc_src_mask);
    c->rc_src_img = c->rc_src_mask = nullptr;
    CK(dalloc(&c->rc_src_img, src_px * 3));
    CK(dalloc(&c->rc_src_mask, src_px));
    c->rc_src_cap = src_px;
  }
  return SB200_OK;
}

// which OpenCV's stereoRectify the calibration half follows by default: the reference links 2.4.5 (SB200_OPENCV_COMPAT=413
// selects the 4.13 behaviour the golden vectors were made with)
static int default_compat() {
  const char* e = getenv("SB200_OPENCV_COMPAT");
  const int v = e ? atoi(e) : 245;
  return v >= 300 ? 413 : 245;
}

int sb200_rectify_calib_compat(const double* K0, const double* Rt0, const double* K1, const double* Rt1, int origin_w, int origin_h,
                               int lowest_w, int pyrm_num, int opencv_compat, double* R_new, double* P_scaled, double* P_final, double* Q,
                               double* R_final, double* T_final) {
  if (!K0 || !Rt0 || !K1 || !Rt1 || !R_new || !P_scaled || !P_final || !Q || !R_final || !T_final || origin_w <= 0 || origin_h <= 0 ||
      lowest_w <= 0 || pyrm_num < 1 || (opencv_compat != 245 && opencv_compat != 413))
    return SB200_ERR_BAD_ARG;
  sb_rectify_calib(K0, Rt0, K1, Rt1, origin_w, origin_h, lowest_w, pyrm_num, R_new, P_scaled, P_final, Q, R_final, T_final, opencv_compat);
  return SB200_OK;
}

int sb200_rectify_calib(const double* K0, const double* Rt0, const double* K1, const double* Rt1, int origin_w, int origin_h,
                        int lowest_w, int pyrm_num, double* R_new, double* P_scaled, double* P_final, double* Q, double* R_final,
                        double* T_final) {
  return sb200_rectify_calib_compat(K0, Rt0, K1, Rt1, origin_w, origin_h, lowest_w, pyrm_num, default_compat(), R_new, P_scaled, P_final, Q,
                                    R_final, T_final);
}

int sb200_stereo_rectify_host(const double* K1, const double* K2, int nx, int ny, const double* R, const double* T, int opencv_compat,
                              double* R1, double* R2, double* P1, double* P2, double* Q) {
  if (!K1 || !K2 || !R || !T || !R1 || !R2 || !P1 || !P2 || !Q || (opencv_compat != 245 && opencv_compat != 413)) return SB200_ERR_BAD_ARG;
  sb_stereo_rectify(K1, K2, nx, ny, R, T, R1, R2, P1, P2, Q, opencv_compat);
  return SB200_OK;
}

int sb200_set_rectify_maps(sb200_ctx* c, const int16_t* map1, const uint16_t* map2) {
  if (!c || !map1 || !map2) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  int rc = rectify_prepare(c, 0);
  if (rc) return rc;
  const Level& top = c->lv[c->L - 1];
  const size_t n = (size_t)top.w * top.h;
  CK(cudaMemcpyAsync(c->rc_map1, map1, n * 4, cudaMemcpyHostToDevice, c->st));
  CK(cudaMemcpyAsync(c->rc_map2, map2, n * 2, cudaMemcpyHostToDevice, c->st));
  CK(sync_ctx(c));
  c->rc_maps_given = true;
  return SB200_OK;
}

int sb200_get_rectify_maps(sb200_ctx* c, int16_t* map1, uint16_t* map2) {
  if (!c || !map1 || !map2) return SB200_ERR_BAD_ARG;
  if (!c->rc_map1) { c->err = "no view rectified yet"; return SB200_ERR_STATE; }
  CK(cudaSetDevice(c->device));
  const Level& top = c->lv[c->L - 1];
  const size_t n = (size_t)top.w * top.h;
  CK(cudaMemcpyAsync(map1, c->rc_map1, n * 4, cudaMemcpyDeviceToHost, c->st));
  CK(cudaMemcpyAsync(map2, c->rc_map2, n * 2, cudaMemcpyDeviceToHost, c->st));
  CK(sync_ctx(c));
  return SB200_OK;
}

int sb200_get_remapped_mask(sb200_ctx* c, uint8_t* out) {
  if (!c || !out) return SB200_ERR_BAD_ARG;
  if (!c->rc_tab) { c->err = "no view rectified yet"; return SB200_ERR_STATE; }
  CK(cudaSetDevice(c->device));
  const Level& top = c->lv[c->L - 1];
  CK(cudaMemcpyAsync(out, c->rc_tab, (size_t)top.w * top.h, cudaMemcpyDeviceToHost, c->st));
  CK(sync_ctx(c));
  return SB200_OK;
}

int sb200_rectify_view(sb200_ctx* c, int view, const uint8_t* src_bgr, const uint8_t* src_mask, int src_w, int src_h, const double* K,
                       const double* R_new, const double* P_scaled, int use_given_maps) {
  if (!c || view < 0 || view > 1 || !src_bgr || !src_mask || src_w <= 0 || src_h <= 0) return SB200_ERR_BAD_ARG;
  if (!use_given_maps && (!K || !R_new || !P_scaled)) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  StageTimer timer(c, 13);
  const size_t spx = (size_t)src_w * src_h;
  int rc = rectify_prepare(c, spx);
  if (rc) return rc;
  Level& top = c->lv[c->L - 1];
  CK(cudaMemcpyAsync(c->rc_src_img, src_bgr, spx * 3, cudaMemcpyHostToDevice, c->st));
  CK(cudaMemcpyAsync(c->rc_src_mask, src_mask, spx, cudaMemcpyHostToDevice, c->st));
  if (use_given_maps) {
    if (!c->rc_maps_given) { c->err = "sb200_set_rectify_maps first"; return SB200_ERR_STATE; }
  } else {
    RectifyView rv;
    if (!sb_rectify_inverse(P_scaled, R_new, rv.iR)) { c->err = "singular rectification matrix"; return SB200_ERR_BAD_ARG; }
    rv.fx = K[0]; rv.fy = K[4]; rv.u0 = K[2]; rv.v0 = K[5];
    c->launches += launch_rectify_maps(top.w, top.h, rv, c->rc_starts, c->rc_map1, c->rc_map2, c->st);  // :144
  }
  c->launches += launch_remap(c->rc_src_img, src_w, src_h, 3, c->rc_map1, c->rc_map2, top.w, top.h, top.img[view], c->st);  // :154
  c->launches += launch_remap(c->rc_src_mask, src_w, src_h, 1, c->rc_map1, c->rc_map2, top.w, top.h, c->rc_tab, c->st);    // :156
  c->launches += launch_erode_ellipse(c->rc_tab, c->rc_levels, top.w, top.h, c->rc_ks, c->rc_ellipse, top.mask[view], c->st);  // :157-158
  CK(cudaGetLastError());
  CK(sync_ctx(c));
  c->uploaded = false;  // the pyramid is stale until sb200_pair_build
  return SB200_OK;
}

int sb200_pair_build(sb200_ctx* c) {
  if (!c) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  StageTimer timer(c, 0);
  return build_pyramid(c);
}

int sb200_pair_upload(sb200_ctx* c, const uint8_t* bgr0, const uint8_t* bgr1, const uint8_t* mask0, const uint8_t* mask1) {
  return pair_stage(c, bgr0, bgr1, mask0, mask1, cudaMemcpyHostToDevice);
}

int sb200_pair_stage_device(sb200_ctx* c, const void* bgr0, const void* bgr1, const void* mask0, const void* mask1) {
  return pair_stage(c, (const uint8_t*)bgr0, (const uint8_t*)bgr1, (const uint8_t*)mask0, (const uint8_t*)mask1,
                    cudaMemcpyDeviceToDevice);
}

int sb200_pair_set_calib(sb200_ctx* c, const double* Q, const double* R_final, const double* T_final) {
  if (!c || !Q || !R_final || !T_final) return SB200_ERR_BAD_ARG;
  memcpy(c->Q, Q, sizeof c->Q);
  memcpy(c->Rf, R_final, sizeof c->Rf);
  memcpy(c->Tf, T_final, sizeof c->Tf);
  c->calib_set = true;
  return SB200_OK;
}

int sb200_run_stage(sb200_ctx* c, int level, int stage) {
  if (!c) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  int rc = run_stage_impl(c, level, stage);
  if (rc) return rc;
  CK(sync_ctx(c));
  return SB200_OK;
}

int sb200_match_one_layer(sb200_ctx* c, int level) {
  if (!c) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  for (int s = 1; s <= 10; s++) {
    int rc = run_stage_impl(c, level, s);
    if (rc) return rc;
  }
  CK(sync_ctx(c));
  return SB200_OK;
}

static int match_pair_async(sb200_ctx* c) {
  for (int i = 0; i < c->L; i++)
    for (int s = 1; s <= 10; s++) {
      int rc = run_stage_impl(c, i, s);
      if (rc) return rc;
    }
  return triangulate_impl(c, false);
}

// match_pair_async captured into a graph and launched once.  Everything match_pair_async enqueues is capturable: kernels,
// memsets, the async copy of the point count, and the event fork / join of K3's side streams.
static int match_pair_graph(sb200_ctx* c) {
  CK(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
  const int rc = match_pair_async(c);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(c->st, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    (void)cudaGetLastError();
    return rc;
  }
  if (e != cudaSuccess || !graph) {
    c->err = std::string("stream capture failed: ") + cudaGetErrorString(e);
    (void)cudaGetLastError();
    return SB200_ERR_CUDA;
  }
  bool fresh = c->gexec == nullptr;
  if (!fresh) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(c->gexec, graph, &info) != cudaSuccess) {  // topology changed (e.g. a level fell back to another path)
      (void)cudaGetLastError();
      cudaGraphExecDestroy(c->gexec);
      c->gexec = nullptr;
      fresh = true;
    }
  }
  if (fresh) {
    const cudaError_t ei = cudaGraphInstantiate(&c->gexec, graph, 0);
    if (ei != cudaSuccess) {
      cudaGraphDestroy(graph);
      c->gexec = nullptr;
      c->err = std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ei);
      return SB200_ERR_CUDA;
    }
    c->graph_rebuilds++;
  }
  cudaGraphDestroy(graph);
  CK(cudaGraphLaunch(c->gexec, c->st));
  c->graph_launches++;
  return SB200_OK;
}

int sb200_match_pair(sb200_ctx* c, int64_t* n_points) {
  if (!c) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  int rc = (c->use_graph && !c->profiling) ? match_pair_graph(c) : match_pair_async(c);
  if (rc) return rc;
  CK(sync_ctx(c));
  c->n_points = *c->h_npoints;
  if (n_points) *n_points = c->n_points;
  return SB200_OK;
}

int sb200_set_refine_iters(sb200_ctx* c, int n) {
  if (!c) return SB200_ERR_BAD_ARG;
  c->refine_override = n;
  return SB200_OK;
}

int sb200_disparity_info(const sb200_ctx* c, int* w, int* h, int* es) {
  if (!c) return SB200_ERR_BAD_ARG;
  if (w) *w = c->dw;
  if (h) *h = c->dh;
  if (es) *es = c->elem;
  return SB200_OK;
}

int sb200_get_disparity(sb200_ctx* c, int dir, void* host_out) {
  if (!c || dir < 0 || dir > 1 || !host_out) return SB200_ERR_BAD_ARG;
  if (c->elem == 0) { c->err = "no disparity yet"; return SB200_ERR_STATE; }
  CK(cudaSetDevice(c->device));
  const size_t n = (size_t)c->dw * c->dh;
  const void* src = c->elem == 2 ? (const void*)c->ds[dir] : (const void*)c->dd[dir];
  CK(cudaMemcpyAsync(host_out, src, n * c->elem, cudaMemcpyDeviceToHost, c->st));
  CK(sync_ctx(c));
  return SB200_OK;
}

int sb200_set_disparity(sb200_ctx* c, int dir, const void* host_in, int width, int height, int es) {
  if (!c || dir < 0 || dir > 1 || !host_in || (es != 2 && es != 8)) return SB200_ERR_BAD_ARG;
  const Level& top = c->lv[c->L - 1];
  if (width <= 0 || height <= 0 || (size_t)width * height > (size_t)top.w * top.h) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  const size_t n = (size_t)width * height;
  if (es == 2) {
    CK(cudaMemcpyAsync(c->ds[dir], host_in, n * 2, cudaMemcpyHostToDevice, c->st));
  } else {
    c->dd[dir] = c->f64buf[2 * dir];
    CK(cudaMemcpyAsync(c->dd[dir], host_in, n * 8, cudaMemcpyHostToDevice, c->st));
  }
  CK(sync_ctx(c));
  c->dw = width; c->dh = height; c->elem = es;
  return SB200_OK;
}

int sb200_get_rematch_bounds(sb200_ctx* c, int dir, int16_t* bl, int16_t* br) {
  if (!c || dir < 0 || dir > 1 || !bl || !br) return SB200_ERR_BAD_ARG;
  if (c->cur_level < 0) { c->err = "no level processed"; return SB200_ERR_STATE; }
  CK(cudaSetDevice(c->device));
  const Level& l = c->lv[c->cur_level];
  const size_t n = (size_t)l.w * l.h;
  CK(cudaMemcpyAsync(bl, c->BL[dir], n * 2, cudaMemcpyDeviceToHost, c->st));
  CK(cudaMemcpyAsync(br, c->BR[dir], n * 2, cudaMemcpyDeviceToHost, c->st));
  CK(sync_ctx(c));
  return SB200_OK;
}

int sb200_get_level(sb200_ctx* c, int level, int view, uint8_t* bgr_out, uint8_t* mask_out) {
  if (!c || level < 0 || level >= c->L || view < 0 || view > 1) return SB200_ERR_BAD_ARG;
  // the top level is readable right after sb200_rectify_view, before the pyramid is built
  if (!c->uploaded && !(level == c->L - 1 && c->rc_map1)) { c->err = "no pair uploaded"; return SB200_ERR_STATE; }
  CK(cudaSetDevice(c->device));
  const Level& l = c->lv[level];
  const size_t n = (size_t)l.w * l.h;
  if (bgr_out) CK(cudaMemcpyAsync(bgr_out, l.img[view], n * 3, cudaMemcpyDeviceToHost, c->st));
  if (mask_out) CK(cudaMemcpyAsync(mask_out, l.mask[view], n, cudaMemcpyDeviceToHost, c->st));
  CK(sync_ctx(c));
  return SB200_OK;
}

int sb200_get_margin(const sb200_ctx* c, int level, int view, sb200_boundary* out) {
  if (!c || level < 0 || level >= c->L || view < 0 || view > 1 || !out) return SB200_ERR_BAD_ARG;
  if (!c->uploaded) return SB200_ERR_STATE;
  const Bound& b = c->lv[level].margin[view];
  out->YL = b.YL; out->YR = b.YR; out->XL = b.XL; out->XR = b.XR; out->width = b.width; out->height = b.height;
  return SB200_OK;
}

int sb200_triangulate(sb200_ctx* c, int64_t* n_points) {
  if (!c) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  int rc = triangulate_impl(c, true);
  if (rc) return rc;
  if (n_points) *n_points = c->n_points;
  return SB200_OK;
}

int sb200_get_points(sb200_ctx* c, double* xyz_out, uint8_t* bgr_out, int32_t* pix_out) {
  if (!c) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  const size_t n = (size_t)c->n_points;
  if (n) {
    if (xyz_out) CK(cudaMemcpyAsync(xyz_out, c->xyz, n * 24, cudaMemcpyDeviceToHost, c->st));
    if (bgr_out) CK(cudaMemcpyAsync(bgr_out, c->bgr, n * 3, cudaMemcpyDeviceToHost, c->st));
    if (pix_out) CK(cudaMemcpyAsync(pix_out, c->pix, n * 4, cudaMemcpyDeviceToHost, c->st));
  }
  CK(sync_ctx(c));
  return SB200_OK;
}

int sb200_points_device(sb200_ctx* c, void** xyz_dev, void** bgr_dev, void** pix_dev, int64_t* n_points) {
  if (!c) return SB200_ERR_BAD_ARG;
  if (xyz_dev) *xyz_dev = c->xyz;
  if (bgr_dev) *bgr_dev = c->bgr;
  if (pix_dev) *pix_dev = c->pix;
  if (n_points) *n_points = c->n_points;
  return SB200_OK;
}

int sb200_match_pair_host(sb200_ctx* c, const uint8_t* bgr0, const uint8_t* bgr1, const uint8_t* mask0, const uint8_t* mask1,
                          const double* Q, const double* R_final, const double* T_final, double* xyz_out, uint8_t* bgr_out,
                          int32_t* pix_out, int64_t capacity, int64_t* n_points) {
  int rc = sb200_pair_upload(c, bgr0, bgr1, mask0, mask1);
  if (rc) return rc;
  rc = sb200_pair_set_calib(c, Q, R_final, T_final);
  if (rc) return rc;
  int64_t n = 0;
  rc = sb200_match_pair(c, &n);
  if (rc) return rc;
  if (n_points) *n_points = n;
  if (n > capacity) { c->err = "point buffers too small"; return SB200_ERR_BAD_ARG; }
  return sb200_get_points(c, xyz_out, bgr_out, pix_out);
}

void* sb200_stream(sb200_ctx* c) { return c ? (void*)c->st : nullptr; }
int64_t sb200_launch_count(const sb200_ctx* c) { return c ? c->launches : 0; }
int sb200_graph_info(const sb200_ctx* c, int64_t* graph_launches, int64_t* graph_instantiations) {
  if (!c) return SB200_ERR_BAD_ARG;
  if (graph_launches) *graph_launches = c->graph_launches;
  if (graph_instantiations) *graph_instantiations = c->graph_rebuilds;
  return SB200_OK;
}

int sb200_set_profiling(sb200_ctx* c, int enable) {
  if (!c) return SB200_ERR_BAD_ARG;
  c->profiling = enable != 0;
  return SB200_OK;
}

int sb200_get_stage_ms(sb200_ctx* c, double* ms16, int reset) {
  if (!c || !ms16) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  CK(sync_ctx(c));
  for (auto& e : c->events) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess && e.stage >= 0 && e.stage < 16) {
      c->stage_ms[e.stage] += ms;
      if (e.level >= 0 && e.level < SB_MAX_LEVELS) c->stage_level_ms[e.stage][e.level] += ms;
      if (e.stage == 12 && e.level >= 0 && e.level < SB_MAX_LEVELS) c->sweep_ms[e.level] += ms;
    } else {
      (void)cudaGetLastError();  // an event pair that was never recorded (no sweep ran)
    }
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  c->events.clear();
  memcpy(ms16, c->stage_ms, sizeof c->stage_ms);
  if (reset) memset(c->stage_ms, 0, sizeof c->stage_ms);
  return SB200_OK;
}

int sb200_get_stage_level_ms(sb200_ctx* c, int stage, int level, double* ms, int reset) {
  if (!c || !ms || stage < 0 || stage >= 16 || level < 0 || level >= SB_MAX_LEVELS) return SB200_ERR_BAD_ARG;
  double tmp[16];
  int rc = sb200_get_stage_ms(c, tmp, 0);
  if (rc) return rc;
  *ms = c->stage_level_ms[stage][level];
  if (reset) c->stage_level_ms[stage][level] = 0;
  return SB200_OK;
}

int sb200_get_refine_profile(sb200_ctx* c, int level, double* sweep_ms, int64_t* sweep_launches, int64_t* px_iters, int reset) {
  if (!c || level >= c->L) return SB200_ERR_BAD_ARG;
  double ms[16];
  int rc = sb200_get_stage_ms(c, ms, 0);
  if (rc) return rc;
  double t = 0;
  int64_t nl = 0, px = 0;
  for (int l = 0; l < c->L; l++) {
    if (level >= 0 && l != level) continue;
    t += c->sweep_ms[l]; nl += c->sweep_launches[l]; px += c->sweep_px_iters[l];
    if (reset) { c->sweep_ms[l] = 0; c->sweep_launches[l] = 0; c->sweep_px_iters[l] = 0; }
  }
  if (sweep_ms) *sweep_ms = t;
  if (sweep_launches) *sweep_launches = nl;
  if (px_iters) *px_iters = px;
  return SB200_OK;
}

int sb200_get_refine_counters(sb200_ctx* c, int64_t* out2, int reset) {
  if (!c || !out2) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  CK(sync_ctx(c));
  unsigned long long h[2];
  CK(cudaMemcpy(h, c->rs[0].counters, sizeof h, cudaMemcpyDeviceToHost));
  out2[1] = (int64_t)h[1];
  if (reset) CK(cudaMemset(c->rs[0].counters, 0, sizeof h));
  CK(cudaMemcpy(h, c->search_counters, sizeof h, cudaMemcpyDeviceToHost));
  out2[0] = (int64_t)h[1];
  if (reset) CK(cudaMemset(c->search_counters, 0, sizeof h));
  return SB200_OK;
}

int sb200_get_search_counters(sb200_ctx* c, int64_t* out2, int reset) {
  if (!c || !out2) return SB200_ERR_BAD_ARG;
  CK(cudaSetDevice(c->device));
  CK(sync_ctx(c));
  unsigned long long h[2];
  CK(cudaMemcpy(h, c->search_counters, sizeof h, cudaMemcpyDeviceToHost));
  out2[0] = (int64_t)h[0];
  out2[1] = (int64_t)h[1];
  if (reset) CK(cudaMemset(c->search_counters, 0, sizeof h));
  return SB200_OK;
}

}  // extern "C"
#pragma GCC visibility pop
