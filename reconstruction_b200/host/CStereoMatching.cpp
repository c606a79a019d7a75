#include "CStereoMatching.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <utility>

#include "../../include/stereo_b200.h"

struct CStereoMatching::PairResult {
  bool ok = false;
  int status = 0;
  std::string error;
  sbcv::Mat Q, Rf, Tf;
  Boundary margin[2];
  std::vector<double, UninitAllocator<double>> xyz;
  std::vector<unsigned char, UninitAllocator<unsigned char>> bgr;
  int64_t n = 0;
};

void CStereoMatching::Init(CManageData* data, CCloudOptimization* CloudOptimization, int radii, double ws, int disparity_offset) {
  m_data = data;
  m_CloudOptimization = CloudOptimization;
  MatchBlockRadius = radii;
  m_ws = ws;
  m_offset = disparity_offset;
  Verbose = 1;
}

static std::string staged_name(const std::string& root, int pair, const char* what) {
  char b[64];
  snprintf(b, sizeof b, "staged/pair%d%s", pair, what);
  return root + b;
}

// Rectify (CStereoMatching.cpp:117-168).  Two sources, tried in this order:
//  (1) NATIVE: the original frames named in config.yml (binary PNM) + the calibration of CManageData::Init: the calibration
//      half runs on the host (sb200_rectify_calib: stereoRectify restatement), the image half on the GPU
//      (sb200_rectify_view: initUndistortRectifyMap + remap + erode); the rectified frames stay in HBM and are copied
//      back only to fill cam[pair][k].image / .mask for the sink.
//  (2) STAGED: <filepath>staged/pairN.yml (Q after the sign flip, R_final, T_final, P0, P1) + pairN_view{0,1}.ppm +
//      pairN_mask{0,1}.pgm, i.e. the RESULTS of Rectify written by reconstruction_b200/stage.py (tools/stage_rig.py).
// On success with (1) the context already holds the pyramid (`staged_on_device`).
bool CStereoMatching::Rectify(sb200_ctx* ctx, int CamPair, sbcv::Mat& Qo, sbcv::Mat& Rf, sbcv::Mat& Tf, bool& staged_on_device) {
  if (Verbose >= 1) printf("\trectifying...\n");
  staged_on_device = false;
  std::vector<camera>& cur = m_data->cam[CamPair];
  const sbcv::Size largest = m_data->m_LowestLevelSize * (1 << (m_data->m_PyrmNum - 1));
  sbcv::FileStorage fs(staged_name(m_data->m_FilePath, CamPair, ".yml"), sbcv::FileStorage::READ);
  if (!fs.isOpened()) {  // ---- (1) native ----
    sbcv::Mat src[2], msk[2];
    const bool pf = prefetch_ && prefetch_->active();  // decoded ahead by the host thread pool (MatchAllLayer), else read here
    for (int j = 0; j < 2; j++) {
      if (!(pf ? prefetch_->get(cur[j].image_name, false, src[j]) : sbcv::imread(cur[j].image_name, src[j], false))) {
        printf("read image %s error\n", cur[j].image_name.c_str());  // :147-151
        return false;
      }
      if (!(pf ? prefetch_->get(cur[j].mask_name, true, msk[j]) : sbcv::imread(cur[j].mask_name, msk[j], true)) || msk[j].cols != src[j].cols ||
          msk[j].rows != src[j].rows) {
        printf("read image %s error\n", cur[j].mask_name.c_str());
        return false;
      }
    }
    double R_new[18], P_scaled[24], P_final[24];
    Qo.create(4, 4, sbcv::SB_64FC1);
    Rf.create(3, 3, sbcv::SB_64FC1);
    Tf.create(3, 1, sbcv::SB_64FC1);
    if (sb200_rectify_calib(cur[0].MatIntrinsics.ptr<double>(), cur[0].MatExtrinsics.ptr<double>(), cur[1].MatIntrinsics.ptr<double>(),
                            cur[1].MatExtrinsics.ptr<double>(), m_data->m_OriginSize.width, m_data->m_OriginSize.height,
                            m_data->m_LowestLevelSize.width, m_data->m_PyrmNum, R_new, P_scaled, P_final, Qo.ptr<double>(),
                            Rf.ptr<double>(), Tf.ptr<double>()) != SB200_OK)
      return false;
    for (int j = 0; j < 2; j++) {
      if (sb200_rectify_view(ctx, j, src[j].data, msk[j].data, src[j].cols, src[j].rows, cur[j].MatIntrinsics.ptr<double>(), R_new + 9 * j,
                             P_scaled + 12 * j, 0) != SB200_OK) {
        printf("rectification of pair %d view %d failed: %s\n", CamPair, j, sb200_last_error(ctx));
        return false;
      }
      cur[j].P.create(3, 4, sbcv::SB_64FC1);
      memcpy(cur[j].P.data, P_final + 12 * j, 12 * sizeof(double));
      cur[j].image.create(largest.height, largest.width, sbcv::SB_8UC3);
      cur[j].mask.create(largest.height, largest.width, sbcv::SB_8UC1);
      if (sb200_get_level(ctx, m_data->m_PyrmNum - 1, j, cur[j].image.data, cur[j].mask.data) != SB200_OK) return false;
      if (m_data->isoutput) {  // the reference writes "<pair>_<camID>.jpg" (:159-166); PNM here
        char filename[64];
        snprintf(filename, sizeof filename, "%d_%d.ppm", CamPair, cur[j].camID);
        sbcv::imwrite_pnm(filename, cur[j].image);
      }
    }
    if (sb200_pair_build(ctx) != SB200_OK) return false;
    staged_on_device = true;
    return true;
  }
  // ---- (2) staged ----
  fs["Q"] >> Qo;
  fs["R_final"] >> Rf;
  fs["T_final"] >> Tf;
  fs["P0"] >> cur[0].P;
  fs["P1"] >> cur[1].P;
  if (Qo.empty() || Qo.rows != 4 || Qo.cols != 4 || Rf.empty() || Rf.rows * Rf.cols != 9 || Tf.empty() || Tf.rows * Tf.cols != 3) {
    printf("staged calibration of pair %d is malformed\n", CamPair);
    return false;
  }
  for (int j = 0; j < 2; j++) {
    char suffix[32];
    snprintf(suffix, sizeof suffix, "_view%d.ppm", j);
    if (!sbcv::imread_pnm(staged_name(m_data->m_FilePath, CamPair, suffix), cur[j].image, false)) {
      printf("read image %s error\n", staged_name(m_data->m_FilePath, CamPair, suffix).c_str());
      return false;
    }
    snprintf(suffix, sizeof suffix, "_mask%d.pgm", j);
    if (!sbcv::imread_pnm(staged_name(m_data->m_FilePath, CamPair, suffix), cur[j].mask, true)) {
      printf("read image %s error\n", staged_name(m_data->m_FilePath, CamPair, suffix).c_str());
      return false;
    }
    if (cur[j].image.cols != largest.width || cur[j].image.rows != largest.height || cur[j].mask.cols != largest.width ||
        cur[j].mask.rows != largest.height) {
      printf("staged pair %d view %d is %dx%d, expected %dx%d (LowestLevelSize << (PyrmNum-1))\n", CamPair, j, cur[j].image.cols,
             cur[j].image.rows, largest.width, largest.height);
      return false;
    }
  }
  return true;
}

bool CStereoMatching::RunPair(sb200_ctx* ctx, int device, int CamPair, PairResult& r, bool keep_on_device) {
  bool on_device = false;
  if (!Rectify(ctx, CamPair, r.Q, r.Rf, r.Tf, on_device)) {
    r.status = SB200_ERR_BAD_ARG;
    r.error = "Rectify failed";
    return false;
  }
  std::vector<camera>& cur = m_data->cam[CamPair];
  const size_t cap = (size_t)cur[0].image.rows * cur[0].image.cols;
  const bool want_bgr = m_data->isoutput != 0;  // colours are only used by the PLY branch (:754-756); InsertPoint takes xyz
  if (!keep_on_device) {
    r.xyz.resize(3 * cap);
    if (want_bgr) r.bgr.resize(3 * cap);
  }
  unsigned char* bgr_out = want_bgr ? r.bgr.data() : nullptr;
  // ConstructPyrm, MatchOneLayer x PyrmNum and DisparityToCloud (CStereoMatching.cpp:21-29) on the device
  int rc;
  if (keep_on_device) {  // the points stay in HBM: the caller hands them to the all-gather (sb200_exchange_submit)
    if (on_device) rc = SB200_OK;
    else rc = sb200_pair_upload(ctx, cur[0].image.data, cur[1].image.data, cur[0].mask.data, cur[1].mask.data);
    if (rc == SB200_OK) rc = sb200_pair_set_calib(ctx, r.Q.ptr<double>(), r.Rf.ptr<double>(), r.Tf.ptr<double>());
    if (rc == SB200_OK) rc = sb200_match_pair(ctx, &r.n);
  } else if (on_device) {  // frames were rectified in HBM, the pyramid is built: match and fetch the points
    rc = sb200_pair_set_calib(ctx, r.Q.ptr<double>(), r.Rf.ptr<double>(), r.Tf.ptr<double>());
    if (rc == SB200_OK) rc = sb200_match_pair(ctx, &r.n);
    if (rc == SB200_OK) rc = sb200_get_points(ctx, r.xyz.data(), bgr_out, nullptr);
  } else {
    rc = sb200_match_pair_host(ctx, cur[0].image.data, cur[1].image.data, cur[0].mask.data, cur[1].mask.data, r.Q.ptr<double>(),
                               r.Rf.ptr<double>(), r.Tf.ptr<double>(), r.xyz.data(), bgr_out, nullptr, (int64_t)cap, &r.n);
  }
  if (rc != SB200_OK) {
    r.status = rc;
    r.error = std::string(sb200_status_string(rc)) + ": " + sb200_last_error(ctx);
    return false;
  }
  if (!keep_on_device) {
    r.xyz.resize(3 * (size_t)r.n);
    if (want_bgr) r.bgr.resize(3 * (size_t)r.n);
  }
  for (int k = 0; k < 2; k++) {  // margin[k] of the top level (:27-28)
    sb200_boundary b;
    sb200_get_margin(ctx, m_data->m_PyrmNum - 1, k, &b);
    r.margin[k] = Boundary{b.YL, b.YR, b.XL, b.XR, b.width, b.height};
  }
  // the sink's per-pair filter (outlier removal, normals, orientation) on this worker's device, now, while other pairs are
  // still matching; CCloudOptimization::filter(CamPair) appends the stored result when the pairs are handed over in order
  if (!keep_on_device && m_CloudOptimization && m_CloudOptimization->sink_enabled && r.n > 0) {
    SinkRecords rec;
    size_t kept = 0;
    double stats[5] = {0, 0, 0, 0, 0};
    std::string err;
    if (m_CloudOptimization->FilterPoints(CamPair, device, r.xyz.data(), (size_t)r.n, rec, kept, stats, err))
      m_CloudOptimization->StoreFiltered(CamPair, std::move(rec), kept, stats);
    // on failure filter(CamPair) retries on the sink's own device and reports
  }
  r.ok = true;
  return true;
}

// Several devices, one communicator per device (one host thread each): pair p -> context g = p mod G, context g
// on device g mod n_dev.  A worker matches its pair and only SNAPSHOTS the points for the exchange (sb200_exchange_submit);
// ticket (seq, k) gathers the pairs seq*G + k*n_dev + d of the devices d = 0 .. n_dev-1 in that order - pair order - on every
// device, and device 0's copy is read back once per ticket.  This replaces the reference's serial pair loop feeding the sink
// (CStereoMatching.cpp:17-33 -> CloudOptimization/CCloudOptimization.cpp:123).  false: NCCL unavailable, nothing was done.
bool CStereoMatching::GatherPairs(std::vector<sb200_ctx*>& ctxs, const std::vector<int>& ctx_dev, int n_dev, std::vector<PairResult>& results) {
  const int G = (int)ctxs.size(), P = m_data->m_CampairNum;
  const int per_dev = (G + n_dev - 1) / n_dev;
  unsigned char uid[SB200_UNIQUE_ID_BYTES];
  if (sb200_comm_unique_id(uid) != SB200_OK) {
    printf("point all-gather unavailable (%s): every worker copies its own points instead\n", sb200_comm_last_error(nullptr));
    return false;
  }
  std::vector<sb200_comm*> comms(n_dev, nullptr);
  {
    std::vector<std::thread> th;  // ncclCommInitRank blocks until every rank has joined: one thread per device
    std::vector<int> rcs(n_dev, 0);
    for (int d = 0; d < n_dev; d++)
      th.emplace_back([&, d]() { rcs[d] = sb200_comm_init(&comms[d], ctx_dev[d], d, n_dev, uid, per_dev, 2); });
    for (auto& t : th) t.join();
    for (int d = 0; d < n_dev; d++)
      if (rcs[d] != SB200_OK) {
        printf("point all-gather unavailable (device %d: %s): every worker copies its own points instead\n", ctx_dev[d],
               comms[d] ? sb200_comm_last_error(comms[d]) : sb200_status_string(rcs[d]));
        for (sb200_comm* c : comms) sb200_comm_destroy(c);
        return false;
      }
  }
  sb200_comm_set_consumer(comms[0], 1);  // device 0's results are read back ticket by ticket: do not let the exchange run ahead of that
  const int64_t n_seq = (P + G - 1) / G, n_tickets = n_seq * per_dev;
  std::mutex io;
  std::vector<std::thread> workers;
  for (int g = 0; g < n_dev * per_dev; g++)
    workers.emplace_back([&, g]() {
      const int d = g % n_dev, k = g / n_dev;
      for (int64_t seq = 0; seq < n_seq; seq++) {
        const int64_t p = seq * G + g;
        sb200_ctx* mine = nullptr;
        if (g < G && p < P) {
          { std::lock_guard<std::mutex> lk(io); printf("processing pair %d on GPU %d: cam %d and cam %d...\n", (int)p + 1, ctx_dev[g], m_data->cam[p][0].camID, m_data->cam[p][1].camID); }
          if (RunPair(ctxs[g], ctx_dev[g], (int)p, results[p], true)) mine = ctxs[g];
        }
        // every device takes part in every ticket; a device without a pair (or whose pair failed) contributes no points
        if (sb200_exchange_submit(comms[d], mine, k, seq) != SB200_OK) {
          std::lock_guard<std::mutex> lk(io);
          printf("exchange_submit failed on GPU %d: %s\n", ctx_dev[d], sb200_comm_last_error(comms[d]));
          return;
        }
      }
    });
  // The sink's per-pair filter runs ahead of the ordered hand-over here too: one host thread per device takes the pairs as the
  // collector below receives them and leaves the records with StoreFiltered(); filter(pair) then only appends them.
  std::mutex fq_mu;
  std::condition_variable fq_cv;
  std::deque<int> fq;
  bool fq_done = false;
  std::vector<std::thread> filterers;
  if (m_CloudOptimization && m_CloudOptimization->sink_enabled)
    for (int d = 0; d < n_dev; d++)
      filterers.emplace_back([&, d]() {
        for (;;) {
          int p;
          {
            std::unique_lock<std::mutex> lk(fq_mu);
            fq_cv.wait(lk, [&] { return fq_done || !fq.empty(); });
            if (fq.empty()) return;
            p = fq.front();
            fq.pop_front();
          }
          PairResult& r = results[p];
          SinkRecords rec;
          size_t kept = 0;
          double stats[5] = {0, 0, 0, 0, 0};
          std::string err;
          if (m_CloudOptimization->FilterPoints(p, ctx_dev[d], r.xyz.data(), (size_t)r.n, rec, kept, stats, err))
            m_CloudOptimization->StoreFiltered(p, std::move(rec), kept, stats);
          // on failure filter(p) retries on the sink's own device and reports
        }
      });
  // collector: device 0's gathered buffers, ticket by ticket, split into the pairs by the gathered counts
  const bool want_bgr = m_data->isoutput != 0;
  std::vector<int64_t> counts(n_dev);
  std::vector<double, UninitAllocator<double>> xyz;
  std::vector<unsigned char, UninitAllocator<unsigned char>> bgr;
  const size_t cap_px = (size_t)(m_data->m_LowestLevelSize.width << (m_data->m_PyrmNum - 1)) * (m_data->m_LowestLevelSize.height << (m_data->m_PyrmNum - 1));
  xyz.resize(3 * cap_px * n_dev);
  if (want_bgr) bgr.resize(3 * cap_px * n_dev);
  bool ok = true;
  for (int64_t t = 0; t < n_tickets && ok; t++) {
    int64_t total = 0;
    const int rc = sb200_exchange_wait(comms[0], t, counts.data(), xyz.data(), want_bgr ? bgr.data() : nullptr, nullptr, (int64_t)(cap_px * n_dev), &total);
    if (rc != SB200_OK) {
      printf("point all-gather failed: %s\n", sb200_comm_last_error(comms[0]));
      ok = false;
      break;
    }
    const int64_t seq = t / per_dev, k = t % per_dev;
    size_t off = 0;
    for (int d = 0; d < n_dev; d++) {
      const int64_t p = seq * G + k * n_dev + d;
      const size_t n = (size_t)counts[d];
      if (p < P && k * n_dev + d < G && results[p].ok) {
        results[p].xyz.assign(xyz.begin() + 3 * off, xyz.begin() + 3 * (off + n));
        if (want_bgr) results[p].bgr.assign(bgr.begin() + 3 * off, bgr.begin() + 3 * (off + n));
        results[p].n = (int64_t)n;
        if (!filterers.empty() && n > 0) {
          { std::lock_guard<std::mutex> lk(fq_mu); fq.push_back((int)p); }
          fq_cv.notify_one();
        }
      }
      off += n;
    }
  }
  for (auto& w : workers) w.join();
  { std::lock_guard<std::mutex> lk(fq_mu); fq_done = true; }
  fq_cv.notify_all();
  for (auto& f : filterers) f.join();
  double ms = 0;
  int64_t bytes = 0, nx = 0;
  sb200_comm_stats(comms[0], &ms, &bytes, &nx, 0);
  printf("point all-gather: %lld exchanges over %d GPUs, %.1f MB received per GPU, %.2f ms inside the collectives\n", (long long)nx, n_dev, bytes / 1e6, ms);
  for (sb200_comm* c : comms) sb200_comm_destroy(c);
  if (!ok)
    for (auto& r : results)
      if (r.ok && r.xyz.empty() && r.n > 0) { r.ok = false; r.status = SB200_ERR_CUDA; r.error = "point all-gather failed"; }
  return true;
}

void CStereoMatching::MatchAllLayer() {
  last_status = 0;
  last_error.clear();
  std::vector<int> devs = devices;
  if (devs.empty()) {
    if (const char* e = getenv("SB200_DEVICES")) {
      for (const char* p = e; *p;) {
        devs.push_back(atoi(p));
        const char* c = strchr(p, ',');
        if (!c) break;
        p = c + 1;
      }
    }
  }
  const int P = m_data->m_CampairNum;
  const int L = m_data->m_PyrmNum;
  // contexts: first one per device (without an explicit list, add devices until creation fails), then further contexts
  // round-robin over those devices up to `contexts_per_device` each: every context is one camera pair in flight (own
  // stream, own host thread), so a device overlaps one pair's latency-bound coarse levels and host<->device copies with
  // another pair's top-level sweeps.  Pairs are independent (:17), so the results do not depend on the split.
  std::vector<sb200_ctx*> ctxs;
  std::vector<int> ctx_dev;
  for (int i = 0; devs.empty() ? (i < 64 && (int)ctxs.size() < P) : (i < (int)devs.size()); i++) {
    sb200_ctx* c = nullptr;
    const int dev = devs.empty() ? i : devs[i];
    const int rc = sb200_ctx_create(&c, dev, L, m_data->m_LowestLevelSize.width, m_data->m_LowestLevelSize.height,
                                    m_data->m_OriginSize.width, m_data->m_OriginSize.height, MatchBlockRadius, m_ws, m_offset);
    if (rc != SB200_OK) {
      if (c) { last_error = sb200_last_error(c); sb200_ctx_destroy(c); }
      if (devs.empty() && i > 0) break;  // ran out of devices
      last_status = rc;
      if (last_error.empty()) last_error = sb200_status_string(rc);
      printf("cannot create a GPU context on device %d: %s (there is no CPU path)\n", dev, last_error.c_str());
      for (sb200_ctx* x : ctxs) sb200_ctx_destroy(x);
      return;
    }
    ctxs.push_back(c);
    ctx_dev.push_back(dev);
  }
  const int n_dev = (int)ctxs.size();
  int per_dev = contexts_per_device;
  if (per_dev <= 0) {
    const char* e = getenv("SB200_CTX_PER_DEVICE");
    per_dev = e ? atoi(e) : 3;
  }
  if (per_dev < 1) per_dev = 1;
  for (int j = n_dev; j < n_dev * per_dev && j < P; j++) {
    sb200_ctx* c = nullptr;
    const int dev = ctx_dev[j % n_dev];
    if (sb200_ctx_create(&c, dev, L, m_data->m_LowestLevelSize.width, m_data->m_LowestLevelSize.height, m_data->m_OriginSize.width,
                         m_data->m_OriginSize.height, MatchBlockRadius, m_ws, m_offset) != SB200_OK) {
      if (c) sb200_ctx_destroy(c);
      break;  // not enough memory for another pair in flight: carry on with what there is
    }
    ctxs.push_back(c);
    ctx_dev.push_back(dev);
  }
  const int G = (int)ctxs.size();
  std::vector<PairResult> results(P);
  const auto t0 = std::chrono::steady_clock::now();
  // native Rectify path (no staged/pairN.yml): decode every distinct frame and mask once, in the order the pairs need them, on
  // a pool of host threads that runs ahead of the GPU workers
  sbcv::ImagePrefetcher prefetcher;
  {
    std::vector<std::pair<std::string, bool>> req;
    for (int p = 0; p < P; p++) {
      FILE* staged = fopen(staged_name(m_data->m_FilePath, p, ".yml").c_str(), "rb");
      if (staged) { fclose(staged); continue; }
      for (int j = 0; j < 2; j++) {
        req.emplace_back(m_data->cam[p][j].image_name, false);
        req.emplace_back(m_data->cam[p][j].mask_name, true);
      }
    }
    int nt = decode_threads;
    if (nt <= 0) {
      const char* e = getenv("SB200_DECODE_THREADS");
      nt = e ? atoi(e) : (int)std::thread::hardware_concurrency();
      if (nt > 16) nt = 16;
    }
    if (!req.empty() && nt > 0) prefetcher.start(req, nt);
  }
  prefetch_ = &prefetcher;
  // 0 = off, 1 = with several devices (default), 2 = always (one device too: the same path end to end, for tests on a one-GPU box)
  const int gather_mode = allgather >= 0 ? allgather : (getenv("SB200_ALLGATHER") ? atoi(getenv("SB200_ALLGATHER")) : 1);
  if ((gather_mode == 2 || (gather_mode == 1 && n_dev > 1 && G > 1)) && GatherPairs(ctxs, ctx_dev, n_dev, results)) {
    // the points of every pair reached the host through ONE path, the NCCL all-gather of the C ABI
  } else if (G == 1) {
    for (int p = 0; p < P; p++) {
      printf("processing pair %d: cam %d and cam %d...\n", p + 1, m_data->cam[p][0].camID, m_data->cam[p][1].camID);
      RunPair(ctxs[0], ctx_dev[0], p, results[p]);
    }
  } else {  // pair p -> context p mod G, context j on device j mod n_dev; each worker owns its context
    std::mutex io;
    std::vector<std::thread> workers;
    for (int g = 0; g < G; g++)
      workers.emplace_back([&, g]() {
        for (int p = g; p < P; p += G) {
          { std::lock_guard<std::mutex> lk(io); printf("processing pair %d on GPU %d: cam %d and cam %d...\n", p + 1, ctx_dev[g], m_data->cam[p][0].camID, m_data->cam[p][1].camID); }
          RunPair(ctxs[g], ctx_dev[g], p, results[p]);
        }
      });
    for (auto& w : workers) w.join();
  }
  prefetch_ = nullptr;
  gpu_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // hand the results to the sink in pair order: InsertPoint per point, then filter(pair) (:29-31,751)
  if (m_CloudOptimization) {
    size_t total = 0;
    for (int p = 0; p < P; p++)
      if (results[p].ok) total += (size_t)results[p].n;
    m_CloudOptimization->Reserve(total);
  }
  for (int p = 0; p < P; p++) {
    PairResult& r = results[p];
    if (!r.ok) {
      last_status = r.status;
      last_error = r.error;
      printf("pair %d failed: %s\n", p + 1, r.error.c_str());
      continue;
    }
    Q = r.Q; R_final = r.Rf; T_final = r.Tf;
    margin[0] = r.margin[0];
    margin[1] = r.margin[1];
    m_data->cam[p][0].bound = margin[0];
    m_data->cam[p][1].bound = margin[1];
    if (Verbose >= 1) printf("\tconverting disparity to cloud %d... %lld points\n", p, (long long)r.n);
    if (m_CloudOptimization) m_CloudOptimization->InsertPoints(r.xyz.data(), r.bgr.empty() ? nullptr : r.bgr.data(), (size_t)r.n);
    if (m_data->isoutput) {
      char filename[32];
      snprintf(filename, sizeof filename, "cloud%d.ply", p);
      WritePlyF32(filename, r.xyz.data(), r.bgr.data(), (size_t)r.n);
    }
    if (m_CloudOptimization) m_CloudOptimization->filter(p);
    r.xyz.clear(); r.xyz.shrink_to_fit();
    r.bgr.clear(); r.bgr.shrink_to_fit();
  }
  if (last_ctx_) sb200_ctx_destroy(last_ctx_);
  last_ctx_ = nullptr;
  for (int g = 0; g < G; g++) {  // keep the context that processed the last pair for FetchPyrm
    if (P > 0 && g == (P - 1) % G) last_ctx_ = ctxs[g]; else sb200_ctx_destroy(ctxs[g]);
  }
}

bool CStereoMatching::FetchPyrm(int CamPair) {
  (void)CamPair;
  if (!last_ctx_ || !m_data->imagePyrm) return false;
  for (int i = 0; i < m_data->m_PyrmNum; i++)
    for (int k = 0; k < 2; k++) {
      const int w = m_data->m_LowestLevelSize.width << i, h = m_data->m_LowestLevelSize.height << i;
      m_data->imagePyrm[i][k].create(h, w, sbcv::SB_8UC3);
      m_data->maskPyrm[i][k].create(h, w, sbcv::SB_8UC1);
      if (sb200_get_level(last_ctx_, i, k, m_data->imagePyrm[i][k].data, m_data->maskPyrm[i][k].data) != SB200_OK) return false;
    }
  return true;
}
