#include "CCloudOptimization.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <new>

#include "../../include/stereo_b200.h"

void CCloudOptimization::Init(int sor_meank, double sor_stdThres, int sor_meank1, double sor_stdThres1, double mls_radius,
                              CManageData* m_data, bool isdelete_) {
  m_sor_meank = sor_meank;
  m_sor_stdThres = sor_stdThres;
  m_outrem_neighbor = sor_meank1;
  m_outrem_radius = sor_stdThres1;
  m_mls_radius = mls_radius;
  m_ImageData = m_data;
  isdelete = isdelete_;
  xyz.clear();
  bgr.clear();
  pair_begin.assign(1, 0);
  pair_index.clear();
  normals.clear();
  kept_per_pair.clear();
  open_begin_ = 0;
  if (const char* e = getenv("SB200_SINK")) sink_enabled = atoi(e) != 0;
  mkdir("tmp", 0777);  // "mkdir tmp" (:55-57)
}

void CCloudOptimization::InsertPoint(sbcv::Mat p) {
  if (p.empty() || p.type() != sbcv::SB_64FC1 || p.rows * p.cols != 3) return;
  const double* d = p.ptr<double>(0);
  xyz.insert(xyz.end(), d, d + 3);
  bgr.insert(bgr.end(), 3, (unsigned char)0);
}

void CCloudOptimization::InsertPoints(const double* p, const unsigned char* c, size_t n) {
  xyz.insert(xyz.end(), p, p + 3 * n);
  if (c) bgr.insert(bgr.end(), c, c + 3 * n);
  else bgr.insert(bgr.end(), 3 * n, (unsigned char)0);
}

void CCloudOptimization::Reserve(size_t more) {
  try {
    xyz.reserve(xyz.size() + 3 * more);
    bgr.reserve(bgr.size() + 3 * more);
    if (sink_enabled) normals.reserve(normals.size() + 7 * more);  // at most every point is kept
  } catch (const std::bad_alloc&) {
  }
}

CCloudOptimization::~CCloudOptimization() {
  FlushFilterPly();
  {
    std::lock_guard<std::mutex> lk(ply_mu_);
    ply_stop_ = true;
  }
  ply_cv_.notify_all();
  if (ply_thread_.joinable()) ply_thread_.join();
}

void CCloudOptimization::QueueFilterPly(SinkRecords&& rec, size_t kept) {
  std::lock_guard<std::mutex> lk(ply_mu_);
  ply_pending_ = std::move(rec);  // supersedes records that were still waiting
  ply_pending_kept_ = kept;
  ply_have_ = true;
  if (!ply_thread_.joinable())
    ply_thread_ = std::thread([this]() {
      std::unique_lock<std::mutex> lk(ply_mu_);
      for (;;) {
        ply_cv_.wait(lk, [this] { return ply_have_ || ply_stop_; });
        if (!ply_have_) return;
        SinkRecords rec = std::move(ply_pending_);
        const size_t kept = ply_pending_kept_;
        ply_pending_ = SinkRecords();
        ply_have_ = false;
        ply_writing_ = true;
        lk.unlock();
        if (!WritePlyPointNormal("tmp/cloud_filter.ply", rec.data(), kept)) printf("cannot write tmp/cloud_filter.ply\n");
        rec = SinkRecords();
        lk.lock();
        ply_writing_ = false;
        ply_idle_.notify_all();
      }
    });
  ply_cv_.notify_all();
}

void CCloudOptimization::FlushFilterPly() {
  std::unique_lock<std::mutex> lk(ply_mu_);
  ply_idle_.wait(lk, [this] { return !ply_have_ && !ply_writing_; });
}

bool CCloudOptimization::FilterPoints(int idx, int device, const double* p, size_t n, SinkRecords& rec, size_t& kept, double stats[5],
                                      std::string& err) const {
  kept = 0;
  if (n == 0) return true;
  // CamCenter[idx] = centre of the pair's first camera (:50-51)
  double cam[3] = {0, 0, 0};
  if (m_ImageData && idx >= 0 && idx < (int)m_ImageData->cam.size() && !m_ImageData->cam[idx][0].CamCenter.empty())
    for (int k = 0; k < 3; k++) cam[k] = m_ImageData->cam[idx][0].CamCenter.at<double>(k, 0);
  rec.resize(7 * n);
  // touch the pages before the device writes into them: a device-to-host copy into never-touched pageable memory faults them in
  // inside the driver, measured 0.1 s per 7 M-point pair slower than faulting them here (profiles/r2_handover_ab_v1.json, _v2.json)
  memset(rec.data(), 0, sizeof(float) * 7 * n);
  int64_t k64 = 0;
  const int rc = sb200_sink_filter(device, p, (int64_t)n, m_sor_meank, m_sor_stdThres, m_mls_radius, cam, rec.data(), nullptr, (int64_t)n, &k64, stats);
  if (rc != SB200_OK) {
    err = std::string(sb200_status_string(rc)) + ": " + sb200_sink_last_error();
    rec.clear();
    return false;
  }
  kept = (size_t)k64;
  rec.resize(7 * kept);
  return true;
}

void CCloudOptimization::StoreFiltered(int idx, SinkRecords&& rec, size_t kept, const double stats[5]) {
  std::lock_guard<std::mutex> lk(ready_mu_);
  Ready& r = ready_[idx];
  r.rec = std::move(rec);
  r.kept = kept;
  for (int k = 0; k < 5; k++) r.stats[k] = stats[k];
}

void CCloudOptimization::filter(int idx) {
  const size_t begin = open_begin_, end = xyz.size() / 3;
  pair_index.push_back(idx);
  open_begin_ = end;
  pair_begin.push_back(open_begin_);
  if (!sink_enabled || end <= begin) { kept_per_pair.push_back(0); return; }
  printf("Initial points: %zu\n", end - begin);
  SinkRecords rec;
  size_t kept = 0;
  double stats[5] = {0, 0, 0, 0, 0};
  bool have = false;
  {
    std::lock_guard<std::mutex> lk(ready_mu_);
    auto it = ready_.find(idx);
    if (it != ready_.end()) {  // a matcher worker already ran the GPU filter on exactly these points
      rec = std::move(it->second.rec);
      kept = it->second.kept;
      for (int k = 0; k < 5; k++) stats[k] = it->second.stats[k];
      ready_.erase(it);
      have = true;
    }
  }
  if (!have) {
    std::string err;
    if (!FilterPoints(idx, sink_device, xyz.data() + 3 * begin, end - begin, rec, kept, stats, err)) {
      last_status = SB200_ERR_CUDA;
      last_error = err;
      printf("sink filter of pair %d failed: %s\n", idx, last_error.c_str());
      kept_per_pair.push_back(0);
      return;
    }
  }
  printf("Cloud after filtering: %zu points (mean distance %.6g, stddev %.6g, threshold %.6g; %.2f ms on the GPU)\n", kept, stats[0],
         stats[1], stats[2], stats[3]);
  normals.insert(normals.end(), rec.begin(), rec.begin() + 7 * kept);  // *cloud_normals += *cloud_normal (:117)
  kept_per_pair.push_back(kept);
  QueueFilterPly(std::move(rec), kept);  // :119 (the mesher's input)
}

bool WritePlyPointNormal(const std::string& path, const float* rec7, size_t n) {
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) return false;
  fprintf(fp, "ply\nformat binary_little_endian 1.0\ncomment PCL generated\nelement vertex %zu\n", n);
  fprintf(fp, "property float x\nproperty float y\nproperty float z\nproperty float normal_x\nproperty float normal_y\nproperty float normal_z\n"
              "property float curvature\nend_header\n");
  const bool ok = n == 0 || fwrite(rec7, 28, n, fp) == n;
  fclose(fp);
  return ok;
}

bool WritePlyF32(const std::string& path, const double* xyz, const unsigned char* bgr, size_t n) {
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) return false;
  fprintf(fp, "ply\nformat binary_little_endian 1.0\nelement vertex %zu\n", n);
  fprintf(fp, "property float x\nproperty float y\nproperty float z\nproperty uchar blue\nproperty uchar green\nproperty uchar red\n");
  fprintf(fp, "end_header\n");
  std::vector<unsigned char> rec(15 * 4096);
  for (size_t i = 0; i < n;) {
    const size_t m = n - i < 4096 ? n - i : 4096;
    for (size_t j = 0; j < m; j++) {
      float f[3] = {(float)xyz[3 * (i + j)], (float)xyz[3 * (i + j) + 1], (float)xyz[3 * (i + j) + 2]};
      unsigned char* r = &rec[15 * j];
      __builtin_memcpy(r, f, 12);
      r[12] = bgr ? bgr[3 * (i + j)] : 0;
      r[13] = bgr ? bgr[3 * (i + j) + 1] : 0;
      r[14] = bgr ? bgr[3 * (i + j) + 2] : 0;
    }
    fwrite(rec.data(), 15, m, fp);
    i += m;
  }
  fclose(fp);
  return true;
}

void CCloudOptimization::run() {
  FlushFilterPly();
  if (!m_ImageData || m_ImageData->outfilename.empty()) return;
  const size_t n = xyz.size() / 3;
  if (WritePlyF32(m_ImageData->outfilename, xyz.data(), bgr.data(), n))
    printf("wrote %zu points of %zu pairs to %s\n", n, pair_index.size(), m_ImageData->outfilename.c_str());
  else
    printf("cannot write %s\n", m_ImageData->outfilename.c_str());
  if (!normals.empty()) {
    const std::string path = m_ImageData->outfilename + ".normals.ply";
    if (WritePlyPointNormal(path, normals.data(), normals.size() / 7)) printf("wrote %zu oriented points to %s\n", normals.size() / 7, path.c_str());
    else printf("cannot write %s\n", path.c_str());
  }
}
