#include "CCloudOptimization.h"

#include <stdio.h>

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
  open_begin_ = 0;
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

void CCloudOptimization::filter(int idx) {
  pair_index.push_back(idx);
  open_begin_ = xyz.size() / 3;
  pair_begin.push_back(open_begin_);
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
  if (!m_ImageData || m_ImageData->outfilename.empty()) return;
  const size_t n = xyz.size() / 3;
  if (WritePlyF32(m_ImageData->outfilename, xyz.data(), bgr.data(), n))
    printf("wrote %zu points of %zu pairs to %s\n", n, pair_index.size(), m_ImageData->outfilename.c_str());
  else
    printf("cannot write %s\n", m_ImageData->outfilename.c_str());
}
