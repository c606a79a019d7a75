// CCloudOptimization.h — stand-in for the sink DLL behind the hot path (import declaration at
// reconstruction/CStereoMatching.h:17-32, implementation CloudOptimization/CCloudOptimization.cpp, PCL-based and
// out of scope here).  Same public methods; it collects the points the matcher emits, keeps the
// per-pair ranges filter() would process, and run() writes the merged cloud as PLY.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "CManageData.h"

class CCloudOptimization {
 public:
  void Init(int sor_meank, double sor_stdThres, int sor_meank1, double sor_stdThres1, double mls_radius, CManageData* m_data,
            bool isdelete_);
  void InsertPoint(sbcv::Mat p);  // 3x1 f64 (CCloudOptimization.cpp:59-62)
  // n InsertPoint calls in one go: xyz = n x 3 f64, bgr = n x 3 u8 (may be null)
  void InsertPoints(const double* xyz, const unsigned char* bgr, size_t n);
  void filter(int idx);  // closes the point range of pair idx (the reference filters + meshes it here)
  void run();            // writes <outfilename> (binary little-endian PLY: float xyz, uchar b g r)

  // what the sink holds after the matcher ran
  std::vector<double> xyz;           // 3 per point, reference order
  std::vector<unsigned char> bgr;    // 3 per point
  std::vector<size_t> pair_begin;    // [pair] -> first point; pair_begin.back() closes the last filter()ed pair
  std::vector<int> pair_index;

 private:
  int m_sor_meank = 0;
  double m_mls_radius = 0;
  int m_outrem_neighbor = 0;
  double m_outrem_radius = 0;
  double m_sor_stdThres = 0;
  bool isdelete = false;
  CManageData* m_ImageData = nullptr;
  size_t open_begin_ = 0;
};

// cloud<idx>.ply as DisparityToCloud writes it when isoutput is set (CStereoMatching.cpp:707-730,753-757)
bool WritePlyF32(const std::string& path, const double* xyz, const unsigned char* bgr, size_t n);
