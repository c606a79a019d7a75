// CCloudOptimization.h — stand-in for the sink DLL behind the hot path (import declaration at
// reconstruction/CStereoMatching.h:17-32, implementation CloudOptimization/CCloudOptimization.cpp, PCL-based).  Same public
// methods.  It collects the points the matcher emits; filter(idx) runs the per-pair point processing of the reference's
// filter() on the GPU through the C ABI (sb200_sink_filter: statistical outlier removal, radius normals, orientation towards
// the pair's first camera) and writes tmp/cloud_filter.ply in the PointNormal layout the reference hands to
// its mesher (:119); run() writes the merged raw cloud and the merged oriented cloud.  Meshing (Poisson, MeshLab, texture
// stitching: external Windows executables) is out of scope.
#pragma once
#include <stdint.h>
#include <map>
#include <mutex>
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
  // The GPU part of filter(idx) on its own (thread-safe, const): outlier removal + normals + orientation of n points of pair idx
  // on `device`.  The matcher's workers call it right after a pair's triangulation — while other pairs are still matching — and
  // hand the result over with StoreFiltered(); filter(idx) then only appends it (same records, same order, same files).
  bool FilterPoints(int idx, int device, const double* xyz, size_t n, std::vector<float>& rec7, size_t& kept, double stats5[5],
                    std::string& err) const;
  void StoreFiltered(int idx, std::vector<float>&& rec7, size_t kept, const double stats5[5]);
  void run();            // writes <outfilename> (binary little-endian PLY: float xyz, uchar b g r) and <outfilename>.normals.ply

  // what the sink holds after the matcher ran
  std::vector<double> xyz;           // 3 per point, reference order
  std::vector<unsigned char> bgr;    // 3 per point
  std::vector<size_t> pair_begin;    // [pair] -> first point; pair_begin.back() closes the last filter()ed pair
  std::vector<int> pair_index;
  // after filter(): kept points with normals, 7 floats each (x y z nx ny nz curvature), all pairs appended (cloud_normals, :117)
  std::vector<float> normals;
  std::vector<size_t> kept_per_pair;
  int sink_device = 0;               // GPU the filter runs on
  bool sink_enabled = true;          // SB200_SINK=0 turns the GPU filter off (points are only collected)
  int last_status = 0;
  std::string last_error;

 private:
  int m_sor_meank = 0;
  double m_mls_radius = 0;
  int m_outrem_neighbor = 0;
  double m_outrem_radius = 0;
  double m_sor_stdThres = 0;
  bool isdelete = false;
  CManageData* m_ImageData = nullptr;
  size_t open_begin_ = 0;
  struct Ready { std::vector<float> rec; size_t kept = 0; double stats[5] = {0, 0, 0, 0, 0}; };
  std::map<int, Ready> ready_;  // pairs filtered ahead of filter(idx)
  std::mutex ready_mu_;
};

// cloud<idx>.ply as DisparityToCloud writes it when isoutput is set (CStereoMatching.cpp:707-730,753-757)
bool WritePlyF32(const std::string& path, const double* xyz, const unsigned char* bgr, size_t n);
// pcl::io::savePLYFileBinary of a PointNormal cloud (:119): float x y z normal_x normal_y normal_z curvature
bool WritePlyPointNormal(const std::string& path, const float* rec7, size_t n);
