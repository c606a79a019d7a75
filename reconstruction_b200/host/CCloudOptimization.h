// CCloudOptimization.h — stand-in for the sink DLL behind the hot path (import declaration at
// reconstruction/CStereoMatching.h:17-32, implementation CloudOptimization/CCloudOptimization.cpp, PCL-based).  Same public
// methods.  It collects the points the matcher emits; filter(idx) runs the per-pair point processing of the reference's
// filter() on the GPU through the C ABI (sb200_sink_filter: statistical outlier removal, radius normals, orientation towards
// the pair's first camera) and writes tmp/cloud_filter.ply in the PointNormal layout the reference hands to
// its mesher (:119); run() writes the merged raw cloud and the merged oriented cloud.  Meshing (Poisson, MeshLab, texture
// stitching: external Windows executables) is out of scope.
#pragma once
#include <stdint.h>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "CManageData.h"

// std::allocator whose construct() default-initialises: resize() of a vector of doubles / floats / bytes allocates without
// touching the pages.  The per-pair point buffers are sized for the worst case (one point per pixel, 302 MB of doubles at
// 4096x3072) before the pair is matched; value-initialising them would touch every page (more host time than the GPU spends on
// the pair), this way only the rows the device copy writes are ever touched.
template <class T>
struct UninitAllocator : std::allocator<T> {
  template <class U> struct rebind { using other = UninitAllocator<U>; };
  template <class U, class... A> void construct(U* p, A&&... a) {
    if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<A>(a)...);
  }
};
typedef std::vector<float, UninitAllocator<float>> SinkRecords;  // 7 floats per kept point: x y z nx ny nz curvature

class CCloudOptimization {
 public:
  CCloudOptimization() = default;
  ~CCloudOptimization();
  void Init(int sor_meank, double sor_stdThres, int sor_meank1, double sor_stdThres1, double mls_radius, CManageData* m_data,
            bool isdelete_);
  void InsertPoint(sbcv::Mat p);  // 3x1 f64 (CCloudOptimization.cpp:59-62)
  // n InsertPoint calls in one go: xyz = n x 3 f64, bgr = n x 3 u8 (may be null)
  void InsertPoints(const double* xyz, const unsigned char* bgr, size_t n);
  void filter(int idx);  // closes the point range of pair idx (the reference filters + meshes it here)
  // The GPU part of filter(idx) on its own (thread-safe, const): outlier removal + normals + orientation of n points of pair idx
  // on `device`.  The matcher's workers call it right after a pair's triangulation — while other pairs are still matching — and
  // hand the result over with StoreFiltered(); filter(idx) then only appends it (same records, same order, same files).
  bool FilterPoints(int idx, int device, const double* xyz, size_t n, SinkRecords& rec7, size_t& kept, double stats5[5],
                    std::string& err) const;
  void StoreFiltered(int idx, SinkRecords&& rec7, size_t kept, const double stats5[5]);
  // the matcher knows how many points all pairs hold before it hands the first one over: size the merged buffers once
  // (address space only; a refused reservation is not an error, the vectors then grow as before)
  void Reserve(size_t more_points);
  // tmp/cloud_filter.ply is rewritten by every filter() call (:119) and only its last state is ever read: a background thread
  // writes the newest records handed to it and drops the ones a newer pair superseded meanwhile.  Flush waits for the file
  // to hold the last pair's records (run() and the destructor call it).
  void FlushFilterPly();
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
  struct Ready { SinkRecords rec; size_t kept = 0; double stats[5] = {0, 0, 0, 0, 0}; };
  std::map<int, Ready> ready_;  // pairs filtered ahead of filter(idx)
  std::mutex ready_mu_;
  void QueueFilterPly(SinkRecords&& rec7, size_t kept);
  std::thread ply_thread_;
  std::mutex ply_mu_;
  std::condition_variable ply_cv_, ply_idle_;
  SinkRecords ply_pending_;
  size_t ply_pending_kept_ = 0;
  bool ply_have_ = false, ply_writing_ = false, ply_stop_ = false;
};

// cloud<idx>.ply as DisparityToCloud writes it when isoutput is set (CStereoMatching.cpp:707-730,753-757)
bool WritePlyF32(const std::string& path, const double* xyz, const unsigned char* bgr, size_t n);
// pcl::io::savePLYFileBinary of a PointNormal cloud (:119): float x y z normal_x normal_y normal_z curvature
bool WritePlyPointNormal(const std::string& path, const float* rec7, size_t n);
