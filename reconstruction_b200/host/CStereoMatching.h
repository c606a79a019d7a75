// CStereoMatching.h — host mirror of the reference's matcher class (reconstruction/CStereoMatching.h:35-69): same
// public members and methods; the per-pair body of MatchAllLayer runs on the GPU through the C ABI
// (include/stereo_b200.h).  `contexts_per_device` contexts per GPU (camera pairs in flight: one stream and one worker
// thread each); the camera pairs are dealt round-robin to the workers and handed to the sink in pair order.
#pragma once
#include <string>
#include <vector>

#include "CCloudOptimization.h"
#include "CManageData.h"

#define NOMATCH -10000  // CStereoMatching.h:9

struct sb200_ctx;

class CStereoMatching {
 public:
  CManageData* m_data = nullptr;
  CCloudOptimization* m_CloudOptimization = nullptr;
  int MatchBlockRadius = 2;
  double m_ws = 0.5;
  int m_offset = 2;
  sbcv::Mat Q, R_final, T_final;  // of the pair processed last (as in the reference)
  Boundary margin[2];
  int Verbose = 1;
  // Initial function, must be called first (CStereoMatching.cpp:5-13)
  void Init(CManageData* data, CCloudOptimization* CloudOptimization, int radii = 2, double ws = 0.5, int disparity_offset = 2);
  void MatchAllLayer();  // CStereoMatching.cpp:15-34

  // --- additions of the mirror (not in the reference) ---
  std::vector<int> devices;        // CUDA devices to use; empty = SB200_DEVICES or every visible device
  int contexts_per_device = 0;     // camera pairs in flight per device; <= 0 = SB200_CTX_PER_DEVICE or 3
  int allgather = -1;              // several devices: collect the per-pair point buffers with the NCCL all-gather of the C ABI
                                   // (sb200_exchange_*) instead of one device-to-host copy per worker; < 0 = SB200_ALLGATHER or 1;
                                   // 2 = take that path with a single device too
  int decode_threads = 0;          // host threads decoding the input files ahead of the GPU; <= 0 = SB200_DECODE_THREADS or the core count (max 16)
  int last_status = 0;             // sb200 status of the last failing call, 0 if none
  std::string last_error;
  double gpu_seconds = 0;          // wall time spent inside the C ABI (all pairs)
  bool FetchPyrm(int CamPair);     // fill m_data->imagePyrm / maskPyrm from the device pyramid of the pair processed last

 private:
  struct PairResult;
  bool Rectify(sb200_ctx* ctx, int CamPair, sbcv::Mat& Q, sbcv::Mat& Rf, sbcv::Mat& Tf, bool& staged_on_device);  // :117-168
  bool RunPair(sb200_ctx* ctx, int device, int CamPair, PairResult& out, bool keep_points_on_device = false);
  bool GatherPairs(std::vector<sb200_ctx*>& ctxs, const std::vector<int>& ctx_dev, int n_dev, std::vector<PairResult>& results);
  sb200_ctx* last_ctx_ = nullptr;
  sbcv::ImagePrefetcher* prefetch_ = nullptr;  // original frames and masks, decoded ahead of the GPU (native Rectify path)
};
