// CManageData.h — host mirror of the reference's data manager (reconstruction/CManageData.h:10-62):
// same struct / member / method names, sbcv:: types standing in for cv:: (see sbcv.h).  Parses config.yml
// and the calibration file exactly as CManageData::Init does (CManageData.cpp:24-79).
#pragma once
#include <string>
#include <vector>

#include "sbcv.h"

struct Boundary {  // CManageData.h:10-14
  int YL, YR, XL, XR;
  int width, height;
};

struct camera {  // CManageData.h:16-27 (bucket, used only by the sink's duplicate removal, is not mirrored)
  sbcv::Mat CamCenter;                       // 3x1 f64 here (the reference converts to CV_32FC1)
  sbcv::Mat P, MatIntrinsics, MatExtrinsics;  // 3x4, 3x3, 3x4 f64
  std::string image_name, mask_name;
  sbcv::Mat image;  // rectified top-level BGR  (filled by CStereoMatching::Rectify)
  sbcv::Mat mask;   // rectified, eroded top-level mask
  Boundary bound;
  int camID;
};

class CManageData {
 public:
  std::vector<std::vector<camera>> cam;  // [pair][0..1]
  int m_CameraNum = 0;
  int m_CampairNum = 0;
  int m_PyrmNum = 0;
  sbcv::Size m_LowestLevelSize;
  std::string m_FilePath;
  int isoutput = 0;
  std::string outfilename;
  sbcv::Size m_OriginSize;
  sbcv::Mat **imagePyrm = nullptr, **maskPyrm = nullptr;  // [level][view]; filled on request (FetchPyrm in CStereoMatching)
  CManageData() {}
  ~CManageData();
  bool Init(sbcv::FileStorage fs);
  // The NCC primitive (CManageData.cpp:81-90), kept for callers outside the hot path (the sink's duplicate
  // test uses it); the matcher itself runs on the GPU and never calls this.
  double WindowToVec(unsigned char* image_ptr[], int x, int window_size, std::vector<double>& u);
  bool SaveMat(sbcv::Mat input, const char* filename);  // CManageData.cpp:94-114, same on-disk layout
};
