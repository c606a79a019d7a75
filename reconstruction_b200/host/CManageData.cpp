#include "CManageData.h"

#include <math.h>
#include <stdio.h>

CManageData::~CManageData() {
  if (imagePyrm != nullptr) {
    for (int i = 0; i < m_PyrmNum; i++) {
      delete[] imagePyrm[i];
      delete[] maskPyrm[i];
    }
    delete[] imagePyrm;
    delete[] maskPyrm;
    imagePyrm = maskPyrm = nullptr;
  }
}

// The reference's tools write Windows paths ("mask\\0001_Cam0.jpg", BatchProcess/main.cpp:68; filepath "D:\\data\\"): on a
// system whose separator is '/', backslashes in the file names of a data set are read as separators.
static std::string native_path(std::string s) {
#ifndef _WIN32
  for (char& ch : s)
    if (ch == '\\') ch = '/';
#endif
  return s;
}

// Keys and order as CManageData::Init (CManageData.cpp:26-76).
bool CManageData::Init(sbcv::FileStorage fs) {
  fs["filepath"] >> m_FilePath;
  m_FilePath = native_path(m_FilePath);
  fs["outfilename"] >> outfilename;
  fs["isoutput"] >> isoutput;
  std::string camera_calib_name;
  fs["camera_calib_name"] >> camera_calib_name;
  int LowestLevelWidth = 0, LowestLevelHeight = 0;
  fs["LowestLevelWidth"] >> LowestLevelWidth;
  fs["LowestLevelHeight"] >> LowestLevelHeight;
  m_LowestLevelSize = sbcv::Size(LowestLevelWidth, LowestLevelHeight);
  sbcv::Mat camID;
  fs["camID"] >> camID;
  if (camID.empty() || camID.cols != 2) {
    printf("config: camID must be an N x 2 matrix\n");
    return false;
  }
  m_CampairNum = camID.rows;
  cam.resize(m_CampairNum);

  std::vector<std::string> imagelist, masklist;
  fs["imagelist"] >> imagelist;
  fs["masklist"] >> masklist;
  for (std::string& s : imagelist) s = native_path(s);
  for (std::string& s : masklist) s = native_path(s);
  camera_calib_name = native_path(camera_calib_name);
  m_CameraNum = (int)imagelist.size();

  sbcv::FileStorage f_calib(m_FilePath + camera_calib_name, sbcv::FileStorage::READ);
  if (f_calib.isOpened() == false) {
    printf("cannot open file %s\n", camera_calib_name.c_str());
    return false;
  }
  for (int i = 0; i < m_CampairNum; i++) {
    cam[i].resize(2);
    for (int k = 0; k < 2; k++) {
      camera& c = cam[i][k];
      c.camID = camID.type() == sbcv::SB_8UC1 ? camID.at<unsigned char>(i, k) : (int)camID.at<double>(i, k);
      if (c.camID < 0 || c.camID >= m_CameraNum || c.camID >= (int)masklist.size()) {
        printf("config: camID %d out of range\n", c.camID);
        return false;
      }
      c.image_name = m_FilePath + imagelist[c.camID];
      c.mask_name = m_FilePath + masklist[c.camID];
      const std::string currentID = std::to_string(c.camID);
      f_calib["intrinsic-" + currentID] >> c.MatIntrinsics;
      f_calib["extrinsic-" + currentID] >> c.MatExtrinsics;
      if (c.MatIntrinsics.empty() || c.MatExtrinsics.empty() || c.MatExtrinsics.cols != 4 || c.MatExtrinsics.rows != 3 ||
          c.MatIntrinsics.cols != 3 || c.MatIntrinsics.rows != 3 ||
          c.MatIntrinsics.type() != sbcv::SB_64FC1 || c.MatExtrinsics.type() != sbcv::SB_64FC1) {
        printf("calibration of camera %d missing or malformed\n", c.camID);
        return false;
      }
      // centre = -R^T t  (CManageData.cpp:61)
      c.CamCenter.create(3, 1, sbcv::SB_64FC1);
      for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int q = 0; q < 3; q++) s += c.MatExtrinsics.at<double>(q, r) * c.MatExtrinsics.at<double>(q, 3);
        c.CamCenter.at<double>(r, 0) = -s;
      }
    }
  }

  fs["PyrmNum"] >> m_PyrmNum;
  if (m_PyrmNum < 1 || LowestLevelWidth <= 0 || LowestLevelHeight <= 0) {
    printf("config: PyrmNum / LowestLevelWidth / LowestLevelHeight invalid\n");
    return false;
  }

  // m_OriginSize = size of the first mask image (CManageData.cpp:68-69); OriginWidth/OriginHeight keys (an
  // extension the stager writes) take precedence so a staged data set needs no original frames.
  int ow = 0, oh = 0;
  fs["OriginWidth"] >> ow;
  fs["OriginHeight"] >> oh;
  if (ow > 0 && oh > 0) {
    m_OriginSize = sbcv::Size(ow, oh);
  } else {
    sbcv::Mat img;
    if (masklist.empty() || !sbcv::imread(m_FilePath + masklist[0], img, true)) {
      printf("cannot read %s to determine the original size\n", masklist.empty() ? "masklist[0]" : masklist[0].c_str());
      return false;
    }
    m_OriginSize = img.size();
  }
  imagePyrm = new sbcv::Mat*[m_PyrmNum];
  maskPyrm = new sbcv::Mat*[m_PyrmNum];
  for (int i = 0; i < m_PyrmNum; i++) {
    imagePyrm[i] = new sbcv::Mat[2];
    maskPyrm[i] = new sbcv::Mat[2];
  }
  return true;
}

// Gather order, two-accumulator sums and the zero-norm rule follow CManageData.cpp:81-90 with Armadillo
// 4.200's mean / norm (arrayops_meat.hpp:902-921, fn_norm.hpp:108-127).
double CManageData::WindowToVec(unsigned char* image_ptr[], int x, int window_size, std::vector<double>& u) {
  const int n = window_size * window_size * 3;
  u.resize(n);
  int k = 0;
  for (int j = x * 3; j < (window_size + x) * 3; j++)
    for (int i = 0; i < window_size; i++) u[k++] = image_ptr[i][j];
  double a1 = 0, a2 = 0;
  int i2 = 0;
  for (; i2 + 1 < n; i2 += 2) { a1 += u[i2]; a2 += u[i2 + 1]; }
  if (i2 < n) a1 += u[i2];
  const double m = (a1 + a2) / double(n);
  for (int i = 0; i < n; i++) u[i] -= m;
  a1 = a2 = 0;
  for (i2 = 0; i2 + 1 < n; i2 += 2) { a1 += u[i2] * u[i2]; a2 += u[i2 + 1] * u[i2 + 1]; }
  if (i2 < n) a1 += u[i2] * u[i2];
  const double normu = sqrt(a1 + a2);
  return normu == 0 ? 1 : normu;
}

bool CManageData::SaveMat(sbcv::Mat input, const char* filename) {
  FILE* fp = fopen(filename, "wb");
  if (fp == NULL) {
    fprintf(stderr, "Create file %s failed...\n", filename);
    return false;
  }
  int v = input.rows;
  fwrite(&v, 4, 1, fp);
  v = input.cols;
  fwrite(&v, 4, 1, fp);
  v = input.channels();
  fwrite(&v, 4, 1, fp);
  v = (int)input.elemSize();
  fwrite(&v, 4, 1, fp);
  for (int r = 0; r < input.rows; r++) fwrite(input.ptr<unsigned char>(r), input.cols * input.elemSize(), 1, fp);
  fclose(fp);
  return true;
}
