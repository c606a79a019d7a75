// sbcv.h — the few OpenCV types the reference's class surface exposes (cv::Mat, cv::Size,
// cv::FileStorage), restated minimally so the host mirror builds without OpenCV (no C++ OpenCV in
// the build image).  A maintainer dropping the mirror into the reference tree can
// `namespace sbcv = cv` style alias these away: member names follow cv:: (rows, cols, ptr<T>, at<T>).
#pragma once
#include <stdint.h>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace sbcv {

struct Size {
  int width = 0, height = 0;
  Size() {}
  Size(int w, int h) : width(w), height(h) {}
};
inline Size operator*(const Size& s, int k) { return Size(s.width * k, s.height * k); }

enum { SB_8UC1 = 0, SB_8UC3 = 1, SB_16SC1 = 2, SB_64FC1 = 3 };

// Dense row-major matrix, reference-counted like cv::Mat (copies share the payload; clone() copies).
class Mat {
 public:
  int rows = 0, cols = 0;
  uint8_t* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type);
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  int channels() const { return type_ == SB_8UC3 ? 3 : 1; }
  size_t elemSize() const { return type_ == SB_8UC1 ? 1 : type_ == SB_8UC3 ? 3 : type_ == SB_16SC1 ? 2 : 8; }
  Size size() const { return Size(cols, rows); }
  size_t total_bytes() const { return (size_t)rows * cols * elemSize(); }
  template <class T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * cols * elemSize()); }
  template <class T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * cols * elemSize()); }
  template <class T> T& at(int y, int x) { return ptr<T>(y)[x]; }
  template <class T> const T& at(int y, int x) const { return ptr<T>(y)[x]; }
  Mat clone() const;
  void release() { store_.reset(); data = nullptr; rows = cols = 0; }

 private:
  int type_ = SB_8UC1;
  std::shared_ptr<std::vector<uint8_t>> store_;
};

// Reader for the OpenCV "%YAML:1.0" dialect the reference's config.yml / calib_camera.yml use
// (CManageData.cpp:26-66; writer example BatchProcess/main.cpp:47-73): scalars (plain or quoted, .Inf / .Nan), string
// sequences (block "- x" or flow "[a, b]", possibly over several lines), nested mappings and !!opencv-matrix nodes
// (dt u / c / w / s / i / f / d with an optional channel count, e.g. "3u").  Pinned against files written by cv::FileStorage
// (tests/test_host_decode.py).
class FileNode {
 public:
  enum Kind { NONE, SCALAR, SEQ, MATRIX, MAP };
  Kind kind = NONE;
  std::string scalar;
  std::vector<std::string> seq;
  int rows = 0, cols = 0, channels = 1;
  std::string dt;
  std::vector<double> values;
  std::vector<std::string> child_keys;  // MAP: children in file order
  std::vector<FileNode> child_nodes;
  bool empty() const { return kind == NONE; }
  const FileNode& operator[](const std::string& key) const;
};
void operator>>(const FileNode& n, int& v);
void operator>>(const FileNode& n, double& v);
void operator>>(const FileNode& n, std::string& v);
void operator>>(const FileNode& n, std::vector<std::string>& v);
void operator>>(const FileNode& n, Mat& m);  // dt u -> 8UC1, 3u -> 8UC3, any other single-channel type -> 64FC1

class FileStorage {
 public:
  enum { READ = 0 };
  FileStorage() {}
  FileStorage(const std::string& path, int /*flags*/) { open(path); }
  bool open(const std::string& path);
  bool isOpened() const { return opened_; }
  const FileNode& operator[](const std::string& key) const;
  const std::string& error() const { return err_; }
  std::vector<std::string> keys() const;  // top-level keys in file order

 private:
  bool opened_ = false;
  std::string err_;
  FileNode root_;
};

// Binary PNM (P5 grey / P6 colour, maxval 255) <-> Mat.  P6 is RGB on disk and BGR in memory (cv::imread
// order).  imread(..., grayscale=true) converts colour input with the BT.601 weights cv::imread uses.
bool imread_pnm(const std::string& path, Mat& out, bool grayscale);
bool imwrite_pnm(const std::string& path, const Mat& m);

// cv::imread stand-in (sbimg.cpp): JPEG (sequential and progressive Huffman), PNG, BMP and PNM by content; BGR or grey 8-bit output, the same bits
// OpenCV produces for the file.  `grayscale` = the CV_LOAD_IMAGE_GRAYSCALE flag the reference reads its masks with.
bool imread(const std::string& path, Mat& out, bool grayscale);
bool imdecode(const uint8_t* bytes, size_t n, Mat& out, bool grayscale);
const std::string& imread_error();  // why the last imread / imdecode of this thread failed

// Background decoding of a list of image files on a pool of host threads (sbimg.cpp).  The matcher needs two frames and two
// masks per camera pair, adjacent pairs share a camera, and a 12-megapixel JPEG takes a few hundred milliseconds to decode —
// ten times the GPU time of the pair — so the host mirror decodes every distinct file once, ahead of the GPU, in request order.
class ImagePrefetcher {
 public:
  ImagePrefetcher();
  ~ImagePrefetcher();
  // requests: (path, grayscale) in the order they will be needed (duplicates allowed: they share one decode)
  void start(const std::vector<std::pair<std::string, bool>>& requests, int n_threads);
  bool active() const;
  // Blocks until the file has been decoded.  Every get() consumes one of the requests made for (path, grayscale); the cached
  // image is dropped with the last one.  False (with the decoder's message in *err) when decoding failed or the file was
  // never requested.
  bool get(const std::string& path, bool grayscale, Mat& out, std::string* err = nullptr);

 private:
  struct Impl;
  Impl* impl_;
  ImagePrefetcher(const ImagePrefetcher&) = delete;
  ImagePrefetcher& operator=(const ImagePrefetcher&) = delete;
};

}  // namespace sbcv
