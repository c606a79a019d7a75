// TEST INFRASTRUCTURE — not product code.
//
// Minimal stand-in for the OpenCV 2.4.5 C++ API surface that the reference's
// reconstruction/CStereoMatching.cpp and reconstruction/CManageData.cpp touch
// (SURVEY.md §8c "Option A").  It exists so those two files can be compiled
// UNMODIFIED, where they lie under /root/reference, into oracle/_ref/ and used
// as the checker for the CUDA path.  Nothing here is copied from OpenCV; the
// few arithmetic entry points the hot path really needs (pyrDown, erode,
// getStructuringElement) are restated from OpenCV's published behaviour and
// pinned against cv2 4.13 by tests/test_oracle_cpu.py (golden vectors under
// tests/golden/).  Everything that belongs to the un-pinned Rectify boundary
// (stereoRectify, initUndistortRectifyMap, remap, imread) is a loud stub.
#ifndef SHIM_OPENCV_CORE_HPP
#define SHIM_OPENCV_CORE_HPP

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ctime>
#include <cassert>
#include <string>
#include <vector>
#include <memory>
#include <algorithm>
#include <deque>
#include <stdexcept>

typedef unsigned char uchar;
typedef unsigned short ushort;

#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC2 CV_MAKETYPE(CV_8U, 2)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_16SC2 CV_MAKETYPE(CV_16S, 2)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_INTER_LINEAR 1
#define CV_LOAD_IMAGE_GRAYSCALE 0

namespace cv {

typedef std::string String;

[[noreturn]] inline void shim_unsupported(const char* what) {
  std::fprintf(stderr, "cv shim: %s is outside the staged parity boundary (SURVEY.md §8c)\n", what);
  std::abort();
}

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
};
inline Size operator*(const Size& s, int k) { return Size(s.width * k, s.height * k); }

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
};
struct Point {
  int x, y;
  Point(int x_ = -1, int y_ = -1) : x(x_), y(y_) {}
};
struct Range {
  int start, end;
  Range(int s, int e) : start(s), end(e) {}
};
struct Scalar {
  double val[4];
  Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) {
    val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3;
  }
};

enum { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };

inline int shim_depth_size(int depth) {
  switch (depth) {
    case CV_8U: case CV_8S: return 1;
    case CV_16U: case CV_16S: return 2;
    case CV_32S: case CV_32F: return 4;
    case CV_64F: return 8;
  }
  return 0;
}

class Mat {
 public:
  int flags;  // holds the type code only
  int rows, cols;
  uchar* data;
  size_t step;  // bytes between rows
  std::shared_ptr<std::vector<uchar> > owner;

  Mat() : flags(0), rows(0), cols(0), data(0), step(0) {}
  Mat(int r, int c, int type) : flags(0), rows(0), cols(0), data(0), step(0) { create(r, c, type); }
  Mat(int r, int c, int type, const Scalar& s) : flags(0), rows(0), cols(0), data(0), step(0) {
    create(r, c, type);
    *this = s;
  }
  Mat(Size sz, int type) : flags(0), rows(0), cols(0), data(0), step(0) { create(sz.height, sz.width, type); }
  Mat(Size sz, int type, const Scalar& s) : flags(0), rows(0), cols(0), data(0), step(0) {
    create(sz.height, sz.width, type);
    *this = s;
  }
  // header over caller-owned memory
  Mat(int r, int c, int type, void* ext) : flags(type), rows(r), cols(c), data((uchar*)ext), step(0) {
    step = (size_t)c * elemSize();
  }

  int type() const { return flags; }
  int depth() const { return flags & 7; }
  int channels() const { return (flags >> 3) + 1; }
  size_t elemSize() const { return (size_t)shim_depth_size(depth()) * channels(); }
  size_t elemSize1() const { return (size_t)shim_depth_size(depth()); }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return data == 0 || rows == 0 || cols == 0; }
  bool isContinuous() const { return step == (size_t)cols * elemSize(); }

  void create(int r, int c, int type) {
    if (data && rows == r && cols == c && flags == type) return;  // same as cv: no-op
    flags = type;
    rows = r;
    cols = c;
    step = (size_t)c * elemSize();
    owner.reset(new std::vector<uchar>((size_t)r * step + 64, 0));
    data = owner->data();
  }
  void release() {
    owner.reset();
    data = 0;
    rows = cols = 0;
    step = 0;
  }

  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <class T> T* ptr(int y = 0) { return (T*)(data + (ptrdiff_t)y * (ptrdiff_t)step); }
  template <class T> const T* ptr(int y = 0) const { return (const T*)(data + (ptrdiff_t)y * (ptrdiff_t)step); }
  template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
  template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }

  Mat& operator=(const Scalar& s) {
    const int cn = channels();
    for (int y = 0; y < rows; y++)
      for (int x = 0; x < cols; x++)
        for (int k = 0; k < cn; k++) set_elem(y, x * cn + k, s.val[k]);
    return *this;
  }

  double get_elem(int y, int xk) const {
    const uchar* p = data + (size_t)y * step;
    switch (depth()) {
      case CV_8U: return ((const uchar*)p)[xk];
      case CV_8S: return ((const signed char*)p)[xk];
      case CV_16U: return ((const ushort*)p)[xk];
      case CV_16S: return ((const short*)p)[xk];
      case CV_32S: return ((const int*)p)[xk];
      case CV_32F: return ((const float*)p)[xk];
      default: return ((const double*)p)[xk];
    }
  }
  static long shim_round(double v) { return (long)std::nearbyint(v); }
  void set_elem(int y, int xk, double v) {
    uchar* p = data + (size_t)y * step;
    switch (depth()) {
      case CV_8U: { long t = shim_round(v); ((uchar*)p)[xk] = (uchar)(t < 0 ? 0 : t > 255 ? 255 : t); break; }
      case CV_8S: { long t = shim_round(v); ((signed char*)p)[xk] = (signed char)(t < -128 ? -128 : t > 127 ? 127 : t); break; }
      case CV_16U: { long t = shim_round(v); ((ushort*)p)[xk] = (ushort)(t < 0 ? 0 : t > 65535 ? 65535 : t); break; }
      case CV_16S: { long t = shim_round(v); ((short*)p)[xk] = (short)(t < -32768 ? -32768 : t > 32767 ? 32767 : t); break; }
      case CV_32S: ((int*)p)[xk] = (int)shim_round(v); break;
      case CV_32F: ((float*)p)[xk] = (float)v; break;
      default: ((double*)p)[xk] = v; break;
    }
  }

  Mat clone() const {
    Mat m;
    copyTo(m);
    return m;
  }
  void copyTo(Mat& dst) const {
    if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.flags == flags) return;
    Mat out;
    if (dst.data && dst.rows == rows && dst.cols == cols && dst.flags == flags) out = dst;
    else out.create(rows, cols, flags);
    const size_t rb = (size_t)cols * elemSize();
    for (int y = 0; y < rows; y++) std::memcpy(out.ptr(y), ptr(y), rb);
    dst = out;
  }
  void convertTo(Mat& dst, int rtype) const {
    const int cn = channels();
    const int t = CV_MAKETYPE(rtype & 7, cn);
    Mat src = *this;  // keeps the source alive when dst aliases *this
    Mat out;
    if (dst.data && dst.rows == rows && dst.cols == cols && dst.flags == t && dst.data != data) out = dst;
    else out.create(rows, cols, t);
    for (int y = 0; y < rows; y++)
      for (int xk = 0; xk < cols * cn; xk++) out.set_elem(y, xk, src.get_elem(y, xk));
    dst = out;
  }

  Mat operator()(const Range& rr, const Range& cr) const {
    Mat m = *this;
    m.rows = rr.end - rr.start;
    m.cols = cr.end - cr.start;
    m.data = data + (size_t)rr.start * step + (size_t)cr.start * elemSize();
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Range(a, b), Range(0, cols)); }
  Mat colRange(int a, int b) const { return (*this)(Range(0, rows), Range(a, b)); }
  Mat col(int c) const { return colRange(c, c + 1); }
  Mat row(int r) const { return rowRange(r, r + 1); }

  Mat t() const {
    Mat m(cols, rows, flags);
    const size_t es = elemSize();
    for (int y = 0; y < rows; y++)
      for (int x = 0; x < cols; x++) std::memcpy(m.ptr(x) + y * es, ptr(y) + x * es, es);
    return m;
  }
  Mat& operator*=(double k) {
    const int cn = channels();
    for (int y = 0; y < rows; y++)
      for (int xk = 0; xk < cols * cn; xk++) set_elem(y, xk, get_elem(y, xk) * k);
    return *this;
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type, Scalar(0, 0, 0, 0)); }
};

// f64 matrix product, plain left-to-right accumulation per output element.
inline Mat operator*(const Mat& a, const Mat& b) {
  if (a.cols != b.rows || a.depth() != CV_64F || b.depth() != CV_64F) shim_unsupported("Mat*Mat of this shape/type");
  Mat c(a.rows, b.cols, CV_64FC1);
  for (int i = 0; i < a.rows; i++)
    for (int j = 0; j < b.cols; j++) {
      double s = 0;
      for (int k = 0; k < a.cols; k++) s += a.at<double>(i, k) * b.at<double>(k, j);
      c.at<double>(i, j) = s;
    }
  return c;
}
inline Mat operator*(const Mat& a, double k) {
  Mat c = a.clone();
  c *= k;
  return c;
}
inline Mat operator+(const Mat& a, const Mat& b) {
  if (a.rows != b.rows || a.cols != b.cols || a.flags != b.flags) shim_unsupported("Mat+Mat of this shape/type");
  Mat c(a.rows, a.cols, a.flags);
  const int cn = a.channels();
  for (int y = 0; y < a.rows; y++)
    for (int xk = 0; xk < a.cols * cn; xk++) c.set_elem(y, xk, a.get_elem(y, xk) + b.get_elem(y, xk));
  return c;
}
inline Mat operator-(const Mat& a) {
  Mat c(a.rows, a.cols, a.flags);
  const int cn = a.channels();
  for (int y = 0; y < a.rows; y++)
    for (int xk = 0; xk < a.cols * cn; xk++) c.set_elem(y, xk, -a.get_elem(y, xk));
  return c;
}

// ---- FileStorage: the harness never parses YAML through the reference ----
class FileNode {
 public:
  void operator>>(int&) const { shim_unsupported("FileStorage"); }
  void operator>>(std::string&) const { shim_unsupported("FileStorage"); }
  void operator>>(Mat&) const { shim_unsupported("FileStorage"); }
  void operator>>(std::vector<String>&) const { shim_unsupported("FileStorage"); }
};
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
};

// ---- un-pinned Rectify boundary: loud stubs ----
inline Mat imread(const std::string&, int = 1) { shim_unsupported("imread"); }
inline bool imwrite(const std::string&, const Mat&) { shim_unsupported("imwrite"); }
inline void remap(const Mat&, Mat&, const Mat&, const Mat&, int) { shim_unsupported("remap"); }
inline void medianBlur(const Mat&, Mat&, int) { shim_unsupported("medianBlur"); }
inline void stereoRectify(const Mat&, const Mat&, const Mat&, const Mat&, Size, const Mat&, const Mat&, Mat&, Mat&,
                          Mat&, Mat&, Mat&, int, double, Size, Rect*, Rect*) {
  shim_unsupported("stereoRectify");
}
inline void initUndistortRectifyMap(const Mat&, const Mat&, const Mat&, const Mat&, Size, int, Mat&, Mat&) {
  shim_unsupported("initUndistortRectifyMap");
}

// ---- arithmetic the hot path needs (restated; pinned against cv2 in tests) ----

// BORDER_REFLECT_101 index fold (gfedcb|abcdefgh|gfedcba).
inline int shim_reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

// pyrDown for 8-bit images, default dst size ((w+1)/2,(h+1)/2), 5x5 kernel
// [1 4 6 4 1]^2 / 256 with rounding (s+128)>>8, BORDER_REFLECT_101.
inline void pyrDown(const Mat& src_, Mat& dst) {
  Mat src = src_;
  if (src.depth() != CV_8U) shim_unsupported("pyrDown on non-8U");
  const int cn = src.channels();
  const int W = src.cols, H = src.rows;
  const int w = (W + 1) / 2, h = (H + 1) / 2;
  Mat out(h, w, src.type());
  std::vector<int> hrow((size_t)5 * w * cn);
  for (int y = 0; y < h; y++) {
    for (int k = 0; k < 5; k++) {
      const int sy = shim_reflect101(2 * y - 2 + k, H);
      const uchar* s = src.ptr(sy);
      int* hr = &hrow[(size_t)k * w * cn];
      for (int x = 0; x < w; x++) {
        const int x0 = shim_reflect101(2 * x - 2, W), x1 = shim_reflect101(2 * x - 1, W), x2 = 2 * x,
                  x3 = shim_reflect101(2 * x + 1, W), x4 = shim_reflect101(2 * x + 2, W);
        for (int c = 0; c < cn; c++)
          hr[x * cn + c] = s[x0 * cn + c] + 4 * s[x1 * cn + c] + 6 * s[x2 * cn + c] + 4 * s[x3 * cn + c] + s[x4 * cn + c];
      }
    }
    uchar* d = out.ptr(y);
    for (int i = 0; i < w * cn; i++) {
      const int v = hrow[i] + 4 * hrow[(size_t)w * cn + i] + 6 * hrow[(size_t)2 * w * cn + i] +
                    4 * hrow[(size_t)3 * w * cn + i] + hrow[(size_t)4 * w * cn + i];
      d[i] = (uchar)((v + 128) >> 8);
    }
  }
  dst = out;
}

// Structuring element; ellipse rows are [c-dx, c+dx] with
// dx = round(c*sqrt((r*r-dy*dy)/r^2)), r=h/2, c=w/2, dy=i-r.
inline Mat getStructuringElement(int shape, Size ksize, Point anchor = Point(-1, -1)) {
  int r = 0, c = 0;
  double inv_r2 = 0;
  if (anchor.x == -1) anchor.x = ksize.width / 2;
  if (anchor.y == -1) anchor.y = ksize.height / 2;
  if (ksize.width == 1 && ksize.height == 1) shape = MORPH_RECT;
  if (shape == MORPH_ELLIPSE) {
    r = ksize.height / 2;
    c = ksize.width / 2;
    inv_r2 = r ? 1. / ((double)r * r) : 0;
  }
  Mat elem(ksize.height, ksize.width, CV_8UC1);
  for (int i = 0; i < ksize.height; i++) {
    uchar* p = elem.ptr(i);
    int j1 = 0, j2 = 0;
    if (shape == MORPH_RECT || (shape == MORPH_CROSS && i == anchor.y)) j2 = ksize.width;
    else if (shape == MORPH_CROSS) { j1 = anchor.x; j2 = j1 + 1; }
    else {
      const int dy = i - r;
      if (std::abs(dy) <= r) {
        const int dx = (int)std::nearbyint(c * std::sqrt((r * r - dy * dy) * inv_r2));
        j1 = std::max(c - dx, 0);
        j2 = std::min(c + dx + 1, ksize.width);
      }
    }
    for (int j = 0; j < ksize.width; j++) p[j] = (uchar)(j >= j1 && j < j2);
  }
  return elem;
}

// Grey-scale erosion, anchor at the element centre, pixels outside the image
// do not constrain the minimum (cv's default morphology border).
inline void erode(const Mat& src_, Mat& dst, const Mat& elem) {
  Mat src = src_;
  if (src.type() != CV_8UC1) shim_unsupported("erode on non-8UC1");
  const int W = src.cols, H = src.rows;
  const int ax = elem.cols / 2, ay = elem.rows / 2;
  Mat out(H, W, CV_8UC1, Scalar(255));
  std::vector<uchar> hmin(W);
  for (int i = 0; i < elem.rows; i++) {
    const uchar* e = elem.ptr(i);
    int j1 = -1, j2 = -1;
    for (int j = 0; j < elem.cols; j++)
      if (e[j]) { if (j1 < 0) j1 = j; j2 = j; }
    if (j1 < 0) continue;
    for (int j = j1; j <= j2; j++)
      if (!e[j]) shim_unsupported("erode with non-convex element rows");
    const int lo = j1 - ax, hi = j2 - ax;  // window [x+lo, x+hi]
    for (int y = 0; y < H; y++) {
      const int sy = y + i - ay;
      if (sy < 0 || sy >= H) continue;
      const uchar* s = src.ptr(sy);
      // sliding-window minimum with a monotone deque
      std::deque<int> dq;
      int next = 0;
      for (int x = 0; x < W; x++) {
        const int a = std::max(x + lo, 0), b = std::min(x + hi, W - 1);
        while (next <= b) {
          while (!dq.empty() && s[dq.back()] >= s[next]) dq.pop_back();
          dq.push_back(next++);
        }
        while (!dq.empty() && dq.front() < a) dq.pop_front();
        hmin[x] = (a <= b && !dq.empty()) ? s[dq.front()] : 255;
      }
      uchar* d = out.ptr(y);
      for (int x = 0; x < W; x++) d[x] = std::min(d[x], hmin[x]);
    }
  }
  dst = out;
}

}  // namespace cv

#endif
