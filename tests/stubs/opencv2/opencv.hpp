// Stand-in for <opencv2/opencv.hpp> (see tests/stubs/README.md): declarations only, of what VO_utility.h,
// math_utility.h and the two shim sources name; PODs that cross the C ABI have OpenCV's layout.  The classes carry
// just enough state for tests/interpose/fake_cv_core.cpp to implement them (a toy cv::Mat), which is what lets
// tests/test_interpose.py link and RUN the cv:: interposer without OpenCV.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)

namespace cv {
typedef unsigned char uchar;
typedef std::string String;

struct Point2f {
  float x, y;
  Point2f() : x(0), y(0) {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct KeyPoint {  // 28 bytes, as in OpenCV
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
};
struct DMatch {  // 16 bytes, as in OpenCV
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch();
  DMatch(int q, int t, int i, float d);
};

namespace Error {
enum Code { StsError = -2, GpuNotSupported = -216 };
}
class Exception : public std::exception {
 public:
  Exception(int code, const String& err, const String& func, const String& file, int line);
  const char* what() const noexcept override;
  String msg;
};
void error(int code, const String& err, const char* func, const char* file, int line);
#define CV_Assert(expr) \
  do {                  \
    if (!(expr)) ::cv::error(::cv::Error::StsError, #expr, __func__, __FILE__, __LINE__); \
  } while (0)

class MatExpr;
class Mat {
 public:
  int rows, cols;
  uchar* data;
  struct MStep {
    size_t bytes = 0;
    operator size_t() const;
  } step;
  Mat();
  Mat(int rows, int cols, int type);
  Mat(int rows, int cols, int type, void* data, size_t step = 0);
  Mat(const MatExpr& e);
  Mat& operator=(const MatExpr& e);
  int type() const;
  bool empty() const;
  bool isContinuous() const;
  void create(int rows, int cols, int type);
  Mat clone() const;
  Mat rowRange(int start, int end) const;
  template <class T> T& at(int i);
  template <class T> const T& at(int i) const;
  template <class T> T& at(int i, int j);
  template <class T> const T& at(int i, int j) const;
  template <class T> T* ptr(int row = 0);
  template <class T> const T* ptr(int row = 0) const;
  template <class T> void push_back(const T& elem);
  void push_back(const Mat& m);
  static MatExpr eye(int rows, int cols, int type);
  static MatExpr zeros(int rows, int cols, int type);

 private:
  int type_ = 0;
  std::shared_ptr<uchar> owner_;  // toy implementation: tests/interpose/fake_cv_core.cpp
};
class MatExpr {
 public:
  operator Mat() const;
};

// proxy argument types of the public cv:: functions (shim/cv_interpose.cpp defines two of those functions)
class _InputArray {
 public:
  _InputArray();
  _InputArray(const Mat& m);
  template <class T> _InputArray(const std::vector<T>& v);
  Mat getMat(int idx = -1) const;
  bool empty() const;

 protected:
  enum Kind { NONE, MAT, VEC_POINT2F, VEC_INT, VEC_DOUBLE };
  int kind_ = NONE;
  void* obj_ = nullptr;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray();
  _OutputArray(Mat& m);
  template <class T> _OutputArray(std::vector<T>& v);
  bool needed() const;
  void create(int rows, int cols, int type, int i = -1, bool allowTransposed = false, int fixedDepthMask = 0) const;
  void release() const;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
const _OutputArray& noArray();
enum SolvePnPMethod { SOLVEPNP_ITERATIVE = 0, SOLVEPNP_EPNP = 1, SOLVEPNP_P3P = 2 };
void Rodrigues(InputArray src, OutputArray dst, OutputArray jacobian = noArray());
void triangulatePoints(InputArray projMatr1, InputArray projMatr2, InputArray projPoints1, InputArray projPoints2,
                       OutputArray points4D);
bool solvePnPRansac(InputArray objectPoints, InputArray imagePoints, InputArray cameraMatrix, InputArray distCoeffs,
                    OutputArray rvec, OutputArray tvec, bool useExtrinsicGuess = false, int iterationsCount = 100,
                    float reprojectionError = 8.0, double confidence = 0.99, OutputArray inliers = noArray(),
                    int flags = SOLVEPNP_ITERATIVE);
}  // namespace cv
