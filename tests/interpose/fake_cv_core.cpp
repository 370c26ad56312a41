// tests/interpose/fake_cv_core.cpp -- a toy implementation of the stand-in OpenCV classes of tests/stubs (cv::Mat as a
// dense owned / borrowed array, the InputArray / OutputArray proxies, cv::Exception).  Test infrastructure: it plays
// "libopencv_core" so that shim/cv_interpose.cpp and a node-like executable can be LINKED AND RUN without OpenCV
// (tests/test_interpose.py).  Nothing here is on the product path.
#include <opencv2/opencv.hpp>

#include <cstring>

namespace cv {
static int elem_size(int type) {
  static const int depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
  return depth_bytes[type & 7] * ((type >> 3) + 1);
}
DMatch::DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(0) {}
DMatch::DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
Exception::Exception(int, const String& err, const String& func, const String&, int) : msg(func + ": " + err) {}
const char* Exception::what() const noexcept { return msg.c_str(); }
void error(int code, const String& err, const char* func, const char* file, int line) {
  throw Exception(code, err, func, file, line);
}
Mat::MStep::operator size_t() const { return bytes; }
Mat::Mat() : rows(0), cols(0), data(nullptr) {}
Mat::Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr) { create(r, c, type); }
Mat::Mat(int r, int c, int type, void* d, size_t st) : rows(r), cols(c), data((uchar*)d), type_(type) {
  step.bytes = st ? st : (size_t)c * elem_size(type);
}
int Mat::type() const { return type_; }
bool Mat::empty() const { return data == nullptr || rows * cols == 0; }
bool Mat::isContinuous() const { return step.bytes == (size_t)cols * elem_size(type_) || rows <= 1; }
void Mat::create(int r, int c, int type) {
  if (data && r == rows && c == cols && type == type_) return;
  rows = r;
  cols = c;
  type_ = type;
  step.bytes = (size_t)c * elem_size(type);
  const size_t bytes = step.bytes * (size_t)r;
  owner_.reset(bytes ? new uchar[bytes]() : nullptr, std::default_delete<uchar[]>());
  data = owner_.get();
}
Mat Mat::clone() const {
  Mat m(rows, cols, type_);
  for (int i = 0; i < rows; i++) std::memcpy(m.data + i * m.step.bytes, data + i * step.bytes, m.step.bytes);
  return m;
}
template <class T> T& Mat::at(int i) { return ((T*)data)[i]; }
template <class T> const T& Mat::at(int i) const { return ((const T*)data)[i]; }
template <class T> T& Mat::at(int i, int j) { return ((T*)(data + i * step.bytes))[j]; }
template <class T> const T& Mat::at(int i, int j) const { return ((const T*)(data + i * step.bytes))[j]; }
template <class T> T* Mat::ptr(int row) { return (T*)(data + row * step.bytes); }
template <class T> const T* Mat::ptr(int row) const { return (const T*)(data + row * step.bytes); }
#define INST(T)                                   \
  template T& Mat::at<T>(int);                    \
  template const T& Mat::at<T>(int) const;        \
  template T& Mat::at<T>(int, int);               \
  template const T& Mat::at<T>(int, int) const;   \
  template T* Mat::ptr<T>(int);                   \
  template const T* Mat::ptr<T>(int) const;
INST(double)
INST(float)
INST(int)
INST(uchar)

_InputArray::_InputArray() {}
_InputArray::_InputArray(const Mat& m) : kind_(MAT), obj_((void*)&m) {}
template <> _InputArray::_InputArray(const std::vector<Point2f>& v) : kind_(VEC_POINT2F), obj_((void*)&v) {}
template <> _InputArray::_InputArray(const std::vector<int>& v) : kind_(VEC_INT), obj_((void*)&v) {}
template <> _InputArray::_InputArray(const std::vector<double>& v) : kind_(VEC_DOUBLE), obj_((void*)&v) {}
Mat _InputArray::getMat(int) const {
  switch (kind_) {
    case MAT: return *(const Mat*)obj_;
    case VEC_POINT2F: {
      auto& v = *(std::vector<Point2f>*)obj_;
      return v.empty() ? Mat() : Mat((int)v.size(), 1, CV_32FC2, v.data());
    }
    case VEC_INT: {
      auto& v = *(std::vector<int>*)obj_;
      return v.empty() ? Mat() : Mat((int)v.size(), 1, CV_32S, v.data());
    }
    case VEC_DOUBLE: {
      auto& v = *(std::vector<double>*)obj_;
      return v.empty() ? Mat() : Mat((int)v.size(), 1, CV_64F, v.data());
    }
    default: return Mat();
  }
}
bool _InputArray::empty() const { return kind_ == NONE || getMat().empty(); }
_OutputArray::_OutputArray() {}
_OutputArray::_OutputArray(Mat& m) : _InputArray(m) {}
template <> _OutputArray::_OutputArray(std::vector<int>& v) {
  kind_ = VEC_INT;
  obj_ = &v;
}
template <> _OutputArray::_OutputArray(std::vector<double>& v) {
  kind_ = VEC_DOUBLE;
  obj_ = &v;
}
bool _OutputArray::needed() const { return kind_ != NONE; }
void _OutputArray::create(int r, int c, int type, int, bool, int) const {
  if (kind_ == MAT) ((Mat*)obj_)->create(r, c, type);
  else if (kind_ == VEC_INT) ((std::vector<int>*)obj_)->resize((size_t)r * c);
  else if (kind_ == VEC_DOUBLE) ((std::vector<double>*)obj_)->resize((size_t)r * c);
}
void _OutputArray::release() const {}
const _OutputArray& noArray() {
  static _OutputArray none;
  return none;
}
}  // namespace cv
