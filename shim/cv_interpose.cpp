// shim/cv_interpose.cpp -- optional interposer for the cv:: calls the UVO node makes DIRECTLY, i.e. not through
// uvo_libraries: cv::triangulatePoints (visual_odometry.h:355, :631), cv::solvePnPRansac
// (visual_odometry.h:647-648) and cv::Rodrigues (visual_odometry.h:673).  The real build needs the OpenCV C++ headers;
// here it is type-checked against tests/stubs (tests/test_shim_compiles.py) and -- linked against a toy OpenCV in the
// node's link order -- RUN by tests/test_interpose.py, which checks with LD_DEBUG=bindings that the node's calls bind
// to these definitions and that unsupported argument shapes fall through to the next definition.  See
// INTEGRATION.md, "Unchanged node".
//
// How it works: this object defines the two functions with OpenCV's own signatures, so an executable that resolves
// them against this object first (link order: -luvo_libraries before ${OpenCV_LIBS}, or LD_PRELOAD of the shim
// library) reaches the GPU without a source change in the node.  Each definition takes the GPU route only for the
// argument shapes the node uses -- 3x4 CV_64F projection matrices with N x 1 CV_32FC2 points; N x 3 CV_64F object
// points, zero distortion, no extrinsic guess, SOLVEPNP_EPNP -- and hands every other call to the next definition
// of the same symbol (OpenCV's), found with dlsym(RTLD_NEXT, <own mangled name>).  cv::Rodrigues (:673) takes the
// library's host-only uvo_rodrigues for the node's shape (a 3-vector in, the matrix out, no Jacobian).
// cv::KeyPoint::convert (:616-617, :640) is left alone: it copies points the node already holds in host vectors.
#include <dlfcn.h>

#include <opencv2/opencv.hpp>

#include "uvo_c.h"

namespace {
uvo_ctx* ictx() {  // same lazy process-wide context idea as VO_utility_shim.cpp; nullptr = no GPU: use OpenCV
  static uvo_ctx* c = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    if (uvo_ctx_create(0, nullptr, &c) != UVO_OK) c = nullptr;
  }
  return c;
}
// the next definition of the function `self` in symbol-resolution order (OpenCV's own)
template <class Fn>
Fn next_definition(Fn self) {
  Dl_info info;
  if (!dladdr(reinterpret_cast<void*>(self), &info) || !info.dli_sname) return nullptr;
  return reinterpret_cast<Fn>(dlsym(RTLD_NEXT, info.dli_sname));
}
bool is_points2f(const cv::Mat& m) {  // vector<Point2f> seen through InputArray::getMat()
  return m.type() == CV_32FC2 && (m.rows == 1 || m.cols == 1) && m.isContinuous();
}
bool is_3x4_f64(const cv::Mat& m) { return m.type() == CV_64F && m.rows == 3 && m.cols == 4 && m.isContinuous(); }
bool all_zero_or_empty(const cv::Mat& m) {
  if (m.empty()) return true;
  if (m.type() != CV_64F || !m.isContinuous()) return false;
  const double* p = m.ptr<double>();
  for (int i = 0; i < m.rows * m.cols; i++)
    if (p[i] != 0.0) return false;
  return true;
}
}  // namespace

namespace cv {

void Rodrigues(InputArray src, OutputArray dst, OutputArray jacobian) {
  typedef void (*Fn)(InputArray, OutputArray, OutputArray);
  const Mat r = src.getMat();
  if (!jacobian.needed() && r.type() == CV_64F && r.rows * r.cols == 3 && r.isContinuous()) {
    double R[9];
    if (uvo_rodrigues(r.ptr<double>(), R) == UVO_OK) {
      dst.create(3, 3, CV_64F);
      Mat out = dst.getMat();
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out.at<double>(i, j) = R[3 * i + j];
      return;
    }
  }
  Fn next = next_definition<Fn>(&cv::Rodrigues);
  if (!next) throw Exception(Error::StsError, "uvo_b200: no OpenCV Rodrigues to fall through to", __func__, __FILE__, __LINE__);
  next(src, dst, jacobian);
}

void triangulatePoints(InputArray projMatr1, InputArray projMatr2, InputArray projPoints1, InputArray projPoints2,
                       OutputArray points4D) {
  typedef void (*Fn)(InputArray, InputArray, InputArray, InputArray, OutputArray);
  const Mat P1 = projMatr1.getMat(), P2 = projMatr2.getMat(), a = projPoints1.getMat(), b = projPoints2.getMat();
  const int n = a.rows * a.cols;
  uvo_ctx* c = ictx();
  if (c && is_3x4_f64(P1) && is_3x4_f64(P2) && is_points2f(a) && is_points2f(b) && b.rows * b.cols == n && n > 0) {
    points4D.create(4, n, CV_32F);
    Mat out = points4D.getMat();
    if (out.isContinuous() &&
        uvo_triangulate_points(c, P1.ptr<double>(), P2.ptr<double>(), a.ptr<float>(), b.ptr<float>(), n,
                               out.ptr<float>()) == UVO_OK)
      return;
  }
  Fn next = next_definition<Fn>(&cv::triangulatePoints);
  if (!next) throw Exception(Error::StsError, "uvo_b200: no OpenCV triangulatePoints to fall through to", __func__, __FILE__, __LINE__);
  next(projMatr1, projMatr2, projPoints1, projPoints2, points4D);
}

bool solvePnPRansac(InputArray objectPoints, InputArray imagePoints, InputArray cameraMatrix, InputArray distCoeffs,
                    OutputArray rvec, OutputArray tvec, bool useExtrinsicGuess, int iterationsCount,
                    float reprojectionError, double confidence, OutputArray inliers, int flags) {
  typedef bool (*Fn)(InputArray, InputArray, InputArray, InputArray, OutputArray, OutputArray, bool, int, float, double,
                     OutputArray, int);
  const Mat X = objectPoints.getMat(), x = imagePoints.getMat(), K = cameraMatrix.getMat(), D = distCoeffs.getMat();
  uvo_ctx* c = ictx();
  // fewer than 5 points: OpenCV switches the minimal solver (P3P at 4) or fails -- its business
  if (c && flags == SOLVEPNP_EPNP && !useExtrinsicGuess && all_zero_or_empty(D) && X.type() == CV_64F && X.cols == 3 &&
      X.isContinuous() && X.rows >= 5 && is_points2f(x) && x.rows * x.cols == X.rows && K.type() == CV_64F &&
      K.rows == 3 && K.cols == 3) {
    const double k[4] = {K.at<double>(0, 0), K.at<double>(1, 1), K.at<double>(0, 2), K.at<double>(1, 2)};
    std::vector<int32_t> inl(X.rows);
    int n_inl = 0, hyps = 0;
    double r[3], t[3];
    if (uvo_solve_pnp_ransac(c, X.ptr<double>(), x.ptr<float>(), X.rows, k, iterationsCount, reprojectionError,
                             confidence, r, t, inl.data(), &n_inl, &hyps) == UVO_OK) {
      if (n_inl <= 0) return false;  // as OpenCV: outputs untouched when no model was found
      rvec.create(3, 1, CV_64F);
      tvec.create(3, 1, CV_64F);
      Mat rm = rvec.getMat(), tm = tvec.getMat();
      for (int i = 0; i < 3; i++) {
        rm.at<double>(i) = r[i];
        tm.at<double>(i) = t[i];
      }
      if (inliers.needed()) {
        inliers.create(n_inl, 1, CV_32S);
        Mat im = inliers.getMat();
        for (int i = 0; i < n_inl; i++) im.at<int>(i) = inl[i];
      }
      return true;
    }
  }
  Fn next = next_definition<Fn>(&cv::solvePnPRansac);
  if (!next) throw Exception(Error::StsError, "uvo_b200: no OpenCV solvePnPRansac to fall through to", __func__, __FILE__, __LINE__);
  return next(objectPoints, imagePoints, cameraMatrix, distCoeffs, rvec, tvec, useExtrinsicGuess, iterationsCount,
              reprojectionError, confidence, inliers, flags);
}

}  // namespace cv
