// tests/interpose/fake_cv_calib3d.cpp -- plays "libopencv_calib3d" in tests/test_interpose.py: exports cv::Rodrigues,
// cv::triangulatePoints and cv::solvePnPRansac with OpenCV's exact signatures (hence the same mangled names the real
// library exports), counts its calls and writes recognisable values, so that the test can tell which definition a
// call reached -- the interposer's (ahead in link order) or this one (through dlsym(RTLD_NEXT, ...)).
#include <opencv2/opencv.hpp>

extern "C" {
int fake_calib3d_calls[3] = {0, 0, 0};  // Rodrigues, triangulatePoints, solvePnPRansac
}

namespace cv {
void Rodrigues(InputArray, OutputArray dst, OutputArray) {
  fake_calib3d_calls[0]++;
  dst.create(3, 3, CV_64F);
  Mat R = dst.getMat();
  for (int i = 0; i < 9; i++) R.at<double>(i) = -7.0;
}
void triangulatePoints(InputArray, InputArray, InputArray projPoints1, InputArray, OutputArray points4D) {
  fake_calib3d_calls[1]++;
  const Mat a = projPoints1.getMat();
  const int n = a.rows * a.cols;
  points4D.create(4, n, CV_32F);
  Mat out = points4D.getMat();
  for (int i = 0; i < 4 * n; i++) out.at<float>(i) = 42.f;
}
bool solvePnPRansac(InputArray, InputArray, InputArray, InputArray, OutputArray rvec, OutputArray tvec, bool, int, float,
                    double, OutputArray, int) {
  fake_calib3d_calls[2]++;
  rvec.create(3, 1, CV_64F);
  tvec.create(3, 1, CV_64F);
  Mat r = rvec.getMat(), t = tvec.getMat();
  for (int i = 0; i < 3; i++) r.at<double>(i) = t.at<double>(i) = 42.0;
  return true;
}
}  // namespace cv
