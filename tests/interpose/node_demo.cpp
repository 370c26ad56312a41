// tests/interpose/node_demo.cpp -- makes the three cv:: calls the UNCHANGED node makes by itself, with the node's
// argument shapes (visual_odometry.h:631 triangulatePoints, :647-648 solvePnPRansac, :673 Rodrigues), then the same
// functions with argument shapes the interposer does not take (CV_32F projection matrices, SOLVEPNP_ITERATIVE), and
// prints one line per call saying which definition answered.  Linked like the node: the interposer library ahead of
// "OpenCV" (tests/interpose/fake_cv_calib3d.cpp).  Usage: node_demo <n_points>
#include <opencv2/opencv.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

extern "C" int fake_calib3d_calls[3];
using namespace cv;

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 200;
  // a synthetic two-view / PnP scene (left camera at the origin, right camera 0.33 m to the side)
  const double fx = 1300, fy = 1300, cx = 640, cy = 512;
  Mat K(3, 3, CV_64F), P1(3, 4, CV_64F), P2(3, 4, CV_64F), dist(4, 1, CV_64F);
  const double k[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
  for (int i = 0; i < 9; i++) K.at<double>(i) = k[i];
  for (int i = 0; i < 4; i++) dist.at<double>(i) = 0;
  const double tR[3] = {-0.33, 0, 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 4; j++) {
      P1.at<double>(i, j) = j < 3 ? k[3 * i + j] : 0.0;
      P2.at<double>(i, j) = j < 3 ? k[3 * i + j] : k[3 * i] * tR[0] + k[3 * i + 1] * tR[1] + k[3 * i + 2] * tR[2];
    }
  std::vector<Point2f> x1(n), x2(n);
  Mat X(n, 3, CV_64F);
  unsigned s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) / 16777216.0; };
  for (int i = 0; i < n; i++) {
    const double p[3] = {rnd() * 6 - 3, rnd() * 4 - 2, 4 + rnd() * 5};
    for (int c = 0; c < 3; c++) X.at<double>(i, c) = p[c];
    x1[i] = Point2f((float)(fx * p[0] / p[2] + cx), (float)(fy * p[1] / p[2] + cy));
    x2[i] = Point2f((float)(fx * (p[0] + tR[0]) / p[2] + cx), (float)(fy * p[1] / p[2] + cy));
  }
  int before[3];
  auto snap = [&] { for (int i = 0; i < 3; i++) before[i] = fake_calib3d_calls[i]; };
  auto who = [&](int i) { return fake_calib3d_calls[i] != before[i] ? "opencv" : "interposer"; };

  // visual_odometry.h:631
  snap();
  Mat X4;
  triangulatePoints(P1, P2, x1, x2, X4);
  double err = 0;
  for (int i = 0; i < n && X4.rows == 4; i++)
    for (int c = 0; c < 3; c++) err = fmax(err, fabs(X4.at<float>(c, i) / X4.at<float>(3, i) - X.at<double>(i, c)));
  printf("triangulatePoints node-shape: %s maxerr=%.3g\n", who(1), err);
  // visual_odometry.h:647-648
  snap();
  Mat rvec, tvec, inliers;
  const bool ok = solvePnPRansac(X, x1, K, dist, rvec, tvec, false, 1000, 1.0f, 0.99, inliers, SOLVEPNP_EPNP);
  printf("solvePnPRansac node-shape: %s ok=%d inliers=%d |r|+|t|=%.3g\n", who(2), (int)ok, inliers.rows,
         ok ? fabs(rvec.at<double>(0)) + fabs(rvec.at<double>(1)) + fabs(rvec.at<double>(2)) + fabs(tvec.at<double>(0)) +
                  fabs(tvec.at<double>(1)) + fabs(tvec.at<double>(2)) : -1.0);
  // visual_odometry.h:673
  snap();
  Mat rv(3, 1, CV_64F), R;
  rv.at<double>(0) = 0.01;
  rv.at<double>(1) = -0.02;
  rv.at<double>(2) = 0.015;
  Rodrigues(rv, R);
  printf("Rodrigues node-shape: %s R00=%.15g R01=%.15g\n", who(0), R.at<double>(0, 0), R.at<double>(0, 1));
  // argument shapes the interposer leaves to OpenCV
  snap();
  Mat P1f(3, 4, CV_32F), P2f(3, 4, CV_32F), X4b;
  for (int i = 0; i < 12; i++) {
    P1f.at<float>(i) = (float)P1.at<double>(i);
    P2f.at<float>(i) = (float)P2.at<double>(i);
  }
  triangulatePoints(P1f, P2f, x1, x2, X4b);
  printf("triangulatePoints f32-projections: %s X4[0]=%g\n", who(1), X4b.at<float>(0));
  snap();
  Mat r2, t2;
  solvePnPRansac(X, x1, K, dist, r2, t2, false, 100, 8.0f, 0.99, noArray(), SOLVEPNP_ITERATIVE);
  printf("solvePnPRansac iterative: %s r[0]=%g\n", who(2), r2.at<double>(0));
  snap();
  Mat R3(3, 3, CV_64F), rv3;
  for (int i = 0; i < 9; i++) R3.at<double>(i) = i % 4 == 0;
  Rodrigues(R3, rv3);  // matrix -> vector: not the node's direction
  printf("Rodrigues matrix-input: %s\n", who(0));
  return 0;
}
