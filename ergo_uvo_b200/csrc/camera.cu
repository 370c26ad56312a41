// camera.cu -- one-time camera set-up on the host (no GPU work): resize_camera_matrix (reference
// VO_utility.cpp:658-675) and the cv::getOptimalNewCameraMatrix(K, D, size, alpha = 0, size, false) it calls, so
// that a shim needs no OpenCV for the camera model either.
//
// getOptimalNewCameraMatrix with alpha = 0 [cv-mem calib3d/src/calibration.cpp icvGetRectangles, pinned against the
// cv2 4.13 wheel in tests/test_cabi.py -- bit-exact, all fp64]: a 9 x 9 grid of pixel positions spanning the image is
// undistorted (cv::undistortPoints, 5 fixed-point iterations), the largest axis-aligned rectangle
// inscribed in the undistorted grid ("inner") is taken from the border points, and the new intrinsics map that
// rectangle onto the full image.
#include <cfloat>
#include <cmath>

#include "uvo_c.h"

namespace {

// cvUndistortPointsInternal for the 4-coefficient model (k1, k2, p1, p2), R = P = identity, criteria = (COUNT, 5)
void undistort_point(double u, double v, double fx, double fy, double cx, double cy, const double D[4], double& ox,
                     double& oy) {
  const double ifx = 1. / fx, ify = 1. / fy;
  double x = (u - cx) * ifx, y = (v - cy) * ify;
  const double x0 = x, y0 = y;
  const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3];
  if (k1 != 0 || k2 != 0 || p1 != 0 || p2 != 0) {
    for (int j = 0; j < 5; j++) {
      const double r2 = x * x + y * y;
      const double icdist = 1. / (1 + ((0 * r2 + k2) * r2 + k1) * r2);
      if (icdist < 0) {
        x = x0;
        y = y0;
        break;
      }
      const double dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
      const double dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
  }
  ox = x;
  oy = y;
}

}  // namespace

extern "C" {

int uvo_optimal_new_camera_matrix(const double K[9], const double D[4], int width, int height, double newK[9]) {
  if (!K || !D || !newK || width < 2 || height < 2 || K[0] == 0 || K[4] == 0) return UVO_ERR_INVALID;
  const int N = 9;
  double iX0 = -DBL_MAX, iX1 = DBL_MAX, iY0 = -DBL_MAX, iY1 = DBL_MAX;
  for (int y = 0; y < N; y++)
    for (int x = 0; x < N; x++) {
      const double px = (double)x * (width - 1) / (N - 1), py = (double)y * (height - 1) / (N - 1);
      double ux, uy;
      undistort_point(px, py, K[0], K[4], K[2], K[5], D, ux, uy);
      if (x == 0) iX0 = fmax(iX0, ux);
      if (x == N - 1) iX1 = fmin(iX1, ux);
      if (y == 0) iY0 = fmax(iY0, uy);
      if (y == N - 1) iY1 = fmin(iY1, uy);
    }
  const double in_x = iX0, in_y = iY0, in_w = iX1 - iX0, in_h = iY1 - iY0;
  if (!(in_w > 0) || !(in_h > 0)) return UVO_ERR_INVALID;
  const double fx0 = (width - 1) / in_w, fy0 = (height - 1) / in_h;
  const double cx0 = -fx0 * in_x, cy0 = -fy0 * in_y;
  // OpenCV writes the four intrinsics into a copy of the input matrix: skew and the last row are carried over
  double M[9];
  for (int i = 0; i < 9; i++) M[i] = K[i];
  M[0] = fx0;
  M[2] = cx0;
  M[4] = fy0;
  M[5] = cy0;
  for (int i = 0; i < 9; i++) newK[i] = M[i];
  return UVO_OK;
}

int uvo_resize_camera_matrix(int original_width, int original_height, int desired_width, double K_inout[9],
                             const double D[4], double newK[9], int* out_width, int* out_height) {
  if (!K_inout || !D || !newK || original_width <= 0 || original_height <= 0 || desired_width <= 0)
    return UVO_ERR_INVALID;
  const double ratio = (double)original_width / (double)desired_width;
  const int desired_height = (int)(original_height / ratio);
  const double skew = K_inout[1];
  const double inv = 1. / ratio;  // cv::Mat / double scales by the reciprocal [cv-mem MatExpr operator/]
  for (int i = 0; i < 9; i++) K_inout[i] *= inv;
  K_inout[1] = skew;
  K_inout[8] = 1;
  if (out_width) *out_width = desired_width;
  if (out_height) *out_height = desired_height;
  return uvo_optimal_new_camera_matrix(K_inout, D, desired_width, desired_height, newK);
}

// cv::Rodrigues(rvec, R) as the node calls it on solvePnPRansac's output (visual_odometry.h:673): rotation vector ->
// 3 x 3 matrix, fp64, the arithmetic of OpenCV's cvRodrigues2 (theta = |r|; R = cos I + (1 - cos) r r^T + sin [r]x).
// Host-only (a 3-vector), here so that the cv:: interposer needs no OpenCV for it; pinned against cv2 in
// tests/test_cabi.py.
int uvo_rodrigues(const double rvec[3], double R[9]) {
  if (!rvec || !R) return UVO_ERR_INVALID;
  const double theta = std::sqrt(rvec[0] * rvec[0] + rvec[1] * rvec[1] + rvec[2] * rvec[2]);
  if (theta < DBL_EPSILON) {
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return UVO_OK;
  }
  const double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c, it = 1. / theta;
  const double x = rvec[0] * it, y = rvec[1] * it, z = rvec[2] * it;
  const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
  const double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  for (int i = 0; i < 9; i++) R[i] = c * (i % 4 == 0 ? 1. : 0.) + c1 * rrt[i] + s * rx[i];
  return UVO_OK;
}

}  // extern "C"
