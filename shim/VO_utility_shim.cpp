// shim/VO_utility_shim.cpp -- reference-side binding (NOT compiled in this repository: it needs the OpenCV C++ and
// ROS headers the reference builds against; see INTEGRATION.md.  tests/test_shim_compiles.py type-checks it against
// the reference's own VO_utility.h with the declaration-only stand-ins of tests/stubs).
//
// Drop-in replacement translation unit for uvo_libraries/src/VO_utility.cpp: it defines the same C++ free functions
// with the same signatures (uvo_libraries/include/uvo_libraries/VO_utility.h:96-117), reads the same header-defined
// parameter globals (VO_utility.h:25-89) at call time, and forwards the arithmetic to libuvo_b200.so through the C ABI
// of include/uvo_c.h.  Outputs are appended (push_back), never cleared, exactly like the reference
// (VO_utility.cpp:538, :216-217).  Link: add_library(uvo_libraries math_utility.cpp VO_utility_shim.cpp) +
// target_link_libraries(uvo_libraries uvo_b200 ${catkin_LIBRARIES}).
#include "uvo_libraries/VO_utility.h"

#include "uvo_c.h"

namespace {
uvo_ctx* ctx() {  // lazy process-wide context: the reference API has no init/teardown
  static uvo_ctx* c = nullptr;
  if (!c && uvo_ctx_create(0, nullptr, &c) != UVO_OK) {
    ROS_FATAL("uvo_b200: no CUDA device (there is no CPU fallback)");
    throw cv::Exception(cv::Error::GpuNotSupported, "uvo_b200: no CUDA device", __func__, __FILE__, __LINE__);
  }
  return c;
}
void check(int rc) {
  if (rc != UVO_OK) throw cv::Exception(cv::Error::StsError, uvo_last_error(ctx()), "uvo_b200", __FILE__, __LINE__);
}
uvo_params params_from_globals() {
  uvo_params p;
  uvo_default_params(1, &p);
  p.clahe = CLAHE_CORRECTION;
  p.clip_limit = CLIP_LIMIT;
  p.distance = DISTANCE;
  p.lowe_ratio = LOWE_RATIO_THRESHOLD;
  p.reprojection_tolerance = REPROJECTION_TOLERANCE;
  p.min_num_features = MIN_NUM_FEATURES;
  p.min_num_3dpoints = MIN_NUM_3DPOINTS;
  p.min_num_inliers = MIN_NUM_INLIERS;
  p.iterations_count = ITERATIONS_COUNT;
  p.reprojection_error = REPROJECTION_ERROR_THRESHOLD;
  p.confidence = CONFIDENCE;
  p.pnp_method_flag = PNP_METHOD_FLAG;
  p.surf_min_hessian = SURF_MIN_HESSIAN;
  p.surf_octaves = SURF_OCTAVES_NUMBER;
  p.surf_octave_layers = SURF_OCTAVES_LAYERS;
  p.surf_extended = SURF_EXTENDED;
  p.surf_upright = SURF_UPRIGHT;
  p.essential_method = ESSENTIAL_OUTLIER_METHOD;
  p.essential_max_iters = ESSENTIAL_MAX_ITERS;
  p.essential_confidence = ESSENTIAL_CONFIDENCE;
  p.essential_threshold = ESSENTIAL_THRESHOLD;
  p.homography_method = HOMOGRAPHY_OUTLIER_METHOD;
  p.homography_max_iters = HOMOGRAPHY_MAX_ITERS;
  p.homography_confidence = HOMOGRAPHY_CONFIDENCE;
  p.homography_threshold = HOMOGRAPHY_THRESHOLD;
  p.homography_distance = HOMOGRAPHY_DISTANCE;
  p.vpf_threshold = VPF_THRESHOLD;
  return p;
}
void K4(const Mat& K, double k[4]) {
  k[0] = K.at<double>(0, 0);
  k[1] = K.at<double>(1, 1);
  k[2] = K.at<double>(0, 2);
  k[3] = K.at<double>(1, 2);
}
}  // namespace

// VO_utility.cpp:337-379: both branches (native size, and INTER_AREA pre-resize to DESIRED_WIDTH) on the GPU
Mat get_image(const Mat& current_img, const Mat& cameraMatrix, const Mat& distortionCoeff, const Mat& newCamMatrix) {
  const Mat& src = current_img;
  CV_Assert(src.type() == CV_8UC3);  // cvtColor(RGB2GRAY) throws on anything else
  const double ratio = (double)src.cols / (double)DESIRED_WIDTH;
  const int desired_height = (int)(src.rows / ratio);
  uvo_camera cam;
  double k[4], nk[4];
  K4(cameraMatrix, k);
  K4(newCamMatrix, nk);
  cam = {k[0], k[1], k[2], k[3], distortionCoeff.at<double>(0), distortionCoeff.at<double>(1),
         distortionCoeff.at<double>(2), distortionCoeff.at<double>(3), nk[0], nk[1], nk[2], nk[3]};
  if (src.cols == DESIRED_WIDTH && src.rows == desired_height) {
    Mat out(src.rows, src.cols, CV_8UC1);
    check(uvo_get_image(ctx(), src.data, src.cols, src.rows, src.step, &cam, CLAHE_CORRECTION, CLIP_LIMIT, out.data,
                        out.step));
    return out;
  }
  Mat out(desired_height, DESIRED_WIDTH, CV_8UC1);
  int ow = 0, oh = 0;
  check(uvo_get_image_resized(ctx(), src.data, src.cols, src.rows, src.step, DESIRED_WIDTH, &cam, CLAHE_CORRECTION,
                              CLIP_LIMIT, out.data, out.step, &ow, &oh));
  return out;
}

// VO_utility.cpp:658-675: host-only (no OpenCV needed: uvo_optimal_new_camera_matrix restates
// getOptimalNewCameraMatrix(alpha = 0) bit-exactly)
void resize_camera_matrix(Mat original_image, Mat& cameraMatrix, Mat distortionCoeff, Mat& newCamMatrix) {
  CV_Assert(cameraMatrix.type() == CV_64F && cameraMatrix.isContinuous());
  double D[4] = {distortionCoeff.at<double>(0), distortionCoeff.at<double>(1), distortionCoeff.at<double>(2),
                 distortionCoeff.at<double>(3)};
  newCamMatrix.create(3, 3, CV_64F);
  int ow = 0, oh = 0;
  check(uvo_resize_camera_matrix(original_image.cols, original_image.rows, DESIRED_WIDTH, cameraMatrix.ptr<double>(), D,
                                 newCamMatrix.ptr<double>(), &ow, &oh));
}

// VO_utility.cpp:91-126 (SURF branch; the other detectors stay on OpenCV)
void detect_features(Mat img, vector<KeyPoint>& keypoints, Mat& descriptors) {
  CV_Assert(FEATURE_DETECTOR == "SURF" && img.type() == CV_8UC1);
  uvo_params p = params_from_globals();
  std::vector<uvo_keypoint> k(p.max_features);
  const int dd = p.surf_extended ? 128 : 64;  // SURF::descriptorSize()
  Mat d(p.max_features, dd, CV_32F);
  int n = 0;
  check(uvo_detect_features(ctx(), img.data, img.cols, img.rows, img.step, &p, k.data(), d.ptr<float>(),
                            p.max_features, &n));
  static_assert(sizeof(uvo_keypoint) == sizeof(KeyPoint), "cv::KeyPoint layout");
  keypoints.assign(reinterpret_cast<KeyPoint*>(k.data()), reinterpret_cast<KeyPoint*>(k.data()) + n);
  descriptors = d.rowRange(0, n).clone();
  ROS_INFO("FEATURES EXTRACTED - CURR. IMAGE: %lu", keypoints.size());
}

// VO_utility.cpp:515-543
void match_features(vector<KeyPoint> keypoints1, vector<KeyPoint> keypoints2, Mat descriptors1, Mat descriptors2,
                    vector<DMatch>& matches) {
  // knnMatch on an empty query set returns no rows and the node carries on into its "TOO LOW ... ASSUMING CONSTANT
  // MOTION" branch (visual_odometry.h:567, :626); an empty cv::Mat has type() == 0, so this comes before the asserts.
  // Fewer than two train descriptors: the reference reads knn[i][1] out of bounds (VO_utility.cpp:536) -- no match here.
  if (descriptors1.rows == 0 || descriptors2.rows < 2) {
    ROS_INFO("MATCHES BEFORE LOWE'S RATIO: %d", descriptors1.rows);
    ROS_INFO("MATCHES AFTER LOWE'S RATIO: %lu", matches.size());
    return;
  }
  std::vector<uvo_dmatch> m(std::max(descriptors1.rows, 1));
  int n = 0;
  CV_Assert(descriptors1.type() == CV_32F && descriptors2.type() == CV_32F && descriptors1.isContinuous() &&
            descriptors2.isContinuous() && (descriptors2.rows == 0 || descriptors1.cols == descriptors2.cols));
  check(uvo_match_features(ctx(), descriptors1.ptr<float>(), descriptors1.rows, descriptors2.ptr<float>(),
                           descriptors2.rows, descriptors1.cols, (float)LOWE_RATIO_THRESHOLD, m.data(), &n));
  ROS_INFO("MATCHES BEFORE LOWE'S RATIO: %d", descriptors1.rows);
  for (int i = 0; i < n; i++) matches.push_back(DMatch(m[i].queryIdx, m[i].trainIdx, m[i].imgIdx, m[i].distance));
  ROS_INFO("MATCHES AFTER LOWE'S RATIO: %lu", matches.size());
}

// VO_utility.cpp:551-573
void match_features(vector<KeyPoint> keypoints1, vector<KeyPoint> keypoints2, Mat descriptors1, Mat descriptors2,
                    vector<DMatch>& matches, vector<Point2f>& keypoints1_conv, vector<Point2f>& keypoints2_conv) {
  const size_t first = matches.size();
  match_features(keypoints1, keypoints2, descriptors1, descriptors2, matches);
  for (size_t i = first; i < matches.size(); i++) {
    keypoints1_conv.push_back(keypoints1[matches[i].queryIdx].pt);
    keypoints2_conv.push_back(keypoints2[matches[i].trainIdx].pt);
  }
}

// VO_utility.cpp:188-237
void extract_3Dpoints(vector<Point2f> keypoints1_conv, vector<Point2f> keypoints2_conv, Mat R1, Mat t1, Mat R2, Mat t2,
                      Mat cameraMatrix1, Mat cameraMatrix2, Mat points4D, Mat& very_good_cam1_points,
                      Mat& very_good_indexes) {
  const int n = (int)keypoints1_conv.size();
  CV_Assert(points4D.type() == CV_32F && points4D.rows == 4 && points4D.isContinuous());
  std::vector<double> pts(3 * (size_t)std::max(n, 1));
  std::vector<int32_t> idx(std::max(n, 1));
  double k1[4], k2[4];
  K4(cameraMatrix1, k1);
  K4(cameraMatrix2, k2);
  int m = 0;
  check(uvo_extract_3dpoints(ctx(), &keypoints1_conv[0].x, &keypoints2_conv[0].x, n, R1.ptr<double>(),
                             t1.ptr<double>(), R2.ptr<double>(), t2.ptr<double>(), k1, k2, points4D.ptr<float>(),
                             REPROJECTION_TOLERANCE, MIN_NUM_3DPOINTS, pts.data(), idx.data(), &m));
  for (int i = 0; i < m; i++) {
    very_good_indexes.push_back(idx[i]);
    very_good_cam1_points.push_back(Mat(1, 3, CV_64F, &pts[3 * i]).clone());
  }
}

// VO_utility.cpp:725-748
bool select_estimation_method(const vector<Point2f>& keypoints1_conv, const vector<Point2f>& keypoints2_conv) {
  int use_e = 1;
  check(uvo_select_estimation_method(ctx(), &keypoints1_conv[0].x, &keypoints2_conv[0].x, (int)keypoints1_conv.size(),
                                     DISTANCE, &use_e));
  if (!use_e) ROS_INFO("BASELINE IS TOO LOW. USING HOMOGRAPHY!");
  return use_e != 0;
}

// VO_utility.cpp:134-180.  findEssentialMat + recoverPose / findHomography + recover_pose_homography, the VPF and
// MIN_NUM_INLIERS gate and the single switch of method all run behind one C call; `use_essential` is the sticky
// global of VO_utility.h:89 and is updated in place exactly as the reference does (:175).
void estimate_relative_pose(vector<Point2f> keypoints1_conv, vector<Point2f> keypoints2_conv, Mat cameraMatrix,
                            Mat& R_currCam_prevCam, Mat& t_currCam_prevCam, vector<Point2f>& inliers1,
                            vector<Point2f>& inliers2, vector<DMatch>& inlier_matches, bool& success) {
  const int n = (int)keypoints1_conv.size();
  double k[4];
  K4(cameraMatrix, k);
  uvo_params p = params_from_globals();
  std::vector<uint8_t> mask(std::max(n, 1));
  int ue = use_essential ? 1 : 0, n_inl = 0, ok = 0;
  if (R_currCam_prevCam.empty()) R_currCam_prevCam = Mat::eye(3, 3, CV_64F);
  if (t_currCam_prevCam.empty()) t_currCam_prevCam = Mat::zeros(3, 1, CV_64F);
  check(uvo_estimate_relative_pose(ctx(), n ? &keypoints1_conv[0].x : nullptr, n ? &keypoints2_conv[0].x : nullptr, n,
                                   k, &p, &ue, R_currCam_prevCam.ptr<double>(), t_currCam_prevCam.ptr<double>(),
                                   mask.data(), &n_inl, &ok));
  if ((ue != 0) != use_essential) ROS_WARN("###### SWITCHING METHOD ######");
  use_essential = ue != 0;
  success = ok != 0;
  if (!success) ROS_WARN("###### BOTH METHODS FAILED ######");
  // extract_inliers (VO_utility.cpp:306-329)
  inliers1.clear();
  inliers2.clear();
  inlier_matches.clear();
  for (int i = 0; i < n; i++)
    if (mask[i]) {
      inliers1.push_back(keypoints1_conv[i]);
      inliers2.push_back(keypoints2_conv[i]);
      DMatch m;
      m.queryIdx = (int)inliers1.size() - 1;
      m.trainIdx = (int)inliers2.size() - 1;
      inlier_matches.push_back(m);
    }
}

// VO_utility.cpp:581-624
int recover_pose_homography(Mat H, vector<Point2f> inliers1, vector<Point2f> inliers2, Mat cameraMatrix, Mat& R, Mat& t) {
  double k[4], Rm[9] = {0}, tv[3] = {0};
  K4(cameraMatrix, k);
  int good = 0, found = 0;
  const int n = (int)inliers1.size();
  check(uvo_recover_pose_homography(ctx(), H.ptr<double>(), n ? &inliers1[0].x : nullptr, n ? &inliers2[0].x : nullptr,
                                    n, k, HOMOGRAPHY_DISTANCE, Rm, tv, &good, &found));
  if (found) {
    R = Mat(3, 3, CV_64F, Rm).clone();
    t = Mat(3, 1, CV_64F, tv).clone();
  }
  return good;
}

// The three cv:: calls the node makes directly (visual_odometry.h:355/:631, :647, :673) are not part of
// uvo_libraries.  Under a strictly unchanged node they stay on OpenCV unless shim/cv_interpose.cpp (which defines
// cv::triangulatePoints and cv::solvePnPRansac themselves) is linked in as well; the one-line node patch is to call
// these instead:
namespace uvo_shim {
void triangulatePoints(const Mat& P1, const Mat& P2, const vector<Point2f>& a, const vector<Point2f>& b, Mat& out4) {
  out4.create(4, (int)a.size(), CV_32F);
  check(uvo_triangulate_points(ctx(), P1.ptr<double>(), P2.ptr<double>(), &a[0].x, &b[0].x, (int)a.size(),
                               out4.ptr<float>()));
}
bool solvePnPRansac(const Mat& X /* Nx3 CV_64F */, const vector<Point2f>& x, const Mat& K, Mat& rvec, Mat& tvec,
                    int iters, float err, double conf, Mat& inliers) {
  double k[4];
  K4(K, k);
  std::vector<int32_t> inl(std::max(X.rows, 1));
  int n = 0, hyps = 0;
  rvec.create(3, 1, CV_64F);
  tvec.create(3, 1, CV_64F);
  check(uvo_solve_pnp_ransac(ctx(), X.ptr<double>(), &x[0].x, X.rows, k, iters, err, conf, rvec.ptr<double>(),
                             tvec.ptr<double>(), inl.data(), &n, &hyps));
  inliers = n ? Mat(n, 1, CV_32S, inl.data()).clone() : Mat();
  return n > 0;
}
}  // namespace uvo_shim

#ifdef UVO_SHIM_GPU_IMAGE_DECODE
// math_utility.cpp:154-173 -- optional (take the definition out of math_utility.cpp when enabling this one): the
// cv::imdecode that cv_bridge::toCvCopy performs on a compressed message, and the demosaic of bayer-format messages, on
// the GPU.  Baseline JPEG only; for anything else (PNG, progressive JPEG) uvo_jpeg_info fails and the caller gets the
// reference's own error behaviour: a ROS_ERROR line and an empty image.  toCvCopy's channel-order conversions for
// encodings other than bgr8 / mono8 / bayer are not reproduced.
Mat from_ros_to_cv_image(const sensor_msgs::CompressedImage::ConstPtr& image) {
  const bool bayer = image->format.find("bayer") != std::string::npos;
  uvo_jpeg_layout lay;
  if (uvo_jpeg_info(image->data.data(), image->data.size(), &lay) != UVO_OK) {
    ROS_ERROR("cv_bridge exception: %s", "uvo_b200: not a baseline JPEG stream");
    return Mat();
  }
  const bool three = lay.components == 3 || bayer;
  Mat out(lay.height, lay.width, three ? CV_8UC3 : CV_8UC1);
  int w = 0, h = 0, c = 0;
  check(uvo_jpeg_decode(ctx(), image->data.data(), image->data.size(), bayer ? 1 : 0, out.data, out.step,
                        (size_t)out.step * out.rows, &w, &h, &c));
  return out;
}
#endif

// compute_projection_matrix, convert_from_homogeneous_coords, extract_inliers, reproject_errors,
// select_desired_*, the parameter loaders and show_matches carry no hot arithmetic and are
// compiled unchanged from the reference's VO_utility.cpp.
