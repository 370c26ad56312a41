/*
 * uvo_oracle.h -- CPU ORACLE for the UVO per-frame hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This library restates, in plain scalar C++ with no dependencies, the arithmetic that the reference
 * (team-ergo-unipi/ergo_uvo) delegates to OpenCV 4.5 + contrib on the path named by BASELINE.json
 * (kernels K1..K12).  The reference itself holds no arithmetic and no golden vectors
 * (SURVEY.md 8c); the algorithm lives in un-vendored third-party OpenCV (find_package(OpenCV 4 REQUIRED),
 * uvo_libraries/CMakeLists.txt:15; README.md:60 pins "OpenCV 4.5" + contrib, no patch version).
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - K1 gray+undistort, K2 CLAHE, K3 integral, INTER_AREA patch resize, BF kNN matcher, cv::RNG subset stream,
 *     triangulatePoints, solvePnPRansac (inlier sets), projectPoints: pinned against the importable
 *     opencv-python-headless 4.13.0 wheel (tests/test_oracle_*.py + fixtures in tests/golden/ made by
 *     tools/make_golden.py).
 *   - K4..K7 SURF (opencv_contrib xfeatures2d/src/surf.cpp): contrib is absent from this image, no network.
 *     PARITY UNPINNED for SURF: the restatement follows SURVEY.md Appendix A (from memory of the 4.5 source);
 *     only its sub-steps (integral, Gaussian tables, INTER_AREA resize) are pinned against cv2.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product (ergo_uvo_b200/) never links, imports or calls it.
 */
#ifndef UVO_ORACLE_H
#define UVO_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- K1: cvtColor(COLOR_RGB2GRAY) + cv::undistort   (VO_utility.cpp:346-350) ---- */
/* 3-channel interleaved u8 -> gray u8, weights applied to channels as stored. */
void orc_gray(const uint8_t* src3, int w, int h, uint8_t* dst);
/* K = (fx,fy,cx,cy), D = (k1,k2,p1,p2), newK = (fx,fy,cx,cy).  Fixed-point maps as cv::initUndistortRectifyMap
 * CV_16SC2: map_xy int16 [h][w][2], map_frac uint16 [h][w]. */
void orc_undistort_map(const double K[4], const double D[4], const double newK[4], int w, int h,
                       int16_t* map_xy, uint16_t* map_frac);
/* remap INTER_LINEAR, BORDER_CONSTANT(0) with the fixed-point map. */
void orc_remap_bilinear(const uint8_t* src, int w, int h, const int16_t* map_xy, const uint16_t* map_frac,
                        uint8_t* dst);
void orc_undistort(const uint8_t* gray, int w, int h, const double K[4], const double D[4], const double newK[4],
                   uint8_t* dst);
/* ---- K2: CLAHE 8x8 tiles (VO_utility.cpp:352-357), in/out may alias ---- */
void orc_clahe(const uint8_t* src, int w, int h, double clip_limit, int tiles_x, int tiles_y, uint8_t* dst);
/* whole get_image (native-size branch, VO_utility.cpp:344-360) */
void orc_get_image(const uint8_t* src3, int w, int h, const double K[4], const double D[4], const double newK[4],
                   int clahe, double clip_limit, uint8_t* dst);
/* ---- K0: resize INTER_AREA (VO_utility.cpp:362-363), u8, cn channels ---- */
void orc_resize_area(const uint8_t* src, int sw, int sh, int cn, uint8_t* dst, int dw, int dh);
/* ---- before K0: cvtColor(COLOR_BayerBGGR2BGR) of from_ros_to_cv_image (math_utility.cpp:161-164); w, h >= 3 ---- */
void orc_bayer_bggr2bgr(const uint8_t* src, int w, int h, uint8_t* dst3);
/* ---- K3: integral(CV_32S): (h+1)x(w+1) ---- */
void orc_integral(const uint8_t* src, int w, int h, int32_t* sum);

/* ---- K4..K7: SURF (VO_utility.cpp:117-118) ---- */
typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} orc_keypoint; /* cv::KeyPoint, 28 bytes */
/* returns number of keypoints written (<= capacity), or -needed if capacity too small. desc: n x 64 (or 128). */
int orc_surf_detect_and_compute(const uint8_t* img, int w, int h, double hessian_threshold, int n_octaves,
                                int n_octave_layers, int extended, int upright, orc_keypoint* kps, float* desc,
                                int capacity);
/* sub-steps exposed for tests */
void orc_surf_det_trace_layer(const int32_t* sum, int w, int h, int size, int step, float* det, float* trace);
void orc_surf_patch(const uint8_t* img, int w, int h, float cx, float cy, float size, float angle_deg, int upright,
                    uint8_t* patch21x21, int* win_size_out);
void orc_gaussian_kernel_f32(int n, double sigma, float* out);
float orc_fast_atan2(float y, float x);

/* ---- K8: BFMatcher(NORM_L2).knnMatch(k=2) + Lowe ratio (VO_utility.cpp:515-573) ---- */
typedef struct {
  int32_t queryIdx, trainIdx, imgIdx;
  float distance;
} orc_dmatch; /* cv::DMatch, 16 bytes */
/* knn: writes nq*2 entries (trainIdx=-1 when nt<k).  */
void orc_knn2(const float* q, int nq, const float* t, int nt, int dim, orc_dmatch* out2);
/* match_features: returns number of matches that pass d0 < ratio*d1 */
int orc_match_features(const float* q, int nq, const float* t, int nt, int dim, float ratio, orc_dmatch* out);

/* ---- cv::RNG + RANSAC subset stream ---- */
void orc_rng_subsets(int count, int model_points, int n_subsets, int32_t* out /* n_subsets*model_points */);
int orc_ransac_update_num_iters(double p, double ep, int model_points, int max_iters);

/* ---- K11: cv::triangulatePoints (visual_odometry.h:355,:631) ---- */
void orc_triangulate_points(const double P1[12], const double P2[12], const float* pts1, const float* pts2, int n,
                            float* out4xn);
/* ---- K12: extract_3Dpoints & friends (VO_utility.cpp:188-237) ---- */
/* points4d: 4 x n f32 (row-major, as triangulatePoints returns). R*,t*: row-major 3x3 / 3.  K*: fx,fy,cx,cy.
 * returns count M'; out_points M' x 3 f64, out_idx M' i32. */
int orc_extract_3dpoints(const float* kp1, const float* kp2, int n, const double R1[9], const double t1[3],
                         const double R2[9], const double t2[3], const double K1[4], const double K2[4],
                         const float* points4d, double reproj_tol, int min_num_3dpoints, double* out_points,
                         int32_t* out_idx);
void orc_project_points(const double* X, int n, const double R[9], const double t[3], const double K[4],
                        double* out2);
void orc_rodrigues_vec2mat(const double r[3], double R[9]);
void orc_rodrigues_mat2vec(const double R[9], double r[3]);
double orc_compute_median(const double* v, int n);
/* convert_3Dpoints_camera + compute_scale_factor (VO_utility.cpp:23-63): returns SF, 0.0 on failure */
double orc_scale_factor(const double* pts_nx3, int n, const double R[9], const double t[3], float range);
int orc_select_estimation_method(const float* p1, const float* p2, int n, int distance);

/* ---- K10c: cv::solvePnPRansac(EPNP) (visual_odometry.h:647-648) ---- */
/* X: n x 3 f64, x: n x 2 f32, K: fx,fy,cx,cy (distortion zero).  Returns number of inliers (0 => failure),
 * inliers ascending.  hyps_evaluated (optional) = iterations the adaptive loop ran. */
int orc_solve_pnp_ransac_epnp(const double* X, const float* x, int n, const double K[4], int iterations,
                              float reproj_err, double confidence, double rvec[3], double tvec[3],
                              int32_t* inliers, int* hyps_evaluated);
/* EPnP on a given set (cv::solvePnP(..., SOLVEPNP_EPNP)); X n x3 f64, x n x2 f64 */
void orc_epnp(const double* X, const double* x, int n, const double K[4], double R[9], double t[3]);

/* ---- ingest: JPEG decode of from_ros_to_cv_image (math_utility.cpp:154-173: cv_bridge::toCvCopy -> cv::imdecode) ---- */
/* baseline / extended-sequential Huffman, 8-bit, 1 or 3 components.  0 on success, -1 corrupt, -2 unsupported. */
int orc_jpeg_info(const uint8_t* data, size_t len, int* w, int* h, int* channels);
/* parity tap: quantised coefficients after entropy decoding (component-major, blocks padded to whole MCUs, natural
 * order); returns their count, negative on error */
long orc_jpeg_coefficients(const uint8_t* data, size_t len, int16_t* out, size_t capacity);
/* out: h x w (1 component) or h x w x 3 BGR, as cv::imdecode(IMREAD_UNCHANGED) lays it out */
int orc_jpeg_decode(const uint8_t* data, size_t len, uint8_t* out);

/* K10a / K10b (findEssentialMat, findHomography, recoverPose, decomposeHomographyMat) are restated in numpy:
 * oracle/twoview.py, pinned to cv2 4.13 (tests/test_oracle_twoview.py, tests/golden/twoview.npz). */

#ifdef __cplusplus
}
#endif
#endif
