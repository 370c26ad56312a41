/*
 * uvo_c.h -- C ABI of the B200-native UVO hot path (libuvo_b200.so).
 *
 * Drop-in boundary for the per-frame feature front-end and pose inner loop of team-ergo-unipi/ergo_uvo.
 * The reference exposes this path as C++ free functions of the catkin library `uvo_libraries`
 * (uvo_libraries/include/uvo_libraries/VO_utility.h:96-117) operating on cv::Mat / std::vector by value, plus three
 * cv:: calls made by the node itself (uvo/include/visual_odometry.h:355,:631,:647,:673).  Each entry point below
 * names the reference interface it replaces.  No OpenCV, ROS or torch types cross this boundary: plain pointers,
 * sizes and POD structs only.  shim/VO_utility_shim.cpp shows the reference-side binding.
 *
 * Conventions
 *   - every function returns UVO_OK (0) or a negative uvo_status; uvo_last_error(ctx) gives the message;
 *   - there is NO CPU fallback: without a CUDA device uvo_ctx_create fails with UVO_ERR_NO_DEVICE;
 *   - "_host" pointers are host memory (pinned or pageable); the call copies H2D/D2H on the ctx stream and returns
 *     after the stream is synchronised; handles (uvo_stereo) keep all per-frame state device-resident;
 *   - images are row-major u8 with an explicit pitch in bytes; matrices are row-major double.
 */
#ifndef UVO_C_H
#define UVO_C_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVO_API __attribute__((visibility("default")))

typedef enum {
  UVO_OK = 0,
  UVO_ERR_NO_DEVICE = -1,
  UVO_ERR_CUDA = -2,
  UVO_ERR_INVALID = -3,
  UVO_ERR_CAPACITY = -4,
  UVO_ERR_UNSUPPORTED = -5
} uvo_status;

typedef struct uvo_ctx uvo_ctx;

/* cv::KeyPoint layout (28 B) -- VO_utility.h:100 `vector<KeyPoint>&` */
typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} uvo_keypoint;

/* cv::DMatch layout (16 B) -- VO_utility.h:109 `vector<DMatch>&` */
typedef struct {
  int32_t queryIdx, trainIdx, imgIdx;
  float distance;
} uvo_dmatch;

/* Camera model: the globals fx,fy,ccx,ccy,k1,k2,p1,p2 (VO_utility.h:31-36) after resize_camera_matrix
 * (VO_utility.cpp:658-675) plus the new camera matrix it returns. */
typedef struct {
  double fx, fy, cx, cy;     /* cameraMatrix */
  double k1, k2, p1, p2;     /* distortionCoeff */
  double nfx, nfy, ncx, ncy; /* newCamMatrix (getOptimalNewCameraMatrix, alpha = 0) */
} uvo_camera;

/* The mutable parameter globals of VO_utility.h:38-89, read by the shim at call time. */
typedef struct {
  int32_t clahe;                  /* CLAHE_CORRECTION */
  int32_t clip_limit;             /* CLIP_LIMIT */
  int32_t distance;               /* DISTANCE */
  double lowe_ratio;              /* LOWE_RATIO_THRESHOLD */
  int32_t essential_method;       /* ESSENTIAL_OUTLIER_METHOD (4 = LMEDS, 8 = RANSAC) */
  double essential_max_iters;     /* ESSENTIAL_MAX_ITERS */
  double essential_confidence;    /* ESSENTIAL_CONFIDENCE */
  double essential_threshold;     /* ESSENTIAL_THRESHOLD */
  int32_t homography_method;      /* HOMOGRAPHY_OUTLIER_METHOD */
  double homography_max_iters;    /* HOMOGRAPHY_MAX_ITERS */
  double homography_confidence;   /* HOMOGRAPHY_CONFIDENCE */
  double homography_threshold;    /* HOMOGRAPHY_THRESHOLD */
  double homography_distance;     /* HOMOGRAPHY_DISTANCE */
  double vpf_threshold;           /* VPF_THRESHOLD */
  double reprojection_tolerance;  /* REPROJECTION_TOLERANCE */
  int32_t min_num_features;       /* MIN_NUM_FEATURES */
  int32_t min_num_3dpoints;       /* MIN_NUM_3DPOINTS */
  int32_t min_num_inliers;        /* MIN_NUM_INLIERS */
  int32_t iterations_count;       /* ITERATIONS_COUNT */
  double reprojection_error;      /* REPROJECTION_ERROR_THRESHOLD */
  double confidence;              /* CONFIDENCE */
  int32_t pnp_method_flag;        /* PNP_METHOD_FLAG (1 = SOLVEPNP_EPNP, the only one implemented) */
  int32_t surf_min_hessian;       /* SURF_MIN_HESSIAN */
  int32_t surf_octaves;           /* SURF_OCTAVES_NUMBER (4) */
  int32_t surf_octave_layers;     /* SURF_OCTAVES_LAYERS (3) */
  int32_t surf_extended;          /* SURF_EXTENDED (0) */
  int32_t surf_upright;           /* SURF_UPRIGHT */
  int32_t max_features;           /* capacity of device keypoint buffers (not in the reference: it has no cap;
                                     exceeding it returns UVO_ERR_CAPACITY instead of truncating) */
  /* Stereo epipolar / disparity gate on the left-right match (visual_odometry.h:558).  NOT in the reference, whose
   * match_features applies the ratio test only (VO_utility.cpp:515-573): default 0 = OFF, which is what parity with
   * the reference requires.  When on, a ratio-test survivor (left keypoint q, right keypoint t) is kept iff
   * |y_q - y_t| <= stereo_max_epipolar_dy and stereo_min_disparity <= x_q - x_t <= stereo_max_disparity
   * (rectified pair, f32 compares). */
  int32_t stereo_gate;
  double stereo_max_epipolar_dy;
  double stereo_min_disparity;
  double stereo_max_disparity;
} uvo_params;

/* ---------------------------------------------------------------------------------------------------- context */
/* One context per (GPU, stream).  `cuda_stream` may be NULL (the context creates its own non-blocking stream) or a
 * cudaStream_t owned by the caller (e.g. torch.cuda.current_stream().cuda_stream). */
UVO_API int uvo_ctx_create(int device, void* cuda_stream, uvo_ctx** out);
UVO_API void uvo_ctx_destroy(uvo_ctx* ctx);
UVO_API const char* uvo_last_error(const uvo_ctx* ctx);
UVO_API const char* uvo_version(void);
UVO_API void* uvo_ctx_stream(uvo_ctx* ctx);
UVO_API int uvo_ctx_synchronize(uvo_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
UVO_API int64_t uvo_ctx_launch_count(const uvo_ctx* ctx);
/* Diagnostics for bench.py's roofline leg: when enabled every kernel launch is bracketed by CUDA events on the
 * context stream; uvo_ctx_kernel_report writes one line per kernel name, "<name> <launches> <total_ms>", and resets. */
UVO_API int uvo_ctx_kernel_timing(uvo_ctx* ctx, int enable);
UVO_API int uvo_ctx_kernel_report(uvo_ctx* ctx, char* buf, size_t buflen);
/* fills every field with the shipped stereo (stereo!=0) or mono YAML values
 * (uvo/config/stereo_VO_parameters.yaml:8-47, mono_VO_parameters.yaml:2-49) */
UVO_API void uvo_default_params(int stereo, uvo_params* out);
/* pinned host memory helpers for callers without a CUDA runtime of their own */
UVO_API void* uvo_host_alloc(size_t bytes);
UVO_API void uvo_host_free(void* p);

/* ---------------------------------------------------------------------------------------------------- camera set-up */
/* cv::getOptimalNewCameraMatrix(K, D, Size(w, h), 0, Size(w, h), 0) as resize_camera_matrix calls it
 * (VO_utility.cpp:674): host-only, no GPU.  K, newK: 3x3 row-major; D = k1, k2, p1, p2. */
UVO_API int uvo_optimal_new_camera_matrix(const double K[9], const double D[4], int width, int height,
                                          double newK[9]);
/* void resize_camera_matrix(Mat original_image, Mat& cameraMatrix, Mat distortionCoeff, Mat& newCamMatrix)
 * -- VO_utility.h:112, VO_utility.cpp:658-675: scales K by DESIRED_WIDTH / original width (skew kept, K[2][2] = 1)
 * in place and returns the new camera matrix for the DESIRED_WIDTH x int(h / ratio) image.  Host-only. */
UVO_API int uvo_resize_camera_matrix(int original_width, int original_height, int desired_width, double K_inout[9],
                                     const double D[4], double newK[9], int* out_width, int* out_height);

/* cv::Rodrigues(rvec, R), rotation vector -> 3x3 row-major matrix, as the node calls it on solvePnPRansac's output
 * (visual_odometry.h:673).  Host-only, no GPU. */
UVO_API int uvo_rodrigues(const double rvec[3], double R[9]);

/* ---------------------------------------------------------------------------------------------------- ingest */
/* cvtColor(image, image, COLOR_BayerBGGR2BGR) -- the demosaic from_ros_to_cv_image applies to bayer-format camera
 * messages (math_utility.h:27, math_utility.cpp:161-164), bit-exact with OpenCV's bilinear demosaic.  bayer: 1-channel
 * u8 (w, h >= 3); bgr: 3-channel interleaved u8.  The stereo handle can take bayer images directly
 * (uvo_stereo_enqueue_host_bayer), which cuts the host-to-device traffic of a frame to a third. */
UVO_API int uvo_demosaic_bggr2bgr(uvo_ctx* ctx, const uint8_t* bayer_host, int width, int height, size_t src_pitch,
                                  uint8_t* bgr_host, size_t dst_pitch);

/* ---------------------------------------------------------------------------------------------------- K1-K3 */
/* Mat get_image(const Mat&, const Mat&, const Mat&, const Mat&)  -- VO_utility.h:105, VO_utility.cpp:337-379.
 * Native-size branch: cvtColor(RGB2GRAY) + undistort + optional CLAHE(8x8).  src is 3-channel interleaved u8. */
UVO_API int uvo_get_image(uvo_ctx* ctx, const uint8_t* src3_host, int width, int height, size_t src_pitch,
                          const uvo_camera* cam, int clahe, int clip_limit, uint8_t* dst_host, size_t dst_pitch);
/* K0: the pre-scaling branch of get_image (VO_utility.cpp:339-342, :362-376): cv::resize(INTER_AREA) to
 * DESIRED_WIDTH x int(h / (w / DESIRED_WIDTH)), then the same gray / undistort / CLAHE chain.  `cam` is the camera
 * AFTER resize_camera_matrix (VO_utility.cpp:658-675).  dst must hold *out_h rows of dpitch bytes. */
UVO_API int uvo_get_image_resized(uvo_ctx* ctx, const uint8_t* src3_host, int width, int height, size_t src_pitch,
                                  int desired_width, const uvo_camera* cam, int clahe, int clip_limit,
                                  uint8_t* dst_host, size_t dst_pitch, int* out_width, int* out_height);
/* cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_AREA) for u8 images with 1 or 3 interleaved channels (shrinking) */
UVO_API int uvo_resize_area(uvo_ctx* ctx, const uint8_t* src_host, int src_width, int src_height, size_t src_pitch,
                            int channels, uint8_t* dst_host, int dst_width, int dst_height, size_t dst_pitch);
/* integral(img, sum, CV_32S) as built inside SURF::detectAndCompute (VO_utility.cpp:118): (h+1) x (w+1) int32 */
UVO_API int uvo_integral(uvo_ctx* ctx, const uint8_t* gray_host, int width, int height, size_t pitch,
                         int32_t* sum_host);

/* ---------------------------------------------------------------------------------------------------- ingest: JPEG */
/* The decode inside Mat from_ros_to_cv_image(const sensor_msgs::CompressedImage::ConstPtr&) -- math_utility.h:27,
 * math_utility.cpp:154-173: cv_bridge::toCvCopy -> cv::imdecode(IMREAD_UNCHANGED) -> libjpeg-turbo (JDCT_ISLOW, fancy
 * upsampling, YCbCr -> BGR).  Baseline / extended-sequential Huffman streams, 8-bit, 1 or 3 components; anything else
 * returns UVO_ERR_UNSUPPORTED.  Entropy decoding runs on the host, the transform on the GPU (csrc/jpeg.cu). */
typedef struct {
  int32_t width, height, components;   /* components: 1 (e.g. a bayer mosaic) or 3 (YCbCr) */
  int32_t h_samp[3], v_samp[3];        /* sampling factors of the frame header */
  int32_t blocks_x[3], blocks_y[3];    /* 8x8 blocks per component row / column, padded to whole MCUs */
  int32_t samples_x[3], samples_y[3];  /* real samples per component row / column (libjpeg's downsampled_width/height) */
  int64_t coeff_offset[3];             /* index of the component's first coefficient in the coefficient buffer */
  int64_t coeff_total;                 /* coefficients in all: sum of blocks_x * blocks_y * 64 */
  uint16_t quant[3][64];               /* quantisation table per component, natural (row-major) order */
} uvo_jpeg_layout;
/* header only (host, no GPU): sizes for the caller's buffers */
UVO_API int uvo_jpeg_info(const uint8_t* jpeg, size_t len, uvo_jpeg_layout* layout);
/* host half (no GPU): Huffman decoding into quantised coefficients, int16, per component blocks_y x blocks_x x 64 in
 * natural order at coeff_offset[c]; `capacity` in coefficients (>= coeff_total, else UVO_ERR_CAPACITY) */
UVO_API int uvo_jpeg_entropy_decode(const uint8_t* jpeg, size_t len, int16_t* coeffs_host, size_t capacity,
                                    uvo_jpeg_layout* layout);
/* the same, in the form uvo_jpeg_decode ships to the GPU: one 32-bit entry per NON-ZERO coefficient,
 * (natural index << 16) | (value & 0xffff), in scan order (capacity in entries; coeff_total always suffices), and per
 * block -- numbered component-major, row-major, coeff_offset[c] / 64 first -- the index of its first entry and its
 * entry count (coeff_total / 64 elements each) */
UVO_API int uvo_jpeg_entropy_decode_sparse(const uint8_t* jpeg, size_t len, uint32_t* entries_host, size_t capacity,
                                           uint32_t* block_first_host, uint8_t* block_count_host, size_t* n_entries,
                                           uvo_jpeg_layout* layout);
/* one compressed camera image in that form: what the GPU half of the decode takes.  Produced by
 * uvo_jpeg_entropy_decode_sparse on any host thread (it touches no GPU state), consumed by
 * uvo_stereo_enqueue_host_sparse; buffers from uvo_host_alloc (pinned) make the upload asynchronous. */
typedef struct {
  const uint32_t* entries;       /* n_entries words */
  size_t n_entries;
  const uint32_t* block_first;   /* layout.coeff_total / 64 words */
  const uint8_t* block_count;    /* layout.coeff_total / 64 bytes */
  uvo_jpeg_layout layout;
} uvo_jpeg_sparse;
/* the whole decode: out_host is height x width (1 component) or height x width x 3 BGR, rows out_pitch bytes apart,
 * exactly what cv::imdecode(IMREAD_UNCHANGED) returns; out_capacity in bytes.  bayer_bggr != 0 mirrors
 * `image->format.find("bayer") != npos` (math_utility.cpp:161-164): a 1-component stream is taken as the BGGR mosaic
 * and cvtColor(COLOR_BayerBGGR2BGR) is applied on the GPU, the output is height x width x 3 (ignored for 3 components) */
UVO_API int uvo_jpeg_decode(uvo_ctx* ctx, const uint8_t* jpeg, size_t len, int bayer_bggr, uint8_t* out_host,
                            size_t out_pitch, size_t out_capacity, int* width, int* height, int* channels);
/* Huffman decoding ON THE GPU (default on): streams made of one interleaved scan without restart intervals -- what
 * cv::imencode and camera encoders write -- are entropy-decoded by a self-synchronising parallel decoder
 * (csrc/jpeg_huff.cuh), so only the scan bytes cross PCIe and the host does no per-coefficient work; other streams take
 * the host decoder.  Results are identical.  enable: 1 / 0, or -1 to leave the setting alone; last_route (nullable):
 * 1 when the previous uvo_jpeg_decode on this context took the GPU decoder; last_rounds: its synchronisation rounds. */
UVO_API int uvo_jpeg_gpu_entropy(uvo_ctx* ctx, int enable, int* last_route, int* last_rounds);
/* diagnostics: nanosecond time stamps (%globaltimer) of the phases of the previous GPU entropy decode -- start, after
 * rounds 1 and 2, after the last round, count, scan, write, DC partial sums ..., end (tools/jh_time.py) */
UVO_API int uvo_jpeg_gpu_entropy_stamps(uvo_ctx* ctx, int64_t stamps[24], int* count);
/* host-only (no GPU needed; tests and the fuzzer): the HOST half of that route -- marker walk, table plan, copy of the
 * scan with the stuffed zeros removed -- run into caller memory.  staging holds uvo_jpeg_gpu_staging_bytes(len) bytes.
 * *qualifies = 1: the stream takes the GPU decoder, *upload_bytes of staging would go to the device and the scan is
 * *scan_bits long; 0: it takes the host decoder.  Malformed streams return an error like uvo_jpeg_info. */
UVO_API size_t uvo_jpeg_gpu_staging_bytes(size_t len);
UVO_API int uvo_jpeg_gpu_plan(const uint8_t* jpeg, size_t len, uint8_t* staging, size_t staging_bytes, int* qualifies,
                              size_t* upload_bytes, uint32_t* scan_bits, uvo_jpeg_layout* layout);
/* the same with the image left in DEVICE memory (no copy back; the work is ordered on the context stream), ready for
 * uvo_stereo_enqueue_device / uvo_mono_frame_device on the same context */
UVO_API int uvo_jpeg_decode_device(uvo_ctx* ctx, const uint8_t* jpeg, size_t len, int bayer_bggr, uint8_t* out_dev,
                                   size_t out_pitch, size_t out_capacity, int* width, int* height, int* channels);

/* ---------------------------------------------------------------------------------------------------- K4-K7 */
/* void detect_features(Mat img, vector<KeyPoint>&, Mat& descriptors) -- VO_utility.h:100, VO_utility.cpp:114-119:
 * SURF::create(hess, octaves, layers, extended, upright)->detectAndCompute.  Keypoints come back in OpenCV's order
 * (response desc, size desc, octave desc, y desc, x asc); descriptors row-major n x 64 f32, or n x 128 when
 * prm->surf_extended is set (SURF_EXTENDED, VO_utility.h:86) -- `desc_host` must hold capacity x 128 floats then. */
UVO_API int uvo_detect_features(uvo_ctx* ctx, const uint8_t* gray_host, int width, int height, size_t pitch,
                                const uvo_params* prm, uvo_keypoint* kps_host, float* desc_host, int capacity,
                                int* count);

/* Parity tap for K6: the total order SURF::detectAndCompute leaves its keypoints in (KeypointGreater: response desc,
 * size desc, octave desc, y desc, x asc; identical keys keep their input order).  `capacity` >= n selects the device
 * buffers' size and with it the sort variant (single-block bitonic sort up to 16 384, rank sort above). */
UVO_API int uvo_sort_keypoints(uvo_ctx* ctx, const uvo_keypoint* kps_in_host, int n, int capacity,
                               uvo_keypoint* kps_out_host);

/* ---------------------------------------------------------------------------------------------------- K8 */
/* void match_features(vector<KeyPoint>, vector<KeyPoint>, Mat d1, Mat d2, vector<DMatch>&) -- VO_utility.h:109-110,
 * VO_utility.cpp:515-573: BFMatcher(NORM_L2).knnMatch(k=2) + Lowe ratio.  Matches in query order. `dim` = 64, or 128
 * for extended SURF descriptors (SURF_EXTENDED, VO_utility.h:86; anything else: UVO_ERR_UNSUPPORTED). */
UVO_API int uvo_match_features(uvo_ctx* ctx, const float* desc1_host, int n1, const float* desc2_host, int n2,
                               int dim, float ratio, uvo_dmatch* matches_host, int* count);
/* match_features followed by the optional stereo epipolar / disparity gate described at uvo_params.stereo_gate
 * (north_star; the reference has no such gate, so the node's drop-in path calls uvo_match_features).  kps1 / kps2
 * are the keypoints of the query (left) / train (right) descriptors. */
UVO_API int uvo_match_features_gated(uvo_ctx* ctx, const uvo_keypoint* kps1_host, const float* desc1_host, int n1,
                                     const uvo_keypoint* kps2_host, const float* desc2_host, int n2, int dim,
                                     float ratio, float max_epipolar_dy, float min_disparity, float max_disparity,
                                     uvo_dmatch* matches_host, int* count);
/* raw knnMatch(k=2) rows (2 per query; trainIdx = -1 where n2 < 2) */
UVO_API int uvo_knn_match2(uvo_ctx* ctx, const float* desc1_host, int n1, const float* desc2_host, int n2, int dim,
                           uvo_dmatch* knn_host);
/* diagnostics: how many queries of the last uvo_match_features / uvo_knn_match2 call on this context were resolved
 * by the exact full scan because the tensor-core candidate set could not be proven complete (results are exact
 * either way; this is a performance counter) */
UVO_API int uvo_match_last_fallbacks(uvo_ctx* ctx, int* count);
/* diagnostics: when enabled, the stage-level matcher calls of this context skip the tensor-core candidate pass and
 * resolve every query by the exact full scan (same results, much slower); tests use it to cross-check the two routes */
UVO_API int uvo_match_exact_only(uvo_ctx* ctx, int enable);

/* ---------------------------------------------------------------------------------------------------- K9, K11, K12 */
/* bool select_estimation_method(const vector<Point2f>&, const vector<Point2f>&) -- VO_utility.cpp:725-748 */
UVO_API int uvo_select_estimation_method(uvo_ctx* ctx, const float* pts1_host, const float* pts2_host, int n,
                                         int distance, int* use_essential);
/* cv::triangulatePoints(P1, P2, pts1, pts2, points4D) -- visual_odometry.h:355, :631; out is 4 x n f32 */
UVO_API int uvo_triangulate_points(uvo_ctx* ctx, const double P1[12], const double P2[12], const float* pts1_host,
                                   const float* pts2_host, int n, float* points4d_host);
/* void extract_3Dpoints(...) -- VO_utility.h:102, VO_utility.cpp:188-237. K1,K2 = fx,fy,cx,cy.
 * out_points: M' x 3 f64, out_idx: M' i32 (capacity n each). */
UVO_API int uvo_extract_3dpoints(uvo_ctx* ctx, const float* kp1_host, const float* kp2_host, int n,
                                 const double R1[9], const double t1[3], const double R2[9], const double t2[3],
                                 const double K1[4], const double K2[4], const float* points4d_host,
                                 double reprojection_tolerance, int min_num_3dpoints, double* out_points_host,
                                 int32_t* out_idx_host, int* count);
/* convert_3Dpoints_camera + compute_scale_factor -- VO_utility.cpp:23-63 (visual_odometry.h:365-368) */
UVO_API int uvo_scale_factor(uvo_ctx* ctx, const double* points_nx3_host, int n, const double R[9],
                             const double t[3], float range, double* scale_factor);
/* the same, also returning how many points convert_3Dpoints_camera kept (the columns of good_currCam_points): the
 * node assigns SF whenever that set is non-empty -- visual_odometry.h:365-374 -- which is not the same as SF != 0
 * (range == 0 before the first altimeter message gives SF = 0 with a non-empty set) */
UVO_API int uvo_scale_factor_front(uvo_ctx* ctx, const double* points_nx3_host, int n, const double R[9],
                                   const double t[3], float range, double* scale_factor, int* n_front);

/* ---------------------------------------------------------------------------------------------------- K10 */
/* cv::solvePnPRansac(X, x, K, 0-dist, rvec, tvec, false, iters, err, conf, inliers, SOLVEPNP_EPNP)
 * -- visual_odometry.h:647-648.  X: n x 3 f64, x: n x 2 f32.  inliers ascending, *n_inliers = 0 on failure. */
UVO_API int uvo_solve_pnp_ransac(uvo_ctx* ctx, const double* X_host, const float* x_host, int n, const double K[4],
                                 int iterations, float reprojection_error, double confidence, double rvec[3],
                                 double tvec[3], int32_t* inliers_host, int* n_inliers, int* hyps_evaluated);
/* diagnostics: when enabled, uvo_solve_pnp_ransac records SM clock stamps (clock64) at the phase boundaries of
 * hypothesis 0 (slots 0-9: start, correspondences read, control points, M^T M, eigenvectors, betas, pose, model
 * stored, scored, bookkeeping) and of the refit (slots 16-23: start, inlier list, control points, M^T M, eigenvectors,
 * betas, poses, done).  `stamps` (nullable) receives the stamps of the previous call. */
UVO_API int uvo_pnp_profile(uvo_ctx* ctx, int enable, int64_t stamps[32]);

/* ---------------------------------------------------------------------------------------------------- K10a / K10b */
/* cv::findHomography(p1, p2, method, ransacReprojThreshold, mask, maxIters, confidence) -- VO_utility.cpp:152.
 * p1, p2: n x 2 f32 (host).  method: 8 = RANSAC, 4 = LMEDS.  H (row-major, H[8] = 1) is the model re-fitted on the
 * inliers of the best hypothesis and LM-refined; mask is the refined model's mask at the threshold (OpenCV 4.13
 * behaviour, see oracle/twoview.py).  *ok = 0 (H zeroed, mask zeroed) when no model is found. */
UVO_API int uvo_find_homography(uvo_ctx* ctx, const float* p1_host, const float* p2_host, int n, int method,
                                double threshold, int max_iters, double confidence, double H[9], uint8_t* mask_host,
                                int* n_inliers, int* hyps_evaluated, int* ok);
/* cv::findEssentialMat(p1, p2, K, method, prob, threshold, maxIters, mask) -- VO_utility.cpp:147.  K = fx, fy, cx, cy.
 * Nister 5-point hypotheses; E is the best hypothesis (unit Frobenius norm), mask its inliers. */
UVO_API int uvo_find_essential_mat(uvo_ctx* ctx, const float* p1_host, const float* p2_host, int n, const double K[4],
                                   int method, double prob, double threshold, int max_iters, double E[9],
                                   uint8_t* mask_host, int* n_inliers, int* hyps_evaluated, int* ok);
/* cv::recoverPose(E, p1, p2, K, R, t, mask) -- VO_utility.cpp:149: cheirality vote (distanceThresh = 50) over the 4
 * decompositions; mask_inout (nullable) is AND-ed in and overwritten with the winner's validity, *good = its count. */
UVO_API int uvo_recover_pose(uvo_ctx* ctx, const double E[9], const float* p1_host, const float* p2_host, int n,
                             const double K[4], uint8_t* mask_inout_host, double R[9], double t[3], int* good);
/* int recover_pose_homography(H, inliers1, inliers2, K, R, t) -- VO_utility.h / VO_utility.cpp:581-624:
 * decomposeHomographyMat + triangulation vote (0 < z < HOMOGRAPHY_DISTANCE), t normalised.  Returns the vote count
 * in *good; *found = 0 (R, t untouched = zero) when no candidate has a good point.  The reference reads the f32
 * depth through at<double> (undefined behaviour); this implements the evident intent. */
UVO_API int uvo_recover_pose_homography(uvo_ctx* ctx, const double H[9], const float* p1_host, const float* p2_host,
                                        int n, const double K[4], double homography_distance, double R[9],
                                        double t[3], int* good, int* found);
/* void estimate_relative_pose(kp1, kp2, K, R, t, inliers1, inliers2, inlier_matches, success) -- VO_utility.h:101,
 * VO_utility.cpp:134-180: essential or homography branch chosen by *use_essential_inout (the sticky global,
 * VO_utility.h:89), VPF / MIN_NUM_INLIERS gate, one switch of method on failure.  inlier_mask_host receives the mask
 * extract_inliers used (the findEssentialMat / findHomography mask, BEFORE recoverPose shrinks it). */
UVO_API int uvo_estimate_relative_pose(uvo_ctx* ctx, const float* p1_host, const float* p2_host, int n,
                                       const double K[4], const uvo_params* prm, int* use_essential_inout,
                                       double R[9], double t[3], uint8_t* inlier_mask_host, int* n_inliers,
                                       int* success);

/* ---------------------------------------------------------------------------------------------------- frames */
/* Device-resident replay of visual_odometry_node::stereo_VO's per-frame body (visual_odometry.h:526-740):
 * one call per stereo pair; previous-frame state (after-stereo-match keypoints/descriptors, last t) lives on the
 * GPU.  The first successful call initialises (visual_odometry.h:474-520) and reports initialised=1, valid=0. */
typedef struct uvo_stereo uvo_stereo;

typedef struct {
  int32_t initialised;          /* vo_initialized */
  int32_t valid;                /* successful_estimate.data */
  int32_t n_left, n_right;      /* SURF keypoints */
  int32_t n_stereo_matches;     /* results_match_curr.size() */
  int32_t n_temporal_matches;   /* results_match_prev_curr.size() */
  int32_t n_3d;                 /* good_prevCam_points.rows */
  int32_t n_inliers;            /* inliers_idx.rows */
  int32_t hyps_evaluated;       /* RANSAC iterations the reference loop would have run */
  int32_t gate;                 /* 0 ok, else which "ASSUMING CONSTANT MOTION" branch fired (1..5) */
  double rvec[3], tvec[3];      /* R_currCam_prevCam_Vec, t_currCam_prevCam (solvePnPRansac output) */
  double t_prev_curr[3];        /* t_prevCam_currCam = -R^T t (stale on gate failure, visual_odometry.h:717) */
  double velocity[3];           /* t_prevCam_currCam / dt (stereo_output_computation, :148-159) */
} uvo_stereo_result;

UVO_API int uvo_stereo_create(uvo_ctx* ctx, int width, int height, const uvo_camera* left, const uvo_camera* right,
                              const double R_right[9], const double t_right[3], const uvo_params* prm,
                              uvo_stereo** out);
UVO_API void uvo_stereo_destroy(uvo_stereo* s);
/* Host images in, 1 result struct out (H2D + D2H inside). */
UVO_API int uvo_stereo_frame(uvo_stereo* s, const uint8_t* left3_host, const uint8_t* right3_host, size_t pitch,
                             double dt, uvo_stereo_result* out);
/* Same, images already resident in device memory (cudaMalloc'd, pitch bytes per row). */
UVO_API int uvo_stereo_frame_device(uvo_stereo* s, const uint8_t* left3_dev, const uint8_t* right3_dev,
                                    size_t pitch, double dt, uvo_stereo_result* out);
/* Asynchronous calls: enqueue a frame without waiting; results are collected in order.  Consecutive frames run on
 * separate CUDA streams inside the library (the front end of frame t+1 does not depend on frame t), so keeping
 * uvo_stereo_lanes() frames in flight is what fills the GPU; at most uvo_stereo_max_in_flight() frames may be pending.
 * _host: images in host memory (pinned for a truly asynchronous copy); the H2D copies are part of the enqueued work and
 * the buffers may be reused once the frame has been collected. */
UVO_API int uvo_stereo_enqueue_device(uvo_stereo* s, const uint8_t* left3_dev, const uint8_t* right3_dev,
                                      size_t pitch, double dt);
UVO_API int uvo_stereo_enqueue_host(uvo_stereo* s, const uint8_t* left3_host, const uint8_t* right3_host,
                                    size_t pitch, double dt);
/* same with 1-channel BGGR bayer images (pitch >= width): demosaiced on the device before get_image, as the node's
 * image callback does on the CPU (visual_odometry.h:88-91 -> from_ros_to_cv_image) */
UVO_API int uvo_stereo_enqueue_host_bayer(uvo_stereo* s, const uint8_t* left1_host, const uint8_t* right1_host,
                                          size_t pitch, double dt);
UVO_API int uvo_stereo_max_in_flight(void);
/* number of lanes: frames whose kernels can be on the GPU at the same time (frames enqueued beyond that wait for a lane;
 * keeping this many in flight gives the highest frame rate) */
UVO_API int uvo_stereo_lanes(void);
/* COMPRESSED input (the step before the path: the node receives sensor_msgs::CompressedImage and from_ros_to_cv_image
 * decodes it, math_utility.cpp:154-173).  The frame travels to the GPU as entropy-decoded sparse coefficients (about
 * a quarter of the raw image's bytes); IDCT, chroma upsampling and colour conversion (or, with bayer_bggr, the demosaic
 * of a 1-component mosaic) run on the frame's lane ahead of get_image.  Results are identical to enqueueing the image
 * cv::imdecode would have produced.  The image size must be the handle's.
 *   _sparse: the caller has run uvo_jpeg_entropy_decode_sparse (typically on several host threads -- the Huffman
 *            decode is the serial part, ~2 ms per 1280x1024 image and core);
 *   _jpeg:   the library does it, the two images of the pair on two host threads, inside the call. */
UVO_API int uvo_stereo_enqueue_host_sparse(uvo_stereo* s, const uvo_jpeg_sparse* left, const uvo_jpeg_sparse* right,
                                           int bayer_bggr, double dt);
UVO_API int uvo_stereo_enqueue_host_jpeg(uvo_stereo* s, const uint8_t* left_jpeg, size_t left_len,
                                         const uint8_t* right_jpeg, size_t right_len, int bayer_bggr, double dt);
/* _jpeg decodes the Huffman data ON THE GPU when both streams qualify (uvo_jpeg_gpu_entropy: one interleaved scan,
 * no restart interval): the host then only walks the markers and copies the scan bytes (~0.5 MB per pair) -- no
 * per-coefficient host work.  enable = 0 forces the host decoder; _frames counts the frames that took the GPU decoder.
 * A frame whose entropy-coded data is corrupt is reported by uvo_stereo_collect (UVO_ERR_INVALID). */
UVO_API int uvo_stereo_set_gpu_entropy(uvo_stereo* s, int enable);
UVO_API int64_t uvo_stereo_gpu_entropy_frames(const uvo_stereo* s);
/* The asynchronous entry points replay each lane's fixed runs of kernels as CUDA graphs (three graph launches + a few
 * direct launches per frame instead of ~30 kernel launches; results are identical).  `enable` = 0 goes back to direct
 * launches (diagnostics / A-B measurements); uvo_stereo_graph_launches counts the graph launches made so far. */
UVO_API int uvo_stereo_set_graphs(uvo_stereo* s, int enable);
UVO_API int64_t uvo_stereo_graph_launches(const uvo_stereo* s);
UVO_API int uvo_stereo_collect(uvo_stereo* s, uvo_stereo_result* out);
/* Debug/parity taps of the last frame (device -> host copies of intermediate products).  Descriptor rows are 64
 * floats, 128 when the handle was created with surf_extended. */
UVO_API int uvo_stereo_last_keypoints(uvo_stereo* s, int right, uvo_keypoint* kps_host, float* desc_host,
                                      int capacity, int* count);
UVO_API int uvo_stereo_last_matches(uvo_stereo* s, int temporal, uvo_dmatch* matches_host, int capacity,
                                    int* count);
UVO_API int uvo_stereo_last_inliers(uvo_stereo* s, int32_t* inliers_host, int capacity, int* count);
/* per-stage device time of the last synchronous frame, in ms (CUDA events on the ctx stream) */
#define UVO_N_STAGES 8
UVO_API int uvo_stereo_stage_ms(uvo_stereo* s, float ms[UVO_N_STAGES]);
UVO_API const char* uvo_stage_name(int i);

/* ---------------------------------------------------------------------------------------------------- mono frames */
/* visual_odometry_node::mono_VO's per-frame body (visual_odometry.h:247-397) behind one handle: get_image ->
 * detect_features -> match_features(7-arg) -> select_estimation_method -> estimate_relative_pose -> triangulatePoints
 * -> extract_3Dpoints -> convert_3Dpoints_camera / compute_scale_factor -> mono_output_computation.  Previous-frame
 * keypoints / descriptors, the sticky use_essential flag and the last R, t, SF ("ASSUMING CONSTANT MOTION",
 * visual_odometry.h:343-382) live in the handle. */
typedef struct uvo_mono uvo_mono;

typedef struct {
  int32_t initialised;     /* vo_initialized */
  int32_t skipped;         /* a "SKIP IMAGE" gate fired: nothing is published for this frame */
  int32_t published;       /* mono_output_computation ran */
  int32_t valid;           /* successful_estimate.data */
  int32_t used_essential;  /* use_essential after estimate_relative_pose */
  int32_t n_keypoints, n_matches, n_inliers, n_3d;
  int32_t reserved;
  double R[9], t[3];       /* R_currCam_prevCam, t_currCam_prevCam (unit norm) */
  double scale_factor;     /* SF */
  double velocity[3];      /* -SF * R^T t / dt */
} uvo_mono_result;

UVO_API int uvo_mono_create(uvo_ctx* ctx, int width, int height, const uvo_camera* cam, const uvo_params* prm,
                            uvo_mono** out);
UVO_API void uvo_mono_destroy(uvo_mono* m);
/* one image (3-channel interleaved u8, host) + the altimeter range of that instant (visual_odometry.h:367) */
UVO_API int uvo_mono_frame(uvo_mono* m, const uint8_t* img3_host, size_t pitch, double dt, double range,
                           uvo_mono_result* out);
UVO_API int uvo_mono_frame_device(uvo_mono* m, const uint8_t* img3_dev, size_t pitch, double dt, double range,
                                  uvo_mono_result* out);

#ifdef __cplusplus
}
#endif
#endif /* UVO_C_H */
