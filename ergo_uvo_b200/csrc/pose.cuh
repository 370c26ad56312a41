// pose.cuh -- K9-K12: triangulation, 3-D point filtering, PnP-RANSAC (EPnP), scale helpers.
#pragma once
#include "common.cuh"
#include "match.cuh"

namespace uvo {

// raw outputs of cv::RNG(0xFFFFFFFFFFFFFFFF): RANSACPointSetRegistrator::run re-seeds with the same constant on
// every call, so the 32-bit stream is a fixed table; only `% count` and the duplicate redraws are data dependent.
constexpr int RNG_TABLE_SIZE = 1 << 18;
const uint32_t* rng_table_device(Ctx& c);  // lazily uploaded, one per device

struct TriangulateArgs {
  double P1[12], P2[12];
  const float* pts1;   // n x 2 (used when matches == nullptr)
  const float* pts2;
  // gather mode (stereo frame): point i = kps1[matches[i].queryIdx].pt / kps2[matches[i].queryIdx].pt
  const uvo_dmatch* matches;
  const uvo_keypoint* kps1;
  const uvo_keypoint* kps2;
  const int* n_dev;    // device count (nullable)
  int n;               // count or capacity
  int min_points;      // gate: triangulate only if n > min_points (-1: always)
  const int* gate_dev; // optional second gate: *gate_dev == 0 => n = 0
  int* n_out;          // optional: effective n written back (device)
  float* out4;         // 4 x stride f32
  int stride;          // row stride of out4 (>= n)
  float* pts1_out;     // optional: gathered points (n x 2) for later stages
  float* pts2_out;
};
void launch_triangulate(Ctx& c, const TriangulateArgs& a);

struct Extract3dArgs {
  const float* kp1;     // n x 2
  const float* kp2;
  const float* p4;      // 4 x stride
  int stride;
  const int* n_dev;
  int n;
  double R1[9], t1[3], R2[9], t2[3], K1[4], K2[4];
  double tol;
  int min3d;
  double* out_pts;      // n x 3
  int32_t* out_idx;     // n
  int* out_count;       // device
  // scratch (n each)
  double* tmp_pts;
  int32_t* tmp_idx;
};
void launch_extract3d(Ctx& c, const Extract3dArgs& a);

// RANSAC bookkeeping carried from one chunk of iterations to the next (device memory, reset by k_pnp_prepare)
struct PnpState {
  int niters;   // current RANSACUpdateNumIters bound
  int best;     // best inlier count so far
  int best_h;   // its hypothesis (-1: none yet)
  int iter;     // iterations replayed so far
  int ticket;   // blocks of the running chunk that have finished
  int done;     // the sequential loop would have stopped: later chunks return at once
};

struct PnpArgs {
  const double* X;       // n x 3 f64 (down-cast to f32 inside, as solvePnPRansac does)
  const float* x;        // n x 2 f32 (used when x_idx == nullptr)
  // gather mode: x[i] = kps[matches[idx[i]].trainIdx].pt
  const int32_t* x_idx;
  const uvo_dmatch* matches;
  const uvo_keypoint* kps;
  const int* n_dev;
  int n;                 // count or capacity
  double K[4];
  int iterations;
  float reproj_err;
  double confidence;
  int min_points;        // gate: run only if n > min_points (stereo frame: MIN_NUM_3DPOINTS); -1 = always
  // outputs (device)
  double* result;        // [0..2] rvec, [3..5] tvec, [6] ok
  int32_t* inliers;      // n
  int* n_inliers;        // device
  int* hyps;             // device: iterations the sequential loop would have run
  // scratch
  int32_t* subsets;      // iterations x 5
  double* hyp_model;     // iterations x 15 (R 9, t 3, rvec 3)
  int* hyp_good;         // iterations
  float* xs;             // n x 2 gathered image points
  float* Xf;             // n x 3 f32 object points
  int* best;             // [0] best hypothesis, [1] best count
  PnpState* state;       // chunk-to-chunk bookkeeping
  long long* prof;       // optional (diagnostics): clock64 stamps of the phases of hypothesis 0 and of the refit
};
// carves subsets / hyp_model / hyp_good / xs / Xf / state / best out of one allocation of pnp_scratch_bytes(n, iterations)
void pnp_bind_scratch(PnpArgs& a, uint8_t* base, int n, int iterations);
void launch_pnp_ransac(Ctx& c, const PnpArgs& a);   // = prepare + solve
// the two halves, for callers that split them over streams: prepare gathers the correspondences and rebuilds the
// subset stream; solve runs hypotheses, scoring, the sequential bookkeeping replay and the refit
void launch_pnp_prepare(Ctx& c, const PnpArgs& a);
void launch_pnp_solve(Ctx& c, const PnpArgs& a);
size_t pnp_scratch_bytes(int n, int iterations);

// compute_median (math_utility.cpp:65-86) by rank selection: out[0] = median (mean of the two middle values for
// even n), 0.0 for n == 0.
void launch_median(Ctx& c, const double* v, const int* n_dev, int n, double* out);
// select_estimation_method (VO_utility.cpp:725-748): writes the n displacements |p1-p2| (f64) to `disp`
void launch_displacements(Ctx& c, const float* p1, const float* p2, int n, double* disp);
// convert_3Dpoints_camera (VO_utility.cpp:46-63) z-filter: keeps the UN-transformed z of points whose transformed z
// is > 0, order preserved; count -> *n_out
void launch_front_z(Ctx& c, const double* pts, int n, const double R[9], const double t[3], double* z_out, int* n_out);

}  // namespace uvo
