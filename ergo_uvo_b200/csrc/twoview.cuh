// twoview.cuh -- K10a / K10b: the mono two-view step: findEssentialMat (5-point, RANSAC / LMedS) +
// recoverPose, findHomography (4-point DLT, RANSAC / LMedS, inlier refit + LM refinement) + decomposeHomographyMat
// vote, as estimate_relative_pose / recover_pose_homography call them (reference VO_utility.cpp:134-180, :581-624).
#pragma once
#include "common.cuh"

namespace uvo {

constexpr int TV_HOMOGRAPHY = 0, TV_ESSENTIAL = 1;
constexpr int TV_RANSAC = 8, TV_LMEDS = 4;
constexpr int TV_LSQ = 0;  // findHomography(method = 0): "a regular method using all the points" (no sampling)
constexpr int TV_MAX_MODELS = 10;  // essential matrices per 5-point hypothesis

struct RobustArgs {
  const float* p1;  // n x 2 f32 (device)
  const float* p2;
  int n;
  int kind;    // TV_HOMOGRAPHY / TV_ESSENTIAL
  int method;  // TV_RANSAC / TV_LMEDS / TV_LSQ (homography only)
  double threshold, confidence;
  int iters;   // hypotheses generated: maxIters (RANSAC) or the fixed LMedS count
  double K[4]; // essential: fx, fy, cx, cy
  // scratch (robust_bind_scratch)
  int* subsets;        // iters x 5
  int* ctl;            // [0] subsets produced [1] best hyp [2] best model [3] best good [4] hyps run [5] ok
  double* models;      // iters x models_per_hyp x 9
  int* n_models;       // iters
  float* score;        // iters x models_per_hyp: inlier count (RANSAC) or median error (LMedS)
  double* q;           // n x 4 f64 normalised correspondences (essential)
  double* best;        // [0..8] best hypothesis model, [9] f32 mask threshold (thr^2 or sigma^2)
  // outputs (device)
  double* model_out;   // 9
  uint8_t* mask;       // n
  int* n_inliers;
};
size_t robust_scratch_bytes(int n, int iters, int kind);
void robust_bind_scratch(RobustArgs& a, void* scratch);
// number of hypotheses OpenCV's driver can run for these parameters (LMedS: RANSACUpdateNumIters(conf, 0.45, m, maxIters))
int robust_iterations(int method, int model_points, double confidence, int max_iters);
// runs the whole estimation on the context stream; results stay on the device
void launch_robust(Ctx& c, const RobustArgs& a);

struct RecoverPoseArgs {
  const float* p1;
  const float* p2;
  int n;
  double K[4];
  int from_homography;        // 0: recoverPose(E, ...) ; 1: recover_pose_homography(H, ...)
  const double* model;        // device: E or H (9)
  const uint8_t* mask_in;     // device, nullable (recoverPose's in/out mask)
  double distance_thresh;     // 50 for recoverPose; HOMOGRAPHY_DISTANCE for the homography vote
  // scratch
  double* cand;               // 4 x 12 candidate [R|t] (+ [48] number of candidates)
  uint8_t* flags;             // n: bit k = point valid for candidate k
  int* counts;                // 4
  // outputs (device)
  double* Rt;                 // [0..8] R, [9..11] t (unit norm), [12] good points, [13] ok
  uint8_t* mask_out;          // n (recoverPose only; nullable)
};
void launch_recover_pose(Ctx& c, const RecoverPoseArgs& a);

}  // namespace uvo
