// match.cuh -- K8: brute-force L2 kNN (k=2) + Lowe ratio test.
// Replaces BFMatcher(NORM_L2).knnMatch(d1, d2, knn, 2) + the ratio loop of match_features
// (reference VO_utility.cpp:515-573).
#pragma once
#include "common.cuh"

namespace uvo {

struct Knn2 {  // per query: best and second-best (distance, train index); idx = -1 when absent
  float d0, d1;
  int i0, i1;
};

constexpr int MATCH_MAX_CHUNKS = 8;  // train-set chunks (grid.y of the tensor-core kernel)
constexpr int MATCH_EX_SLICES = 16;  // train-set slices per flagged query in the exact scan
constexpr int MATCH_TOPK = 4;        // candidates kept per (query, chunk, column quarter)
// |tf32-pass similarity - exact| <= 2^-9 |q||t| (both operands truncated to 10 mantissa bits) => 2^-8 on d^2; 1 % slack
constexpr double MATCH_TF32_EPS = 0.00390625 * 1.01;

struct MatchArgs {
  const float* q;      // nq x dim, 16-byte aligned
  const float* t;      // nt x dim
  const int* nq_dev;   // device counts (nullable: use nq/nt)
  const int* nt_dev;
  int nq, nt;          // host-known counts or upper bounds (capacity) when *_dev is set
  int dim = 64;        // floats per descriptor row: 64 (SURF) or 128 (extended SURF)
  int exact_only = 0;  // diagnostics: skip the tensor-core pass, every query takes the exact full scan
  float ratio;
  // optional stereo epipolar / disparity gate applied with the ratio test (off when gate_kq == nullptr):
  // keep iff |y_q - y_t| <= gate_dy and gate_dmin <= x_q - x_t <= gate_dmax
  const uvo_keypoint* gate_kq;
  const uvo_keypoint* gate_kt;
  float gate_dy, gate_dmin, gate_dmax;
  // scratch (match_bind_scratch)
  float* cand;         // capacity x 32 lists x MATCH_TOPK candidate keys (similarity with the column in the low bits)
  float* hb;           // train capacity: -|t|^2 / 2
  Knn2* knn;           // capacity: exact result per query (always written)
  int* fb_list;        // capacity: queries whose candidate set could not be proven complete (this call)
  int* n_flagged;      // ... and how many
  int* ex_done;        // capacity: per flagged query, slices finished (self-cleaning)
  Knn2* ex_part;       // capacity x MATCH_EX_SLICES partial results of the exact scan
  int* n_fallback;     // cumulative n_flagged (statistics)
  unsigned* tn2max;    // float bits of max |t|^2 of the current call (reset by k_knn_compact)
  // outputs
  uvo_dmatch* matches; // capacity: ratio-test survivors in query order
  int* n_matches;      // device counter
};

// bytes of scratch for `cap` queries; the buffer must be zero-filled once after allocation
size_t match_scratch_bytes(int cap_q, int cap_t);
void match_bind_scratch(MatchArgs& a, void* scratch, int cap_q, int cap_t);
void launch_match(Ctx& c, const MatchArgs& a);

}  // namespace uvo
