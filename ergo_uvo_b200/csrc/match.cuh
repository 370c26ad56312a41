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

constexpr int MATCH_SPLITS = 8;

struct MatchArgs {
  const float* q;      // nq x 64
  const float* t;      // nt x 64
  const int* nq_dev;   // device counts (nullable: use nq/nt)
  const int* nt_dev;
  int nq, nt;          // host-known counts or upper bounds (capacity) when *_dev is set
  float ratio;
  Knn2* partial;       // MATCH_SPLITS x capacity scratch
  Knn2* knn;           // capacity: merged result (always written)
  uvo_dmatch* matches; // capacity: ratio-test survivors in query order
  int* n_matches;      // device counter
};

void launch_match(Ctx& c, const MatchArgs& a);

}  // namespace uvo
