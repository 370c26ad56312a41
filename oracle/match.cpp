// oracle/match.cpp -- CPU ORACLE (test infrastructure only; see uvo_oracle.h).
// Restates BFMatcher(NORM_L2, crossCheck=false).knnMatch(d1, d2, k=2) + Lowe's ratio test as used by match_features
// (reference VO_utility.cpp:515-573).  OpenCV sources restated: modules/core/src/batch_distance.cpp
// (batchDistL2_32f, BatchDistInvoker k-insertion) and modules/core/src/norm.cpp (normL2Sqr_).
// PINNED: bit-equal distances and identical indices against cv2 4.13 (tests/golden/matcher_300x400.npz and the
// live cv2 comparison in tests/test_oracle_match.py).  The f32 summation order below (4 accumulators of 4 SSE lanes,
// no FMA, ((d0+d1)+d2)+d3, then (s0+s2)+(s1+s3)) was identified by probing: it is OpenCV's baseline (SSE) build of
// normL2Sqr_, which the AVX2/AVX-512 capable hosts here also run.
#include "uvo_oracle.h"

#include <cmath>
#if defined(__SSE2__)
#include <immintrin.h>
#endif
#include <limits>
#include <thread>
#include <vector>

static inline float norm_l2_sqr(const float* a, const float* b, int n) {
  int j = 0;
  float s[4];
#if defined(__SSE2__)
  // the same operations in the same order as the scalar branch below, four lanes at a time (this IS the baseline
  // SSE build of normL2Sqr_; separate multiply and add, no FMA): bit-identical, about twice as fast
  __m128 acc0 = _mm_setzero_ps(), acc1 = acc0, acc2 = acc0, acc3 = acc0;
  for (; j <= n - 16; j += 16) {
    __m128 t0 = _mm_sub_ps(_mm_loadu_ps(a + j), _mm_loadu_ps(b + j));
    __m128 t1 = _mm_sub_ps(_mm_loadu_ps(a + j + 4), _mm_loadu_ps(b + j + 4));
    __m128 t2 = _mm_sub_ps(_mm_loadu_ps(a + j + 8), _mm_loadu_ps(b + j + 8));
    __m128 t3 = _mm_sub_ps(_mm_loadu_ps(a + j + 12), _mm_loadu_ps(b + j + 12));
    acc0 = _mm_add_ps(acc0, _mm_mul_ps(t0, t0));
    acc1 = _mm_add_ps(acc1, _mm_mul_ps(t1, t1));
    acc2 = _mm_add_ps(acc2, _mm_mul_ps(t2, t2));
    acc3 = _mm_add_ps(acc3, _mm_mul_ps(t3, t3));
  }
  _mm_storeu_ps(s, _mm_add_ps(_mm_add_ps(_mm_add_ps(acc0, acc1), acc2), acc3));
#else
  float acc[4][4] = {};
  for (; j <= n - 16; j += 16)
    for (int k = 0; k < 4; k++)
      for (int l = 0; l < 4; l++) {
        float t = a[j + 4 * k + l] - b[j + 4 * k + l];
        acc[k][l] = acc[k][l] + t * t;
      }
  for (int l = 0; l < 4; l++) s[l] = ((acc[0][l] + acc[1][l]) + acc[2][l]) + acc[3][l];
#endif
  float d = (s[0] + s[2]) + (s[1] + s[3]);
  for (; j < n; j++) {
    float t = a[j] - b[j];
    d += t * t;
  }
  return d;
}

extern "C" void orc_knn2(const float* q, int nq, const float* t, int nt, int dim, orc_dmatch* out) {
  auto work = [&](int i0, int i1) {
    for (int i = i0; i < i1; i++) {
      float bd[2] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
      int bi[2] = {-1, -1};
      for (int j = 0; j < nt; j++) {
        const float d = std::sqrt(norm_l2_sqr(q + (size_t)i * dim, t + (size_t)j * dim, dim));
        if (d < bd[1]) {  // BatchDistInvoker: insert, shifting entries that are strictly greater
          int k = 0;
          if (bd[0] > d) {
            bd[1] = bd[0];
            bi[1] = bi[0];
            k = 0;
          } else {
            k = 1;
          }
          bd[k] = d;
          bi[k] = j;
        }
      }
      for (int k = 0; k < 2; k++) {
        out[2 * (size_t)i + k].queryIdx = i;
        out[2 * (size_t)i + k].trainIdx = bi[k];
        out[2 * (size_t)i + k].imgIdx = 0;
        out[2 * (size_t)i + k].distance = bi[k] >= 0 ? bd[k] : 0.f;
      }
    }
  };
  unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if ((long)nq * nt < 100000) nthr = 1;
  std::vector<std::thread> th;
  for (unsigned k = 0; k < nthr; k++) th.emplace_back(work, (int)((long)nq * k / nthr), (int)((long)nq * (k + 1) / nthr));
  for (auto& x : th) x.join();
}

// match_features (VO_utility.cpp:533-540).  With fewer than two train descriptors the reference indexes
// knn_matches[i][1] out of bounds; the restatement emits no match in that case.
extern "C" int orc_match_features(const float* q, int nq, const float* t, int nt, int dim, float ratio,
                                  orc_dmatch* out) {
  std::vector<orc_dmatch> knn((size_t)std::max(nq, 1) * 2);
  orc_knn2(q, nq, t, nt, dim, knn.data());
  int n = 0;
  if (nt < 2) return 0;
  for (int i = 0; i < nq; i++)
    if (knn[2 * i].distance < ratio * knn[2 * i + 1].distance) out[n++] = knn[2 * i];
  return n;
}
