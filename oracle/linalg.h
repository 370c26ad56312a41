// oracle/linalg.h -- small dense fp64 linear algebra for the CPU ORACLE (test infrastructure only).
// One-sided Jacobi SVD in the style of OpenCV's JacobiSVDImpl_ (modules/core/src/lapack.cpp); the cv2 wheel itself
// links LAPACK, so solver outputs are validated against cv2 to tolerance, never bit-for-bit.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <utility>
#include <vector>

namespace orc {

// A: m x n row-major (m >= n not required).  Computes A = U diag(w) Vt with w descending.
// U: m x n (thin, columns = left singular vectors), Vt: n x n.  Works on the transpose like OpenCV.
// round_robin = false: OpenCV's cyclic pair order (i, j), i < j.  round_robin = true: the circle-method order of
// jacobi_eigh below -- rounds of index-disjoint column pairs over NP = n rounded up to even (a pair with the padding
// index is skipped); the pairs of a round touch disjoint columns, so the GPU rotates them side by side
// (csrc/linalg.cuh: jacobi_sweeps_rr) and obtains the same bits.  Used for EPnP's 6 x {3, 4, 5} beta systems only.
inline void jacobi_svd(const double* A, int m, int n, double* w, double* U, double* Vt, bool round_robin = false) {
  // At: n rows of length m (columns of A)
  std::vector<double> At((size_t)n * m), V((size_t)n * n, 0.0), W(n);
  for (int i = 0; i < n; i++)
    for (int k = 0; k < m; k++) At[(size_t)i * m + k] = A[(size_t)k * n + i];
  for (int i = 0; i < n; i++) {
    double sd = 0;
    for (int k = 0; k < m; k++) sd += At[(size_t)i * m + k] * At[(size_t)i * m + k];
    W[i] = sd;
    V[(size_t)i * n + i] = 1;
  }
  const double eps = DBL_EPSILON * 10;
  const int max_iter = std::max(m, 30);
  // the pair list of one sweep
  std::vector<std::pair<int, int>> order;
  if (!round_robin) {
    for (int i = 0; i < n - 1; i++)
      for (int j = i + 1; j < n; j++) order.emplace_back(i, j);
  } else {
    const int NP = (n + 1) & ~1;
    for (int r = 0; r < NP - 1; r++)
      for (int k = 0; k < NP / 2; k++) {
        const int a = k == 0 ? NP - 1 : (r + k) % (NP - 1), b = k == 0 ? r : (r - k + NP - 1) % (NP - 1);
        const int i = std::min(a, b), j = std::max(a, b);
        if (j < n) order.emplace_back(i, j);
      }
  }
  for (int iter = 0; iter < max_iter; iter++) {
    bool changed = false;
    for (const auto& pr : order) {
        const int i = pr.first, j = pr.second;
        double* Ai = &At[(size_t)i * m];
        double* Aj = &At[(size_t)j * m];
        double a = W[i], p = 0, b = W[j];
        for (int k = 0; k < m; k++) p += Ai[k] * Aj[k];
        if (std::abs(p) <= eps * std::sqrt(a * b)) continue;
        p *= 2;
        double beta = a - b, c, s;
        if (round_robin) {
          // the same rotation with gamma = sqrt(p^2 + beta^2) written out and the two branches folded into one
          // expression (0.5 and 2 are exact factors: (g - b) 0.5 / g = (g + |b|) / (2 g) for b < 0) -- the form the
          // GPU evaluates without a branch, kept identical here so that the two agree bit for bit
          const double gamma = std::sqrt(p * p + beta * beta);
          const double r1 = std::sqrt((gamma + std::fabs(beta)) / (gamma * 2)), r2 = p / (gamma * r1 * 2);
          c = beta < 0 ? r2 : r1;
          s = beta < 0 ? r1 : r2;
        } else {
          const double gamma = std::hypot(p, beta);
          if (beta < 0) {
            double delta = (gamma - beta) * 0.5;
            s = std::sqrt(delta / gamma);
            c = p / (gamma * s * 2);
          } else {
            c = std::sqrt((gamma + beta) / (gamma * 2));
            s = p / (gamma * c * 2);
          }
        }
        a = b = 0;
        for (int k = 0; k < m; k++) {
          double t0 = c * Ai[k] + s * Aj[k];
          double t1 = -s * Ai[k] + c * Aj[k];
          Ai[k] = t0;
          Aj[k] = t1;
          a += t0 * t0;
          b += t1 * t1;
        }
        W[i] = a;
        W[j] = b;
        changed = true;
        double* Vi = &V[(size_t)i * n];
        double* Vj = &V[(size_t)j * n];
        for (int k = 0; k < n; k++) {
          double t0 = c * Vi[k] + s * Vj[k];
          double t1 = -s * Vi[k] + c * Vj[k];
          Vi[k] = t0;
          Vj[k] = t1;
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; i++) {
    double sd = 0;
    for (int k = 0; k < m; k++) sd += At[(size_t)i * m + k] * At[(size_t)i * m + k];
    W[i] = std::sqrt(sd);
  }
  // selection sort, descending
  for (int i = 0; i < n - 1; i++) {
    int j = i;
    for (int k = i + 1; k < n; k++)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      std::swap(W[i], W[j]);
      for (int k = 0; k < m; k++) std::swap(At[(size_t)i * m + k], At[(size_t)j * m + k]);
      for (int k = 0; k < n; k++) std::swap(V[(size_t)i * n + k], V[(size_t)j * n + k]);
    }
  }
  // left vectors u_i = A v_i / w_i; for (numerically) zero singular values complete U to an orthonormal set by
  // Gram-Schmidt on the coordinate axes (OpenCV uses random vectors there, LAPACK its own completion: the choice
  // is arbitrary by construction).
  const double minval = W[0] * DBL_EPSILON * 4 + DBL_MIN * 100;
  for (int i = 0; i < n; i++) {
    w[i] = W[i];
    if (Vt)
      for (int k = 0; k < n; k++) Vt[(size_t)i * n + k] = V[(size_t)i * n + k];
    if (!U) continue;
    if (W[i] > minval) {
      double s = 1 / W[i];
      for (int k = 0; k < m; k++) U[(size_t)k * n + i] = At[(size_t)i * m + k] * s;
    } else {
      bool done = false;
      for (int ax = 0; ax < m && !done; ax++) {
        std::vector<double> v(m, 0.0);
        v[ax] = 1;
        for (int rep = 0; rep < 2; rep++)
          for (int j = 0; j < i; j++) {
            double d = 0;
            for (int k = 0; k < m; k++) d += v[k] * U[(size_t)k * n + j];
            for (int k = 0; k < m; k++) v[k] -= d * U[(size_t)k * n + j];
          }
        double nn = 0;
        for (int k = 0; k < m; k++) nn += v[k] * v[k];
        if (nn > 1e-6) {
          nn = 1 / std::sqrt(nn);
          for (int k = 0; k < m; k++) U[(size_t)k * n + i] = v[k] * nn;
          done = true;
        }
      }
      if (!done)
        for (int k = 0; k < m; k++) U[(size_t)k * n + i] = 0;
    }
  }
}

// Eigen-decomposition of a symmetric N x N matrix (N even) by the two-sided Jacobi method in ROUND-ROBIN order:
// A = Vt^T diag(w) Vt, w descending, rows of Vt = eigenvectors.  Used for EPnP's 12 x 12 M^T M, whose 2-dimensional
// null space (5-point minimal sets) keeps a Hestenes SVD with a relative stopping rule rotating noise for all 30
// sweeps; with the absolute threshold eps * trace this converges in 6-8 sweeps.  The cv2 wheel's LAPACK produces yet
// another basis of that null space, so no choice is bit-comparable with OpenCV; the CUDA kernels
// (csrc/linalg.cuh: jacobi_eigh_rr_warp) and this routine run the same sequence of IEEE operations (no FMA, sqrt and
// divide correctly rounded on both), so the two agree bit for bit.
//
// Order: a sweep is N-1 rounds of N/2 index-disjoint pairs (the circle method: index N-1 stays, the others rotate --
// round r pairs N-1 with r and (r+k) mod (N-1) with (r-k) mod (N-1), k = 1..N/2-1).  Disjoint rotations commute, so
// the N/2 rotation angles of a round are all taken from the matrix as it stands at the start of the round and the
// round is applied as one congruence S' = J^T S J with J = the product of the round's rotations.  Element (i, j), i < j,
// of S' is evaluated as: columns first (the rotation of j's pair applied to rows i and partner(i)), then rows (the
// rotation of i's pair); the lower triangle mirrors the upper one.  The N/2 rotations of a round have no data
// dependence on each other, which is what the GPU uses: one lane per rotation, one lane per element.
template <int N>
struct RoundRobin {
  // partner of index x in round r
  static int partner(int x, int r) {
    if (x == N - 1) return r;
    if (x == r) return N - 1;
    return ((2 * r - x) % (N - 1) + (N - 1)) % (N - 1);
  }
};

template <int N>
inline void jacobi_eigh(const double* A, double* w, double* Vt) {
  static_assert(N % 2 == 0, "round-robin order needs an even dimension");
  // S: full symmetric working copy, Vr: rows converge to the eigenvectors
  double S[N * N], T[N * N], Vr[N * N], V2[N * N], W[N];
  for (int i = 0; i < N * N; i++) S[i] = A[i];
  for (int i = 0; i < N; i++)
    for (int k = 0; k < N; k++) Vr[i * N + k] = (k == i) ? 1.0 : 0.0;
  double tr = 0;
  for (int i = 0; i < N; i++) tr += std::fabs(S[i * N + i]);
  const double thr = tr * DBL_EPSILON;  // absolute: an off-diagonal entry below eps * trace is left alone
  for (int sweep = 0; sweep < 30; sweep++) {
    bool changed = false;
    for (int r = 0; r < N - 1; r++) {
      int pt[N];
      bool low[N], rot[N];   // low: the index is the smaller one (p) of its pair; rot: its pair is rotated this round
      double cc[N], ss[N];
      bool any = false;
      for (int x = 0; x < N; x++) {
        pt[x] = RoundRobin<N>::partner(x, r);
        low[x] = x < pt[x];
      }
      for (int p = 0; p < N; p++) {
        if (!low[p]) continue;
        const int q = pt[p];
        const double apq = S[p * N + q];
        double c = 1.0, s = 0.0;
        const bool rt = std::fabs(apq) > thr;
        if (rt) {
          const double app = S[p * N + p], aqq = S[q * N + q];
          // c = u / h, s = sgn(d) x / h with d = aqq - app, x = 2 apq, u = |d| + sqrt(d^2 + x^2), h^2 = x^2 + u^2
          // (t = s / c = sgn(theta) / (|theta| + sqrt(theta^2 + 1)), theta = d / x): two square roots and one
          // reciprocal on the dependent chain
          const double d = aqq - app, x = 2 * apq;
          const double rr = std::sqrt(d * d + x * x);
          const double u = std::fabs(d) + rr, xs = d >= 0 ? x : -x;
          const double ih = 1 / std::sqrt(x * x + u * u);
          c = u * ih;
          s = xs * ih;
          changed = true;
        }
        cc[p] = cc[q] = c;
        ss[p] = ss[q] = s;
        rot[p] = rot[q] = rt;
        any = any || rt;
      }
      if (!any) continue;  // a round without a rotation leaves S and Vr as they are
      for (int i = 0; i < N; i++)
        for (int j = i; j < N; j++) {
          double v;
          if (i == j) {
            // the pair's own diagonal: a'pp = c^2 app - 2 c s apq + s^2 aqq, a'qq = s^2 app + 2 c s apq + c^2 aqq
            const int i2 = pt[i];
            const double aii = S[i * N + i], a22 = S[i2 * N + i2], a12 = S[i * N + i2], ci = cc[i], si = ss[i];
            v = !rot[i] ? aii
                        : (low[i] ? (ci * ci) * aii - (2 * ci * si) * a12 + (si * si) * a22
                                  : (si * si) * a22 + (2 * ci * si) * a12 + (ci * ci) * aii);
          } else if (pt[i] == j) {
            v = rot[i] ? 0.0 : S[i * N + j];
          } else {
            const int i2 = pt[i], j2 = pt[j];
            // columns j / j2 of rows i and i2, then rows i / i2
            const double bi = low[j] ? cc[j] * S[i * N + j] - ss[j] * S[i * N + j2]
                                     : ss[j] * S[i * N + j2] + cc[j] * S[i * N + j];
            const double bi2 = low[j] ? cc[j] * S[i2 * N + j] - ss[j] * S[i2 * N + j2]
                                      : ss[j] * S[i2 * N + j2] + cc[j] * S[i2 * N + j];
            v = low[i] ? cc[i] * bi - ss[i] * bi2 : ss[i] * bi2 + cc[i] * bi;
          }
          T[i * N + j] = v;
          T[j * N + i] = v;
        }
      for (int x = 0; x < N; x++)
        for (int k = 0; k < N; k++) {
          const double a = Vr[x * N + k], b = Vr[pt[x] * N + k];
          V2[x * N + k] = low[x] ? cc[x] * a - ss[x] * b : ss[x] * b + cc[x] * a;
        }
      for (int i = 0; i < N * N; i++) {
        S[i] = T[i];
        Vr[i] = V2[i];
      }
    }
    if (!changed) break;
  }
  for (int i = 0; i < N; i++) W[i] = S[i * N + i];
  for (int i = 0; i < N - 1; i++) {  // selection sort, descending
    int j = i;
    for (int k = i + 1; k < N; k++)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      double tmp = W[i];
      W[i] = W[j];
      W[j] = tmp;
      for (int k = 0; k < N; k++) {
        tmp = Vr[i * N + k];
        Vr[i * N + k] = Vr[j * N + k];
        Vr[j * N + k] = tmp;
      }
    }
  }
  for (int i = 0; i < N; i++) {
    w[i] = W[i];
    for (int k = 0; k < N; k++) Vt[i * N + k] = Vr[i * N + k];
  }
}

// least-squares / pseudo-inverse solve of A x = b via SVD (cvSolve(..., CV_SVD)); A m x n, b m, x n
inline void svd_solve(const double* A, int m, int n, const double* b, double* x, bool round_robin = false) {
  std::vector<double> w(n), U((size_t)m * n), Vt((size_t)n * n);
  jacobi_svd(A, m, n, w.data(), U.data(), Vt.data(), round_robin);
  double thr = 0;
  for (int i = 0; i < n; i++) thr += w[i];
  thr *= DBL_EPSILON * 2;
  for (int k = 0; k < n; k++) x[k] = 0;
  for (int i = 0; i < n; i++) {
    if (w[i] <= thr) continue;
    double s = 0;
    for (int k = 0; k < m; k++) s += U[(size_t)k * n + i] * b[k];
    s /= w[i];
    for (int k = 0; k < n; k++) x[k] += s * Vt[(size_t)i * n + k];
  }
}

// pseudo-inverse of a 3x3 (cvInvert(CV_SVD))
inline void svd_invert3(const double A[9], double Ainv[9]) {
  for (int c = 0; c < 3; c++) {
    double e[3] = {0, 0, 0}, x[3];
    e[c] = 1;
    svd_solve(A, 3, 3, e, x);
    for (int r = 0; r < 3; r++) Ainv[r * 3 + c] = x[r];
  }
}

inline void mat3_mul(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
inline double det3(const double m[9]) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

}  // namespace orc
