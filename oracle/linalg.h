// oracle/linalg.h -- small dense fp64 linear algebra for the CPU ORACLE (test infrastructure only).
// One-sided Jacobi SVD in the style of OpenCV's JacobiSVDImpl_ (modules/core/src/lapack.cpp); the cv2 wheel itself
// links LAPACK, so solver outputs are validated against cv2 to tolerance, never bit-for-bit.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

namespace orc {

// A: m x n row-major (m >= n not required).  Computes A = U diag(w) Vt with w descending.
// U: m x n (thin, columns = left singular vectors), Vt: n x n.  Works on the transpose like OpenCV.
inline void jacobi_svd(const double* A, int m, int n, double* w, double* U, double* Vt) {
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
  for (int iter = 0; iter < max_iter; iter++) {
    bool changed = false;
    for (int i = 0; i < n - 1; i++)
      for (int j = i + 1; j < n; j++) {
        double* Ai = &At[(size_t)i * m];
        double* Aj = &At[(size_t)j * m];
        double a = W[i], p = 0, b = W[j];
        for (int k = 0; k < m; k++) p += Ai[k] * Aj[k];
        if (std::abs(p) <= eps * std::sqrt(a * b)) continue;
        p *= 2;
        double beta = a - b, gamma = std::hypot(p, beta), c, s;
        if (beta < 0) {
          double delta = (gamma - beta) * 0.5;
          s = std::sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = std::sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
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

// Eigen-decomposition of a symmetric N x N matrix by the cyclic two-sided Jacobi method: A = Vt^T diag(w) Vt, w
// descending, rows of Vt = eigenvectors.  Used for EPnP's 12 x 12 M^T M, whose 2-dimensional null space (5-point
// minimal sets) keeps a Hestenes SVD with a relative stopping rule rotating noise for all 30 sweeps; with the
// absolute threshold eps * trace this converges in 6-8 sweeps.  The cv2 wheel's LAPACK produces yet another basis of
// that null space, so no choice is bit-comparable with OpenCV; the CUDA kernels and the CPU oracle
// run this same sequence of IEEE operations (no FMA, sqrt and divide correctly rounded on both).
template <int N>
inline void jacobi_eigh(const double* A, double* w, double* Vt) {
  // S: full symmetric working copy, Vr: rows converge to the eigenvectors
  double S[N * N], Vr[N * N], W[N];
  for (int i = 0; i < N * N; i++) S[i] = A[i];
  for (int i = 0; i < N; i++)
    for (int k = 0; k < N; k++) Vr[i * N + k] = (k == i) ? 1.0 : 0.0;
  double tr = 0;
  for (int i = 0; i < N; i++) tr += std::fabs(S[i * N + i]);
  const double thr = tr * DBL_EPSILON;  // absolute: an off-diagonal entry below eps * trace is left alone
  for (int sweep = 0; sweep < 30; sweep++) {
    bool changed = false;
    for (int p = 0; p < N - 1; p++)
      for (int q = p + 1; q < N; q++) {
        const double apq = S[p * N + q];
        if (std::fabs(apq) <= thr) continue;
        const double app = S[p * N + p], aqq = S[q * N + q];
        // t = sgn(theta) / (|theta| + std::sqrt(theta^2 + 1)), c = 1 / std::sqrt(t^2 + 1), s = t c with theta = d / x, written
        // so that only two square roots and one reciprocal are on the dependent chain
        const double d = aqq - app, x = 2 * apq;
        const double rr = std::sqrt(d * d + x * x);
        const double u = std::fabs(d) + rr, xs = d >= 0 ? x : -x;
        const double ih = 1 / std::sqrt(x * x + u * u);
        const double t = xs / u, c = u * ih, s = xs * ih;
        S[p * N + p] = app - t * apq;
        S[q * N + q] = aqq + t * apq;
        S[p * N + q] = 0;
        S[q * N + p] = 0;
        for (int k = 0; k < N; k++) {
          if (k == p || k == q) continue;
          const double skp = S[k * N + p], skq = S[k * N + q];
          const double np_ = c * skp - s * skq, nq_ = s * skp + c * skq;
          S[k * N + p] = np_;
          S[p * N + k] = np_;
          S[k * N + q] = nq_;
          S[q * N + k] = nq_;
        }
        for (int k = 0; k < N; k++) {
          const double vp = Vr[p * N + k], vq = Vr[q * N + k];
          Vr[p * N + k] = c * vp - s * vq;
          Vr[q * N + k] = s * vp + c * vq;
        }
        changed = true;
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
inline void svd_solve(const double* A, int m, int n, const double* b, double* x) {
  std::vector<double> w(n), U((size_t)m * n), Vt((size_t)n * n);
  jacobi_svd(A, m, n, w.data(), U.data(), Vt.data());
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
