// linalg.cuh -- small dense fp64 linear algebra used by the pose kernels (one problem per thread).
// One-sided (Hestenes) Jacobi SVD with the sweep order, rotation formulas and stopping rule of OpenCV's
// JacobiSVDImpl_; everything is plain IEEE fp64 without FMA contraction (-fmad=false).
#pragma once
#include <cfloat>
#include <cuda_runtime.h>

namespace uvo {

// A: m x n row-major, m,n <= MAXD.  A = U diag(w) Vt, w descending.  U: m x n (may be null), Vt: n x n.
template <int MAXD>
__device__ void jacobi_svd(const double* A, int m, int n, double* w, double* U, double* Vt) {
  double At[MAXD * MAXD], V[MAXD * MAXD], W[MAXD];
  for (int i = 0; i < n; i++)
    for (int k = 0; k < m; k++) At[i * m + k] = A[k * n + i];
  for (int i = 0; i < n; i++) {
    double sd = 0;
    for (int k = 0; k < m; k++) sd += At[i * m + k] * At[i * m + k];
    W[i] = sd;
    for (int k = 0; k < n; k++) V[i * n + k] = (k == i) ? 1.0 : 0.0;
  }
  const double eps = DBL_EPSILON * 10;
  const int max_iter = m > 30 ? m : 30;
  for (int iter = 0; iter < max_iter; iter++) {
    bool changed = false;
    for (int i = 0; i < n - 1; i++)
      for (int j = i + 1; j < n; j++) {
        double* Ai = At + i * m;
        double* Aj = At + j * m;
        double a = W[i], p = 0, b = W[j];
        for (int k = 0; k < m; k++) p += Ai[k] * Aj[k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) {
          const double delta = (gamma - beta) * 0.5;
          s = sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        a = b = 0;
        for (int k = 0; k < m; k++) {
          const double t0 = c * Ai[k] + s * Aj[k];
          const double t1 = -s * Ai[k] + c * Aj[k];
          Ai[k] = t0;
          Aj[k] = t1;
          a += t0 * t0;
          b += t1 * t1;
        }
        W[i] = a;
        W[j] = b;
        changed = true;
        double* Vi = V + i * n;
        double* Vj = V + j * n;
        for (int k = 0; k < n; k++) {
          const double t0 = c * Vi[k] + s * Vj[k];
          const double t1 = -s * Vi[k] + c * Vj[k];
          Vi[k] = t0;
          Vj[k] = t1;
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; i++) {
    double sd = 0;
    for (int k = 0; k < m; k++) sd += At[i * m + k] * At[i * m + k];
    W[i] = sqrt(sd);
  }
  for (int i = 0; i < n - 1; i++) {
    int j = i;
    for (int k = i + 1; k < n; k++)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      double t = W[i];
      W[i] = W[j];
      W[j] = t;
      for (int k = 0; k < m; k++) {
        t = At[i * m + k];
        At[i * m + k] = At[j * m + k];
        At[j * m + k] = t;
      }
      for (int k = 0; k < n; k++) {
        t = V[i * n + k];
        V[i * n + k] = V[j * n + k];
        V[j * n + k] = t;
      }
    }
  }
  const double minval = W[0] * DBL_EPSILON * 4 + DBL_MIN * 100;
  for (int i = 0; i < n; i++) {
    w[i] = W[i];
    if (Vt)
      for (int k = 0; k < n; k++) Vt[i * n + k] = V[i * n + k];
    if (!U) continue;
    if (W[i] > minval) {
      const double s = 1 / W[i];
      for (int k = 0; k < m; k++) U[k * n + i] = At[i * m + k] * s;
    } else {
      // complete U with Gram-Schmidt on the coordinate axes (arbitrary by construction; same rule as the oracle)
      bool done = false;
      for (int ax = 0; ax < m && !done; ax++) {
        double v[MAXD];
        for (int k = 0; k < m; k++) v[k] = (k == ax) ? 1.0 : 0.0;
        for (int rep = 0; rep < 2; rep++)
          for (int j = 0; j < i; j++) {
            double d = 0;
            for (int k = 0; k < m; k++) d += v[k] * U[k * n + j];
            for (int k = 0; k < m; k++) v[k] -= d * U[k * n + j];
          }
        double nn = 0;
        for (int k = 0; k < m; k++) nn += v[k] * v[k];
        if (nn > 1e-6) {
          nn = 1 / sqrt(nn);
          for (int k = 0; k < m; k++) U[k * n + i] = v[k] * nn;
          done = true;
        }
      }
      if (!done)
        for (int k = 0; k < m; k++) U[k * n + i] = 0;
    }
  }
}

// least-squares / pseudo-inverse solve via SVD (cvSolve CV_SVD); m <= 6, n <= 6
__device__ inline void svd_solve6(const double* A, int m, int n, const double* b, double* x) {
  double w[6], U[36], Vt[36];
  jacobi_svd<6>(A, m, n, w, U, Vt);
  double thr = 0;
  for (int i = 0; i < n; i++) thr += w[i];
  thr *= DBL_EPSILON * 2;
  for (int k = 0; k < n; k++) x[k] = 0;
  for (int i = 0; i < n; i++) {
    if (w[i] <= thr) continue;
    double s = 0;
    for (int k = 0; k < m; k++) s += U[k * n + i] * b[k];
    s /= w[i];
    for (int k = 0; k < n; k++) x[k] += s * Vt[i * n + k];
  }
}

__device__ inline void svd_invert3(const double A[9], double Ainv[9]) {
  for (int c = 0; c < 3; c++) {
    double e[3] = {0, 0, 0}, x[3];
    e[c] = 1;
    svd_solve6(A, 3, 3, e, x);
    for (int r = 0; r < 3; r++) Ainv[r * 3 + c] = x[r];
  }
}

__device__ inline void mat3_mul(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

// cv::Rodrigues, both directions
__device__ inline void rodrigues_vec2mat(const double r[3], double R[9]) {
  const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (theta < DBL_EPSILON) {
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
  const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
  const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
  const double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  for (int i = 0; i < 9; i++) R[i] = c * (i % 4 == 0 ? 1. : 0.) + c1 * rrt[i] + s * rx[i];
}

__device__ inline void rodrigues_mat2vec(const double Rin[9], double rv[3]) {
  double w[3], U[9], Vt[9], R[9];
  jacobi_svd<3>(Rin, 3, 3, w, U, Vt);
  mat3_mul(U, Vt, R);
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = acos(c);
  if (s < 1e-5) {
    if (c > 0) {
      rx = ry = rz = 0;
    } else {
      double t = (R[0] + 1) * 0.5;
      rx = sqrt(fmax(t, 0.));
      t = (R[4] + 1) * 0.5;
      ry = sqrt(fmax(t, 0.)) * (R[1] < 0 ? -1. : 1.);
      t = (R[8] + 1) * 0.5;
      rz = sqrt(fmax(t, 0.)) * (R[2] < 0 ? -1. : 1.);
      if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
      theta /= sqrt(rx * rx + ry * ry + rz * rz);
      rx *= theta;
      ry *= theta;
      rz *= theta;
    }
  } else {
    double vth = 1 / (2 * s);
    vth *= theta;
    rx *= vth;
    ry *= vth;
    rz *= vth;
  }
  rv[0] = rx;
  rv[1] = ry;
  rv[2] = rz;
}

// cv::projectPoints with zero distortion (reciprocal multiply, as cvProjectPoints2)
__device__ __forceinline__ void project1(const double X[3], const double R[9], const double t[3], const double K[4],
                                         double m[2]) {
  double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
  z = z ? 1. / z : 1;
  x *= z;
  y *= z;
  m[0] = x * K[0] + K[2];
  m[1] = y * K[1] + K[3];
}

}  // namespace uvo
