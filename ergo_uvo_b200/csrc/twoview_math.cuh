// twoview_math.cuh -- fp64 building blocks of the mono two-view step (K10a / K10b), one problem per thread:
//   symmetric Jacobi eigen-decomposition (cv::eigen as used by the homography DLT and by LMSolver's DECOMP_EIG solves),
//   normalised 4..n point homography DLT pieces (fundam.cpp HomographyEstimatorCallback::runKernel),
//   Nister's 5-point essential-matrix solver (five-point.cpp EMEstimatorCallback::runKernel): null space of the 5 x 9
//   epipolar system, the ten cubic constraints by polynomial arithmetic, Gauss-Jordan on the 10 x 20 matrix, the
//   degree-10 determinant polynomial, Durand-Kerner roots (cv::solvePoly's iteration and starting points), back
//   substitution through the 3 x 3 null vector,
//   decomposeEssentialMat / decomposeHomographyMat (Malis-Vargas) closed forms.
// Model values agree with OpenCV to rounding (~1e-12): like the CPU library itself they are not bit-reproducible
// across LAPACK builds; masks are compared exactly in tests/test_gpu_twoview.py.
#pragma once
#include <cfloat>

#include "linalg.cuh"

namespace uvo {

// RANSACUpdateNumIters (ptsetreg.cpp)
__device__ inline int tv_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmin(fmax(p, 0.), 1.);
  ep = fmin(fmax(ep, 0.), 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : __double2int_rn(num / denom);
}

// ---------------------------------------------------------------------------------------------- symmetric eigen
// A (N x N, symmetric, row-major) is destroyed; w[i] unsorted eigenvalues, V rows = eigenvectors (V[i*N + k]).
template <int N>
__device__ void jacobi_eigen_sym(double* A, double* w, double* V) {
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) V[i * N + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < N; i++) {
      diag += fabs(A[i * N + i]);
      for (int j = i + 1; j < N; j++) off += fabs(A[i * N + j]);
    }
    if (off <= DBL_EPSILON * 1e-3 * diag || off == 0) break;
    for (int p = 0; p < N - 1; p++)
      for (int q = p + 1; q < N; q++) {
        const double apq = A[p * N + q];
        if (fabs(apq) <= DBL_MIN) continue;
        const double app = A[p * N + p], aqq = A[q * N + q];
        if (fabs(apq) < DBL_EPSILON * 1e-4 * (fabs(app) + fabs(aqq))) {
          A[p * N + q] = A[q * N + p] = 0;
          continue;
        }
        const double theta = (aqq - app) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < N; k++) {  // columns p, q
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; k++) {  // rows p, q
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; k++) {
          const double vp = V[p * N + k], vq = V[q * N + k];
          V[p * N + k] = c * vp - s * vq;
          V[q * N + k] = s * vp + c * vq;
        }
      }
  }
  for (int i = 0; i < N; i++) w[i] = A[i * N + i];
}

// x = V diag(1/w) V^T b with eigenvalues below 2 eps * sum|w| dropped (cv::solve(..., DECOMP_EIG) on symmetric A)
template <int N>
__device__ void solve_eig_sym(const double* A, const double* b, double* x, double* diag_inv /* nullable: diag(A^-1) */) {
  double M[N * N], w[N], V[N * N];
  for (int i = 0; i < N * N; i++) M[i] = A[i];
  jacobi_eigen_sym<N>(M, w, V);
  double thr = 0;
  for (int i = 0; i < N; i++) thr += fabs(w[i]);
  thr *= DBL_EPSILON * 2;
  for (int k = 0; k < N; k++) {
    x[k] = 0;
    if (diag_inv) diag_inv[k] = 0;
  }
  for (int i = 0; i < N; i++) {
    if (!(fabs(w[i]) > thr)) continue;
    double t = 0;
    for (int k = 0; k < N; k++) t += V[i * N + k] * b[k];
    t /= w[i];
    for (int k = 0; k < N; k++) {
      x[k] += V[i * N + k] * t;
      if (diag_inv) diag_inv[k] += V[i * N + k] * V[i * N + k] / w[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------- homography DLT
struct HNorm {  // the normalisation of HomographyEstimatorCallback::runKernel
  double cMx, cMy, cmx, cmy, sMx, sMy, smx, smy;
};

// accumulates the upper triangle of LtL (45 values, row-major j <= k) for one correspondence
__device__ __forceinline__ void dlt_accumulate(const HNorm& nm, double Mx, double My, double mx, double my, double* ltl45) {
  const double x = (mx - nm.cmx) * nm.smx, y = (my - nm.cmy) * nm.smy;
  const double X = (Mx - nm.cMx) * nm.sMx, Y = (My - nm.cMy) * nm.sMy;
  const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
  const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
  int o = 0;
  for (int j = 0; j < 9; j++)
    for (int k = j; k < 9; k++) ltl45[o++] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
}

// smallest eigenvector of LtL -> H = invHnorm * H0 * Hnorm2, scaled so that H[8] == 1
__device__ void dlt_solve(const HNorm& nm, const double* ltl45, double H[9]) {
  double A[81], w[9], V[81];
  int o = 0;
  for (int j = 0; j < 9; j++)
    for (int k = j; k < 9; k++) A[j * 9 + k] = A[k * 9 + j] = ltl45[o++];
  jacobi_eigen_sym<9>(A, w, V);
  int best = 0;
  for (int i = 1; i < 9; i++)
    if (w[i] < w[best]) best = i;
  const double* h = V + best * 9;
  const double inv[9] = {1. / nm.smx, 0, nm.cmx, 0, 1. / nm.smy, nm.cmy, 0, 0, 1};
  const double n2[9] = {nm.sMx, 0, -nm.cMx * nm.sMx, 0, nm.sMy, -nm.cMy * nm.sMy, 0, 0, 1};
  double T[9];
  mat3_mul(inv, h, T);
  mat3_mul(T, n2, H);
  const double s = 1. / H[8];
  for (int i = 0; i < 9; i++) H[i] *= s;
}

// 4-point (minimal) kernel; false when the normalisation degenerates
__device__ bool homography_from4(const float* M, const float* m, double H[9]) {
  HNorm nm{};
  for (int i = 0; i < 4; i++) {
    nm.cmx += m[2 * i];
    nm.cmy += m[2 * i + 1];
    nm.cMx += M[2 * i];
    nm.cMy += M[2 * i + 1];
  }
  nm.cmx /= 4, nm.cmy /= 4, nm.cMx /= 4, nm.cMy /= 4;
  for (int i = 0; i < 4; i++) {
    nm.smx += fabs(m[2 * i] - nm.cmx);
    nm.smy += fabs(m[2 * i + 1] - nm.cmy);
    nm.sMx += fabs(M[2 * i] - nm.cMx);
    nm.sMy += fabs(M[2 * i + 1] - nm.cMy);
  }
  if (fabs(nm.smx) < DBL_EPSILON || fabs(nm.smy) < DBL_EPSILON || fabs(nm.sMx) < DBL_EPSILON || fabs(nm.sMy) < DBL_EPSILON)
    return false;
  nm.smx = 4 / nm.smx, nm.smy = 4 / nm.smy, nm.sMx = 4 / nm.sMx, nm.sMy = 4 / nm.sMy;
  double ltl[45];
  for (int i = 0; i < 45; i++) ltl[i] = 0;
  for (int i = 0; i < 4; i++) dlt_accumulate(nm, M[2 * i], M[2 * i + 1], m[2 * i], m[2 * i + 1], ltl);
  dlt_solve(nm, ltl, H);
  return true;
}

// HomographyEstimatorCallback::checkSubset for a complete 4-point subset (f32 points, f64 arithmetic)
__device__ inline bool h_collinear_last(const float* p) {
  const int i = 3;
  for (int j = 0; j < i; j++) {
    const double dx1 = (double)p[2 * j] - p[2 * i], dy1 = (double)p[2 * j + 1] - p[2 * i + 1];
    for (int k = 0; k < j; k++) {
      const double dx2 = (double)p[2 * k] - p[2 * i], dy2 = (double)p[2 * k + 1] - p[2 * i + 1];
      if (fabs(dx2 * dy1 - dy2 * dx1) <= (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
    }
  }
  return false;
}
__device__ inline double det3_pts(const float* p, int a, int b, int c) {
  const double A[9] = {p[2 * a], p[2 * a + 1], 1., p[2 * b], p[2 * b + 1], 1., p[2 * c], p[2 * c + 1], 1.};
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
__device__ inline bool h_check_subset(const float* ms1, const float* ms2) {
  if (h_collinear_last(ms1) || h_collinear_last(ms2)) return false;
  const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
  int negative = 0;
  for (int i = 0; i < 4; i++)
    negative += (det3_pts(ms1, tt[i][0], tt[i][1], tt[i][2]) * det3_pts(ms2, tt[i][0], tt[i][1], tt[i][2]) < 0) ? 1 : 0;
  return negative == 0 || negative == 4;
}

// HomographyEstimatorCallback::computeError for one point: pure f32, no contraction
__device__ __forceinline__ float h_error_f32(const float Hf[8], float Mx, float My, float mx, float my) {
  const float ww = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(Hf[6], Mx), __fmul_rn(Hf[7], My)), 1.f));
  const float dx = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[0], Mx), __fmul_rn(Hf[1], My)), Hf[2]), ww), mx);
  const float dy = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[3], Mx), __fmul_rn(Hf[4], My)), Hf[5]), ww), my);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// ---------------------------------------------------------------------------------------------- 5-point
// polynomial arithmetic in (x, y, z): linear forms (x, y, z, 1), quadratics (x2, y2, z2, xy, xz, yz, x, y, z, 1) and
// cubics in Nister's order (x3, y3, x2y, xy2, x2z, x2, y2z, y2, xyz, xy | xz2, xz, x, yz2, yz, y, z3, z2, z, 1)
__device__ __constant__ const signed char kQI[4][4] = {{0, 3, 4, 6}, {3, 1, 5, 7}, {4, 5, 2, 8}, {6, 7, 8, 9}};
__device__ __constant__ const signed char kCI[10][4] = {{0, 2, 4, 5},     {3, 1, 6, 7},     {10, 13, 16, 17}, {2, 3, 8, 9},
                                                        {4, 8, 10, 11},   {8, 6, 13, 14},   {5, 9, 11, 12},   {9, 7, 14, 15},
                                                        {11, 14, 17, 18}, {12, 15, 18, 19}};

__device__ __forceinline__ void lin_mul_acc(const double* a, const double* b, double sgn, double* q) {  // q += sgn a b
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) q[kQI[i][j]] += sgn * a[i] * b[j];
}
__device__ __forceinline__ void quad_lin_acc(const double* q, const double* l, double sgn, double* c) {  // c += sgn q l
  for (int i = 0; i < 10; i++)
    for (int j = 0; j < 4; j++) c[kCI[i][j]] += sgn * q[i] * l[j];
}

// complex helpers for Durand-Kerner
struct Cx {
  double re, im;
};
__device__ __forceinline__ Cx cmul(Cx a, Cx b) { return Cx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cx csub(Cx a, Cx b) { return Cx{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cx cdiv(Cx a, Cx b) {
  const double t = 1. / (b.re * b.re + b.im * b.im);
  return Cx{(a.re * b.re + a.im * b.im) * t, (a.im * b.re - a.re * b.im) * t};
}

// cv::solvePoly's Durand-Kerner iteration (same starting points (1+i)^k); coeffs[k] multiplies z^k, degree n <= 10.
// OpenCV iterates to an exact fixed point (or 300 n sweeps); here the sweep stops when no root moves by more than
// 1e-15 relative, which is the same fixed point to rounding.
__device__ int solve_poly_dk(const double* coeffs, int n, Cx* roots) {
  while (n > 0 && coeffs[n] == 0) n--;  // OpenCV trims vanishing leading coefficients
  if (n <= 0) return 0;
  Cx p{1, 0};
  const Cx r{1, 1};
  for (int i = 0; i < n; i++) {
    roots[i] = p;
    p = cmul(p, r);
  }
  for (int iter = 0; iter < 600; iter++) {
    double max_diff = 0, scale = 0;
    for (int i = 0; i < n; i++) {
      p = roots[i];
      Cx num{coeffs[n], 0}, den{coeffs[n], 0};
      for (int j = 0; j < n; j++) {
        num = cmul(num, p);
        num.re += coeffs[n - j - 1];
        if (j != i) {
          const Cx d = csub(p, roots[j]);
          if (d.re != 0 || d.im != 0) den = cmul(den, d);
        }
      }
      num = cdiv(num, den);
      roots[i] = csub(p, num);
      max_diff = fmax(max_diff, fmax(fabs(num.re), fabs(num.im)));
      scale = fmax(scale, fmax(fabs(p.re), fabs(p.im)));
    }
    if (!(max_diff > 1e-15 * fmax(scale, 1e-3))) break;
  }
  return n;
}

// The same iteration with one root per lane of a group (`roots` in shared memory, `gl` = lane in the group, every lane
// of the group calls).  All roots of a sweep are updated from the previous sweep's values (the simultaneous
// Weierstrass form; cv::solvePoly updates in place), which reaches the same fixed point: the roots agree with the
// sequential iteration to rounding, and a sweep costs one root's latency instead of ten.  Polynomials with a
// near-vanishing leading coefficient (a root at infinity) take hundreds of sweeps either way; that tail is why the
// one-thread form spent 90 % of the 5-point kernel here.
__device__ int solve_poly_dk_group(const double* coeffs, int n, Cx* roots, int gl, unsigned gmask) {
  while (n > 0 && coeffs[n] == 0) n--;
  if (n <= 0) return 0;
  if (gl == 0) {
    Cx p{1, 0};
    const Cx r{1, 1};
    for (int i = 0; i < n; i++) {
      roots[i] = p;
      p = cmul(p, r);
    }
  }
  __syncwarp(gmask);
  for (int iter = 0; iter < 600; iter++) {
    Cx p{0, 0}, num{0, 0};
    if (gl < n) {
      p = roots[gl];
      num = Cx{coeffs[n], 0};
      Cx den{coeffs[n], 0};
      for (int j = 0; j < n; j++) {
        num = cmul(num, p);
        num.re += coeffs[n - j - 1];
        if (j != gl) {
          const Cx d = csub(p, roots[j]);
          if (d.re != 0 || d.im != 0) den = cmul(den, d);
        }
      }
      num = cdiv(num, den);
    }
    double diff = gl < n ? fmax(fabs(num.re), fabs(num.im)) : 0.0;
    double scale = gl < n ? fmax(fabs(p.re), fabs(p.im)) : 0.0;
    __syncwarp(gmask);  // every lane has read the previous sweep's roots
    if (gl < n) roots[gl] = csub(p, num);
#pragma unroll
    for (int o = 8; o; o >>= 1) {
      diff = fmax(diff, __shfl_xor_sync(gmask, diff, o));
      scale = fmax(scale, __shfl_xor_sync(gmask, scale, o));
    }
    __syncwarp(gmask);
    if (!(diff > 1e-15 * fmax(scale, 1e-3))) break;
  }
  return n;
}

// p (degree da) * q (degree db) -> out (degree da + db); arrays indexed by power (lowest first)
__device__ __forceinline__ void poly_mul(const double* p, int da, const double* q, int db, double* out) {
  for (int i = 0; i <= da + db; i++) out[i] = 0;
  for (int i = 0; i <= da; i++)
    for (int j = 0; j <= db; j++) out[i + j] += p[i] * q[j];
}

// EMEstimatorCallback::runKernel, first half: q1, q2 = 5 normalised correspondences (x, y) -> the null-space basis EE
// (4 x 9), the 3 x 13 matrix B(z) and the degree-10 determinant polynomial c11 (lowest power first).  Returns false
// when the 10 x 10 elimination meets a vanishing pivot (no model).  One thread.
__device__ bool five_point_poly(const double* q1, const double* q2, double* EE_out, double (*B)[13], double* c11) {
  // epipolar rows x2^T E x1 = 0 with E row-major
  double Q[45];
  for (int i = 0; i < 5; i++) {
    const double a[3] = {q2[2 * i], q2[2 * i + 1], 1.}, b[3] = {q1[2 * i], q1[2 * i + 1], 1.};
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Q[i * 9 + 3 * r + c] = a[r] * b[c];
  }
  double w9[9], Vt[81];
  jacobi_svd<9, 5, 9, true>(Q, 5, 9, w9, nullptr, Vt);
  const double* EE = Vt + 5 * 9;  // null-space basis E0..E3 (rows 5..8 of Vt)
  for (int i = 0; i < 36; i++) EE_out[i] = EE[i];
  // linear forms of the nine entries
  double L[9][4];
  for (int e = 0; e < 9; e++)
    for (int k = 0; k < 4; k++) L[e][k] = EE[k * 9 + e];
  double A[10][20];
  for (int r = 0; r < 10; r++)
    for (int c = 0; c < 20; c++) A[r][c] = 0;
  {
    // det(E)
    double q[10];
    auto minor2 = [&](int a, int b, int c, int d) {  // L[a] L[b] - L[c] L[d]
      for (int i = 0; i < 10; i++) q[i] = 0;
      lin_mul_acc(L[a], L[b], 1.0, q);
      lin_mul_acc(L[c], L[d], -1.0, q);
    };
    minor2(4, 8, 5, 7);
    quad_lin_acc(q, L[0], 1.0, A[0]);
    minor2(3, 8, 5, 6);
    quad_lin_acc(q, L[1], -1.0, A[0]);
    minor2(3, 7, 4, 6);
    quad_lin_acc(q, L[2], 1.0, A[0]);
  }
  {
    // 2 E E^T E - trace(E E^T) E
    double EEt[3][3][10], tr[10];
    for (int i = 0; i < 3; i++)
      for (int j = i; j < 3; j++) {
        for (int k = 0; k < 10; k++) EEt[i][j][k] = 0;
        for (int k = 0; k < 3; k++) lin_mul_acc(L[3 * i + k], L[3 * j + k], 1.0, EEt[i][j]);
        if (j != i)
          for (int k = 0; k < 10; k++) EEt[j][i][k] = EEt[i][j][k];
      }
    for (int k = 0; k < 10; k++) tr[k] = EEt[0][0][k] + EEt[1][1][k] + EEt[2][2][k];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double* row = A[1 + 3 * i + j];
        for (int k = 0; k < 3; k++) quad_lin_acc(EEt[i][k], L[3 * k + j], 2.0, row);
        quad_lin_acc(tr, L[3 * i + j], -1.0, row);
      }
  }
  // A <- inv(A[:, 0:10]) A[:, 10:20]: Gauss-Jordan with partial pivoting on the augmented 10 x 20 matrix
  for (int col = 0; col < 10; col++) {
    int piv = col;
    for (int r = col + 1; r < 10; r++)
      if (fabs(A[r][col]) > fabs(A[piv][col])) piv = r;
    if (fabs(A[piv][col]) < DBL_MIN) return false;
    if (piv != col)
      for (int c = 0; c < 20; c++) {
        const double t = A[col][c];
        A[col][c] = A[piv][c];
        A[piv][c] = t;
      }
    const double inv = 1. / A[col][col];
    for (int c = col; c < 20; c++) A[col][c] *= inv;
    for (int r = 0; r < 10; r++) {
      if (r == col) continue;
      const double f = A[r][col];
      if (f == 0) continue;
      for (int c = col; c < 20; c++) A[r][c] -= f * A[col][c];
    }
  }
  // B (3 x 13): <x^2 z> - z <x^2>, <y^2 z> - z <y^2>, <x y z> - z <x y>; columns [x z3, x z2, x z, x | y z3 .. y |
  // z4, z3, z2, z, 1]
  for (int i = 0; i < 3; i++) {
    const double* r1 = A[2 * i + 4] + 10;
    const double* r2 = A[2 * i + 5] + 10;
    double row1[13], row2[13];
    for (int k = 0; k < 13; k++) row1[k] = row2[k] = 0;
    for (int k = 0; k < 3; k++) {
      row1[1 + k] = r1[k];
      row1[5 + k] = r1[3 + k];
      row2[k] = r2[k];
      row2[4 + k] = r2[3 + k];
    }
    for (int k = 0; k < 4; k++) {
      row1[9 + k] = r1[6 + k];
      row2[8 + k] = r2[6 + k];
    }
    for (int k = 0; k < 13; k++) B[i][k] = row1[k] - row2[k];
  }
  // det B(z): entries as polynomials in z, lowest power first
  double P[3][3][5];
  for (int i = 0; i < 3; i++) {
    for (int k = 0; k < 4; k++) {
      P[i][0][k] = B[i][3 - k];
      P[i][1][k] = B[i][7 - k];
    }
    P[i][0][4] = P[i][1][4] = 0;
    for (int k = 0; k < 5; k++) P[i][2][k] = B[i][12 - k];
  }
  for (int k = 0; k < 11; k++) c11[k] = 0;
  {
    double m[8], t[11];
    auto add_term = [&](const double* a, int da, const double* b1, const double* b2, int db1, int db2, const double* c1,
                        const double* c2, double sgn) {
      // sgn * a * (b1 * b2 - c1 * c2), deg(b1 b2) = deg(c1 c2) = db1 + db2
      double m2[8];
      poly_mul(b1, db1, b2, db2, m);
      poly_mul(c1, db1, c2, db2, m2);  // same degrees by construction of the calls below
      for (int k = 0; k <= db1 + db2; k++) m[k] -= m2[k];
      poly_mul(a, da, m, db1 + db2, t);
      for (int k = 0; k <= da + db1 + db2; k++) c11[k] += sgn * t[k];
    };
    // P00 (P11 P22 - P12 P21) - P01 (P10 P22 - P12 P20) + P02 (P10 P21 - P11 P20)
    add_term(P[0][0], 3, P[1][1], P[2][2], 3, 4, P[2][1], P[1][2], 1.0);
    add_term(P[0][1], 3, P[1][0], P[2][2], 3, 4, P[2][0], P[1][2], -1.0);
    add_term(P[0][2], 4, P[1][0], P[2][1], 3, 3, P[1][1], P[2][0], 1.0);
  }
  return true;
}

// second half, per root z of det B(z): back substitution through the null vector of B(z) (SVD::solveZ) to one
// essential matrix of unit Frobenius norm.  Returns false for complex roots and degenerate null vectors.
__device__ bool five_point_model(Cx root, const double (*B)[13], const double* EE, double* Ev_out) {
  if (fabs(root.im) > 1e-10) return false;
  const double z1 = root.re, z2 = z1 * z1, z3 = z2 * z1, z4 = z3 * z1;
  double Bz[9];
  for (int j = 0; j < 3; j++) {
    const double* br = B[j];
    Bz[3 * j + 0] = br[0] * z3 + br[1] * z2 + br[2] * z1 + br[3];
    Bz[3 * j + 1] = br[4] * z3 + br[5] * z2 + br[6] * z1 + br[7];
    Bz[3 * j + 2] = br[8] * z4 + br[9] * z3 + br[10] * z2 + br[11] * z1 + br[12];
  }
  double w3[3], vt3[9];
  jacobi_svd<3>(Bz, 3, 3, w3, nullptr, vt3);  // SVD::solveZ: right singular vector of the smallest singular value
  const double* xy1 = vt3 + 6;
  if (fabs(xy1[2]) < 1e-10) return false;
  const double xs = xy1[0] / xy1[2], ys = xy1[1] / xy1[2];
  double nrm = 0, Ev[9];
  for (int e = 0; e < 9; e++) {
    Ev[e] = EE[e] * xs + EE[9 + e] * ys + EE[18 + e] * z1 + EE[27 + e];
    nrm += Ev[e] * Ev[e];
  }
  nrm = 1. / sqrt(nrm);
  for (int e = 0; e < 9; e++) Ev_out[e] = Ev[e] * nrm;
  return true;
}

// EMEstimatorCallback::computeError for one normalised correspondence: f64 Sampson distance cast to f32
__device__ __forceinline__ float sampson_f32(const double* E, double x1, double y1, double x2, double y2) {
  const double Ex1[3] = {E[0] * x1 + E[1] * y1 + E[2], E[3] * x1 + E[4] * y1 + E[5], E[6] * x1 + E[7] * y1 + E[8]};
  const double Etx2[2] = {E[0] * x2 + E[3] * y2 + E[6], E[1] * x2 + E[4] * y2 + E[7]};
  const double x2tEx1 = x2 * Ex1[0] + y2 * Ex1[1] + Ex1[2];
  const double a = Ex1[0] * Ex1[0], b = Ex1[1] * Ex1[1], c = Etx2[0] * Etx2[0], d = Etx2[1] * Etx2[1];
  return (float)(x2tEx1 * x2tEx1 / (a + b + c + d));
}

__device__ inline double det3(const double* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}

// cv::decomposeEssentialMat: R1 = U W Vt, R2 = U W^T Vt, t = U[:, 2]
__device__ void decompose_essential(const double* E, double R1[9], double R2[9], double t[3]) {
  double w[3], U[9], Vt[9];
  jacobi_svd<3>(E, 3, 3, w, U, Vt);
  if (det3(U) < 0)
    for (int i = 0; i < 9; i++) U[i] = -U[i];
  if (det3(Vt) < 0)
    for (int i = 0; i < 9; i++) Vt[i] = -Vt[i];
  const double W[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}, Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
  double T[9];
  mat3_mul(U, W, T);
  mat3_mul(T, Vt, R1);
  mat3_mul(U, Wt, T);
  mat3_mul(T, Vt, R2);
  t[0] = U[2], t[1] = U[5], t[2] = U[8];
}

// cv::decomposeHomographyMat (HomographyDecompInria, Malis-Vargas closed form): 4 motions (R, t) -- or 1 (pure
// rotation).  Returns the count.
__device__ int decompose_homography(const double* H, const double K[4], double R_out[4][9], double t_out[4][3]) {
  const double Km[9] = {K[0], 0, K[2], 0, K[1], K[3], 0, 0, 1};
  const double Ki[9] = {1. / K[0], 0, -K[2] / K[0], 0, 1. / K[1], -K[3] / K[1], 0, 0, 1};
  double T[9], Hn[9];
  mat3_mul(Ki, H, T);
  mat3_mul(T, Km, Hn);
  double w[3];
  jacobi_svd<3>(Hn, 3, 3, w, nullptr, nullptr);
  for (int i = 0; i < 9; i++) Hn[i] /= w[1];
  double S[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double a = 0;
      for (int k = 0; k < 3; k++) a += Hn[k * 3 + i] * Hn[k * 3 + j];
      S[i * 3 + j] = a - (i == j ? 1.0 : 0.0);
    }
  double smax = 0;
  for (int i = 0; i < 9; i++) smax = fmax(smax, fabs(S[i]));
  if (smax < 0.001) {
    for (int i = 0; i < 9; i++) R_out[0][i] = Hn[i];
    t_out[0][0] = t_out[0][1] = t_out[0][2] = 0;
    return 1;
  }
  auto opp_minor = [&](int row, int col) {
    const int x1 = col == 0 ? 1 : 0, x2 = col == 2 ? 1 : 2, y1 = row == 0 ? 1 : 0, y2 = row == 2 ? 1 : 2;
    return S[y1 * 3 + x2] * S[y2 * 3 + x1] - S[y1 * 3 + x1] * S[y2 * 3 + x2];
  };
  auto signd = [](double x) { return x >= 0 ? 1.0 : -1.0; };
  const double M00 = opp_minor(0, 0), M11 = opp_minor(1, 1), M22 = opp_minor(2, 2);
  const double rtM00 = sqrt(M00), rtM11 = sqrt(M11), rtM22 = sqrt(M22);
  const double e12 = signd(opp_minor(1, 2)), e02 = signd(opp_minor(0, 2)), e01 = signd(opp_minor(0, 1));
  const double nS00 = fabs(S[0]), nS11 = fabs(S[4]), nS22 = fabs(S[8]);
  int indx = 0;
  if (nS00 < nS11) {
    indx = 1;
    if (nS11 < nS22) indx = 2;
  } else if (nS00 < nS22) {
    indx = 2;
  }
  double npa[3], npb[3];
  if (indx == 0) {
    npa[0] = S[0], npb[0] = S[0];
    npa[1] = S[1] + rtM22, npb[1] = S[1] - rtM22;
    npa[2] = S[2] + e12 * rtM11, npb[2] = S[2] - e12 * rtM11;
  } else if (indx == 1) {
    npa[0] = S[1] + rtM22, npb[0] = S[1] - rtM22;
    npa[1] = S[4], npb[1] = S[4];
    npa[2] = S[5] - e02 * rtM00, npb[2] = S[5] + e02 * rtM00;
  } else {
    npa[0] = S[2] + e01 * rtM11, npb[0] = S[2] - e01 * rtM11;
    npa[1] = S[5] + rtM00, npb[1] = S[5] - rtM00;
    npa[2] = S[8], npb[2] = S[8];
  }
  const double traceS = S[0] + S[4] + S[8];
  const double v = 2.0 * sqrt(1 + traceS - M00 - M11 - M22);
  const double ESii = signd(S[indx * 3 + indx]);
  const double r = sqrt(2 + traceS + v), n_t = sqrt(2 + traceS - v);
  const double na_n = 1. / sqrt(npa[0] * npa[0] + npa[1] * npa[1] + npa[2] * npa[2]);
  const double nb_n = 1. / sqrt(npb[0] * npb[0] + npb[1] * npb[1] + npb[2] * npb[2]);
  double na[3], nb[3], ta_s[3], tb_s[3];
  for (int i = 0; i < 3; i++) {
    na[i] = npa[i] * na_n;
    nb[i] = npb[i] * nb_n;
  }
  const double half_nt = 0.5 * n_t, esii_t_r = ESii * r;
  for (int i = 0; i < 3; i++) {
    ta_s[i] = half_nt * (esii_t_r * nb[i] - n_t * na[i]);
    tb_s[i] = half_nt * (esii_t_r * na[i] - n_t * nb[i]);
  }
  auto rmat = [&](const double* ts, const double* n, double* R) {
    double Mx[9];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Mx[i * 3 + j] = (i == j ? 1.0 : 0.0) - (2.0 / v) * ts[i] * n[j];
    mat3_mul(Hn, Mx, R);
    if (det3(R) < 0)
      for (int i = 0; i < 9; i++) R[i] = -R[i];
  };
  double Ra[9], Rb[9], ta[3], tb[3];
  rmat(ta_s, na, Ra);
  rmat(tb_s, nb, Rb);
  for (int i = 0; i < 3; i++) {
    ta[i] = Ra[3 * i] * ta_s[0] + Ra[3 * i + 1] * ta_s[1] + Ra[3 * i + 2] * ta_s[2];
    tb[i] = Rb[3 * i] * tb_s[0] + Rb[3 * i + 1] * tb_s[1] + Rb[3 * i + 2] * tb_s[2];
  }
  for (int i = 0; i < 9; i++) {
    R_out[0][i] = R_out[1][i] = Ra[i];
    R_out[2][i] = R_out[3][i] = Rb[i];
  }
  for (int i = 0; i < 3; i++) {
    t_out[0][i] = ta[i];
    t_out[1][i] = -ta[i];
    t_out[2][i] = tb[i];
    t_out[3][i] = -tb[i];
  }
  return 4;
}

}  // namespace uvo
