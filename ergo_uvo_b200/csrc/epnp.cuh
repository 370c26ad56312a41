// epnp.cuh -- EPnP (Lepetit, Moreno-Noguer, Fua) as OpenCV's calib3d/src/epnp.cpp runs it inside
// cv::solvePnP(..., SOLVEPNP_EPNP), split into point-independent pieces (small dense algebra on one thread) and
// per-point pieces, so that the same code serves the 5-point RANSAC kernel (one hypothesis per thread) and the
// block-cooperative refit on all inliers.  Replaces the solver behind cv::solvePnPRansac
// (reference visual_odometry.h:647-648).
#pragma once
#include "linalg.cuh"

namespace uvo {

struct EpnpCam {
  double fu, fv, uc, vc;
};

__device__ __forceinline__ double dot3(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
__device__ __forceinline__ double dist2_3(const double* a, const double* b) {
  return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}

// choose_control_points + the inverse used by compute_barycentric_coordinates.
// c0 = centroid, cov = sum (p-c0)(p-c0)^T over the n points.
__device__ inline void epnp_control_points(const double c0[3], const double cov[9], int n, double cws[4][3],
                                           double ci[9]) {
  double dc[3], Vt[9];
  jacobi_svd<3, 3, 3>(cov, 3, 3, dc, nullptr, Vt);
  for (int j = 0; j < 3; j++) cws[0][j] = c0[j];
  for (int i = 1; i < 4; i++) {
    const double k = sqrt(dc[i - 1] / n);
    for (int j = 0; j < 3; j++) cws[i][j] = cws[0][j] + k * Vt[(i - 1) * 3 + j];
  }
  double cc[9];
  for (int i = 0; i < 3; i++)
    for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
  svd_invert3(cc, ci);
}

__device__ __forceinline__ void epnp_alphas(const double p[3], const double cws[4][3], const double ci[9],
                                            double a[4]) {
  for (int j = 0; j < 3; j++)
    a[1 + j] = ci[3 * j] * (p[0] - cws[0][0]) + ci[3 * j + 1] * (p[1] - cws[0][1]) + ci[3 * j + 2] * (p[2] - cws[0][2]);
  a[0] = 1.0f - a[1] - a[2] - a[3];
}

__device__ __forceinline__ void epnp_m_rows(const double a[4], double u, double v, const EpnpCam& cam, double M1[12],
                                            double M2[12]) {
  for (int k = 0; k < 4; k++) {
    M1[3 * k] = a[k] * cam.fu;
    M1[3 * k + 1] = 0.0;
    M1[3 * k + 2] = a[k] * (cam.uc - u);
    M2[3 * k] = 0.0;
    M2[3 * k + 1] = a[k] * cam.fv;
    M2[3 * k + 2] = a[k] * (cam.vc - v);
  }
}

// Householder QR least squares, a transcription of epnp::qr_solve (6 x 4)
__device__ inline void epnp_qr_solve(double* A, int nr, int nc, double* b, double* X) {
  double A1[6], A2[6];
  double* ppAkk = A;
  for (int k = 0; k < nc; k++) {
    double* ppAik1 = ppAkk;
    double eta = fabs(*ppAik1);
    for (int i = k + 1; i < nr; i++) {
      const double elt = fabs(*ppAik1);
      if (eta < elt) eta = elt;
      ppAik1 += nc;
    }
    if (eta == 0) {
      A1[k] = A2[k] = 0.0;
      return;
    }
    double* ppAik2 = ppAkk;
    double sum2 = 0.0;
    const double inv_eta = 1. / eta;
    for (int i = k; i < nr; i++) {
      *ppAik2 *= inv_eta;
      sum2 += *ppAik2 * *ppAik2;
      ppAik2 += nc;
    }
    double sigma = sqrt(sum2);
    if (*ppAkk < 0) sigma = -sigma;
    *ppAkk += sigma;
    A1[k] = sigma * *ppAkk;
    A2[k] = -eta * sigma;
    for (int j = k + 1; j < nc; j++) {
      double* ppAik = ppAkk;
      double sum = 0;
      for (int i = k; i < nr; i++) {
        sum += *ppAik * ppAik[j - k];
        ppAik += nc;
      }
      const double tau = sum / A1[k];
      ppAik = ppAkk;
      for (int i = k; i < nr; i++) {
        ppAik[j - k] -= tau * *ppAik;
        ppAik += nc;
      }
    }
    ppAkk += nc + 1;
  }
  double* ppAjj = A;
  for (int j = 0; j < nc; j++) {
    double* ppAij = ppAjj;
    double tau = 0;
    for (int i = j; i < nr; i++) {
      tau += *ppAij * b[i];
      ppAij += nc;
    }
    tau /= A1[j];
    ppAij = ppAjj;
    for (int i = j; i < nr; i++) {
      b[i] -= tau * *ppAij;
      ppAij += nc;
    }
    ppAjj += nc + 1;
  }
  X[nc - 1] = b[nc - 1] / A2[nc - 1];
  for (int i = nc - 2; i >= 0; i--) {
    double* ppAij = A + i * nc + (i + 1);
    double sum = 0;
    for (int j = i + 1; j < nc; j++) {
      sum += *ppAij * X[j];
      ppAij++;
    }
    X[i] = (b[i] - sum) / A2[i];
  }
}

// The three beta candidates (which = 1, 2, 3: the N = 1, 2, 3 approximations) on one warp: candidate w is solved by
// lane w, with lanes w + 3 and w + 6 helping in the sweeps of its 6 x {4, 3, 5} least-squares SVD
// (jacobi_sweeps_rr); each is then refined by 5 Gauss-Newton steps on its own lane.  Every lane of the warp calls;
// lanes 0..2 receive the betas of candidates 1..3.
struct BetaSmem {
  double At[3][30], V[3][25], W[3][5];
};

__device__ inline void epnp_betas_warp(const double* l, const double* rho, BetaSmem& bs, int lane, double betas[4],
                                       long long* prof = nullptr) {
  const int cand = lane % 3, which = cand + 1;
  const int nc = which == 1 ? 4 : which == 2 ? 3 : 5;
  const int cols1[4] = {0, 1, 3, 6};
  double* At = bs.At[cand];
  double* V = bs.V[cand];
  double* W = bs.W[cand];
  if (lane < 3) {
    for (int j = 0; j < nc; j++) {
      double sd = 0;
      for (int i = 0; i < 6; i++) {
        const double v = l[10 * i + (which == 1 ? cols1[j] : j)];
        At[j * 6 + i] = v;
        sd += v * v;
      }
      W[j] = sd;
      for (int k = 0; k < nc; k++) V[j * nc + k] = (k == j) ? 1.0 : 0.0;
    }
  }
  __syncwarp();
  if (prof && lane == 2) prof[0] = clock64();
  jacobi_sweeps_rr(At, V, W, 6, nc, lane < 9 ? lane / 3 : -1);
  if (prof && lane == 2) prof[1] = clock64();
  if (lane >= 3) return;
  double x[5];
  {  // cvSolve(CV_SVD): x = V diag(1 / w) U^T rho over the singular values above the threshold
    double w[6], U[36], Vt[36];
    jacobi_svd_finish<6>(At, V, W, 6, nc, w, U, Vt);
    double thr = 0;
    for (int i = 0; i < nc; i++) thr += w[i];
    thr *= DBL_EPSILON * 2;
    for (int k = 0; k < nc; k++) x[k] = 0;
    for (int i = 0; i < nc; i++) {
      if (w[i] <= thr) continue;
      double sacc = 0;
      for (int k = 0; k < 6; k++) sacc += U[k * nc + i] * rho[k];
      sacc /= w[i];
      for (int k = 0; k < nc; k++) x[k] += sacc * Vt[i * nc + k];
    }
  }
  if (prof && lane == 2) prof[2] = clock64();
  if (which == 1) {
    if (x[0] < 0) {
      betas[0] = sqrt(-x[0]);
      betas[1] = -x[1] / betas[0];
      betas[2] = -x[2] / betas[0];
      betas[3] = -x[3] / betas[0];
    } else {
      betas[0] = sqrt(x[0]);
      betas[1] = x[1] / betas[0];
      betas[2] = x[2] / betas[0];
      betas[3] = x[3] / betas[0];
    }
  } else {
    if (x[0] < 0) {
      betas[0] = sqrt(-x[0]);
      betas[1] = (x[2] < 0) ? sqrt(-x[2]) : 0.0;
    } else {
      betas[0] = sqrt(x[0]);
      betas[1] = (x[2] > 0) ? sqrt(x[2]) : 0.0;
    }
    if (x[1] < 0) betas[0] = -betas[0];
    betas[2] = which == 3 ? x[3] / betas[0] : 0.0;
    betas[3] = 0.0;
  }
  // gauss_newton, 5 iterations
  for (int it = 0; it < 5; it++) {
    double GA[24], gb[6], gx[4] = {0, 0, 0, 0};
    for (int i = 0; i < 6; i++) {
      const double* rowL = l + i * 10;
      double* rowA = GA + i * 4;
      rowA[0] = 2 * rowL[0] * betas[0] + rowL[1] * betas[1] + rowL[3] * betas[2] + rowL[6] * betas[3];
      rowA[1] = rowL[1] * betas[0] + 2 * rowL[2] * betas[1] + rowL[4] * betas[2] + rowL[7] * betas[3];
      rowA[2] = rowL[3] * betas[0] + rowL[4] * betas[1] + 2 * rowL[5] * betas[2] + rowL[8] * betas[3];
      rowA[3] = rowL[6] * betas[0] + rowL[7] * betas[1] + rowL[8] * betas[2] + 2 * rowL[9] * betas[3];
      gb[i] = rho[i] - (rowL[0] * betas[0] * betas[0] + rowL[1] * betas[0] * betas[1] + rowL[2] * betas[1] * betas[1] +
                        rowL[3] * betas[0] * betas[2] + rowL[4] * betas[1] * betas[2] + rowL[5] * betas[2] * betas[2] +
                        rowL[6] * betas[0] * betas[3] + rowL[7] * betas[1] * betas[3] + rowL[8] * betas[2] * betas[3] +
                        rowL[9] * betas[3] * betas[3]);
    }
    epnp_qr_solve(GA, 6, 4, gb, gx);
    for (int i = 0; i < 4; i++) betas[i] += gx[i];
  }
}

__device__ inline void epnp_ccs(const double betas[4], const double* ut, double ccs[4][3]) {
  for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
  for (int i = 0; i < 4; i++) {
    const double* v = ut + 12 * (11 - i);
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 3; k++) ccs[j][k] += betas[i] * v[3 * j + k];
  }
}

__device__ __forceinline__ void epnp_pc(const double a[4], const double ccs[4][3], double sign, double pc[3]) {
  for (int j = 0; j < 3; j++) pc[j] = sign * (a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j]);
}

// estimate_R_and_t from the centroids and ABt = sum (pc-pc0)(pw-pw0)^T
__device__ inline void epnp_rt_from_abt(const double abt[9], const double pc0[3], const double pw0[3], double R[9],
                                        double t[3]) {
  double d[3], U[9], Vt[9];
  jacobi_svd<3, 3, 3>(abt, 3, 3, d, U, Vt);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i * 3 + j] = U[i * 3] * Vt[j] + U[i * 3 + 1] * Vt[3 + j] + U[i * 3 + 2] * Vt[6 + j];
  const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] -
                     R[1] * R[3] * R[8] - R[0] * R[5] * R[7];
  if (det < 0) {
    R[6] = -R[6];
    R[7] = -R[7];
    R[8] = -R[8];
  }
  t[0] = pc0[0] - dot3(R + 0, pw0);
  t[1] = pc0[1] - dot3(R + 3, pw0);
  t[2] = pc0[2] - dot3(R + 6, pw0);
}

__device__ __forceinline__ double epnp_reproj1(const double R[9], const double t[3], const double pw[3], double u,
                                               double v, const EpnpCam& cam) {
  const double Xc = dot3(R + 0, pw) + t[0], Yc = dot3(R + 3, pw) + t[1], inv_Zc = 1.0 / (dot3(R + 6, pw) + t[2]);
  const double ue = cam.uc + cam.fu * Xc * inv_Zc, ve = cam.vc + cam.fv * Yc * inv_Zc;
  return sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
}

// compute_R_and_t for one beta candidate on a small set (n <= 5): camera-frame points from the control points,
// sign fix, absolute orientation, mean reprojection error.  Returns the error; Rw / tw receive the pose.
__device__ inline double epnp_candidate(const double* pws, const double* us, int n, const double (*alphas)[4],
                                        const double betas[4], const double* ut, const EpnpCam& cam, double Rw[9],
                                        double tw[3]) {
  double ccs[4][3], pcs[5][3];
  epnp_ccs(betas, ut, ccs);
  for (int i = 0; i < n; i++) epnp_pc(alphas[i], ccs, 1.0, pcs[i]);
  if (pcs[0][2] < 0.0)  // solve_for_sign
    for (int i = 0; i < n; i++)
      for (int j = 0; j < 3; j++) pcs[i][j] = -pcs[i][j];
  double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 3; j++) {
      pc0[j] += pcs[i][j];
      pw0[j] += pws[3 * i + j];
    }
  for (int j = 0; j < 3; j++) {
    pc0[j] /= n;
    pw0[j] /= n;
  }
  double abt[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; i++) {
    const double* pw = pws + 3 * i;
    for (int j = 0; j < 3; j++) {
      abt[3 * j] += (pcs[i][j] - pc0[j]) * (pw[0] - pw0[0]);
      abt[3 * j + 1] += (pcs[i][j] - pc0[j]) * (pw[1] - pw0[1]);
      abt[3 * j + 2] += (pcs[i][j] - pc0[j]) * (pw[2] - pw0[2]);
    }
  }
  epnp_rt_from_abt(abt, pc0, pw0, Rw, tw);
  double sum2 = 0.0;
  for (int i = 0; i < n; i++) sum2 += epnp_reproj1(Rw, tw, pws + 3 * i, us[2 * i], us[2 * i + 1], cam);
  return sum2 / n;
}

// centroid, covariance, control points and barycentric coordinates of a small set (n <= 5)
__device__ inline void epnp_small_setup(const double* pws, int n, double cws[4][3], double ci[9], double (*alphas)[4]) {
  double c0[3] = {0, 0, 0};
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 3; j++) c0[j] += pws[3 * i + j];
  for (int j = 0; j < 3; j++) c0[j] /= n;
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; i++) {
    const double d[3] = {pws[3 * i] - c0[0], pws[3 * i + 1] - c0[1], pws[3 * i + 2] - c0[2]};
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) cov[a * 3 + b] += d[a] * d[b];
  }
  epnp_control_points(c0, cov, n, cws, ci);
  for (int i = 0; i < n; i++) epnp_alphas(pws + 3 * i, cws, ci, alphas[i]);
}

// row i (0..5) of compute_L_6x10: control-point pair (a, b) in the order (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
__device__ inline void epnp_L_row(const double* ut, int i, double* row) {
  const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
  const int a = pa[i], b = pb[i];
  double dv[4][3];
  for (int x = 0; x < 4; x++) {
    const double* v = ut + 12 * (11 - x);
    dv[x][0] = v[3 * a] - v[3 * b];
    dv[x][1] = v[3 * a + 1] - v[3 * b + 1];
    dv[x][2] = v[3 * a + 2] - v[3 * b + 2];
  }
  row[0] = dot3(dv[0], dv[0]);
  row[1] = 2.0f * dot3(dv[0], dv[1]);
  row[2] = dot3(dv[1], dv[1]);
  row[3] = 2.0f * dot3(dv[0], dv[2]);
  row[4] = 2.0f * dot3(dv[1], dv[2]);
  row[5] = dot3(dv[2], dv[2]);
  row[6] = 2.0f * dot3(dv[0], dv[3]);
  row[7] = 2.0f * dot3(dv[1], dv[3]);
  row[8] = 2.0f * dot3(dv[2], dv[3]);
  row[9] = dot3(dv[3], dv[3]);
}
__device__ inline double epnp_rho_entry(const double cws[4][3], int i) {
  const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
  return dist2_3(cws[pa[i]], cws[pb[i]]);
}

// element `a` (0..11) of the two rows of M a point contributes (epnp_m_rows, one element at a time)
__device__ __forceinline__ void epnp_m_elem(const double al[4], double u, double v, const EpnpCam& cam, int a,
                                            double& m1, double& m2) {
  const int k = a / 3, r = a - 3 * k;
  m1 = r == 0 ? al[k] * cam.fu : (r == 1 ? 0.0 : al[k] * (cam.uc - u));
  m2 = r == 0 ? 0.0 : (r == 1 ? al[k] * cam.fv : al[k] * (cam.vc - v));
}

}  // namespace uvo
