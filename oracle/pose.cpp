// oracle/pose.cpp -- CPU ORACLE (test infrastructure only; see uvo_oracle.h).
// Restates the pose-estimation inner loop of the stereo path and the shared geometry helpers:
//   cv::RNG + RANSAC subset stream + RANSACUpdateNumIters  (OpenCV calib3d/src/ptsetreg.cpp, core/src/rand.cpp;
//                                                           SURVEY B.1, B.9, C.6)
//   cv::solvePnPRansac(..., SOLVEPNP_EPNP)                 (reference visual_odometry.h:647-648; OpenCV
//                                                           calib3d/src/solvepnp.cpp, epnp.cpp)
//   cv::triangulatePoints                                  (visual_odometry.h:355, :631; calib3d/src/triangulate.cpp)
//   cv::projectPoints (no distortion), cv::Rodrigues       (VO_utility.cpp:636, visual_odometry.h:673)
//   extract_3Dpoints, reproject_errors, compute_mean_and_variance, convert_3Dpoints_camera, compute_scale_factor,
//   compute_median, select_estimation_method               (VO_utility.cpp:23-63, :188-237, :632-651, :725-748;
//                                                           math_utility.cpp:35-86)
// Pinning: RNG stream / stopping rule exact; triangulation, projection, Rodrigues, EPnP and the RANSAC result are
// validated against cv2 4.13 to tolerance (the wheel's SVD is LAPACK, so minimal-set models are not bit-equal;
// SURVEY 7.2-4).  tests/test_oracle_pose.py + tests/golden/pose_600.npz.
#include <cstring>

#include "linalg.h"
#include "uvo_oracle.h"

using namespace orc;

// ------------------------------------------------------------------------------------------------ RNG / RANSAC
namespace {
struct CvRng {
  uint64_t state;
  explicit CvRng(uint64_t s) : state(s ? s : 0xffffffffULL) {}
  unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

void get_subset(CvRng& rng, int count, int model_points, int* idx) {
  for (int i = 0; i < model_points; i++) {
    int v;
    for (;;) {
      v = rng.uniform(0, count);
      bool dup = false;
      for (int k = 0; k < i; k++) dup |= (idx[k] == v);
      if (!dup) break;
    }
    idx[i] = v;
  }
}
}  // namespace

extern "C" void orc_rng_subsets(int count, int model_points, int n_subsets, int32_t* out) {
  CvRng rng((uint64_t)-1);
  for (int s = 0; s < n_subsets; s++) get_subset(rng, count, model_points, out + (size_t)s * model_points);
}

extern "C" int orc_ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = std::max(p, 0.);
  p = std::min(p, 1.);
  ep = std::max(ep, 0.);
  ep = std::min(ep, 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::nearbyint(num / denom);
}

// ------------------------------------------------------------------------------------------------ Rodrigues
extern "C" void orc_rodrigues_vec2mat(const double r[3], double R[9]) {
  double theta = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (theta < DBL_EPSILON) {
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0);
    return;
  }
  double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c, it = 1. / theta;
  double x = r[0] * it, y = r[1] * it, z = r[2] * it;
  const double rrt[9] = {x * x, x * y, x * z, x * y, y * y, y * z, x * z, y * z, z * z};
  const double rx[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  for (int i = 0; i < 9; i++) R[i] = c * (i % 4 == 0 ? 1. : 0.) + c1 * rrt[i] + s * rx[i];
}

extern "C" void orc_rodrigues_mat2vec(const double Rin[9], double rv[3]) {
  double w[3], U[9], Vt[9], R[9];
  jacobi_svd(Rin, 3, 3, w, U, Vt);
  mat3_mul(U, Vt, R);
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1) * 0.5;
  c = c > 1. ? 1. : c < -1. ? -1. : c;
  double theta = std::acos(c);
  if (s < 1e-5) {
    if (c > 0) {
      rx = ry = rz = 0;
    } else {
      double t = (R[0] + 1) * 0.5;
      rx = std::sqrt(std::max(t, 0.));
      t = (R[4] + 1) * 0.5;
      ry = std::sqrt(std::max(t, 0.)) * (R[1] < 0 ? -1. : 1.);
      t = (R[8] + 1) * 0.5;
      rz = std::sqrt(std::max(t, 0.)) * (R[2] < 0 ? -1. : 1.);
      if (std::fabs(rx) < std::fabs(ry) && std::fabs(rx) < std::fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
      theta /= std::sqrt(rx * rx + ry * ry + rz * rz);
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

// ------------------------------------------------------------------------------------------------ projectPoints
// zero distortion: m = (x/z * fx + cx, y/z * fy + cy) with the reciprocal 1/z multiplied in (cvProjectPoints2)
static inline void project1(const double X[3], const double R[9], const double t[3], const double K[4], double m[2]) {
  double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
  z = z ? 1. / z : 1;
  x *= z;
  y *= z;
  m[0] = x * K[0] + K[2];
  m[1] = y * K[1] + K[3];
}

extern "C" void orc_project_points(const double* X, int n, const double R[9], const double t[3], const double K[4],
                                   double* out2) {
  for (int i = 0; i < n; i++) project1(X + 3 * i, R, t, K, out2 + 2 * i);
}

// ------------------------------------------------------------------------------------------------ EPnP
namespace {
struct Epnp {
  double fu, fv, uc, vc;
  int n;
  std::vector<double> pws, us, alphas, pcs;
  double cws[4][3], ccs[4][3];

  static double dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
  static double dist2(const double* a, const double* b) {
    return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
  }

  void choose_control_points() {
    cws[0][0] = cws[0][1] = cws[0][2] = 0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < 3; j++) cws[0][j] += pws[3 * i + j];
    for (int j = 0; j < 3; j++) cws[0][j] /= n;
    double m[9] = {};
    for (int i = 0; i < n; i++) {
      double d[3] = {pws[3 * i] - cws[0][0], pws[3 * i + 1] - cws[0][1], pws[3 * i + 2] - cws[0][2]};
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) m[a * 3 + b] += d[a] * d[b];
    }
    double dc[3], U[9], Vt[9];
    jacobi_svd(m, 3, 3, dc, U, Vt);  // symmetric PSD: rows of V^T (== U^T) are the principal axes
    for (int i = 1; i < 4; i++) {
      double k = std::sqrt(dc[i - 1] / n);
      for (int j = 0; j < 3; j++) cws[i][j] = cws[0][j] + k * Vt[(i - 1) * 3 + j];
    }
  }

  void compute_barycentric() {
    double cc[9], ci[9];
    for (int i = 0; i < 3; i++)
      for (int j = 1; j < 4; j++) cc[3 * i + j - 1] = cws[j][i] - cws[0][i];
    svd_invert3(cc, ci);
    alphas.resize(4 * (size_t)n);
    for (int i = 0; i < n; i++) {
      const double* pi = &pws[3 * i];
      double* a = &alphas[4 * i];
      for (int j = 0; j < 3; j++)
        a[1 + j] = ci[3 * j] * (pi[0] - cws[0][0]) + ci[3 * j + 1] * (pi[1] - cws[0][1]) +
                   ci[3 * j + 2] * (pi[2] - cws[0][2]);
      a[0] = 1.0f - a[1] - a[2] - a[3];
    }
  }

  void compute_ccs(const double* betas, const double* ut) {
    for (int i = 0; i < 4; i++) ccs[i][0] = ccs[i][1] = ccs[i][2] = 0.0;
    for (int i = 0; i < 4; i++) {
      const double* v = ut + 12 * (11 - i);
      for (int j = 0; j < 4; j++)
        for (int k = 0; k < 3; k++) ccs[j][k] += betas[i] * v[3 * j + k];
    }
  }
  void compute_pcs() {
    pcs.resize(3 * (size_t)n);
    for (int i = 0; i < n; i++) {
      const double* a = &alphas[4 * i];
      for (int j = 0; j < 3; j++)
        pcs[3 * i + j] = a[0] * ccs[0][j] + a[1] * ccs[1][j] + a[2] * ccs[2][j] + a[3] * ccs[3][j];
    }
  }
  void solve_for_sign() {
    if (pcs[2] < 0.0) {
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 3; j++) ccs[i][j] = -ccs[i][j];
      for (size_t i = 0; i < pcs.size(); i++) pcs[i] = -pcs[i];
    }
  }
  void estimate_R_and_t(double R[3][3], double t[3]) {
    double pc0[3] = {}, pw0[3] = {};
    for (int i = 0; i < n; i++)
      for (int j = 0; j < 3; j++) {
        pc0[j] += pcs[3 * i + j];
        pw0[j] += pws[3 * i + j];
      }
    for (int j = 0; j < 3; j++) {
      pc0[j] /= n;
      pw0[j] /= n;
    }
    double abt[9] = {};
    for (int i = 0; i < n; i++) {
      const double* pc = &pcs[3 * i];
      const double* pw = &pws[3 * i];
      for (int j = 0; j < 3; j++) {
        abt[3 * j] += (pc[j] - pc0[j]) * (pw[0] - pw0[0]);
        abt[3 * j + 1] += (pc[j] - pc0[j]) * (pw[1] - pw0[1]);
        abt[3 * j + 2] += (pc[j] - pc0[j]) * (pw[2] - pw0[2]);
      }
    }
    double d[3], U[9], Vt[9];
    jacobi_svd(abt, 3, 3, d, U, Vt);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) R[i][j] = U[i * 3] * Vt[j] + U[i * 3 + 1] * Vt[3 + j] + U[i * 3 + 2] * Vt[6 + j];
    const double det = R[0][0] * R[1][1] * R[2][2] + R[0][1] * R[1][2] * R[2][0] + R[0][2] * R[1][0] * R[2][1] -
                       R[0][2] * R[1][1] * R[2][0] - R[0][1] * R[1][0] * R[2][2] - R[0][0] * R[1][2] * R[2][1];
    if (det < 0) {
      R[2][0] = -R[2][0];
      R[2][1] = -R[2][1];
      R[2][2] = -R[2][2];
    }
    t[0] = pc0[0] - dot(R[0], pw0);
    t[1] = pc0[1] - dot(R[1], pw0);
    t[2] = pc0[2] - dot(R[2], pw0);
  }
  double reprojection_error(const double R[3][3], const double t[3]) {
    double sum2 = 0.0;
    for (int i = 0; i < n; i++) {
      const double* pw = &pws[3 * i];
      double Xc = dot(R[0], pw) + t[0], Yc = dot(R[1], pw) + t[1], inv_Zc = 1.0 / (dot(R[2], pw) + t[2]);
      double ue = uc + fu * Xc * inv_Zc, ve = vc + fv * Yc * inv_Zc;
      double u = us[2 * i], v = us[2 * i + 1];
      sum2 += std::sqrt((u - ue) * (u - ue) + (v - ve) * (v - ve));
    }
    return sum2 / n;
  }
  double compute_R_and_t(const double* ut, const double* betas, double R[3][3], double t[3]) {
    compute_ccs(betas, ut);
    compute_pcs();
    solve_for_sign();
    estimate_R_and_t(R, t);
    return reprojection_error(R, t);
  }
  static void compute_L_6x10(const double* ut, double* l) {
    const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
    double dv[4][6][3];
    for (int i = 0; i < 4; i++) {
      int a = 0, b = 1;
      for (int j = 0; j < 6; j++) {
        dv[i][j][0] = v[i][3 * a] - v[i][3 * b];
        dv[i][j][1] = v[i][3 * a + 1] - v[i][3 * b + 1];
        dv[i][j][2] = v[i][3 * a + 2] - v[i][3 * b + 2];
        b++;
        if (b > 3) {
          a++;
          b = a + 1;
        }
      }
    }
    for (int i = 0; i < 6; i++) {
      double* row = l + 10 * i;
      row[0] = dot(dv[0][i], dv[0][i]);
      row[1] = 2.0f * dot(dv[0][i], dv[1][i]);
      row[2] = dot(dv[1][i], dv[1][i]);
      row[3] = 2.0f * dot(dv[0][i], dv[2][i]);
      row[4] = 2.0f * dot(dv[1][i], dv[2][i]);
      row[5] = dot(dv[2][i], dv[2][i]);
      row[6] = 2.0f * dot(dv[0][i], dv[3][i]);
      row[7] = 2.0f * dot(dv[1][i], dv[3][i]);
      row[8] = 2.0f * dot(dv[2][i], dv[3][i]);
      row[9] = dot(dv[3][i], dv[3][i]);
    }
  }
  void compute_rho(double* rho) {
    rho[0] = dist2(cws[0], cws[1]);
    rho[1] = dist2(cws[0], cws[2]);
    rho[2] = dist2(cws[0], cws[3]);
    rho[3] = dist2(cws[1], cws[2]);
    rho[4] = dist2(cws[1], cws[3]);
    rho[5] = dist2(cws[2], cws[3]);
  }
  static void betas_approx(const double* l, const double* rho, int which, double* betas) {
    static const int cols1[4] = {0, 1, 3, 6}, cols2[3] = {0, 1, 2}, cols3[5] = {0, 1, 2, 3, 4};
    const int* cols = which == 1 ? cols1 : which == 2 ? cols2 : cols3;
    const int nc = which == 1 ? 4 : which == 2 ? 3 : 5;
    double A[6 * 5], x[5];
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < nc; j++) A[i * nc + j] = l[10 * i + cols[j]];
    svd_solve(A, 6, nc, rho, x, /*round_robin=*/true);
    if (which == 1) {
      if (x[0] < 0) {
        betas[0] = std::sqrt(-x[0]);
        betas[1] = -x[1] / betas[0];
        betas[2] = -x[2] / betas[0];
        betas[3] = -x[3] / betas[0];
      } else {
        betas[0] = std::sqrt(x[0]);
        betas[1] = x[1] / betas[0];
        betas[2] = x[2] / betas[0];
        betas[3] = x[3] / betas[0];
      }
    } else {
      if (x[0] < 0) {
        betas[0] = std::sqrt(-x[0]);
        betas[1] = (x[2] < 0) ? std::sqrt(-x[2]) : 0.0;
      } else {
        betas[0] = std::sqrt(x[0]);
        betas[1] = (x[2] > 0) ? std::sqrt(x[2]) : 0.0;
      }
      if (x[1] < 0) betas[0] = -betas[0];
      betas[2] = which == 3 ? x[3] / betas[0] : 0.0;
      betas[3] = 0.0;
    }
  }
  // Householder QR least squares exactly as epnp::qr_solve (including its row-range quirk when scanning for eta)
  static void qr_solve(double* A, int nr, int nc, double* b, double* X) {
    double A1[6], A2[6];
    double* ppAkk = A;
    for (int k = 0; k < nc; k++) {
      double* ppAik1 = ppAkk;
      double eta = std::fabs(*ppAik1);
      for (int i = k + 1; i < nr; i++) {
        double elt = std::fabs(*ppAik1);
        if (eta < elt) eta = elt;
        ppAik1 += nc;
      }
      if (eta == 0) {
        A1[k] = A2[k] = 0.0;
        return;
      }
      double* ppAik2 = ppAkk;
      double sum2 = 0.0, inv_eta = 1. / eta;
      for (int i = k; i < nr; i++) {
        *ppAik2 *= inv_eta;
        sum2 += *ppAik2 * *ppAik2;
        ppAik2 += nc;
      }
      double sigma = std::sqrt(sum2);
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
        double tau = sum / A1[k];
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
  static void gauss_newton(const double* l, const double* rho, double betas[4]) {
    for (int it = 0; it < 5; it++) {
      double A[24], b[6], x[4] = {0, 0, 0, 0};
      for (int i = 0; i < 6; i++) {
        const double* rowL = l + i * 10;
        double* rowA = A + i * 4;
        rowA[0] = 2 * rowL[0] * betas[0] + rowL[1] * betas[1] + rowL[3] * betas[2] + rowL[6] * betas[3];
        rowA[1] = rowL[1] * betas[0] + 2 * rowL[2] * betas[1] + rowL[4] * betas[2] + rowL[7] * betas[3];
        rowA[2] = rowL[3] * betas[0] + rowL[4] * betas[1] + 2 * rowL[5] * betas[2] + rowL[8] * betas[3];
        rowA[3] = rowL[6] * betas[0] + rowL[7] * betas[1] + rowL[8] * betas[2] + 2 * rowL[9] * betas[3];
        b[i] = rho[i] - (rowL[0] * betas[0] * betas[0] + rowL[1] * betas[0] * betas[1] + rowL[2] * betas[1] * betas[1] +
                         rowL[3] * betas[0] * betas[2] + rowL[4] * betas[1] * betas[2] + rowL[5] * betas[2] * betas[2] +
                         rowL[6] * betas[0] * betas[3] + rowL[7] * betas[1] * betas[3] + rowL[8] * betas[2] * betas[3] +
                         rowL[9] * betas[3] * betas[3]);
      }
      qr_solve(A, 6, 4, b, x);
      for (int i = 0; i < 4; i++) betas[i] += x[i];
    }
  }

  void compute_pose(double Rout[9], double tout[3]) {
    choose_control_points();
    compute_barycentric();
    // MtM = M^T M, M is 2n x 12
    double mtm[144] = {};
    for (int i = 0; i < n; i++) {
      const double* as = &alphas[4 * i];
      double M1[12], M2[12];
      for (int k = 0; k < 4; k++) {
        M1[3 * k] = as[k] * fu;
        M1[3 * k + 1] = 0.0;
        M1[3 * k + 2] = as[k] * (uc - us[2 * i]);
        M2[3 * k] = 0.0;
        M2[3 * k + 1] = as[k] * fv;
        M2[3 * k + 2] = as[k] * (vc - us[2 * i + 1]);
      }
      for (int a = 0; a < 12; a++)
        for (int b = 0; b < 12; b++) mtm[a * 12 + b] += M1[a] * M1[b] + M2[a] * M2[b];
    }
    // CV_SVD_U_T of the symmetric PSD M^T M: rows = eigenvectors by descending eigenvalue (linalg.h: jacobi_eigh; the
    // null space -- rank(M) = 10 for 5 points -- comes out as a proper orthonormal basis)
    double d[12], ut[144];
    jacobi_eigh<12>(mtm, d, ut);
    double l[60], rho[6];
    compute_L_6x10(ut, l);
    compute_rho(rho);
    double Betas[4][4], rep[4], Rs[4][3][3], ts[4][3];
    for (int w = 1; w <= 3; w++) {
      betas_approx(l, rho, w, Betas[w]);
      gauss_newton(l, rho, Betas[w]);
      rep[w] = compute_R_and_t(ut, Betas[w], Rs[w], ts[w]);
    }
    int N = 1;
    if (rep[2] < rep[1]) N = 2;
    if (rep[3] < rep[N]) N = 3;
    for (int i = 0; i < 3; i++) {
      tout[i] = ts[N][i];
      for (int j = 0; j < 3; j++) Rout[i * 3 + j] = Rs[N][i][j];
    }
  }
};

// solvePnP(..., SOLVEPNP_EPNP) on n points.  `minimal` mirrors the f32 round trip of undistortPoints inside the
// RANSAC kernel (points arrive as CV_32F there); the final refit runs on f64 copies.
void solve_pnp_epnp(const double* X, const double* x, int n, const double K[4], bool f32_normalised, double rvec[3],
                    double tvec[3], double Rout[9]) {
  Epnp e;
  e.fu = K[0];
  e.fv = K[1];
  e.uc = K[2];
  e.vc = K[3];
  e.n = n;
  e.pws.assign(X, X + 3 * (size_t)n);
  e.us.resize(2 * (size_t)n);
  const double ifx = 1. / K[0], ify = 1. / K[1];
  for (int i = 0; i < n; i++) {
    double xn = (x[2 * i] - K[2]) * ifx, yn = (x[2 * i + 1] - K[3]) * ify;  // undistortPoints, zero distortion
    if (f32_normalised) {
      xn = (double)(float)xn;
      yn = (double)(float)yn;
    }
    e.us[2 * i] = xn * K[0] + K[2];
    e.us[2 * i + 1] = yn * K[1] + K[3];
  }
  double R[9];
  e.compute_pose(R, tvec);
  orc_rodrigues_mat2vec(R, rvec);
  if (Rout) std::memcpy(Rout, R, sizeof(R));
}
}  // namespace

extern "C" void orc_epnp(const double* X, const double* x, int n, const double K[4], double R[9], double t[3]) {
  double rvec[3];
  solve_pnp_epnp(X, x, n, K, false, rvec, t, R);
}

// ------------------------------------------------------------------------------------------------ solvePnPRansac
extern "C" int orc_solve_pnp_ransac_epnp(const double* Xd, const float* x, int n, const double K[4], int iterations,
                                         float reproj_err, double confidence, double rvec[3], double tvec[3],
                                         int32_t* inliers, int* hyps_evaluated) {
  if (hyps_evaluated) *hyps_evaluated = 0;
  const int model_points = 5;
  if (n < model_points) return 0;  // (npoints == 4 switches OpenCV to P3P: outside the reference's configuration)
  // object points are down-cast to f32 (solvepnp.cpp: opoints0.convertTo(opoints, CV_32F))
  std::vector<float> Xf(3 * (size_t)n);
  for (size_t i = 0; i < Xf.size(); i++) Xf[i] = (float)Xd[i];
  std::vector<uint8_t> mask(n), best_mask(n, 0);
  double best_r[3] = {0, 0, 0}, best_t[3] = {0, 0, 0};
  int max_good = 0;
  if (n == model_points) {
    // RANSACPointSetRegistrator::run with count == modelPoints: single kernel call, every point an inlier
    double Xs[15], xs[10];
    for (int k = 0; k < 5; k++) {
      for (int c = 0; c < 3; c++) Xs[3 * k + c] = Xf[3 * k + c];
      xs[2 * k] = x[2 * k];
      xs[2 * k + 1] = x[2 * k + 1];
    }
    // solvePnPRansac: `if (model_points == npoints)` -> plain solvePnP on the f32 data, every point an inlier,
    // no refit
    solve_pnp_epnp(Xs, xs, 5, K, true, rvec, tvec, nullptr);
    for (int i = 0; i < n; i++) inliers[i] = i;
    if (hyps_evaluated) *hyps_evaluated = 1;
    return n;
  } else {
    CvRng rng((uint64_t)-1);
    int niters = std::max(iterations, 1);
    const float thr = (float)((double)reproj_err * (double)reproj_err);
    int iter = 0;
    for (; iter < niters; iter++) {
      int idx[5];
      get_subset(rng, n, model_points, idx);
      double Xs[15], xs[10], r[3], t[3];
      for (int k = 0; k < 5; k++) {
        for (int c = 0; c < 3; c++) Xs[3 * k + c] = Xf[3 * idx[k] + c];
        xs[2 * k] = x[2 * idx[k]];
        xs[2 * k + 1] = x[2 * idx[k] + 1];
      }
      solve_pnp_epnp(Xs, xs, 5, K, true, r, t, nullptr);
      // computeError: projectPoints(opoints f32, rvec, tvec) -> f32 points; err = |ip - pp|^2 in f32
      double R[9];
      orc_rodrigues_vec2mat(r, R);
      int good = 0;
      for (int i = 0; i < n; i++) {
        double P[3] = {Xf[3 * i], Xf[3 * i + 1], Xf[3 * i + 2]}, m[2];
        project1(P, R, t, K, m);
        float dx = x[2 * i] - (float)m[0], dy = x[2 * i + 1] - (float)m[1];
        float e = dx * dx + dy * dy;
        int f = e <= thr;
        mask[i] = (uint8_t)f;
        good += f;
      }
      if (good > std::max(max_good, model_points - 1)) {
        std::swap(mask, best_mask);
        std::memcpy(best_r, r, sizeof(r));
        std::memcpy(best_t, t, sizeof(t));
        max_good = good;
        niters = orc_ransac_update_num_iters(confidence, (double)(n - good) / n, model_points, niters);
      }
    }
    if (hyps_evaluated) *hyps_evaluated = iter;
  }
  if (max_good <= 0) {
    std::memcpy(rvec, best_r, sizeof(best_r));
    std::memcpy(tvec, best_t, sizeof(best_t));
    return 0;
  }
  // final refit on all inliers, f64 copies of the f32 data
  std::vector<double> Xi, xi;
  int m = 0;
  for (int i = 0; i < n; i++)
    if (best_mask[i]) {
      for (int c = 0; c < 3; c++) Xi.push_back((double)Xf[3 * i + c]);
      xi.push_back((double)x[2 * i]);
      xi.push_back((double)x[2 * i + 1]);
      inliers[m++] = i;
    }
  solve_pnp_epnp(Xi.data(), xi.data(), m, K, false, rvec, tvec, nullptr);
  return m;
}

// ------------------------------------------------------------------------------------------------ triangulatePoints
extern "C" void orc_triangulate_points(const double P1[12], const double P2[12], const float* pts1, const float* pts2,
                                       int n, float* out) {
  for (int i = 0; i < n; i++) {
    double A[16];
    const double* P[2] = {P1, P2};
    const float* p[2] = {pts1 + 2 * i, pts2 + 2 * i};
    for (int j = 0; j < 2; j++) {
      double x = p[j][0], y = p[j][1];
      for (int k = 0; k < 4; k++) {
        A[(j * 2 + 0) * 4 + k] = x * P[j][8 + k] - P[j][k];
        A[(j * 2 + 1) * 4 + k] = y * P[j][8 + k] - P[j][4 + k];
      }
    }
    double w[4], U[16], Vt[16];
    jacobi_svd(A, 4, 4, w, U, Vt);
    for (int k = 0; k < 4; k++) out[(size_t)k * n + i] = (float)Vt[12 + k];
  }
}

// ------------------------------------------------------------------------------------------------ helpers (K12)
extern "C" double orc_compute_median(const double* v, int n) {  // math_utility.cpp:65-86
  if (n == 0) return 0.0;
  std::vector<double> s(v, v + n);
  std::sort(s.begin(), s.end());
  if (n % 2 == 0) return (s[n / 2 - 1] + s[n / 2]) / 2.0;
  return s[n / 2];
}

extern "C" int orc_select_estimation_method(const float* p1, const float* p2, int n, int distance) {
  std::vector<double> d(n);  // VO_utility.cpp:725-748
  for (int i = 0; i < n; i++) {
    double dx = p1[2 * i] - p2[2 * i], dy = p1[2 * i + 1] - p2[2 * i + 1];
    d[i] = std::sqrt(dx * dx + dy * dy);
  }
  return orc_compute_median(d.data(), n) < distance ? 0 : 1;
}

extern "C" int orc_extract_3dpoints(const float* kp1, const float* kp2, int n, const double R1[9], const double t1[3],
                                    const double R2[9], const double t2[3], const double K1[4], const double K2[4],
                                    const float* p4, double tol, int min3d, double* out_pts, int32_t* out_idx) {
  // convertPointsFromHomogeneous(points4D.t()) in f32 (scale = 1/w, or 1 when w == 0), then convertTo(CV_64F)
  std::vector<double> cam1(3 * (size_t)n);
  for (int i = 0; i < n; i++) {
    float W = p4[(size_t)3 * n + i];
    float scale = W != 0.f ? 1.f / W : 1.f;
    for (int c = 0; c < 3; c++) cam1[3 * i + c] = (double)(p4[(size_t)c * n + i] * scale);
  }
  std::vector<int> good_idx;
  std::vector<double> good;
  if (n >= min3d) {
    for (int i = 0; i < n; i++) {
      double m1[2], m2[2];
      project1(&cam1[3 * i], R1, t1, K1, m1);
      project1(&cam1[3 * i], R2, t2, K2, m2);
      double dx = kp1[2 * i] - m1[0], dy = kp1[2 * i + 1] - m1[1];
      double e1 = std::sqrt(dx * dx + dy * dy);
      dx = kp2[2 * i] - m2[0];
      dy = kp2[2 * i + 1] - m2[1];
      double e2 = std::sqrt(dx * dx + dy * dy);
      double mean = (e1 + e2) / 2.0;
      if (mean < tol && cam1[3 * i + 2] > 0) {
        good_idx.push_back(i);
        for (int c = 0; c < 3; c++) good.push_back(cam1[3 * i + c]);
      }
    }
  }
  int m = 0;
  const int ng = (int)good_idx.size();
  if (ng >= min3d && ng > 0) {
    double sum = 0, sq = 0;  // compute_mean_and_variance, math_utility.cpp:35-56 (population variance)
    for (int i = 0; i < ng; i++) {
      double z = good[3 * i + 2];
      sum += z;
      sq += z * z;
    }
    double mean = sum / ng, var = (sq / ng) - (mean * mean);
    double sd = std::sqrt(var);  // NaN if var < 0 => every comparison false (App. D-5)
    for (int i = 0; i < ng; i++) {
      double z = good[3 * i + 2];
      if (z <= mean + 3.0 * sd && z >= mean - 3.0 * sd) {
        out_idx[m] = good_idx[i];
        for (int c = 0; c < 3; c++) out_pts[3 * m + c] = good[3 * i + c];
        m++;
      }
    }
  }
  return m;
}

// convert_3Dpoints_camera (keeps the UN-transformed row when the transformed z > 0, App. D-3) + compute_scale_factor
// convert_3Dpoints_camera + compute_scale_factor; *n_front (nullable) = columns of good_currCam_points, the set whose
// emptiness -- not the value of SF -- decides whether the node assigns SF (visual_odometry.h:365-374)
extern "C" double orc_scale_factor_front(const double* pts, int n, const double R[9], const double t[3], float range,
                                         int* n_front) {
  std::vector<double> z;
  for (int i = 0; i < n; i++) {
    const double* p = pts + 3 * i;
    double zc = R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + t[2];
    if (zc > 0) z.push_back(p[2]);
  }
  if (n_front) *n_front = (int)z.size();
  if (z.empty()) return 0.0;  // caller keeps the previous SF (visual_odometry.h:366-374)
  return (double)range / orc_compute_median(z.data(), (int)z.size());
}
extern "C" double orc_scale_factor(const double* pts, int n, const double R[9], const double t[3], float range) {
  return orc_scale_factor_front(pts, n, R, t, range, nullptr);
}
