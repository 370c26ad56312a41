// linalg.cuh -- small dense fp64 linear algebra used by the pose kernels (one problem per thread).
// One-sided (Hestenes) Jacobi SVD with the sweep order, rotation formulas and stopping rule of OpenCV's
// JacobiSVDImpl_; everything is plain IEEE fp64 without FMA contraction (-fmad=false).
#pragma once
#include <cfloat>
#include <cuda_runtime.h>

namespace uvo {

// The part of jacobi_svd after the sweeps: singular values from the column norms, descending selection sort, U = A V
// / w with the completion rule for (numerically) zero singular values.  At (n rows of length m), V (n x n), W (n) are
// the sweep state and may live in registers, local or shared memory.
template <int MAXD>
__device__ __forceinline__ void jacobi_svd_finish(double* At, double* V, double* W, int m, int n, double* w, double* U, double* Vt) {
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

// A: m x n row-major, m,n <= MAXD.  A = U diag(w) Vt, w descending.  U: m x n (may be null), Vt: n x n.
// CM / CN: compile-time copies of m / n (0 = run-time).  With both known the k-loops unroll, so the loads of a pair's two
// rows are issued back to back instead of one per trip of a serial loop; the arithmetic and its order are unchanged.
// NULL_FLOOR: additionally leave a pair alone when |p| <= 16 eps^2 max|column|^2, i.e. when both columns are numerically
// null.  A rank-deficient input (the 5 x 9 epipolar system: 4 null columns) otherwise keeps rotating round-off among
// its null columns for all 30 sweeps -- the relative test never fires there -- without changing the null space they
// span.  Off (the OpenCV rule alone) wherever the oracle restates the same SVD bit for bit.
template <int MAXD, int CM = 0, int CN = 0, bool NULL_FLOOR = false>
__device__ void jacobi_svd(const double* A, int m_rt, int n_rt, double* w, double* U, double* Vt) {
  const int m = CM ? CM : m_rt, n = CN ? CN : n_rt;
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
  double null_floor = 0;
  if (NULL_FLOOR) {
    for (int i = 0; i < n; i++) null_floor = fmax(null_floor, W[i]);
    null_floor *= 16 * DBL_EPSILON * DBL_EPSILON;
  }
  const int max_iter = m > 30 ? m : 30;
  for (int iter = 0; iter < max_iter; iter++) {
    bool changed = false;
    for (int i = 0; i < n - 1; i++)
      for (int j = i + 1; j < n; j++) {
        double* Ai = At + i * m;
        double* Aj = At + j * m;
        double a = W[i], p = 0, b = W[j];
#pragma unroll
        for (int k = 0; k < m; k++) p += Ai[k] * Aj[k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        if (NULL_FLOOR && fabs(p) <= null_floor) continue;
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
#pragma unroll
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
#pragma unroll
        for (int k = 0; k < n; k++) {
          const double t0 = c * Vi[k] + s * Vj[k];
          const double t1 = -s * Vi[k] + c * Vj[k];
          Vi[k] = t0;
          Vj[k] = t1;
        }
      }
    if (!changed) break;
  }
  jacobi_svd_finish<MAXD>(At, V, W, m, n, w, U, Vt);
}

// The sweeps of jacobi_svd in ROUND-ROBIN pair order (oracle/linalg.h: jacobi_svd(..., round_robin = true)) for up to
// three small systems at once on one warp.  A lane works on the system its At / V / W pointers name; `slot` (0..2, or
// -1 for a lane that only keeps the warp's barriers company) is the pair of the round the lane rotates.  The pairs of a
// round touch disjoint columns, so rotating them side by side gives the bits of rotating them one after the other; a
// sweep costs NP - 1 rotation latencies instead of n (n - 1) / 2.  A system that has converged sees no rotation pass
// its test in the extra sweeps the slowest system of the warp still needs, so its state does not change.
__device__ inline void jacobi_sweeps_rr(double* At, double* V, double* W, int m, int n, int slot) {
  const double eps = DBL_EPSILON * 10;
  const int NP = (n + 1) & ~1;
  const int max_iter = m > 30 ? m : 30;
  for (int iter = 0; iter < max_iter; iter++) {
    bool changed = false;
    for (int r = 0; r < 5; r++) {  // NP <= 6
      if (slot >= 0 && r < NP - 1 && slot < NP / 2) {
        const int ia = slot == 0 ? NP - 1 : (r + slot) % (NP - 1), ib = slot == 0 ? r : (r - slot + NP - 1) % (NP - 1);
        const int i = min(ia, ib), j = max(ia, ib);
        if (j < n) {
          // OpenCV's rotation with the two branches of its c / s formulas folded into one expression (the factors 0.5
          // and 2 are exact, so (g - b) 0.5 / g for b < 0 and (g + b) / (2 g) for b >= 0 are both (g + |b|) / (2 g)),
          // gamma = sqrt(p^2 + beta^2) written out, and the convergence test evaluated beside the rotation instead of
          // in front of it: the dependent chain is dot product, sqrt, divide, sqrt, divide.  The pair's two columns
          // stay in registers.  (m <= 6)
          double* Ai = At + i * m;
          double* Aj = At + j * m;
          double ai[6], aj[6];
#pragma unroll
          for (int k = 0; k < 6; k++) {
            ai[k] = k < m ? Ai[k] : 0.0;
            aj[k] = k < m ? Aj[k] : 0.0;
          }
          double a = W[i], b = W[j], p = 0;
#pragma unroll
          for (int k = 0; k < 6; k++)
            if (k < m) p += ai[k] * aj[k];
          const bool rot = !(fabs(p) <= eps * sqrt(a * b));
          p *= 2;
          const double beta = a - b, gamma = sqrt(p * p + beta * beta);
          const double r1 = sqrt((gamma + fabs(beta)) / (gamma * 2)), r2 = p / (gamma * r1 * 2);
          const double c = beta < 0 ? r2 : r1, s = beta < 0 ? r1 : r2;
          if (rot) {
            a = b = 0;
#pragma unroll
            for (int k = 0; k < 6; k++)
              if (k < m) {
                const double t0 = c * ai[k] + s * aj[k];
                const double t1 = -s * ai[k] + c * aj[k];
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
        }
      }
      __syncwarp();
    }
    if (!__any_sync(0xffffffffu, changed)) break;
  }
}

// Eigen-decomposition of a symmetric N x N matrix (N even) by the two-sided Jacobi method in round-robin order:
// A = Vt^T diag(w) Vt, w descending, rows of Vt = eigenvectors.  Used for EPnP's 12 x 12 M^T M, whose 2-dimensional
// null space (5-point minimal sets) keeps a Hestenes SVD with a relative stopping rule rotating noise for all 30
// sweeps; with the absolute threshold eps * trace this converges in 6-8 sweeps.  The cv2 wheel's LAPACK produces yet
// another basis of that null space, so no choice is bit-comparable with OpenCV; this routine and the
// CPU oracle (oracle/linalg.h: jacobi_eigh, which documents the order) run the same sequence of IEEE operations (no
// FMA, sqrt and divide correctly rounded on both) and agree bit for bit.
//
// A sweep is N-1 rounds of N/2 index-disjoint rotations whose angles are all taken from the matrix at the start of the
// round: lanes p < partner(p) compute the N/2 rotations side by side (the dependent chain of a round is ONE rotation's
// two square roots and reciprocal, not N/2 of them -- the cyclic order this replaces paid 66 chains per sweep, this
// pays 11), then every thread evaluates its share of the N(N+1)/2 upper-triangle elements of J^T S J and of the N^2
// elements of J^T Vr.
template <int N>
__device__ __forceinline__ int rr_partner(int x, int r) {
  if (x == N - 1) return r;
  if (x == r) return N - 1;
  int v = 2 * r - x;
  v += (v < 0) ? (N - 1) : 0;
  v -= (v >= N - 1) ? (N - 1) : 0;
  return v;
}

// partners of index x over the N-1 rounds of a sweep, 4 bits per round (N <= 16)
template <int N>
__device__ __forceinline__ unsigned long long rr_partners_packed(int x) {
  unsigned long long v = 0;
  for (int r = 0; r < N - 1; r++) v |= (unsigned long long)rr_partner<N>(x, r) << (4 * r);
  return v;
}

// NT = 32: one warp per matrix (tid = lane, __syncwarp); NT = a whole block of 64 / 128 / 256 threads per matrix
// (tid = threadIdx.x, __syncthreads): the element updates of a round then spread over every warp of the block, which
// matters because a single warp runs this code at one dependent instruction per ~8 cycles.  The result does not
// depend on NT.  Shared memory: S and Vr hold TWO N x N buffers each (the round reads one and writes the other; A
// arrives in the first buffer of S), rot 2 N doubles, rflag 2 words.
template <int N, int NT>
__device__ void jacobi_eigh_rr(double* S, double* Vr, double* w, double* Vt, double* rot, unsigned* rflag, int tid,
                               long long* prof = nullptr) {
  static_assert(N % 2 == 0 && N <= 16, "round-robin order: even dimension, partners packed 4 bits per round");
  constexpr int NV = N * N;
  constexpr int NE = N * (N + 1) / 2;          // upper triangle incl. diagonal
  constexpr int EPT = (NE + NT - 1) / NT;      // elements of S per thread
  constexpr int VPT = (NV + NT - 1) / NT;      // elements of Vr per thread
  auto sync = [] {
    if (NT == 32) __syncwarp();
    else __syncthreads();
  };
  long long t_rot = 0, t_apply = 0, t0 = 0, n_rounds = 0;
  double* cc = rot;
  double* ss = rot + N;
  double* Sc = S;
  double* Sn = S + NV;
  double* Vc = Vr;
  double* Vn = Vr + NV;
  int ei[EPT], ej[EPT];
  unsigned long long pi[EPT], pj[EPT];
#pragma unroll
  for (int m = 0; m < EPT; m++) {
    // elements 0 .. NE-N-1: the strict upper triangle, row-major; the last N: the diagonal (kept together so that the
    // three formulas below diverge in one warp only)
    int e = tid + NT * m, i = -1, j = -1;
    if (e < NE - N) {
      i = 0;
      while (e >= N - 1 - i) {
        e -= N - 1 - i;
        i++;
      }
      j = i + 1 + e;
    } else if (e < NE) {
      i = j = e - (NE - N);
    }
    ei[m] = i;
    ej[m] = j;
    pi[m] = i < 0 ? 0ull : rr_partners_packed<N>(i);
    pj[m] = i < 0 ? 0ull : rr_partners_packed<N>(j);
  }
  unsigned long long pv[VPT];
#pragma unroll
  for (int m = 0; m < VPT; m++) {
    const int v = (NT - 1 - tid) + NT * m;
    pv[m] = v < NV ? rr_partners_packed<N>(v / N) : 0ull;
  }
  const unsigned long long prot = tid < N ? rr_partners_packed<N>(tid) : 0ull;
  for (int v = tid; v < NV; v += NT) Vc[v] = (v / N == v % N) ? 1.0 : 0.0;
  double tr = 0;
  for (int i = 0; i < N; i++) tr += fabs(Sc[i * N + i]);
  const double thr = tr * DBL_EPSILON;
  sync();
  int slot = 0;  // rflag is double-buffered over ALL rounds (a sweep has an odd number of them: r & 1 would reuse a
                 // slot across the sweep boundary while a slow thread still reads it)
  for (int sweep = 0; sweep < 30; sweep++) {
    bool changed = false;
    for (int r = 0; r < N - 1; r++) {
      slot ^= 1;
      if (prof) t0 = clock64();
      if (tid < 32) {
        bool rt = false;
        const int p = tid, q = (int)(prot >> (4 * r)) & 15;
        if (tid < N && p < q) {
          // c = u / h, s = sgn(d) x / h with d = aqq - app, x = 2 apq, u = |d| + sqrt(d^2 + x^2), h^2 = x^2 + u^2:
          // two square roots and one reciprocal on the dependent chain.  Evaluated unconditionally (apq = 0 gives
          // h = u = 2 |d| or, for d = 0 too, 0 / 0, which the select below discards) so that the code has no branch.
          const double apq = Sc[p * N + q], app = Sc[p * N + p], aqq = Sc[q * N + q];
          rt = fabs(apq) > thr;
          const double d = aqq - app, x = 2 * apq;
          const double rr = sqrt(d * d + x * x);
          const double u = fabs(d) + rr, xs = d >= 0 ? x : -x;
          const double ih = 1 / sqrt(x * x + u * u);
          const double c = rt ? u * ih : 1.0, s = rt ? xs * ih : 0.0;
          cc[p] = cc[q] = c;
          ss[p] = ss[q] = s;
        }
        const unsigned m = __ballot_sync(0xffffffffu, rt);  // bit p: the pair whose smaller index is p rotates
        if (tid == 0) rflag[slot] = m;
      }
      sync();
      const unsigned rmask = rflag[slot];
      if (!rmask) continue;  // uniform over the NT threads
      changed = true;
      if (prof) {
        const long long t1 = clock64();
        t_rot += t1 - t0;
        t0 = t1;
        n_rounds++;
      }
#pragma unroll
      for (int m = 0; m < EPT; m++) {
        const int i = ei[m], j = ej[m];
        if (i < 0) continue;
        const int i2 = (int)(pi[m] >> (4 * r)) & 15;
        const bool lowi = i < i2, roti = (rmask >> (lowi ? i : i2)) & 1u;
        const double ci = cc[i], si = ss[i];
        double v;
        if (i == j) {
          const double aii = Sc[i * N + i], a22 = Sc[i2 * N + i2], a12 = Sc[i * N + i2];
          v = !roti ? aii
                    : (lowi ? (ci * ci) * aii - (2 * ci * si) * a12 + (si * si) * a22
                            : (si * si) * a22 + (2 * ci * si) * a12 + (ci * ci) * aii);
        } else if (i2 == j) {
          v = roti ? 0.0 : Sc[i * N + j];
        } else {
          const int j2 = (int)(pj[m] >> (4 * r)) & 15;
          const bool lowj = j < j2;
          const double cj = cc[j], sj = ss[j];
          const double a0 = Sc[i * N + j], a1 = Sc[i * N + j2], b0 = Sc[i2 * N + j], b1 = Sc[i2 * N + j2];
          const double bi = lowj ? cj * a0 - sj * a1 : sj * a1 + cj * a0;
          const double bi2 = lowj ? cj * b0 - sj * b1 : sj * b1 + cj * b0;
          v = lowi ? ci * bi - si * bi2 : si * bi2 + ci * bi;
        }
        Sn[i * N + j] = v;
        Sn[j * N + i] = v;
      }
#pragma unroll
      for (int m = 0; m < VPT; m++) {
        const int v = (NT - 1 - tid) + NT * m;  // the threads with the fewest S elements take the most Vr elements
        if (v >= NV) continue;
        const int x = v / N, k = v - N * x, x2 = (int)(pv[m] >> (4 * r)) & 15;
        const double a = Vc[x * N + k], b = Vc[x2 * N + k];
        Vn[v] = (x < x2) ? cc[x] * a - ss[x] * b : ss[x] * b + cc[x] * a;
      }
      sync();
      {
        double* t = Sc;
        Sc = Sn;
        Sn = t;
        t = Vc;
        Vc = Vn;
        Vn = t;
      }
      if (prof) t_apply += clock64() - t0;
    }
    if (!changed) break;
  }
  if (prof && tid == 0) {
    prof[0] = t_rot;
    prof[1] = t_apply;
    prof[2] = 0;
    prof[3] = n_rounds;
  }
  // selection sort (descending) of the eigenvalues; the row swaps are applied as one permutation
  if (tid < 32) {
    int perm[N];
    double W[N];
    for (int i = 0; i < N; i++) {
      W[i] = Sc[i * N + i];
      perm[i] = i;
    }
    for (int i = 0; i < N - 1; i++) {
      int j = i;
      for (int k = i + 1; k < N; k++)
        if (W[j] < W[k]) j = k;
      if (i != j) {
        const double tw = W[i];
        W[i] = W[j];
        W[j] = tw;
        const int tp = perm[i];
        perm[i] = perm[j];
        perm[j] = tp;
      }
    }
    if (tid == 0)
      for (int i = 0; i < N; i++) w[i] = W[i];
    if (tid < N)
      for (int i = 0; i < N; i++) Vt[i * N + tid] = Vc[perm[i] * N + tid];
  }
  sync();
}

// least-squares / pseudo-inverse solve via SVD (cvSolve CV_SVD); m <= 6, n <= 6 (one thread, cyclic pair order)
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
  // cvInvert(CV_SVD) column by column (x_c = pinv(A) e_c); the decomposition of A is computed once and reused -- it
  // is a deterministic function of A, so this equals three independent svd_solve6 calls bit for bit
  double w[3], U[9], Vt[9];
  jacobi_svd<3, 3, 3>(A, 3, 3, w, U, Vt);
  double thr = 0;
  for (int i = 0; i < 3; i++) thr += w[i];
  thr *= DBL_EPSILON * 2;
  for (int c = 0; c < 3; c++) {
    double x[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) {
      if (w[i] <= thr) continue;
      double s = 0;
      for (int k = 0; k < 3; k++) s += U[k * 3 + i] * (k == c ? 1.0 : 0.0);
      s /= w[i];
      for (int k = 0; k < 3; k++) x[k] += s * Vt[i * 3 + k];
    }
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
