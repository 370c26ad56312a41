// pose.cu -- K9-K12 on sm_100a: cv::triangulatePoints, extract_3Dpoints, cv::solvePnPRansac(EPNP) and the median
// helpers.  Reference call sites: visual_odometry.h:631-648 (stereo), :355-368 (mono),
// VO_utility.cpp:188-237, :23-63, :725-748.
//
// RANSAC is evaluated as a batch: the reference-matching subset stream is rebuilt from the fixed cv::RNG table, all
// `iterations` minimal-set hypotheses are solved (one EPnP per thread) and scored (one block per hypothesis) in
// parallel, and a warp then replays OpenCV's sequential bookkeeping -- `if good > max(best, 4) {best = ...;
// niters = RANSACUpdateNumIters(...)}` -- over the per-hypothesis inlier counts to find the hypothesis the CPU loop
// would have kept and the iteration at which it would have stopped.  Hypotheses past that point are wasted work,
// never a different answer.
#include <cstring>
#include <mutex>

#include "epnp.cuh"
#include "pose.cuh"

namespace uvo {

// ------------------------------------------------------------------------------------------------ RNG table
static std::mutex g_rng_mutex;
static uint32_t* g_rng_dev[64] = {};

const uint32_t* rng_table_device(Ctx& c) {
  std::lock_guard<std::mutex> lock(g_rng_mutex);
  UVO_REQUIRE(c.device < 64, "device index too large");
  if (g_rng_dev[c.device]) return g_rng_dev[c.device];
  std::vector<uint32_t> h(RNG_TABLE_SIZE);
  uint64_t state = 0xFFFFFFFFFFFFFFFFull;  // RNG rng((uint64)-1)  (ptsetreg.cpp)
  for (int i = 0; i < RNG_TABLE_SIZE; i++) {
    state = (uint64_t)(uint32_t)state * 4164903690U + (uint32_t)(state >> 32);
    h[i] = (uint32_t)state;
  }
  uint32_t* d = nullptr;
  UVO_CUDA(cudaMalloc((void**)&d, sizeof(uint32_t) * RNG_TABLE_SIZE));
  UVO_CUDA(cudaMemcpy(d, h.data(), sizeof(uint32_t) * RNG_TABLE_SIZE, cudaMemcpyHostToDevice));
  g_rng_dev[c.device] = d;
  return d;
}

// ------------------------------------------------------------------------------------------------ block helpers
template <int K>
__device__ __forceinline__ void block_reduce_sum(double (&v)[K], double* s_red /* >= 32*K */, double* out /* K */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < K; k++)
#pragma unroll
    for (int o = 16; o; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if (lane == 0)
    for (int k = 0; k < K; k++) s_red[wid * K + k] = v[k];
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0;
    for (int w = 0; w < nw; w++) s += s_red[w * K + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// ordered (stable) compaction step for one chunk of blockDim.x candidates; returns destination or -1
__device__ __forceinline__ int ordered_slot(bool keep, int* s_warp /*32*/, int* s_base) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) s_warp[wid] = __popc(bal);
  __syncthreads();
  int off = *s_base;
  for (int k = 0; k < wid; k++) off += s_warp[k];
  const int slot = keep ? off + __popc(bal & ((1u << lane) - 1)) : -1;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int k = 0; k < nw; k++) tot += s_warp[k];
    *s_base += tot;
  }
  __syncthreads();
  return slot;
}

// ------------------------------------------------------------------------------------------------ K11 triangulate
__global__ void __launch_bounds__(64) k_triangulate(const __grid_constant__ TriangulateArgs a) {
  int n = a.n_dev ? min(*a.n_dev, a.n) : a.n;
  if (a.gate_dev && *a.gate_dev == 0) n = 0;
  if (a.min_points >= 0 && n <= a.min_points) n = 0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && a.n_out) *a.n_out = n;
  if (i >= n) return;
  float p1[2], p2[2];
  if (a.matches) {
    const int q = a.matches[i].queryIdx;
    p1[0] = a.kps1[q].x;
    p1[1] = a.kps1[q].y;
    p2[0] = a.kps2[q].x;
    p2[1] = a.kps2[q].y;
  } else {
    p1[0] = a.pts1[2 * i];
    p1[1] = a.pts1[2 * i + 1];
    p2[0] = a.pts2[2 * i];
    p2[1] = a.pts2[2 * i + 1];
  }
  if (a.pts1_out) {
    a.pts1_out[2 * i] = p1[0];
    a.pts1_out[2 * i + 1] = p1[1];
    a.pts2_out[2 * i] = p2[0];
    a.pts2_out[2 * i + 1] = p2[1];
  }
  double A[16];
  {
    const double x = p1[0], y = p1[1];
    for (int k = 0; k < 4; k++) {
      A[0 * 4 + k] = x * a.P1[8 + k] - a.P1[k];
      A[1 * 4 + k] = y * a.P1[8 + k] - a.P1[4 + k];
    }
  }
  {
    const double x = p2[0], y = p2[1];
    for (int k = 0; k < 4; k++) {
      A[2 * 4 + k] = x * a.P2[8 + k] - a.P2[k];
      A[3 * 4 + k] = y * a.P2[8 + k] - a.P2[4 + k];
    }
  }
  double w[4], Vt[16];
  jacobi_svd<4>(A, 4, 4, w, nullptr, Vt);
  for (int k = 0; k < 4; k++) a.out4[(size_t)k * a.stride + i] = (float)Vt[12 + k];
}

void launch_triangulate(Ctx& c, const TriangulateArgs& a) {
  if (a.n <= 0) return;
  UVO_KERNEL(c, "k_triangulate");
  k_triangulate<<<div_up(a.n, 64), 64, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------------ K12 extract_3Dpoints
__global__ void __launch_bounds__(1024) k_extract3d(const __grid_constant__ Extract3dArgs a) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  __shared__ double s_red[64], s_out[2];
  const int n = a.n_dev ? min(*a.n_dev, a.n) : a.n;
  const int tid = threadIdx.x;
  if (tid == 0) s_base = 0;
  __syncthreads();
  // pass 1: dehomogenise (f32), reproject into both views, keep mean error < tol and z > 0
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + tid;
    bool keep = false;
    double P[3] = {0, 0, 0};
    if (i < n && n >= a.min3d) {
      const float W = a.p4[(size_t)3 * a.stride + i];
      const float scale = W != 0.f ? __fdiv_rn(1.f, W) : 1.f;
      for (int cc = 0; cc < 3; cc++) P[cc] = (double)__fmul_rn(a.p4[(size_t)cc * a.stride + i], scale);
      double m1[2], m2[2];
      project1(P, a.R1, a.t1, a.K1, m1);
      project1(P, a.R2, a.t2, a.K2, m2);
      double dx = (double)a.kp1[2 * i] - m1[0], dy = (double)a.kp1[2 * i + 1] - m1[1];
      const double e1 = sqrt(dx * dx + dy * dy);
      dx = (double)a.kp2[2 * i] - m2[0];
      dy = (double)a.kp2[2 * i + 1] - m2[1];
      const double e2 = sqrt(dx * dx + dy * dy);
      const double mean = (e1 + e2) / 2.0;
      keep = (mean < a.tol) && (P[2] > 0);
    }
    const int slot = ordered_slot(keep, s_warp, &s_base);
    if (slot >= 0) {
      a.tmp_idx[slot] = i;
      a.tmp_pts[3 * slot] = P[0];
      a.tmp_pts[3 * slot + 1] = P[1];
      a.tmp_pts[3 * slot + 2] = P[2];
    }
  }
  __syncthreads();
  const int ng = s_base;
  __syncthreads();
  if (tid == 0) s_base = 0;
  __syncthreads();
  if (ng >= a.min3d && ng > 0) {
    // compute_mean_and_variance (population variance E[z^2] - E[z]^2), then the 3-sigma depth gate
    double acc[2] = {0, 0};
    for (int i = tid; i < ng; i += blockDim.x) {
      const double z = a.tmp_pts[3 * i + 2];
      acc[0] += z;
      acc[1] += z * z;
    }
    block_reduce_sum<2>(acc, s_red, s_out);
    const double mean = s_out[0] / ng, var = (s_out[1] / ng) - (mean * mean);
    const double sd = sqrt(var);  // NaN when var < 0: every comparison below is then false
    for (int base = 0; base < ng; base += blockDim.x) {
      const int i = base + tid;
      bool keep = false;
      double z = 0;
      if (i < ng) {
        z = a.tmp_pts[3 * i + 2];
        keep = (z <= mean + 3.0 * sd) && (z >= mean - 3.0 * sd);
      }
      const int slot = ordered_slot(keep, s_warp, &s_base);
      if (slot >= 0) {
        a.out_idx[slot] = a.tmp_idx[i];
        a.out_pts[3 * slot] = a.tmp_pts[3 * i];
        a.out_pts[3 * slot + 1] = a.tmp_pts[3 * i + 1];
        a.out_pts[3 * slot + 2] = z;
      }
    }
  }
  __syncthreads();
  if (tid == 0) *a.out_count = s_base;
}

void launch_extract3d(Ctx& c, const Extract3dArgs& a) {
  UVO_KERNEL(c, "k_extract3d");
  k_extract3d<<<1, 1024, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------------ K10c PnP RANSAC
__device__ __forceinline__ int pnp_n(const PnpArgs& a) {
  int n = a.n_dev ? min(*a.n_dev, a.n) : a.n;
  if (a.min_points >= 0 && n <= a.min_points) return 0;
  return n;
}

__global__ void __launch_bounds__(256) k_pnp_prepare(const __grid_constant__ PnpArgs a) {
  const int n = pnp_n(a);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (a.x_idx) {
      const uvo_keypoint& k = a.kps[a.matches[a.x_idx[i]].trainIdx];
      a.xs[2 * i] = k.x;
      a.xs[2 * i + 1] = k.y;
    } else {
      a.xs[2 * i] = a.x[2 * i];
      a.xs[2 * i + 1] = a.x[2 * i + 1];
    }
    for (int cc = 0; cc < 3; cc++) a.Xf[3 * i + cc] = (float)a.X[3 * i + cc];  // opoints0.convertTo(CV_32F)
  }
}

// getSubset stream for all iterations (one warp).  Lane l speculatively takes draws [p+5l, p+5l+5); a subset with a
// repeated index is rebuilt sequentially by lane 0 with OpenCV's redraw rule and the stream position re-aligned.
__global__ void __launch_bounds__(32) k_pnp_subsets(const __grid_constant__ PnpArgs a,
                                                    const uint32_t* __restrict__ rng) {
  const int n = pnp_n(a);
  const int iters = max(a.iterations, 1);
  const int lane = threadIdx.x;
  if (n < 5) return;
  if (n == 5) {
    if (lane < 5) a.subsets[lane] = lane;
    return;
  }
  int p = 0, s = 0;
  while (s < iters) {
    int idx[5];
    bool dup = false;
    const int q = p + 5 * lane;
    const bool in_range = (s + lane < iters) && (q + 5 <= RNG_TABLE_SIZE);
    if (in_range) {
#pragma unroll
      for (int k = 0; k < 5; k++) idx[k] = (int)(rng[q + k] % (unsigned)n);
#pragma unroll
      for (int k = 1; k < 5; k++)
#pragma unroll
        for (int j = 0; j < k; j++) dup |= (idx[k] == idx[j]);
    }
    const unsigned bad = __ballot_sync(0xffffffffu, in_range && dup);
    const unsigned act = __ballot_sync(0xffffffffu, in_range);
    const int n_act = __popc(act);
    if (n_act == 0) break;  // RNG table exhausted (cannot happen for iterations*5*redraws < 2^18)
    const int first_bad = bad ? (__ffs(bad) - 1) : n_act;
    if (in_range && lane < first_bad)
      for (int k = 0; k < 5; k++) a.subsets[(size_t)(s + lane) * 5 + k] = idx[k];
    s += first_bad;
    p += 5 * first_bad;
    if (bad) {
      if (lane == 0) {
        int sub[5];
        for (int i = 0; i < 5; i++) {
          int v;
          for (;;) {
            v = (int)(rng[min(p, RNG_TABLE_SIZE - 1)] % (unsigned)n);
            p++;
            bool d2 = false;
            for (int k = 0; k < i; k++) d2 |= (sub[k] == v);
            if (!d2) break;
          }
          sub[i] = v;
        }
        for (int k = 0; k < 5; k++) a.subsets[(size_t)s * 5 + k] = sub[k];
      }
      p = __shfl_sync(0xffffffffu, p, 0);
      s += 1;
    }
  }
}

// One EPnP hypothesis (PnPRansacCallback::runKernel) per group of HYP_GL lanes.  The 5 correspondences arrive as f32,
// image points are normalised with undistortPoints (f32 round trip), the pose goes R -> rvec (Rodrigues) and is stored
// together with the rotation matrix projectPoints rebuilds from rvec.  Inside a group: the small set-up algebra is
// done redundantly by every lane, M^T M is accumulated 9 entries per lane, the 12 x 12 eigen-decomposition is
// lane-parallel (jacobi_eigh_group), lanes 0..5 / 6..11 build L / rho, and lanes 0..2 each refine one of the three
// beta candidates and its pose.  Every value is produced by the same IEEE operations in the same order as the
// one-thread restatement in the oracle, so the split changes latency, not bits.
constexpr int HYP_GL = 16;  // (one warp per hypothesis was measured slower: 305 vs 281 us, and it crowds the SMs)
constexpr int HYP_PER_BLOCK = 8;

__global__ void __launch_bounds__(HYP_GL* HYP_PER_BLOCK) k_pnp_hyp(const __grid_constant__ PnpArgs a) {
  __shared__ double s_S[HYP_PER_BLOCK][144], s_Vr[HYP_PER_BLOCK][144], s_ut[HYP_PER_BLOCK][144];
  __shared__ double s_w[HYP_PER_BLOCK][12], s_l[HYP_PER_BLOCK][60], s_rho[HYP_PER_BLOCK][6];
  __shared__ double s_al[HYP_PER_BLOCK][5][4];
  const int n = pnp_n(a);
  if (n < 5) return;
  const int iters = (n == 5) ? 1 : max(a.iterations, 1);
  const int g = threadIdx.x / HYP_GL, gl = threadIdx.x % HYP_GL;
  const int h = blockIdx.x * HYP_PER_BLOCK + g;
  if (h >= iters) return;  // whole group
  const int lane = threadIdx.x & 31;
  const int gbase = lane & ~(HYP_GL - 1);
  const unsigned gmask = (HYP_GL == 32 ? 0xffffffffu : ((1u << (HYP_GL & 31)) - 1u)) << gbase;
  EpnpCam cam{a.K[0], a.K[1], a.K[2], a.K[3]};
  const double ifx = 1. / a.K[0], ify = 1. / a.K[1];
  double pws[15], us[10];
  for (int k = 0; k < 5; k++) {
    const int i = a.subsets[(size_t)h * 5 + k];
    for (int cc = 0; cc < 3; cc++) pws[3 * k + cc] = (double)a.Xf[3 * i + cc];
    const double xn = (double)(float)(((double)a.xs[2 * i] - a.K[2]) * ifx);
    const double yn = (double)(float)(((double)a.xs[2 * i + 1] - a.K[3]) * ify);
    us[2 * k] = xn * cam.fu + cam.uc;
    us[2 * k + 1] = yn * cam.fv + cam.vc;
  }
  double cws[4][3], ci[9];
  double(*alphas)[4] = s_al[g];
  {
    double al[5][4];
    epnp_small_setup(pws, 5, cws, ci, al);  // every lane computes the same values
    if (gl == 0)
      for (int i = 0; i < 5; i++)
        for (int k = 0; k < 4; k++) alphas[i][k] = al[i][k];
  }
  __syncwarp(gmask);
  // M^T M, each entry summed over the points in index order
  for (int e = gl; e < 144; e += HYP_GL) {
    const int ra = e / 12, rb = e - 12 * ra;
    double acc = 0;
    for (int i = 0; i < 5; i++) {
      double m1a, m2a, m1b, m2b;
      epnp_m_elem(alphas[i], us[2 * i], us[2 * i + 1], cam, ra, m1a, m2a);
      epnp_m_elem(alphas[i], us[2 * i], us[2 * i + 1], cam, rb, m1b, m2b);
      acc += m1a * m1b + m2a * m2b;
    }
    s_S[g][e] = acc;
  }
  __syncwarp(gmask);
  jacobi_eigh_group<12>(s_S[g], s_Vr[g], s_w[g], s_ut[g], gl, gmask);
  if (gl < 6) epnp_L_row(s_ut[g], gl, s_l[g] + 10 * gl);
  else if (gl < 12) s_rho[g][gl - 6] = epnp_rho_entry(cws, gl - 6);
  __syncwarp(gmask);
  double err = 0, Rw[9], tw[3];
  if (gl < 3) {
    double betas[4];
    epnp_betas_which(s_l[g], s_rho[g], gl + 1, betas);
    err = epnp_candidate(pws, us, 5, alphas, betas, s_ut[g], cam, Rw, tw);
  }
  // N = 1; if (rep[2] < rep[1]) N = 2; if (rep[3] < rep[N]) N = 3;
  const double e1 = __shfl_sync(gmask, err, gbase + 0), e2 = __shfl_sync(gmask, err, gbase + 1),
               e3 = __shfl_sync(gmask, err, gbase + 2);
  int N = 1;
  if (e2 < e1) N = 2;
  if (e3 < (N == 1 ? e1 : e2)) N = 3;
  if (gl == N - 1) {
    double rvec[3], R2[9];
    rodrigues_mat2vec(Rw, rvec);
    rodrigues_vec2mat(rvec, R2);
    double* m = a.hyp_model + (size_t)h * 15;
    for (int i = 0; i < 9; i++) m[i] = R2[i];
    for (int i = 0; i < 3; i++) m[9 + i] = tw[i];
    for (int i = 0; i < 3; i++) m[12 + i] = rvec[i];
  }
}

// PnPRansacCallback::computeError + findInliers for one point
__device__ __forceinline__ bool pnp_is_inlier(const double* m, const float* Xf, const float* xs, int i,
                                              const double K[4], float thr) {
  const double P[3] = {(double)Xf[3 * i], (double)Xf[3 * i + 1], (double)Xf[3 * i + 2]};
  double pr[2];
  project1(P, m, m + 9, K, pr);
  const float dx = __fsub_rn(xs[2 * i], (float)pr[0]), dy = __fsub_rn(xs[2 * i + 1], (float)pr[1]);
  const float e = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  return e <= thr;
}

__global__ void __launch_bounds__(256) k_pnp_score(const __grid_constant__ PnpArgs a) {
  __shared__ int s_cnt[8];
  const int n = pnp_n(a);
  if (n < 5) return;
  const int iters = (n == 5) ? 1 : max(a.iterations, 1);
  const int h = blockIdx.x;
  if (h >= iters) return;
  const float thr = (float)((double)a.reproj_err * (double)a.reproj_err);
  const double* m = a.hyp_model + (size_t)h * 15;
  int cnt = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) cnt += pnp_is_inlier(m, a.Xf, a.xs, i, a.K, thr) ? 1 : 0;
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 8; w++) tot += s_cnt[w];
    a.hyp_good[h] = tot;
  }
}

__device__ inline int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmax(p, 0.);
  p = fmin(p, 1.);
  ep = fmax(ep, 0.);
  ep = fmin(ep, 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : __double2int_rn(num / denom);
}

// sequential RANSAC bookkeeping replayed by one warp over the per-hypothesis inlier counts
__global__ void __launch_bounds__(32) k_pnp_scan(const __grid_constant__ PnpArgs a) {
  const int n = pnp_n(a);
  const int lane = threadIdx.x;
  if (n < 5) {
    if (lane == 0) {
      a.best[0] = -1;
      a.best[1] = 0;
      *a.hyps = 0;
    }
    return;
  }
  if (n == 5) {  // count == modelPoints: a single kernel call, every point an inlier
    if (lane == 0) {
      a.best[0] = 0;
      a.best[1] = 5;
      *a.hyps = 1;
    }
    return;
  }
  int niters = max(a.iterations, 1), best = 0, best_h = -1, iter = 0;
  while (iter < niters) {
    const int idx = iter + lane;
    const int g = idx < niters ? a.hyp_good[idx] : -1;
    const unsigned mask = __ballot_sync(0xffffffffu, g > max(best, 4));
    if (!mask) {
      iter = min(iter + 32, niters);
      continue;
    }
    const int f = __ffs(mask) - 1;
    best = __shfl_sync(0xffffffffu, g, f);
    best_h = iter + f;
    niters = ransac_update_num_iters(a.confidence, (double)(n - best) / n, 5, niters);
    iter = best_h + 1;
  }
  if (lane == 0) {
    a.best[0] = best_h;
    a.best[1] = best;
    *a.hyps = iter;
  }
}

// inlier mask of the winning hypothesis -> ascending inlier list, then the EPnP refit on all inliers
// (solvePnPRansac's final solvePnP on f64 copies of the f32 data).  One block.
constexpr int REFIT_THREADS = 256;
constexpr int REFIT_TILE = 64;

__global__ void __launch_bounds__(REFIT_THREADS) k_pnp_finalize(const __grid_constant__ PnpArgs a) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  __shared__ double s_red[32 * 27], s_out[27];
  __shared__ double s_cws[4][3], s_ci[9], s_ut[144], s_betas[3][4], s_ccs3[3][4][3], s_R3[3][9], s_t3[3][3], s_sign3[3];
  __shared__ double s_rows[REFIT_TILE][24];
  __shared__ double s_mtm[144], s_Vr[144], s_w[12], s_l[60], s_rho[6];
  const int tid = threadIdx.x;
  const int n = pnp_n(a);
  const int best_h = a.best[0];
  if (n < 5 || best_h < 0) {
    if (tid == 0) {
      *a.n_inliers = 0;
      for (int i = 0; i < 6; i++) a.result[i] = 0.0;
      a.result[6] = 0.0;
      if (n >= 5 && best_h < 0 && a.iterations > 0) {  // no model beat 4 inliers: cv2 returns the last rvec/tvec
      }
    }
    return;
  }
  const float thr = (float)((double)a.reproj_err * (double)a.reproj_err);
  const double* m = a.hyp_model + (size_t)best_h * 15;
  if (n == 5) {  // solvePnPRansac: model_points == npoints -> plain solvePnP, all points inliers, no refit
    if (tid < 5) a.inliers[tid] = tid;
    if (tid == 0) {
      *a.n_inliers = 5;
      for (int i = 0; i < 3; i++) a.result[i] = m[12 + i];
      for (int i = 0; i < 3; i++) a.result[3 + i] = m[9 + i];
      a.result[6] = 1.0;
    }
    return;
  }
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + tid;
    const bool keep = i < n && pnp_is_inlier(m, a.Xf, a.xs, i, a.K, thr);
    const int slot = ordered_slot(keep, s_warp, &s_base);
    if (slot >= 0) a.inliers[slot] = i;
  }
  __syncthreads();
  const int ni = s_base;
  if (tid == 0) *a.n_inliers = ni;
  // ---------------- EPnP on the inliers ----------------
  const EpnpCam cam{a.K[0], a.K[1], a.K[2], a.K[3]};
  const double ifx = 1. / a.K[0], ify = 1. / a.K[1];
  auto point = [&](int k, double pw[3], double& u, double& v) {
    const int i = a.inliers[k];
    for (int cc = 0; cc < 3; cc++) pw[cc] = (double)a.Xf[3 * i + cc];
    u = (((double)a.xs[2 * i] - cam.uc) * ifx) * cam.fu + cam.uc;       // undistortPoints (f64) then init_points
    v = (((double)a.xs[2 * i + 1] - cam.vc) * ify) * cam.fv + cam.vc;
  };
  double pw[3], u, v;
  {
    double acc[3] = {0, 0, 0};
    for (int k = tid; k < ni; k += blockDim.x) {
      point(k, pw, u, v);
      for (int cc = 0; cc < 3; cc++) acc[cc] += pw[cc];
    }
    block_reduce_sum<3>(acc, s_red, s_out);
  }
  const double c0[3] = {s_out[0] / ni, s_out[1] / ni, s_out[2] / ni};
  __syncthreads();
  {
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = tid; k < ni; k += blockDim.x) {
      point(k, pw, u, v);
      const double d[3] = {pw[0] - c0[0], pw[1] - c0[1], pw[2] - c0[2]};
      for (int p = 0; p < 3; p++)
        for (int q = 0; q < 3; q++) acc[p * 3 + q] += d[p] * d[q];
    }
    block_reduce_sum<9>(acc, s_red, s_out);
  }
  if (tid == 0) {
    double cov[9];
    for (int i = 0; i < 9; i++) cov[i] = s_out[i];
    epnp_control_points(c0, cov, ni, s_cws, s_ci);
  }
  __syncthreads();
  // MtM, summed over points in index order (entry (p,q) owned by thread p*12+q)
  double mt = 0;
  for (int base = 0; base < ni; base += REFIT_TILE) {
    const int cntp = min(REFIT_TILE, ni - base);
    if (tid < cntp) {
      double al[4];
      point(base + tid, pw, u, v);
      epnp_alphas(pw, s_cws, s_ci, al);
      epnp_m_rows(al, u, v, cam, &s_rows[tid][0], &s_rows[tid][12]);
    }
    __syncthreads();
    if (tid < 144) {
      const int p = tid / 12, q = tid % 12;
      for (int k = 0; k < cntp; k++) mt += s_rows[k][p] * s_rows[k][q] + s_rows[k][12 + p] * s_rows[k][12 + q];
    }
    __syncthreads();
  }
  if (tid < 144) s_mtm[tid] = mt;
  __syncthreads();
  // eigenvectors of M^T M lane-parallel on warp 0, L / rho on lanes 0..11, one beta candidate per lane 0..2
  if (tid < 32) {
    jacobi_eigh_group<12>(s_mtm, s_Vr, s_w, s_ut, tid, 0xffffffffu);
    if (tid < 6) epnp_L_row(s_ut, tid, s_l + 10 * tid);
    else if (tid < 12) s_rho[tid - 6] = epnp_rho_entry(s_cws, tid - 6);
    __syncwarp();
    if (tid < 3) epnp_betas_which(s_l, s_rho, tid + 1, s_betas[tid]);
  }
  __syncthreads();
  // the three beta candidates go through compute_R_and_t together: one pass over the inliers accumulates all three
  // sums (the barycentric coordinates of a point are shared), the one-thread steps run on lanes 0..2 of warp 0.  Per
  // candidate the operations and their order are those of a loop over the candidates.
  if (tid < 3) {
    const int w = tid;
    epnp_ccs(s_betas[w], s_ut, s_ccs3[w]);
    double al[4], pc[3];
    point(0, pw, u, v);
    epnp_alphas(pw, s_cws, s_ci, al);
    epnp_pc(al, s_ccs3[w], 1.0, pc);
    s_sign3[w] = pc[2] < 0.0 ? -1.0 : 1.0;  // solve_for_sign
  }
  __syncthreads();
  double pc0[3][3];
  {
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = tid; k < ni; k += blockDim.x) {
      double al[4], pc[3];
      point(k, pw, u, v);
      epnp_alphas(pw, s_cws, s_ci, al);
#pragma unroll
      for (int w = 0; w < 3; w++) {
        epnp_pc(al, s_ccs3[w], s_sign3[w], pc);
        for (int cc = 0; cc < 3; cc++) acc[3 * w + cc] += pc[cc];
      }
    }
    block_reduce_sum<9>(acc, s_red, s_out);
    for (int w = 0; w < 3; w++)
      for (int cc = 0; cc < 3; cc++) pc0[w][cc] = s_out[3 * w + cc] / ni;
  }
  __syncthreads();
  {
    double acc[27];
    for (int i = 0; i < 27; i++) acc[i] = 0;
    for (int k = tid; k < ni; k += blockDim.x) {
      double al[4], pc[3];
      point(k, pw, u, v);
      epnp_alphas(pw, s_cws, s_ci, al);
#pragma unroll
      for (int w = 0; w < 3; w++) {
        epnp_pc(al, s_ccs3[w], s_sign3[w], pc);
        for (int j = 0; j < 3; j++) {
          acc[9 * w + 3 * j] += (pc[j] - pc0[w][j]) * (pw[0] - c0[0]);
          acc[9 * w + 3 * j + 1] += (pc[j] - pc0[w][j]) * (pw[1] - c0[1]);
          acc[9 * w + 3 * j + 2] += (pc[j] - pc0[w][j]) * (pw[2] - c0[2]);
        }
      }
    }
    block_reduce_sum<27>(acc, s_red, s_out);
  }
  if (tid < 3) {
    double abt[9];
    for (int i = 0; i < 9; i++) abt[i] = s_out[9 * tid + i];
    epnp_rt_from_abt(abt, pc0[tid], c0, s_R3[tid], s_t3[tid]);
  }
  __syncthreads();
  {
    double acc[3] = {0, 0, 0};
    for (int k = tid; k < ni; k += blockDim.x) {
      point(k, pw, u, v);
#pragma unroll
      for (int w = 0; w < 3; w++) acc[w] += epnp_reproj1(s_R3[w], s_t3[w], pw, u, v, cam);
    }
    block_reduce_sum<3>(acc, s_red, s_out);
  }
  double best_err = 0, bestR[9], bestt[3];
  for (int w = 0; w < 3; w++) {
    const double err = s_out[w] / ni;
    if (w == 0 || err < best_err) {
      best_err = err;
      for (int i = 0; i < 9; i++) bestR[i] = s_R3[w][i];
      for (int i = 0; i < 3; i++) bestt[i] = s_t3[w][i];
    }
  }
  if (tid == 0) {
    double rvec[3];
    rodrigues_mat2vec(bestR, rvec);
    for (int i = 0; i < 3; i++) a.result[i] = rvec[i];
    for (int i = 0; i < 3; i++) a.result[3 + i] = bestt[i];
    a.result[6] = 1.0;
  }
}

// ------------------------------------------------------------------------------------------------ K9 / scale helpers
// rank selection: element i has rank #{j : v_j < v_i or (v_j == v_i and j < i)}; ranks are a permutation, so exactly
// one element lands on each middle rank.  O(n^2) compares spread over the grid; n is a match count (<= capacity).
__global__ void __launch_bounds__(256) k_median_rank(const double* __restrict__ v, const int* n_dev, int n_host,
                                                     double* mid /* [2] */) {
  __shared__ double s_v[256];
  const int n = n_dev ? min(*n_dev, n_host) : n_host;
  if ((int)(blockIdx.x * blockDim.x) >= n) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double me = i < n ? v[i] : 0.0;
  int rank = 0;
  for (int base = 0; base < n; base += 256) {
    if (base + threadIdx.x < n) s_v[threadIdx.x] = v[base + threadIdx.x];
    __syncthreads();
    const int m = min(256, n - base);
    if (i < n)
      for (int q = 0; q < m; q++) rank += (s_v[q] < me || (s_v[q] == me && base + q < i)) ? 1 : 0;
    __syncthreads();
  }
  if (i < n) {
    if (rank == n / 2) mid[1] = me;
    if (rank == n / 2 - 1) mid[0] = me;
  }
}
__global__ void k_median_finish(const int* n_dev, int n_host, const double* mid, double* out) {
  const int n = n_dev ? min(*n_dev, n_host) : n_host;
  if (n == 0) out[0] = 0.0;
  else if (n % 2 == 0) out[0] = (mid[0] + mid[1]) / 2.0;
  else out[0] = mid[1];
}

void launch_median(Ctx& c, const double* v, const int* n_dev, int n, double* out) {
  // out[0] = median, out[1..2] = scratch for the two middle order statistics
  if (n > 0) {
    UVO_KERNEL(c, "k_median_rank");
    k_median_rank<<<div_up(n, 256), 256, 0, c.stream>>>(v, n_dev, n, out + 1);
    UVO_LAUNCH_CHECK(c);
  }
  UVO_KERNEL(c, "k_median_finish");
  k_median_finish<<<1, 1, 0, c.stream>>>(n_dev, n, out + 1, out);
  UVO_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(256) k_displacements(const float* __restrict__ p1, const float* __restrict__ p2,
                                                       int n, double* __restrict__ d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // `double dx = keypoints1_conv[i].x - keypoints2_conv[i].x` : the subtraction is f32, then widened
  const double dx = (double)__fsub_rn(p1[2 * i], p2[2 * i]), dy = (double)__fsub_rn(p1[2 * i + 1], p2[2 * i + 1]);
  d[i] = sqrt(dx * dx + dy * dy);
}
void launch_displacements(Ctx& c, const float* p1, const float* p2, int n, double* disp) {
  if (n <= 0) return;
  UVO_KERNEL(c, "k_displacements");
  k_displacements<<<div_up(n, 256), 256, 0, c.stream>>>(p1, p2, n, disp);
  UVO_LAUNCH_CHECK(c);
}

struct FrontZArgs {
  double R[9], t[3];
};
__global__ void __launch_bounds__(1024) k_front_z(const double* __restrict__ pts, int n, FrontZArgs a, double* z_out,
                                                  int* n_out) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    bool keep = false;
    double z = 0;
    if (i < n) {
      const double* p = pts + 3 * i;
      z = p[2];
      keep = (a.R[6] * p[0] + a.R[7] * p[1] + a.R[8] * p[2] + a.t[2]) > 0;
    }
    const int slot = ordered_slot(keep, s_warp, &s_base);
    if (slot >= 0) z_out[slot] = z;
  }
  __syncthreads();
  if (threadIdx.x == 0) *n_out = s_base;
}
void launch_front_z(Ctx& c, const double* pts, int n, const double R[9], const double t[3], double* z_out, int* n_out) {
  FrontZArgs a;
  memcpy(a.R, R, sizeof(a.R));
  memcpy(a.t, t, sizeof(a.t));
  UVO_KERNEL(c, "k_front_z");
  k_front_z<<<1, 1024, 0, c.stream>>>(pts, n, a, z_out, n_out);
  UVO_LAUNCH_CHECK(c);
}

size_t pnp_scratch_bytes(int n, int iterations) {
  const size_t it = (size_t)std::max(iterations, 1);
  return it * 5 * sizeof(int32_t) + it * 15 * sizeof(double) + it * sizeof(int) + (size_t)n * 2 * sizeof(float) +
         (size_t)n * 3 * sizeof(float) + 64;
}

void launch_pnp_prepare(Ctx& c, const PnpArgs& a) {
  const uint32_t* rng = rng_table_device(c);
  const int iters = std::max(a.iterations, 1);
  UVO_REQUIRE((size_t)iters * 8 < (size_t)RNG_TABLE_SIZE, "solvePnPRansac: iterationsCount too large for the RNG table");
  if (a.n > 0) {
    UVO_KERNEL(c, "k_pnp_prepare");
    k_pnp_prepare<<<std::min(div_up(a.n, 256), 4 * c.sm_count), 256, 0, c.stream>>>(a);
    UVO_LAUNCH_CHECK(c);
  }
  UVO_KERNEL(c, "k_pnp_subsets");
  k_pnp_subsets<<<1, 32, 0, c.stream>>>(a, rng);
  UVO_LAUNCH_CHECK(c);
}

void launch_pnp_solve(Ctx& c, const PnpArgs& a) {
  const int iters = std::max(a.iterations, 1);
  UVO_KERNEL(c, "k_pnp_hyp");
  k_pnp_hyp<<<div_up(iters, HYP_PER_BLOCK), HYP_GL * HYP_PER_BLOCK, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_pnp_score");
  k_pnp_score<<<iters, 256, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_pnp_scan");
  k_pnp_scan<<<1, 32, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_pnp_finalize");
  k_pnp_finalize<<<1, REFIT_THREADS, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
}

void launch_pnp_ransac(Ctx& c, const PnpArgs& a) {
  launch_pnp_prepare(c, a);
  launch_pnp_solve(c, a);
}

}  // namespace uvo
