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

__device__ __forceinline__ void pnp_prof(const PnpArgs& a, int slot) {
  if (a.prof) a.prof[slot] = clock64();
}

// getSubset stream for all iterations (one block; the draws are made by its first warp).  The fixed RNG table is
// staged through a shared-memory window by the whole block, so a draw costs a shared-memory read instead of an L2
// round trip.  Lane l speculatively takes draws [p+5l, p+5l+5); a subset with a repeated index is rebuilt
// sequentially by lane 0 with OpenCV's redraw rule and the stream position re-aligned.
constexpr int RNG_WINDOW = 4096;  // words; one warp step consumes 160 + the redraws of one subset

__device__ void pnp_subsets_block(const PnpArgs& a, const uint32_t* __restrict__ rng, int n) {
  __shared__ uint32_t s_rng[RNG_WINDOW];
  __shared__ int s_ps[2];
  const int iters = max(a.iterations, 1);
  const int lane = threadIdx.x;
  if (n < 5) return;
  if (n == 5) {
    if (lane < 5) a.subsets[lane] = lane;
    return;
  }
  int p = 0, s = 0;  // stream position, subsets done (block-uniform at the top of the loop)
  while (s < iters && p + 5 <= RNG_TABLE_SIZE) {
    const int w0 = p;
    for (int i = threadIdx.x; i < RNG_WINDOW; i += blockDim.x) s_rng[i] = rng[min(w0 + i, RNG_TABLE_SIZE - 1)];
    __syncthreads();
    if (threadIdx.x < 32) {
      // one step needs at most 160 draws for the speculation and a handful for a sequential rebuild
      while (s < iters && p + 256 <= w0 + RNG_WINDOW) {
        int idx[5];
        bool dup = false;
        const int q = p + 5 * lane;
        const bool in_range = (s + lane < iters) && (q + 5 <= RNG_TABLE_SIZE);
        if (in_range) {
#pragma unroll
          for (int k = 0; k < 5; k++) idx[k] = (int)(s_rng[q + k - w0] % (unsigned)n);
#pragma unroll
          for (int k = 1; k < 5; k++)
#pragma unroll
            for (int j = 0; j < k; j++) dup |= (idx[k] == idx[j]);
        }
        const unsigned bad = __ballot_sync(0xffffffffu, in_range && dup);
        const unsigned act = __ballot_sync(0xffffffffu, in_range);
        const int n_act = __popc(act);
        if (n_act == 0) {  // RNG table exhausted (cannot happen for iterations * 8 < 2^18)
          s = iters;
          break;
        }
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
                const int pp = min(p, RNG_TABLE_SIZE - 1);
                const uint32_t rv = (pp - w0 < RNG_WINDOW) ? s_rng[pp - w0] : rng[pp];
                v = (int)(rv % (unsigned)n);
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
      if (lane == 0) {
        s_ps[0] = p;
        s_ps[1] = s;
      }
    }
    __syncthreads();
    p = s_ps[0];
    s = s_ps[1];
    __syncthreads();
  }
}

// Gathers the correspondences (f32 copies, as solvePnPRansac makes them), rebuilds the subset stream (the last block's
// first warp) and resets the RANSAC bookkeeping the chunk kernels carry.
__global__ void __launch_bounds__(256) k_pnp_prepare(const __grid_constant__ PnpArgs a,
                                                     const uint32_t* __restrict__ rng) {
  const int n = pnp_n(a);
  if (blockIdx.x == gridDim.x - 1) {
    pnp_subsets_block(a, rng, n);
    if (threadIdx.x == 32) {
      PnpState* st = a.state;
      st->niters = (n == 5) ? 1 : max(a.iterations, 1);
      st->best = 0;
      st->best_h = -1;
      st->iter = 0;
      st->ticket = 0;
      st->done = n < 5 ? 1 : 0;
      a.best[0] = -1;
      a.best[1] = 0;
      *a.hyps = 0;
    }
    return;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (gridDim.x - 1) * blockDim.x) {
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

// One EPnP hypothesis (PnPRansacCallback::runKernel) on a group of NT = 32 NW threads (NW warps).  The 5
// correspondences arrive as f32, image points are normalised with undistortPoints (f32 round trip), the pose goes
// R -> rvec (Rodrigues) and is stored together with the rotation matrix projectPoints rebuilds from rvec.  Inside the
// group: the small set-up algebra runs on the group's first warp (every lane computes the same values), M^T M and the
// 12 x 12 eigen-decomposition (jacobi_eigh_rr) are spread over all NT threads, lanes 0..5 / 6..11 of the first warp
// build L / rho, and lanes 0..2 each refine one of the three beta candidates and its pose.  Every value is produced
// by the same IEEE operations in the same order as the one-thread restatement in the oracle, so the split changes
// latency, not bits.
struct HypSmem {
  double S[2 * 144], Vr[2 * 144], ut[144], w[12], l[60], rho[6], al[5][4], us[10], rot[24], model[15];
  BetaSmem beta;
  unsigned rflag[2];
  int cnt[8];
};

template <int NT>
__device__ __forceinline__ void group_sync() {
  if (NT == 32) __syncwarp();
  else __syncthreads();
}

template <int NT>
__device__ void pnp_hypothesis(const PnpArgs& a, int h, HypSmem& sm, int tid, bool prof) {
  const EpnpCam cam{a.K[0], a.K[1], a.K[2], a.K[3]};
  double pws[15], us[10], cws[4][3];
  if (tid < 32) {
    const double ifx = 1. / a.K[0], ify = 1. / a.K[1];
    for (int k = 0; k < 5; k++) {
      const int i = a.subsets[(size_t)h * 5 + k];
      for (int cc = 0; cc < 3; cc++) pws[3 * k + cc] = (double)a.Xf[3 * i + cc];
      const double xn = (double)(float)(((double)a.xs[2 * i] - a.K[2]) * ifx);
      const double yn = (double)(float)(((double)a.xs[2 * i + 1] - a.K[3]) * ify);
      us[2 * k] = xn * cam.fu + cam.uc;
      us[2 * k + 1] = yn * cam.fv + cam.vc;
    }
    if (prof && tid == 0) pnp_prof(a, 1);
    double ci[9], al[5][4];
    epnp_small_setup(pws, 5, cws, ci, al);  // every lane computes the same values
    if (tid == 0) {
      for (int i = 0; i < 5; i++)
        for (int k = 0; k < 4; k++) sm.al[i][k] = al[i][k];
      for (int i = 0; i < 10; i++) sm.us[i] = us[i];
    }
    if (prof && tid == 0) pnp_prof(a, 2);
  }
  group_sync<NT>();
  // M^T M, each entry summed over the points in index order
  for (int e = tid; e < 144; e += NT) {
    const int ra = e / 12, rb = e - 12 * ra;
    double acc = 0;
    for (int i = 0; i < 5; i++) {
      double m1a, m2a, m1b, m2b;
      epnp_m_elem(sm.al[i], sm.us[2 * i], sm.us[2 * i + 1], cam, ra, m1a, m2a);
      epnp_m_elem(sm.al[i], sm.us[2 * i], sm.us[2 * i + 1], cam, rb, m1b, m2b);
      acc += m1a * m1b + m2a * m2b;
    }
    sm.S[e] = acc;
  }
  group_sync<NT>();
  if (prof && tid == 0) pnp_prof(a, 3);
  jacobi_eigh_rr<12, NT>(sm.S, sm.Vr, sm.w, sm.ut, sm.rot, sm.rflag, tid, prof ? a.prof + 10 : nullptr);
  if (prof && tid == 0) pnp_prof(a, 4);
  if (tid < 32) {
    const int lane = tid;
    if (lane < 6) epnp_L_row(sm.ut, lane, sm.l + 10 * lane);
    else if (lane < 12) sm.rho[lane - 6] = epnp_rho_entry(cws, lane - 6);
    __syncwarp();
    double err = 0, Rw[9], tw[3], betas[4];
    epnp_betas_warp(sm.l, sm.rho, sm.beta, lane, betas, prof ? a.prof + 24 : nullptr);
    if (lane < 3) {
      if (prof && lane == 2) pnp_prof(a, 5);
      err = epnp_candidate(pws, us, 5, sm.al, betas, sm.ut, cam, Rw, tw);
      if (prof && lane == 2) pnp_prof(a, 6);
    }
    // N = 1; if (rep[2] < rep[1]) N = 2; if (rep[3] < rep[N]) N = 3;
    const double e1 = __shfl_sync(0xffffffffu, err, 0), e2 = __shfl_sync(0xffffffffu, err, 1),
                 e3 = __shfl_sync(0xffffffffu, err, 2);
    int N = 1;
    if (e2 < e1) N = 2;
    if (e3 < (N == 1 ? e1 : e2)) N = 3;
    if (lane == N - 1) {
      double rvec[3], R2[9];
      rodrigues_mat2vec(Rw, rvec);
      rodrigues_vec2mat(rvec, R2);
      double* m = a.hyp_model + (size_t)h * 15;
      for (int i = 0; i < 9; i++) sm.model[i] = m[i] = R2[i];
      for (int i = 0; i < 3; i++) sm.model[9 + i] = m[9 + i] = tw[i];
      for (int i = 0; i < 3; i++) sm.model[12 + i] = m[12 + i] = rvec[i];
    }
    if (prof && lane == 0) pnp_prof(a, 7);
  }
  group_sync<NT>();
}

// One chunk [lo, hi) of the RANSAC iterations.  A group of NW warps solves one hypothesis (pnp_hypothesis) and counts
// that model's inliers; a block of 4 warps holds 4 / NW groups.  The last block of the chunk to finish replays
// OpenCV's sequential bookkeeping -- `if (good > max(best, 4)) { best = good; niters = RANSACUpdateNumIters(...) }` --
// over the chunk's counts, carrying (niters, best, best_h, iter) in a.state from chunk to chunk.  Once the replay
// reaches niters the state is marked done and every later chunk kernel returns at once: hypotheses the CPU loop would
// never have drawn are not solved (the reference configuration stops after a handful -- 97 % inliers give niters = 2
// -- so the first chunk of 32 is normally the only one that runs).  Exactness is untouched: the replay sees the same
// counts in the same order.  NW = 8 (one hypothesis per block of 8 warps: one S / Vr element per thread in a Jacobi
// round, 256 threads on the scoring pass) for the small leading chunks, where the latency of ONE hypothesis is the
// latency of the stage; NW = 1 (four one-warp hypotheses per block) for the large ones, where throughput counts.
template <int NW>
struct ChunkThreads { static constexpr int value = NW == 1 ? 128 : 32 * NW; };
constexpr int CHUNK_NW = 8;

template <int NW>
__global__ void __launch_bounds__(ChunkThreads<NW>::value) k_pnp_chunk(const __grid_constant__ PnpArgs a, int lo, int hi) {
  constexpr int NT = 32 * NW, HPB = ChunkThreads<NW>::value / NT;
  __shared__ HypSmem s_hyp[HPB];
  __shared__ int s_last;
  PnpState* st = a.state;
  if (st->done) return;  // written by an earlier kernel of the stream
  const int n = pnp_n(a);
  const int iters = (n == 5) ? 1 : max(a.iterations, 1);
  const int grp = threadIdx.x / NT, tid = threadIdx.x % NT, lane = threadIdx.x & 31;
  const int h = lo + blockIdx.x * HPB + grp;
  HypSmem& sm = s_hyp[grp];
  const bool prof = a.prof && h == 0;
  if (h < min(hi, iters)) {  // uniform over the group
    if (prof && tid == 0) pnp_prof(a, 0);
    pnp_hypothesis<NT>(a, h, sm, tid, prof);
    const float thr = (float)((double)a.reproj_err * (double)a.reproj_err);
    int cnt = 0;
#pragma unroll 4
    for (int i = tid; i < n; i += NT) cnt += pnp_is_inlier(sm.model, a.Xf, a.xs, i, a.K, thr) ? 1 : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if (NW > 1) {
      if (lane == 0) sm.cnt[tid >> 5] = cnt;
      group_sync<NT>();
      if (tid == 0) {
        cnt = 0;
        for (int w = 0; w < NW; w++) cnt += sm.cnt[w];
      }
    }
    if (tid == 0) {
      a.hyp_good[h] = cnt;
      if (prof) pnp_prof(a, 8);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&st->ticket, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x >= 32) return;
  __threadfence();
  // the sequential RANSAC bookkeeping, continued over this chunk's counts by one warp
  int niters = st->niters, best = st->best, best_h = st->best_h, iter = st->iter;
  if (n == 5) {  // count == modelPoints: a single kernel call, every point an inlier
    best_h = 0;
    best = 5;
    iter = 1;
    niters = 1;
  }
  for (;;) {
    const int limit = min(niters, hi);
    if (iter >= limit) break;
    const int idx = iter + lane;
    const int g = idx < limit ? __ldcg(a.hyp_good + idx) : -1;
    const unsigned mask = __ballot_sync(0xffffffffu, g > max(best, 4));
    if (!mask) {
      iter = min(iter + 32, limit);
      continue;
    }
    const int f = __ffs(mask) - 1;
    best = __shfl_sync(0xffffffffu, g, f);
    best_h = iter + f;
    niters = ransac_update_num_iters(a.confidence, (double)(n - best) / n, 5, niters);
    iter = best_h + 1;
  }
  if (lane == 0) {
    st->niters = niters;
    st->best = best;
    st->best_h = best_h;
    st->iter = iter;
    st->ticket = 0;
    if (iter >= niters) {
      st->done = 1;
      a.best[0] = best_h;
      a.best[1] = best;
      *a.hyps = iter;
    }
    if (a.prof && lo == 0) pnp_prof(a, 9);
  }
}

// inlier mask of the winning hypothesis -> ascending inlier list, then the EPnP refit on all inliers
// (solvePnPRansac's final solvePnP on f64 copies of the f32 data).  One block.
constexpr int REFIT_THREADS = 256;
constexpr int REFIT_MAX_WORDS = 2048;  // inlier-mask words: up to 65 536 correspondences

__global__ void __launch_bounds__(REFIT_THREADS) k_pnp_finalize(const __grid_constant__ PnpArgs a) {
  __shared__ double s_red[8 * 40], s_out[40];
  __shared__ double s_cws[4][3], s_ci[9], s_ut[144], s_betas[3][4], s_ccs3[3][4][3], s_R3[3][9], s_t3[3][3], s_sign3[3];
  __shared__ double s_mtm[2 * 144], s_Vr[2 * 144], s_w[12], s_l[60], s_rho[6], s_rot[24];
  __shared__ BetaSmem s_beta;
  __shared__ unsigned s_rflag[2];
  __shared__ unsigned s_mask[REFIT_MAX_WORDS];
  __shared__ int s_off[REFIT_MAX_WORDS];
  __shared__ int s_total;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n = pnp_n(a);
  const int best_h = a.best[0];
  if (a.prof && tid == 0) pnp_prof(a, 16);
  if (n < 5 || best_h < 0) {
    if (tid == 0) {
      *a.n_inliers = 0;
      for (int i = 0; i < 6; i++) a.result[i] = 0.0;
      a.result[6] = 0.0;
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
  // ascending inlier list: one ballot per 32 consecutive points, an exclusive scan of the popcounts by warp 0, then
  // every warp writes its words' inliers in place
  const int words = (n + 31) >> 5;
#pragma unroll 4
  for (int wd = wid; wd < words; wd += REFIT_THREADS / 32) {
    const int i = 32 * wd + lane;
    const unsigned b = __ballot_sync(0xffffffffu, i < n && pnp_is_inlier(m, a.Xf, a.xs, i, a.K, thr));
    if (lane == 0) s_mask[wd] = b;
  }
  __syncthreads();
  if (wid == 0) {
    int run = 0;
    for (int base = 0; base < words; base += 32) {
      const int wd = base + lane;
      const int c = wd < words ? __popc(s_mask[wd]) : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (wd < words) s_off[wd] = run + incl - c;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_total = run;
  }
  __syncthreads();
  for (int wd = wid; wd < words; wd += REFIT_THREADS / 32) {
    const unsigned b = s_mask[wd];
    if ((b >> lane) & 1u) a.inliers[s_off[wd] + __popc(b & ((1u << lane) - 1u))] = 32 * wd + lane;
  }
  const int ni = s_total;
  if (tid == 0) *a.n_inliers = ni;
  __syncthreads();  // the list is read back below by other threads
  if (a.prof && tid == 0) pnp_prof(a, 17);
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
  // (the covariance is summed from centred points: it must stay positive semi-definite to rounding -- a planar scene
  // has a zero eigenvalue there, and a one-pass S2 - n c0 c0^T would replace it by cancellation noise)
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
  if (a.prof && tid == 0) pnp_prof(a, 18);
  // M^T M.  A point contributes rows M1 = [a_k fu, 0, a_k (uc - u)]_k, M2 = [0, a_k fv, a_k (vc - v)]_k, so the 12 x 12
  // sum is (a a^T) (x) G(u, v): four weighted sums of the ten products a_k a_k' -- weights 1, uc - u, vc - v,
  // (uc - u)^2 + (vc - v)^2 -- give every entry.  Each thread accumulates its points in 40 registers, one block
  // reduction adds them (a fixed-shape tree; the refit is compared with the oracle to 1e-9, not bit for bit).
  {
    double acc[40];
#pragma unroll
    for (int i = 0; i < 40; i++) acc[i] = 0;
    for (int k = tid; k < ni; k += blockDim.x) {
      double al[4];
      point(k, pw, u, v);
      epnp_alphas(pw, s_cws, s_ci, al);
      const double du = cam.uc - u, dv = cam.vc - v, dd = du * du + dv * dv;
      int e = 0;
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int q = p; q < 4; q++) {
          const double aa = al[p] * al[q];
          acc[e] += aa;
          acc[10 + e] += aa * du;
          acc[20 + e] += aa * dv;
          acc[30 + e] += aa * dd;
          e++;
        }
    }
    block_reduce_sum<40>(acc, s_red, s_out);
    if (tid < 144) {
      const int ra = tid / 12, rb = tid - 12 * ra;
      const int ka = ra / 3, ta = ra - 3 * ka, kb = rb / 3, tb = rb - 3 * kb;
      const int p = min(ka, kb), q = max(ka, kb);
      const int e = p * 4 - (p * (p - 1)) / 2 + (q - p);  // index of (p, q), p <= q, in the row-major upper triangle
      const double P = s_out[e], Q = s_out[10 + e], R = s_out[20 + e], T = s_out[30 + e];
      double val = 0.0;
      if (ta == 0 && tb == 0) val = (cam.fu * cam.fu) * P;
      else if (ta == 1 && tb == 1) val = (cam.fv * cam.fv) * P;
      else if (ta == 2 && tb == 2) val = T;
      else if ((ta == 0 && tb == 2) || (ta == 2 && tb == 0)) val = cam.fu * Q;
      else if ((ta == 1 && tb == 2) || (ta == 2 && tb == 1)) val = cam.fv * R;
      s_mtm[tid] = val;
    }
    __syncthreads();
  }
  if (a.prof && tid == 0) pnp_prof(a, 19);
  // eigenvectors of M^T M over the whole block, L / rho on lanes 0..11, one beta candidate per lane 0..2
  jacobi_eigh_rr<12, REFIT_THREADS>(s_mtm, s_Vr, s_w, s_ut, s_rot, s_rflag, tid);
  if (a.prof && tid == 0) pnp_prof(a, 20);
  if (tid < 32) {
    if (tid < 6) epnp_L_row(s_ut, tid, s_l + 10 * tid);
    else if (tid < 12) s_rho[tid - 6] = epnp_rho_entry(s_cws, tid - 6);
    __syncwarp();
    double betas[4];
    epnp_betas_warp(s_l, s_rho, s_beta, tid, betas);
    if (tid < 3)
      for (int i = 0; i < 4; i++) s_betas[tid][i] = betas[i];
  }
  __syncthreads();
  if (a.prof && tid == 0) pnp_prof(a, 21);
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
  // centroids of the camera-frame points and ABt = sum (pc - pc0)(pw - pw0)^T = sum pc pw^T - n pc0 pw0^T for the three
  // candidates from one pass over the inliers (36 sums)
  double pc0[3][3];
  {
    double acc[36];
    for (int i = 0; i < 36; i++) acc[i] = 0;
    for (int k = tid; k < ni; k += blockDim.x) {
      double al[4], pc[3];
      point(k, pw, u, v);
      epnp_alphas(pw, s_cws, s_ci, al);
#pragma unroll
      for (int w = 0; w < 3; w++) {
        epnp_pc(al, s_ccs3[w], s_sign3[w], pc);
        for (int j = 0; j < 3; j++) {
          acc[3 * w + j] += pc[j];
          acc[9 + 9 * w + 3 * j] += pc[j] * pw[0];
          acc[9 + 9 * w + 3 * j + 1] += pc[j] * pw[1];
          acc[9 + 9 * w + 3 * j + 2] += pc[j] * pw[2];
        }
      }
    }
    block_reduce_sum<36>(acc, s_red, s_out);
    for (int w = 0; w < 3; w++)
      for (int cc = 0; cc < 3; cc++) pc0[w][cc] = s_out[3 * w + cc] / ni;
  }
  if (tid < 3) {
    double abt[9];
    for (int j = 0; j < 3; j++)
      for (int q = 0; q < 3; q++) abt[3 * j + q] = s_out[9 + 9 * tid + 3 * j + q] - ni * pc0[tid][j] * c0[q];
    epnp_rt_from_abt(abt, pc0[tid], c0, s_R3[tid], s_t3[tid]);
  }
  __syncthreads();
  if (a.prof && tid == 0) pnp_prof(a, 22);
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
    if (a.prof) pnp_prof(a, 23);
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

static inline size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

size_t pnp_scratch_bytes(int n, int iterations) {
  const size_t it = (size_t)std::max(iterations, 1);
  return al256(it * 5 * sizeof(int32_t)) + al256(it * 15 * sizeof(double)) + al256(it * sizeof(int)) +
         al256((size_t)n * 2 * sizeof(float)) + al256((size_t)n * 3 * sizeof(float)) + al256(sizeof(PnpState)) +
         al256(2 * sizeof(int)) + 256;
}

void pnp_bind_scratch(PnpArgs& a, uint8_t* base, int n, int iterations) {
  const size_t it = (size_t)std::max(iterations, 1);
  uint8_t* b = (uint8_t*)al256((size_t)base);
  auto take = [&](size_t bytes) {
    uint8_t* r = b;
    b += al256(bytes);
    return r;
  };
  a.subsets = (int32_t*)take(it * 5 * sizeof(int32_t));
  a.hyp_model = (double*)take(it * 15 * sizeof(double));
  a.hyp_good = (int*)take(it * sizeof(int));
  a.xs = (float*)take((size_t)n * 2 * sizeof(float));
  a.Xf = (float*)take((size_t)n * 3 * sizeof(float));
  a.state = (PnpState*)take(sizeof(PnpState));
  a.best = (int*)take(2 * sizeof(int));
}

void launch_pnp_prepare(Ctx& c, const PnpArgs& a) {
  UVO_REQUIRE(a.n <= 32 * REFIT_MAX_WORDS, "solvePnPRansac: more than 65 536 correspondences");
  const uint32_t* rng = rng_table_device(c);
  const int iters = std::max(a.iterations, 1);
  UVO_REQUIRE((size_t)iters * 8 < (size_t)RNG_TABLE_SIZE, "solvePnPRansac: iterationsCount too large for the RNG table");
  UVO_KERNEL(c, "k_pnp_prepare");
  k_pnp_prepare<<<std::min(div_up(std::max(a.n, 1), 256), 4 * c.sm_count) + 1, 256, 0, c.stream>>>(a, rng);
  UVO_LAUNCH_CHECK(c);
}

void launch_pnp_solve(Ctx& c, const PnpArgs& a) {
  const int iters = std::max(a.iterations, 1);
  // chunks of the iteration range: small first (the reference configuration stops within it), then growing
  const int bounds[4] = {32, 128, 512, iters};
  int lo = 0;
  for (int k = 0; k < 4 && lo < iters; k++) {
    const int hi = std::min(bounds[k], iters);
    if (hi <= lo) continue;
    // timed under one name per chunk: the first runs hypotheses, the later ones normally return at once
    static const char* const kChunkNames[4] = {"k_pnp_chunk[0:32]", "k_pnp_chunk[32:128]", "k_pnp_chunk[128:512]",
                                               "k_pnp_chunk[512:]"};
    UVO_KERNEL(c, kChunkNames[k]);
    if (hi <= 128) k_pnp_chunk<CHUNK_NW><<<hi - lo, ChunkThreads<CHUNK_NW>::value, 0, c.stream>>>(a, lo, hi);
    else k_pnp_chunk<1><<<div_up(hi - lo, 4), ChunkThreads<1>::value, 0, c.stream>>>(a, lo, hi);
    UVO_LAUNCH_CHECK(c);
    lo = hi;
  }
  UVO_KERNEL(c, "k_pnp_finalize");
  k_pnp_finalize<<<1, REFIT_THREADS, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
}

void launch_pnp_ransac(Ctx& c, const PnpArgs& a) {
  launch_pnp_prepare(c, a);
  launch_pnp_solve(c, a);
}

}  // namespace uvo
