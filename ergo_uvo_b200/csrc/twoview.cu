// twoview.cu -- K10a / K10b on the GPU (see twoview.cuh).  OpenCV's RANSAC / LMedS drivers are sequential loops with
// a data-dependent stopping rule; here every hypothesis the loop could reach is generated, solved and scored in
// parallel from the same sample stream, and the loop's bookkeeping is then replayed over the per-hypothesis scores
// (k_tv_scan), which yields the model, the mask and the iteration count the CPU loop would have produced.
//
//   k_tv_prepare   essential: f32 pixels -> f64 normalised coordinates
//   k_tv_subsets   getSubset stream (cv::RNG(-1) table, % count, duplicate redraw, homography checkSubset, 1000-attempt
//                  rule) for all iterations, one warp, 32 attempts speculated at a time
//   k_tv_hyp       one minimal-set solve per thread (4-point DLT / Nister 5-point, twoview_math.cuh)
//   k_tv_score     one block per (hypothesis, model): f32 errors of all points -> inlier count (RANSAC) or the
//                  count/2-th smallest error by radix select (LMedS)
//   k_tv_scan      sequential bookkeeping replay, one warp
//   k_tv_finalize  mask of the winner; homography: DLT refit on the inliers + 10 LM iterations + final mask
//   k_rp_*         recoverPose / recover_pose_homography: decomposition, per-point triangulation vote, selection
#include "twoview.cuh"

#include <algorithm>
#include <cmath>

#include "pose.cuh"
#include "twoview_math.cuh"

namespace uvo {

__device__ __forceinline__ int tv_model_points(const RobustArgs& a) { return a.kind == TV_ESSENTIAL ? 5 : 4; }
__device__ __forceinline__ int tv_models_per_hyp(const RobustArgs& a) { return a.kind == TV_ESSENTIAL ? TV_MAX_MODELS : 1; }

// ------------------------------------------------------------------------------------------------ prepare
__global__ void __launch_bounds__(256) k_tv_prepare(const __grid_constant__ RobustArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  a.q[4 * i + 0] = ((double)a.p1[2 * i] - a.K[2]) / a.K[0];
  a.q[4 * i + 1] = ((double)a.p1[2 * i + 1] - a.K[3]) / a.K[1];
  a.q[4 * i + 2] = ((double)a.p2[2 * i] - a.K[2]) / a.K[0];
  a.q[4 * i + 3] = ((double)a.p2[2 * i + 1] - a.K[3]) / a.K[1];
}

// ------------------------------------------------------------------------------------------------ subsets
template <int MP>
__device__ void tv_subsets(const RobustArgs& a, const uint32_t* __restrict__ rng) {
  const int n = a.n, iters = a.iters, lane = threadIdx.x;
  const bool check = a.kind == TV_HOMOGRAPHY;
  if (n < MP) {
    if (lane == 0) a.ctl[0] = 0;
    return;
  }
  if (n == MP) {
    if (lane < MP) a.subsets[lane] = lane;
    if (lane == 0) a.ctl[0] = 1;
    return;
  }
  auto accept = [&](const int* idx) -> bool {
    if (!check) return true;
    float ms1[8], ms2[8];
    for (int k = 0; k < 4; k++) {
      ms1[2 * k] = a.p1[2 * idx[k]];
      ms1[2 * k + 1] = a.p1[2 * idx[k] + 1];
      ms2[2 * k] = a.p2[2 * idx[k]];
      ms2[2 * k + 1] = a.p2[2 * idx[k] + 1];
    }
    return h_check_subset(ms1, ms2);
  };
  int p = 0, s = 0, rej = 0;
  bool failed = false;
  while (s < iters && !failed) {
    int idx[MP];
    bool dup = false;
    const int qd = p + MP * lane;
    const bool in_range = qd + MP <= RNG_TABLE_SIZE;
    if (in_range) {
#pragma unroll
      for (int k = 0; k < MP; k++) idx[k] = (int)(rng[qd + k] % (unsigned)n);
#pragma unroll
      for (int k = 1; k < MP; k++)
#pragma unroll
        for (int j = 0; j < k; j++) dup |= (idx[k] == idx[j]);
    }
    const unsigned bad = __ballot_sync(0xffffffffu, in_range && dup);
    const unsigned act = __ballot_sync(0xffffffffu, in_range);
    const int n_act = __popc(act);
    if (n_act == 0) break;  // RNG table exhausted
    const int first_bad = bad ? (__ffs(bad) - 1) : n_act;
    const bool ok = in_range && lane < first_bad && accept(idx);
    const unsigned accm = __ballot_sync(0xffffffffu, ok);
    if (accm == 0) {
      rej += first_bad;
      if (rej >= 1000) failed = true;
    } else {
      const int f0 = __ffs(accm) - 1;
      if (rej + f0 >= 1000) {
        failed = true;
      } else {
        const int slot = s + __popc(accm & ((1u << lane) - 1));
        if (ok && slot < iters)
          for (int k = 0; k < MP; k++) a.subsets[(size_t)slot * MP + k] = idx[k];
        s += __popc(accm);
        rej = first_bad - 1 - (31 - __clz(accm));
      }
    }
    p += MP * first_bad;
    if (bad && !failed && s < iters) {  // one attempt with OpenCV's redraw rule, sequentially
      int acc2 = 0;
      if (lane == 0) {
        int sub[MP];
        for (int i = 0; i < MP; i++) {
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
        acc2 = accept(sub) ? 1 : 0;
        if (acc2)
          for (int k = 0; k < MP; k++) a.subsets[(size_t)s * MP + k] = sub[k];
      }
      p = __shfl_sync(0xffffffffu, p, 0);
      acc2 = __shfl_sync(0xffffffffu, acc2, 0);
      if (acc2) {
        s++;
        rej = 0;
      } else if (++rej >= 1000) {
        failed = true;
      }
    }
  }
  if (lane == 0) a.ctl[0] = min(s, iters);
}

__global__ void __launch_bounds__(32) k_tv_subsets(const __grid_constant__ RobustArgs a, const uint32_t* __restrict__ rng) {
  if (a.kind == TV_ESSENTIAL)
    tv_subsets<5>(a, rng);
  else
    tv_subsets<4>(a, rng);
}

// ------------------------------------------------------------------------------------------------ hypotheses
__global__ void __launch_bounds__(32) k_tv_hyp(const __grid_constant__ RobustArgs a) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= a.ctl[0]) return;
  if (a.kind == TV_HOMOGRAPHY) {
    float M[8], m[8];
    for (int k = 0; k < 4; k++) {
      const int i = a.subsets[(size_t)h * 4 + k];
      M[2 * k] = a.p1[2 * i];
      M[2 * k + 1] = a.p1[2 * i + 1];
      m[2 * k] = a.p2[2 * i];
      m[2 * k + 1] = a.p2[2 * i + 1];
    }
    double H[9];
    const bool ok = homography_from4(M, m, H);
    a.n_models[h] = ok ? 1 : 0;
    if (ok)
      for (int k = 0; k < 9; k++) a.models[(size_t)h * 9 + k] = H[k];
  }
}

// Nister 5-point hypotheses: one hypothesis per group of TVE_GL lanes.  Lane 0 builds the constraint polynomial
// (null space of the 5 x 9 system, the ten cubic constraints, Gauss-Jordan, det B(z)); the ten roots are then iterated
// one per lane, and each real root is turned into its essential matrix on its own lane; models are stored in root
// order, as the one-thread loop did.
constexpr int TVE_GL = 16;
constexpr int TVE_PER_BLOCK = 4;

__global__ void __launch_bounds__(TVE_GL* TVE_PER_BLOCK) k_tv_hyp_e(const __grid_constant__ RobustArgs a) {
  __shared__ double s_EE[TVE_PER_BLOCK][36], s_B[TVE_PER_BLOCK][3][13], s_c11[TVE_PER_BLOCK][11];
  __shared__ Cx s_roots[TVE_PER_BLOCK][10];
  __shared__ int s_ok[TVE_PER_BLOCK];
  const int g = threadIdx.x / TVE_GL, gl = threadIdx.x % TVE_GL;
  const int h = blockIdx.x * TVE_PER_BLOCK + g;
  if (h >= a.ctl[0]) return;  // whole group
  const int lane = threadIdx.x & 31, gbase = lane & ~(TVE_GL - 1);
  const unsigned gmask = ((1u << TVE_GL) - 1u) << gbase;
  if (gl == 0) {
    double q1[10], q2[10];
    for (int k = 0; k < 5; k++) {
      const int i = a.subsets[(size_t)h * 5 + k];
      q1[2 * k] = a.q[4 * i];
      q1[2 * k + 1] = a.q[4 * i + 1];
      q2[2 * k] = a.q[4 * i + 2];
      q2[2 * k + 1] = a.q[4 * i + 3];
    }
    s_ok[g] = five_point_poly(q1, q2, s_EE[g], s_B[g], s_c11[g]) ? 1 : 0;
  }
  __syncwarp(gmask);
  if (!s_ok[g]) {
    if (gl == 0) a.n_models[h] = 0;
    return;
  }
  const int nr = solve_poly_dk_group(s_c11[g], 10, s_roots[g], gl, gmask);
  double Ev[9];
  const bool have = gl < nr && five_point_model(s_roots[g][gl], s_B[g], s_EE[g], Ev);
  const unsigned bal = (__ballot_sync(gmask, have) >> gbase) & ((1u << TVE_GL) - 1u);
  if (have) {
    const int slot = __popc(bal & ((1u << gl) - 1u));
    double* out = a.models + ((size_t)h * TV_MAX_MODELS + slot) * 9;
    for (int e = 0; e < 9; e++) out[e] = Ev[e];
  }
  if (gl == 0) a.n_models[h] = __popc(bal);
}

// ------------------------------------------------------------------------------------------------ scoring
struct TvModel {
  float Hf[8];
  double E[9];
};
__device__ __forceinline__ void tv_load_model(const RobustArgs& a, const double* m, TvModel& md) {
  if (a.kind == TV_HOMOGRAPHY) {
    for (int k = 0; k < 8; k++) md.Hf[k] = (float)m[k];
  } else {
    for (int k = 0; k < 9; k++) md.E[k] = m[k];
  }
}
__device__ __forceinline__ float tv_error(const RobustArgs& a, const TvModel& md, int i) {
  if (a.kind == TV_HOMOGRAPHY) return h_error_f32(md.Hf, a.p1[2 * i], a.p1[2 * i + 1], a.p2[2 * i], a.p2[2 * i + 1]);
  return sampson_f32(md.E, a.q[4 * i], a.q[4 * i + 1], a.q[4 * i + 2], a.q[4 * i + 3]);
}
// the squared f32 threshold findInliers compares against
__device__ __forceinline__ float tv_thr2(const RobustArgs& a) { return (float)(a.threshold * a.threshold); }

__global__ void __launch_bounds__(256) k_tv_score(const __grid_constant__ RobustArgs a) {
  __shared__ int s_cnt[8];
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_prefix;
  __shared__ int s_k;
  const int h = blockIdx.x, m = blockIdx.y, mph = tv_models_per_hyp(a);
  if (h >= a.ctl[0] || m >= a.n_models[h]) return;
  const int n = a.n;
  TvModel md;
  tv_load_model(a, a.models + ((size_t)h * mph + m) * 9, md);
  if (a.method == TV_RANSAC) {
    const float thr = tv_thr2(a);
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) cnt += tv_error(a, md, i) <= thr ? 1 : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; w++) tot += s_cnt[w];
      a.score[(size_t)h * mph + m] = (float)tot;
    }
    return;
  }
  // LMedS: element count/2 of the ascending f32 errors (std::nth_element on the int views: errors are >= 0 or NaN,
  // whose positive bit pattern sorts last), by an 8-bit radix select that recomputes the errors in every pass
  if (threadIdx.x == 0) {
    s_prefix = 0;
    s_k = n / 2;
  }
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 24 - 8 * pass;
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned b = __float_as_uint(tv_error(a, md, i));
      if ((b & pmask) == prefix) atomicAdd(&s_hist[(b >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int k = s_k, bin = 0;
      for (; bin < 255; bin++) {
        if (k < (int)s_hist[bin]) break;
        k -= (int)s_hist[bin];
      }
      s_k = k;
      s_prefix = prefix | ((unsigned)bin << shift);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) a.score[(size_t)h * mph + m] = __uint_as_float(s_prefix);
}

// ------------------------------------------------------------------------------------------------ bookkeeping replay
__global__ void __launch_bounds__(32) k_tv_scan(const __grid_constant__ RobustArgs a) {
  const int lane = threadIdx.x, n = a.n, mp = tv_model_points(a), mph = tv_models_per_hyp(a);
  const int nsub = a.ctl[0];
  int best_h = -1, best_m = -1, best_good = 0, run = 0;
  double mask_thr = 0;
  if (n == mp) {  // count == modelPoints: a single kernel call on all points, every point an inlier
    if (nsub > 0 && a.n_models[0] > 0) {
      best_h = 0, best_m = 0, best_good = n;
    }
    run = 1;
    mask_thr = INFINITY;
  } else if (a.method == TV_RANSAC) {
    int niters = max(a.iters, 1);
    const int total = nsub * mph;
    int e = 0, cur = -1;  // cur: the iteration whose remaining models are still examined after niters shrank below it
    while (e < total && ((e / mph) < niters || (e / mph) == cur)) {
      const int idx = e + lane;
      const int hh = idx / mph, mm = idx - hh * mph;
      const bool valid = idx < total && (hh < niters || hh == cur) && mm < a.n_models[hh];
      const int g = valid ? (int)a.score[idx] : -1;
      const unsigned mk = __ballot_sync(0xffffffffu, g > max(best_good, mp - 1));
      if (!mk) {
        e += 32;
        continue;
      }
      const int f = __ffs(mk) - 1;
      best_good = __shfl_sync(0xffffffffu, g, f);
      best_h = (e + f) / mph;
      best_m = (e + f) - best_h * mph;
      cur = best_h;
      niters = tv_update_num_iters(a.confidence, (double)(n - best_good) / n, mp, niters);
      e = e + f + 1;
    }
    run = min(max(niters, best_h + 1), nsub);  // the loop leaves when iter >= niters, after finishing iteration best_h
    mask_thr = (double)tv_thr2(a);
  } else {
    // LMedS: the first strictly smaller median wins -> minimum of (median, entry index)
    unsigned long long key = ~0ull;
    for (int idx = lane; idx < nsub * mph; idx += 32) {
      const int hh = idx / mph, mm = idx - hh * mph;
      if (mm >= a.n_models[hh]) continue;
      const float med = a.score[idx];
      if (!(med == med)) continue;  // NaN median never compares smaller
      const unsigned long long k2 = ((unsigned long long)__float_as_uint(med) << 32) | (unsigned)idx;
      key = k2 < key ? k2 : key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long x = __shfl_xor_sync(0xffffffffu, key, o);
      key = x < key ? x : key;
    }
    run = nsub;
    if (key != ~0ull) {
      const int idx = (int)(unsigned)key;
      best_h = idx / mph;
      best_m = idx - best_h * mph;
      const double min_median = (double)__uint_as_float((unsigned)(key >> 32));
      double sigma = 2.5 * 1.4826 * (1 + 5. / (n - mp)) * sqrt(min_median);
      sigma = fmax(sigma, 0.001);
      mask_thr = (double)(float)(sigma * sigma);
    }
  }
  if (lane == 0) {
    a.ctl[1] = best_h;
    a.ctl[2] = best_m;
    a.ctl[3] = best_good;
    a.ctl[4] = run;
    a.best[9] = mask_thr;
    if (best_h >= 0)
      for (int k = 0; k < 9; k++) a.best[k] = a.models[((size_t)best_h * mph + best_m) * 9 + k];
  }
}

// ------------------------------------------------------------------------------------------------ finalize
constexpr int FIN_THREADS = 256;

// block reduction of K doubles (sum) into s_out[0..K); every thread passes its partials in v[]
template <int K>
__device__ void block_sum(double* v, double* s_red /* 8 x K */, double* s_out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll 1
  for (int k = 0; k < K; k++) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[wid * K + k] = x;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double t = 0;
    for (int w = 0; w < FIN_THREADS / 32; w++) t += s_red[w * K + k];
    s_out[k] = t;
  }
  __syncthreads();
}

// HomographyRefineCallback::compute over the inliers: S = sum r^2, max |r|, and (with_j) J^T J (upper, 36) and J^T r
template <bool WITH_J>
__device__ void h_refine_eval(const RobustArgs& a, const double* h, double* s_red, double* s_out, double* s_rmax) {
  double acc[46];
  for (int k = 0; k < 46; k++) acc[k] = 0;
  double rmax = 0;
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    if (!a.mask[i]) continue;
    const double Mx = a.p1[2 * i], My = a.p1[2 * i + 1];
    double ww = h[6] * Mx + h[7] * My + 1.;
    ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
    const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww, yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
    const double r0 = xi - (double)a.p2[2 * i], r1 = yi - (double)a.p2[2 * i + 1];
    acc[45] += r0 * r0 + r1 * r1;
    rmax = fmax(rmax, fmax(fabs(r0), fabs(r1)));
    if (WITH_J) {
      const double J0[8] = {Mx * ww, My * ww, ww, 0, 0, 0, -Mx * ww * xi, -My * ww * xi};
      const double J1[8] = {0, 0, 0, Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi};
      int o = 0;
      for (int p = 0; p < 8; p++)
        for (int q = p; q < 8; q++) acc[o++] += J0[p] * J0[q] + J1[p] * J1[q];
      for (int p = 0; p < 8; p++) acc[36 + p] += J0[p] * r0 + J1[p] * r1;
    }
  }
  // max |r| through the same machinery: every warp's maximum, then the block's
#pragma unroll
  for (int o = 16; o; o >>= 1) rmax = fmax(rmax, __shfl_down_sync(0xffffffffu, rmax, o));
  __shared__ double s_wmax[FIN_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s_wmax[threadIdx.x >> 5] = rmax;
  block_sum<46>(acc, s_red, s_out);  // contains the barriers that also publish s_wmax
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < FIN_THREADS / 32; w++) t = fmax(t, s_wmax[w]);
    *s_rmax = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FIN_THREADS) k_tv_finalize(const __grid_constant__ RobustArgs a) {
  __shared__ double s_red[8 * 46], s_out[46], s_h[9], s_xd[8], s_rmax;
  __shared__ int s_cnt[8], s_flag;
  const int n = a.n, mp = tv_model_points(a);
  const int best_h = a.ctl[1];
  const int tid = threadIdx.x;
  auto fail = [&]() {
    for (int i = tid; i < n; i += blockDim.x) a.mask[i] = 0;
    if (tid == 0) {
      for (int k = 0; k < 9; k++) a.model_out[k] = 0;
      *a.n_inliers = 0;
      a.ctl[5] = 0;
    }
  };
  // findHomography(method = 0) (and, in OpenCV, npoints == 4): no sampling -- every point is in the set the kernel
  // and the refinement run on (fundam.cpp: `tempMask = Mat::ones(...)`, `result = cb->runKernel(src, dst, H) > 0`)
  const bool lsq = a.method == TV_LSQ;
  if (best_h < 0 && !lsq) {
    fail();
    return;
  }
  TvModel md;
  float thr = 0.f;
  if (!lsq) {
    tv_load_model(a, a.best, md);
    thr = (float)a.best[9];
  }
  // mask of the winning hypothesis
  int cnt = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const bool in = lsq || (n == mp) || tv_error(a, md, i) <= thr;
    a.mask[i] = in ? 1 : 0;
    cnt += in ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  if ((tid & 31) == 0) s_cnt[tid >> 5] = cnt;
  __syncthreads();
  int total = 0;
  for (int w = 0; w < FIN_THREADS / 32; w++) total += s_cnt[w];
  __syncthreads();
  // RANSAC: result = maxGoodCount > 0; LMedS: result = inliers >= modelPoints
  if (a.method == TV_LMEDS && n != mp && total < mp) {
    fail();
    return;
  }
  if (!lsq && (a.kind == TV_ESSENTIAL || n <= 4 || total == 0)) {
    if (tid == 0) {
      for (int k = 0; k < 9; k++) a.model_out[k] = a.best[k];
      *a.n_inliers = total;
      a.ctl[5] = 1;
    }
    return;
  }
  // ---- homography: runKernel on the inliers, then LMSolver(HomographyRefineCallback, 10) ----
  HNorm nm{};
  {
    double v[4] = {0, 0, 0, 0};
    for (int i = tid; i < n; i += blockDim.x)
      if (a.mask[i]) {
        v[0] += a.p2[2 * i];
        v[1] += a.p2[2 * i + 1];
        v[2] += a.p1[2 * i];
        v[3] += a.p1[2 * i + 1];
      }
    block_sum<4>(v, s_red, s_out);
    nm.cmx = s_out[0] / total, nm.cmy = s_out[1] / total, nm.cMx = s_out[2] / total, nm.cMy = s_out[3] / total;
    __syncthreads();
    double u[4] = {0, 0, 0, 0};
    for (int i = tid; i < n; i += blockDim.x)
      if (a.mask[i]) {
        u[0] += fabs((double)a.p2[2 * i] - nm.cmx);
        u[1] += fabs((double)a.p2[2 * i + 1] - nm.cmy);
        u[2] += fabs((double)a.p1[2 * i] - nm.cMx);
        u[3] += fabs((double)a.p1[2 * i + 1] - nm.cMy);
      }
    block_sum<4>(u, s_red, s_out);
    nm.smx = s_out[0], nm.smy = s_out[1], nm.sMx = s_out[2], nm.sMy = s_out[3];
    __syncthreads();
  }
  const bool degenerate = fabs(nm.smx) < DBL_EPSILON || fabs(nm.smy) < DBL_EPSILON || fabs(nm.sMx) < DBL_EPSILON ||
                          fabs(nm.sMy) < DBL_EPSILON;
  if (!degenerate) {
    nm.smx = total / nm.smx, nm.smy = total / nm.smy, nm.sMx = total / nm.sMx, nm.sMy = total / nm.sMy;
    double ltl[45];
    for (int k = 0; k < 45; k++) ltl[k] = 0;
    for (int i = tid; i < n; i += blockDim.x)
      if (a.mask[i]) dlt_accumulate(nm, a.p1[2 * i], a.p1[2 * i + 1], a.p2[2 * i], a.p2[2 * i + 1], ltl);
    block_sum<45>(ltl, s_red, s_out);
    if (tid == 0) dlt_solve(nm, s_out, s_h);
  } else if (lsq) {  // runKernel returned 0 and there is no hypothesis to keep: findHomography returns an empty H
    fail();
    return;
  } else if (tid == 0) {
    for (int k = 0; k < 9; k++) s_h[k] = a.best[k];  // runKernel returned 0: H keeps the hypothesis value
  }
  __syncthreads();
  if (lsq && n <= 4) {  // exactly determined: no refinement (`npoints > 4` guards it), every point an inlier
    if (tid == 0) {
      for (int k = 0; k < 9; k++) a.model_out[k] = s_h[k];
      *a.n_inliers = n;
      a.ctl[4] = 1;
      a.ctl[5] = 1;
    }
    return;
  }
  // Levenberg-Marquardt exactly as cv::LMSolverImpl::run drives it (levmarq.cpp); small algebra on thread 0
  __shared__ double s_A[64], s_v[8], s_D[8], s_d[8], s_S, s_lambda, s_lc, s_rcur, s_dmax;
  h_refine_eval<true>(a, s_h, s_red, s_out, &s_rmax);
  if (tid == 0) {
    int o = 0;
    for (int p = 0; p < 8; p++)
      for (int q = p; q < 8; q++) s_A[p * 8 + q] = s_A[q * 8 + p] = s_out[o++];
    for (int p = 0; p < 8; p++) {
      s_v[p] = s_out[36 + p];
      s_D[p] = s_A[p * 8 + p];
    }
    s_S = s_out[45];
    s_rcur = s_rmax;  // norm(r, INF) of the current parameters
    s_lambda = 1.0;
    s_lc = 0.75;
  }
  __syncthreads();
  for (int iter = 0;;) {
    if (tid == 0) {
      double Ap[64];
      for (int k = 0; k < 64; k++) Ap[k] = s_A[k];
      for (int k = 0; k < 8; k++) Ap[k * 8 + k] += s_lambda * s_D[k];
      solve_eig_sym<8>(Ap, s_v, s_d, nullptr);
      for (int k = 0; k < 8; k++) s_xd[k] = s_h[k] - s_d[k];
    }
    __syncthreads();
    h_refine_eval<false>(a, s_xd, s_red, s_out, &s_rmax);
    if (tid == 0) {
      const double Sd = s_out[45];
      double dS = 0, dv = 0, dmax = 0;
      for (int p = 0; p < 8; p++) {
        double Ad = 0;
        for (int q = 0; q < 8; q++) Ad += s_A[p * 8 + q] * s_d[q];
        dS += s_d[p] * (-Ad + 2 * s_v[p]);
        dv += s_d[p] * s_v[p];
        dmax = fmax(dmax, fabs(s_d[p]));
      }
      const double S = s_S;
      const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
      if (R > 0.75) {
        s_lambda *= 0.5;
        if (s_lambda < s_lc) s_lambda = 0;
      } else if (R < 0.25) {
        double nu = (Sd - S) / (fabs(dv) > DBL_EPSILON ? dv : 1) + 2;
        nu = fmin(fmax(nu, 2.), 10.);
        if (s_lambda == 0) {
          double x8[8], dinv[8], zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          solve_eig_sym<8>(s_A, zero, x8, dinv);  // diag(invert(A, DECOMP_EIG))
          double maxval = DBL_EPSILON;
          for (int k = 0; k < 8; k++) maxval = fmax(maxval, fabs(dinv[k]));
          s_lambda = s_lc = 1. / maxval;
          nu *= 0.5;
        }
        s_lambda *= nu;
      }
      s_flag = Sd < S ? 1 : 0;
      if (s_flag) {
        s_S = Sd;
        for (int k = 0; k < 8; k++) s_h[k] = s_xd[k];
      }
      s_dmax = dmax;
    }
    __syncthreads();
    if (s_flag) {  // accepted: Jacobian, J^T J, J^T r and residual at the new parameters
      h_refine_eval<true>(a, s_h, s_red, s_out, &s_rmax);
      if (tid == 0) {
        int o = 0;
        for (int p = 0; p < 8; p++)
          for (int q = p; q < 8; q++) s_A[p * 8 + q] = s_A[q * 8 + p] = s_out[o++];
        for (int p = 0; p < 8; p++) s_v[p] = s_out[36 + p];
        s_rcur = s_rmax;
      }
      __syncthreads();
    }
    iter++;
    // proceed = iter < maxIters && norm(d, INF) >= epsx && norm(r, INF) >= epsf  (epsx = epsf = FLT_EPSILON)
    if (!(iter < 10 && s_dmax >= (double)FLT_EPSILON && s_rcur >= (double)FLT_EPSILON)) break;
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) s_h[8] = 1.0;
  __syncthreads();
  // the mask cv2 returns: reprojection error of the refined model at the caller's threshold
  TvModel fin;
  for (int k = 0; k < 8; k++) fin.Hf[k] = (float)s_h[k];
  const float t2 = tv_thr2(a);
  cnt = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    const bool in = tv_error(a, fin, i) <= t2;
    a.mask[i] = in ? 1 : 0;
    cnt += in ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  if ((tid & 31) == 0) s_cnt[tid >> 5] = cnt;
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
    for (int w = 0; w < FIN_THREADS / 32; w++) tot += s_cnt[w];
    for (int k = 0; k < 9; k++) a.model_out[k] = s_h[k];
    *a.n_inliers = tot;
    if (lsq) a.ctl[4] = 1;  // one kernel call
    a.ctl[5] = 1;
  }
}

// ------------------------------------------------------------------------------------------------ recoverPose
// candidates: [R|t] of the 4 decompositions.  Essential: projection matrices in normalised coordinates; homography:
// K [R|t] in pixels (recover_pose_homography triangulates pixel coordinates, VO_utility.cpp:595).
__global__ void k_rp_setup(const __grid_constant__ RecoverPoseArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double R[4][9], t[4][3];
  int nc = 4;
  if (a.from_homography) {
    nc = decompose_homography(a.model, a.K, R, t);
  } else {
    double R1[9], R2[9], tt[3];
    decompose_essential(a.model, R1, R2, tt);
    for (int k = 0; k < 9; k++) {
      R[0][k] = R[2][k] = R1[k];
      R[1][k] = R[3][k] = R2[k];
    }
    for (int k = 0; k < 3; k++) {
      t[0][k] = t[1][k] = tt[k];
      t[2][k] = t[3][k] = -tt[k];
    }
  }
  for (int c = 0; c < 4; c++)
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) a.cand[c * 12 + i * 4 + j] = c < nc ? R[c][i * 3 + j] : 0;
      a.cand[c * 12 + i * 4 + 3] = c < nc ? t[c][i] : 0;
    }
  a.cand[48] = nc;
  for (int c = 0; c < 4; c++) a.counts[c] = 0;
}

__global__ void __launch_bounds__(128) k_rp_vote(const __grid_constant__ RecoverPoseArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nc = (int)a.cand[48];
  int flags = 0;
  if (i < a.n) {
    double x1, y1, x2, y2;
    if (a.from_homography) {
      x1 = a.p1[2 * i], y1 = a.p1[2 * i + 1], x2 = a.p2[2 * i], y2 = a.p2[2 * i + 1];
    } else {
      x1 = ((double)a.p1[2 * i] - a.K[2]) / a.K[0], y1 = ((double)a.p1[2 * i + 1] - a.K[3]) / a.K[1];
      x2 = ((double)a.p2[2 * i] - a.K[2]) / a.K[0], y2 = ((double)a.p2[2 * i + 1] - a.K[3]) / a.K[1];
    }
    const double Kp[12] = {a.K[0], 0, a.K[2], 0, 0, a.K[1], a.K[3], 0, 0, 0, 1, 0};
    const double I0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    const double* P0 = a.from_homography ? Kp : I0;
    for (int c = 0; c < nc; c++) {
      double P1[12];
      if (a.from_homography) {  // K [R|t]
        const double* Rt = a.cand + c * 12;
        for (int j = 0; j < 4; j++) {
          P1[j] = a.K[0] * Rt[j] + a.K[2] * Rt[8 + j];
          P1[4 + j] = a.K[1] * Rt[4 + j] + a.K[3] * Rt[8 + j];
          P1[8 + j] = Rt[8 + j];
        }
      } else {
        for (int j = 0; j < 12; j++) P1[j] = a.cand[c * 12 + j];
      }
      double A[16], w[4], Vt[16];
      for (int k = 0; k < 4; k++) {
        A[k] = x1 * P0[8 + k] - P0[k];
        A[4 + k] = y1 * P0[8 + k] - P0[4 + k];
        A[8 + k] = x2 * P1[8 + k] - P1[k];
        A[12 + k] = y2 * P1[8 + k] - P1[4 + k];
      }
      jacobi_svd<4>(A, 4, 4, w, nullptr, Vt);
      const double* Q = Vt + 12;
      bool good;
      if (a.from_homography) {
        // triangulatePoints returns f32 for f32 input; convert_from_homogeneous_coords divides in f32
        const float z = __fdiv_rn((float)Q[2], (float)Q[3]);
        good = (double)z > 0 && (double)z < a.distance_thresh;
      } else {
        good = Q[2] * Q[3] > 0;
        const double X = Q[0] / Q[3], Y = Q[1] / Q[3], Z = Q[2] / Q[3], W = Q[3] / Q[3];
        good = good && Z < a.distance_thresh;
        const double z2 = P1[8] * X + P1[9] * Y + P1[10] * Z + P1[11] * W;
        good = good && z2 > 0 && z2 < a.distance_thresh;
        if (a.mask_in) good = good && a.mask_in[i] != 0;
      }
      if (good) flags |= 1 << c;
    }
    a.flags[i] = (uint8_t)flags;
  }
  for (int c = 0; c < 4; c++) {
    const unsigned b = __ballot_sync(0xffffffffu, (flags >> c) & 1);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&a.counts[c], __popc(b));
  }
}

__global__ void __launch_bounds__(256) k_rp_pick(const __grid_constant__ RecoverPoseArgs a) {
  const int nc = (int)a.cand[48];
  int k = -1, good = 0;
  if (a.from_homography) {  // first strictly larger count wins (VO_utility.cpp:607-611)
    for (int c = 0; c < nc; c++)
      if (a.counts[c] > good) {
        k = c;
        good = a.counts[c];
      }
  } else {  // recoverPose's cascade: (R1,t), (R2,t), (R1,-t), (R2,-t) with >= comparisons
    const int g1 = a.counts[0], g2 = a.counts[1], g3 = a.counts[2], g4 = a.counts[3];
    if (g1 >= g2 && g1 >= g3 && g1 >= g4) k = 0;
    else if (g2 >= g1 && g2 >= g3 && g2 >= g4) k = 1;
    else if (g3 >= g1 && g3 >= g2 && g3 >= g4) k = 2;
    else k = 3;
    good = a.counts[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (k >= 0) {
      const double* Rt = a.cand + k * 12;
      double tn = 1.0;
      if (a.from_homography) tn = sqrt(Rt[3] * Rt[3] + Rt[7] * Rt[7] + Rt[11] * Rt[11]);
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) a.Rt[i * 3 + j] = Rt[i * 4 + j];
        a.Rt[9 + i] = Rt[i * 4 + 3] / tn;
      }
    } else {
      for (int i = 0; i < 12; i++) a.Rt[i] = 0;
    }
    a.Rt[12] = good;
    a.Rt[13] = k >= 0 ? 1 : 0;
  }
  if (a.mask_out && k >= 0)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x)
      a.mask_out[i] = (a.flags[i] >> k) & 1;
}

// ------------------------------------------------------------------------------------------------ host side
static int update_num_iters_host(double p, double ep, int model_points, int max_iters) {
  p = std::min(std::max(p, 0.), 1.);
  ep = std::min(std::max(ep, 0.), 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::nearbyint(num / denom);
}

int robust_iterations(int method, int model_points, double confidence, int max_iters) {
  if (method == TV_LSQ) return 1;
  max_iters = std::max(max_iters, 1);
  if (method == TV_LMEDS) return std::max(update_num_iters_host(confidence, 0.45, model_points, max_iters), 3);
  return max_iters;
}

static size_t al256(size_t n) { return (n + 255) & ~(size_t)255; }

size_t robust_scratch_bytes(int n, int iters, int kind) {
  const int mph = kind == TV_ESSENTIAL ? TV_MAX_MODELS : 1;
  return al256((size_t)iters * 5 * sizeof(int)) + al256(16 * sizeof(int)) + al256((size_t)iters * mph * 9 * sizeof(double)) +
         al256((size_t)iters * sizeof(int)) + al256((size_t)iters * mph * sizeof(float)) +
         al256((size_t)std::max(n, 1) * 4 * sizeof(double)) + al256(16 * sizeof(double));
}

void robust_bind_scratch(RobustArgs& a, void* scratch) {
  const int mph = a.kind == TV_ESSENTIAL ? TV_MAX_MODELS : 1;
  uint8_t* p = (uint8_t*)scratch;
  a.subsets = (int*)p;
  p += al256((size_t)a.iters * 5 * sizeof(int));
  a.ctl = (int*)p;
  p += al256(16 * sizeof(int));
  a.models = (double*)p;
  p += al256((size_t)a.iters * mph * 9 * sizeof(double));
  a.n_models = (int*)p;
  p += al256((size_t)a.iters * sizeof(int));
  a.score = (float*)p;
  p += al256((size_t)a.iters * mph * sizeof(float));
  a.q = (double*)p;
  p += al256((size_t)std::max(a.n, 1) * 4 * sizeof(double));
  a.best = (double*)p;
}

void launch_robust(Ctx& c, const RobustArgs& a) {
  if (a.method == TV_LSQ) {  // no hypotheses: the kernel on all points + the refinement, both inside k_tv_finalize
    UVO_REQUIRE(a.kind == TV_HOMOGRAPHY, "method 0 exists for findHomography only");
    UVO_KERNEL(c, "k_tv_finalize");
    k_tv_finalize<<<1, FIN_THREADS, 0, c.stream>>>(a);
    UVO_LAUNCH_CHECK(c);
    return;
  }
  const uint32_t* rng = rng_table_device(c);
  const int mph = a.kind == TV_ESSENTIAL ? TV_MAX_MODELS : 1;
  if (a.kind == TV_ESSENTIAL && a.n > 0) {
    UVO_KERNEL(c, "k_tv_prepare");
    k_tv_prepare<<<div_up(a.n, 256), 256, 0, c.stream>>>(a);
    UVO_LAUNCH_CHECK(c);
  }
  UVO_KERNEL(c, "k_tv_subsets");
  k_tv_subsets<<<1, 32, 0, c.stream>>>(a, rng);
  UVO_LAUNCH_CHECK(c);
  if (a.kind == TV_ESSENTIAL) {
    UVO_KERNEL(c, "k_tv_hyp_e");
    k_tv_hyp_e<<<div_up(a.iters, TVE_PER_BLOCK), TVE_GL * TVE_PER_BLOCK, 0, c.stream>>>(a);
  } else {
    UVO_KERNEL(c, "k_tv_hyp");
    k_tv_hyp<<<div_up(a.iters, 32), 32, 0, c.stream>>>(a);
  }
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_tv_score");
  k_tv_score<<<dim3(a.iters, mph), 256, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_tv_scan");
  k_tv_scan<<<1, 32, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_tv_finalize");
  k_tv_finalize<<<1, FIN_THREADS, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
}

void launch_recover_pose(Ctx& c, const RecoverPoseArgs& a) {
  UVO_KERNEL(c, "k_rp_setup");
  k_rp_setup<<<1, 32, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
  if (a.n > 0) {
    UVO_KERNEL(c, "k_rp_vote");
    k_rp_vote<<<div_up(a.n, 128), 128, 0, c.stream>>>(a);
    UVO_LAUNCH_CHECK(c);
  }
  UVO_KERNEL(c, "k_rp_pick");
  k_rp_pick<<<std::max(1, std::min(div_up(a.n, 256), 64)), 256, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
}

}  // namespace uvo
