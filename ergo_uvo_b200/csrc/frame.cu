// frame.cu -- device-resident replay of visual_odometry_node::stereo_VO's per-frame body
// (reference uvo/include/visual_odometry.h:474-520 initialisation, :526-740 main loop).
//
// One uvo_stereo_frame call = get_image x2 -> detect_features x2 -> match_features (stereo) -> gathers ->
// match_features (prev-left-after-stereo vs curr-left) -> triangulatePoints -> extract_3Dpoints ->
// solvePnPRansac -> Rodrigues / t_prevCam_currCam -> velocity.  Nothing returns to the host between stages: every
// "if (count < MIN...) ASSUMING CONSTANT MOTION" gate of the node is evaluated on the device by zeroing the element
// count the next stage reads, and one 160-byte result record comes back per frame.
#include <cstring>
#include <deque>

#include "capi_internal.cuh"
#include "linalg.cuh"
#include "match.cuh"
#include "pose.cuh"

using namespace uvo;

namespace uvo {

struct FrameCtrl {  // device-side control block, one per in-flight frame parity
  int nq_stereo;    // query count of the stereo match (0 when gate 1 fails)
  int n_stereo;     // results_match_curr.size()
  int n_as;         // size of the *_after_stereo_match sets produced by this frame
  int nq_temporal;  // query count of the temporal match
  int n_temporal;   // results_match_prev_curr.size()
  int n_tri;        // points triangulated
  int n_3d;         // good_prevCam_points.rows
  int n_inliers, hyps;
  int overflow;
  int nL, nR;       // SURF keypoint counts (copied here: the front-end counters are reused by the next frame)
  int was_init;     // vo_initialized before this frame
};

struct FrameState {  // persists across frames (device)
  int vo_init;
  double t_prev_curr[3];
};

struct GateParams {
  int min_features, min_3d, min_inliers, capacity;
};

// after SURF: gate 1 (visual_odometry.h:556) and capacity check
__global__ void k_gate_features(const int* cL, const int* cR, FrameCtrl* ctrl, GateParams g) {
  const int nL = min(cL[1], g.capacity), nR = min(cR[1], g.capacity);
  ctrl->nL = nL;
  ctrl->nR = nR;
  ctrl->overflow = (cL[0] > g.capacity || cR[0] > g.capacity) ? 1 : 0;
  ctrl->nq_stereo = (nL >= g.min_features && nR >= g.min_features) ? nL : 0;
}

// select_desired_descriptors / select_desired_keypoints by the stereo match (visual_odometry.h:569-579), gate 2
// (:567), and the query count of the triangular match (:592)
__global__ void __launch_bounds__(256) k_gather_after_stereo(const uvo_dmatch* __restrict__ m, FrameCtrl* ctrl,
                                                             const FrameCtrl* prev_ctrl, FrameState* st,
                                                             const uvo_keypoint* __restrict__ kL,
                                                             const uvo_keypoint* __restrict__ kR,
                                                             const float* __restrict__ dL, uvo_keypoint* kL_as,
                                                             uvo_keypoint* kR_as, float* dL_as, GateParams g) {
  const int ns = ctrl->nq_stereo > 0 ? ctrl->n_stereo : 0;
  const int n_as = ns > g.min_features ? ns : 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctrl->n_stereo = ns;
    ctrl->n_as = n_as;
    // the triangular match only runs inside `if (results_match_curr.size() > MIN_NUM_FEATURES)` and once the node is
    // initialised
    const int was_init = st->vo_init;
    ctrl->was_init = was_init;
    ctrl->nq_temporal = (n_as > 0 && was_init) ? prev_ctrl->n_as : 0;
    if (n_as > 0) st->vo_init = 1;  // the pose stage runs on a side stream: the next frame must already see this
  }
  const int total = n_as * 16;  // float4 chunks of the descriptors
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int i = e >> 4, q = e & 15;
    const uvo_dmatch mm = m[i];
    reinterpret_cast<float4*>(dL_as)[e] = reinterpret_cast<const float4*>(dL)[(size_t)mm.queryIdx * 16 + q];
    if (q == 0) {
      kL_as[i] = kL[mm.queryIdx];
      kR_as[i] = kR[mm.trainIdx];
    }
  }
}

struct ResultParams {
  GateParams g;
  double dt;
};

// gates 3-5 (visual_odometry.h:626, :634, :665), Rodrigues + t_prevCam_currCam = -R^T t (:673-675), velocity (:148-159)
__global__ void k_frame_result(FrameCtrl* ctrl, FrameState* st, const double* pnp_result, const int* n_inl_dev,
                               const int* hyps_dev, uvo_stereo_result* out, ResultParams p) {
  const GateParams& g = p.g;
  uvo_stereo_result r;
  memset(&r, 0, sizeof(r));
  const int was_init = ctrl->was_init;
  r.n_left = ctrl->nL;
  r.n_right = ctrl->nR;
  r.n_stereo_matches = ctrl->n_stereo;
  r.n_temporal_matches = ctrl->nq_temporal > 0 ? ctrl->n_temporal : 0;
  const bool tri = r.n_temporal_matches > g.min_features;
  r.n_3d = tri ? ctrl->n_3d : 0;
  const bool pnp = tri && r.n_3d > g.min_3d;
  r.n_inliers = pnp ? *n_inl_dev : 0;
  r.hyps_evaluated = pnp ? *hyps_dev : 0;
  int gate = 0;
  if (ctrl->overflow) gate = -1;
  else if (ctrl->nq_stereo == 0) gate = 1;
  else if (ctrl->n_as == 0) gate = 2;
  else if (!tri) gate = 3;
  else if (!pnp) gate = 4;
  else if (r.n_inliers < g.min_inliers) gate = 5;
  if (pnp) {
    for (int i = 0; i < 3; i++) {
      r.rvec[i] = pnp_result[i];
      r.tvec[i] = pnp_result[3 + i];
    }
  }
  if (was_init) {
    r.gate = gate;
    r.valid = gate == 0 ? 1 : 0;
    if (gate == 0) {
      double R[9];
      rodrigues_vec2mat(r.rvec, R);
      for (int i = 0; i < 3; i++) st->t_prev_curr[i] = -(R[i] * r.tvec[0] + R[3 + i] * r.tvec[1] + R[6 + i] * r.tvec[2]);
    }
  } else {
    r.gate = gate == -1 ? -1 : 0;
    r.valid = 0;
  }
  r.initialised = (was_init || ctrl->n_as > 0) ? 1 : 0;
  for (int i = 0; i < 3; i++) {
    r.t_prev_curr[i] = st->t_prev_curr[i];  // stale value is re-published on gate failure (:717)
    r.velocity[i] = st->t_prev_curr[i] / p.dt;
  }
  *out = r;
}

}  // namespace uvo

static const char* kStageNames[UVO_N_STAGES] = {"h2d",        "get_image",    "surf", "match_stereo", "match_temporal",
                                                "triangulate+extract3d", "pnp_ransac", "result+d2h"};

struct uvo_stereo {
  uvo_ctx* ctx = nullptr;
  int w = 0, h = 0, cap = 0;
  uvo_camera cam[2];
  uvo_params prm;
  double R_right[9], t_right[3];
  double P_left[12], P_right[12];  // P_eye_using_left_as_world, P_using_left_as_world (visual_odometry.h:460-462)
  FrontEnd fe;
  DevBuf<uint8_t> src[2];
  size_t src_pitch = 0;
  // double-buffered after-stereo sets: [parity]
  DevBuf<uvo_keypoint> kL_as[2], kR_as[2];
  DevBuf<float> dL_as[2];
  DevBuf<FrameCtrl> ctrl;  // 2
  DevBuf<FrameState> state;
  DevBuf<uvo_dmatch> m_stereo, m_temporal;
  DevBuf<Knn2> knn_scratch;
  DevBuf<float> pts1, pts2, X4;
  DevBuf<double> good_pts, tmp_pts;
  DevBuf<int32_t> good_idx, tmp_idx;
  // pose stage (side stream), double-buffered by frame parity
  DevBuf<double> pnp_result[2];
  DevBuf<int32_t> inliers[2];
  DevBuf<int> small[2];  // [0] n_inliers [1] hyps [2..3] best
  DevBuf<uint8_t> pnp_scratch[2];
  cudaStream_t side = nullptr;
  cudaEvent_t ev_main[2] = {}, ev_side[2] = {};  // hand-over main -> side, and side done, per parity
  bool side_used[2] = {false, false};
  DevBuf<uvo_stereo_result> d_result;  // ring
  PinnedBuf<uvo_stereo_result> h_result;
  static constexpr int RING = 8;
  int parity = 0;
  long frame_no = 0;
  std::deque<std::pair<int, cudaEvent_t>> pending;  // (slot, done event)
  cudaEvent_t ev[UVO_N_STAGES + 1] = {};
  bool timing = false;
  float stage_ms[UVO_N_STAGES] = {};
  bool has_timing = false;

  ~uvo_stereo() {
    for (auto& p : pending) cudaEventDestroy(p.second);
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
      if (ev_main[i]) cudaEventDestroy(ev_main[i]);
      if (ev_side[i]) cudaEventDestroy(ev_side[i]);
    }
    if (side) cudaStreamDestroy(side);
  }
};

static void stereo_init(uvo_stereo* s, uvo_ctx* ctx, int w, int h, const uvo_camera* left, const uvo_camera* right,
                        const double R_right[9], const double t_right[3], const uvo_params* prm) {
  UVO_REQUIRE(w > 0 && h > 0 && left && right && R_right && t_right && prm, "uvo_stereo_create: bad argument");
  UVO_REQUIRE(prm->pnp_method_flag == 1, "only SOLVEPNP_EPNP (pnp_method_flag = 1) is implemented");
  UVO_REQUIRE(prm->max_features >= 64, "max_features too small");
  Ctx& c = ctx->c;
  UVO_CUDA(cudaSetDevice(c.device));
  s->ctx = ctx;
  s->w = w;
  s->h = h;
  s->cap = prm->max_features;
  s->cam[0] = *left;
  s->cam[1] = *right;
  s->prm = *prm;
  memcpy(s->R_right, R_right, sizeof(s->R_right));
  memcpy(s->t_right, t_right, sizeof(s->t_right));
  // compute_projection_matrix (VO_utility.cpp:9-15): K * [R | t]
  auto proj = [](const double K[4], const double R[9], const double t[3], double P[12]) {
    const double Km[9] = {K[0], 0, K[2], 0, K[1], K[3], 0, 0, 1};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 4; j++) {
        double acc = 0;
        for (int k = 0; k < 3; k++) acc += Km[i * 3 + k] * (j < 3 ? R[k * 3 + j] : t[k]);
        P[i * 4 + j] = acc;
      }
  };
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z[3] = {0, 0, 0};
  const double KL[4] = {left->nfx, left->nfy, left->ncx, left->ncy};
  const double KR[4] = {right->nfx, right->nfy, right->ncx, right->ncy};
  proj(KL, I, z, s->P_left);
  proj(KR, R_right, t_right, s->P_right);
  const int cap = s->cap;
  s->fe.init(w, h, 2, cap);
  s->src_pitch = ((size_t)3 * w + 15) & ~(size_t)15;
  for (int i = 0; i < 2; i++) {
    s->src[i].ensure(s->src_pitch * h);
    s->kL_as[i].ensure(cap);
    s->kR_as[i].ensure(cap);
    s->dL_as[i].ensure((size_t)cap * 64);
  }
  s->ctrl.ensure(2);
  s->state.ensure(1);
  UVO_CUDA(cudaMemsetAsync(s->ctrl.get(), 0, 2 * sizeof(FrameCtrl), c.stream));
  UVO_CUDA(cudaMemsetAsync(s->state.get(), 0, sizeof(FrameState), c.stream));
  s->m_stereo.ensure(cap);
  s->m_temporal.ensure(cap);
  s->knn_scratch.ensure((size_t)(MATCH_SPLITS + 1) * cap);
  s->pts1.ensure(2 * (size_t)cap);
  s->pts2.ensure(2 * (size_t)cap);
  s->X4.ensure(4 * (size_t)cap);
  s->good_pts.ensure(3 * (size_t)cap);
  s->tmp_pts.ensure(3 * (size_t)cap);
  s->good_idx.ensure(cap);
  s->tmp_idx.ensure(cap);
  for (int i = 0; i < 2; i++) {
    s->inliers[i].ensure(cap);
    s->pnp_result[i].ensure(8);
    s->small[i].ensure(8);
    UVO_CUDA(cudaMemsetAsync(s->small[i].get(), 0, 8 * sizeof(int), c.stream));
    UVO_CUDA(cudaMemsetAsync(s->pnp_result[i].get(), 0, 8 * sizeof(double), c.stream));
    s->pnp_scratch[i].ensure(pnp_scratch_bytes(cap, prm->iterations_count) + 4096);
    UVO_CUDA(cudaEventCreateWithFlags(&s->ev_main[i], cudaEventDisableTiming));
    UVO_CUDA(cudaEventCreateWithFlags(&s->ev_side[i], cudaEventDisableTiming));
  }
  UVO_CUDA(cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking));
  s->d_result.ensure(uvo_stereo::RING);
  s->h_result.ensure(uvo_stereo::RING);
  for (auto& e : s->ev) UVO_CUDA(cudaEventCreate(&e));
  rng_table_device(c);
  UVO_CUDA(cudaStreamSynchronize(c.stream));
}

// enqueue every kernel of one frame on the ctx stream; images already on the device
static void stereo_enqueue(uvo_stereo* s, const uint8_t* dL, const uint8_t* dR, size_t pitch, double dt, int slot) {
  Ctx& c = s->ctx->c;
  const uvo_params& p = s->prm;
  const int cap = s->cap;
  const int cur = s->parity, prv = s->parity ^ 1;
  FrameCtrl* ctrl = s->ctrl.get() + cur;
  FrameCtrl* pctrl = s->ctrl.get() + prv;
  const GateParams g{p.min_num_features, p.min_num_3dpoints, p.min_num_inliers, cap};
  auto mark = [&](int i) {
    if (s->timing) UVO_CUDA(cudaEventRecord(s->ev[i], c.stream));
  };
  // frame t reuses the control block and pose scratch of frame t-2: its pose stage (side stream) must be done
  if (s->side_used[cur]) UVO_CUDA(cudaStreamWaitEvent(c.stream, s->ev_side[cur], 0));
  mark(1);
  // 1. get_image x2 (visual_odometry.h:542-543)
  s->fe.prep(c, 0, dL, pitch, s->cam[0], p.clahe, p.clip_limit);
  s->fe.prep(c, 1, dR, pitch, s->cam[1], p.clahe, p.clip_limit);
  mark(2);
  // 2. detect_features x2 (:548-549), both images batched through each kernel
  s->fe.surf(c, 0, 2, p);
  const int* cL = s->fe.counters.get();
  const int* cR = s->fe.counters.get() + 4;
  UVO_KERNEL(c, "k_gate_features");
  k_gate_features<<<1, 1, 0, c.stream>>>(cL, cR, ctrl, g);
  UVO_LAUNCH_CHECK(c);
  mark(3);
  // 3. match_features(curr_left, curr_right) (:558)
  MatchArgs ms{};
  ms.q = s->fe.desc[0].get();
  ms.t = s->fe.desc[1].get();
  ms.nq_dev = &ctrl->nq_stereo;
  ms.nt_dev = cR + 1;
  ms.nq = cap;
  ms.nt = cap;
  ms.ratio = (float)p.lowe_ratio;
  ms.partial = s->knn_scratch.get();
  ms.knn = s->knn_scratch.get() + (size_t)MATCH_SPLITS * cap;
  ms.matches = s->m_stereo.get();
  ms.n_matches = &ctrl->n_stereo;
  launch_match(c, ms);
  // 4. gathers (:569-579)
  UVO_KERNEL(c, "k_gather_after_stereo");
  k_gather_after_stereo<<<2 * c.sm_count, 256, 0, c.stream>>>(s->m_stereo.get(), ctrl, pctrl, s->state.get(),
                                                            s->fe.kps[0].get(), s->fe.kps[1].get(),
                                                            s->fe.desc[0].get(), s->kL_as[cur].get(),
                                                            s->kR_as[cur].get(), s->dL_as[cur].get(), g);
  UVO_LAUNCH_CHECK(c);
  mark(4);
  // 5. triangular match: prev-left-after-stereo (query) vs all current left features (train) (:592)
  MatchArgs mt = ms;
  mt.q = s->dL_as[prv].get();
  mt.t = s->fe.desc[0].get();
  mt.nq_dev = &ctrl->nq_temporal;
  mt.nt_dev = cL + 1;
  mt.matches = s->m_temporal.get();
  mt.n_matches = &ctrl->n_temporal;
  launch_match(c, mt);
  mark(5);
  // 6-8. triangulatePoints(prev left, prev right) (:631) + extract_3Dpoints (:632)
  TriangulateArgs ta{};
  memcpy(ta.P1, s->P_left, sizeof(ta.P1));
  memcpy(ta.P2, s->P_right, sizeof(ta.P2));
  ta.matches = s->m_temporal.get();
  ta.kps1 = s->kL_as[prv].get();
  ta.kps2 = s->kR_as[prv].get();
  ta.n_dev = &ctrl->n_temporal;
  ta.n = cap;
  ta.min_points = p.min_num_features;  // if (results_match_prev_curr.size() > MIN_NUM_FEATURES)
  ta.gate_dev = &ctrl->nq_temporal;
  ta.out4 = s->X4.get();
  ta.stride = cap;
  ta.pts1_out = s->pts1.get();
  ta.pts2_out = s->pts2.get();
  ta.n_out = &ctrl->n_tri;
  launch_triangulate(c, ta);
  Extract3dArgs ea{};
  ea.kp1 = s->pts1.get();
  ea.kp2 = s->pts2.get();
  ea.p4 = s->X4.get();
  ea.stride = cap;
  ea.n_dev = &ctrl->n_tri;
  ea.n = cap;
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z3[3] = {0, 0, 0};
  memcpy(ea.R1, I, sizeof(I));
  memcpy(ea.t1, z3, sizeof(z3));
  memcpy(ea.R2, s->R_right, sizeof(I));
  memcpy(ea.t2, s->t_right, sizeof(z3));
  const double KL[4] = {s->cam[0].nfx, s->cam[0].nfy, s->cam[0].ncx, s->cam[0].ncy};
  const double KR[4] = {s->cam[1].nfx, s->cam[1].nfy, s->cam[1].ncx, s->cam[1].ncy};
  memcpy(ea.K1, KL, sizeof(KL));
  memcpy(ea.K2, KR, sizeof(KR));
  ea.tol = p.reprojection_tolerance;
  ea.min3d = p.min_num_3dpoints;
  ea.out_pts = s->good_pts.get();
  ea.out_idx = s->good_idx.get();
  ea.out_count = &ctrl->n_3d;
  ea.tmp_pts = s->tmp_pts.get();
  ea.tmp_idx = s->tmp_idx.get();
  launch_extract3d(c, ea);
  mark(6);
  // 9-10. solvePnPRansac(good_prevCam_points, curr left keypoints of the surviving matches) (:638-648)
  PnpArgs pa{};
  pa.X = s->good_pts.get();
  pa.x_idx = s->good_idx.get();
  pa.matches = s->m_temporal.get();
  pa.kps = s->fe.kps[0].get();
  pa.n_dev = &ctrl->n_3d;
  pa.n = cap;
  memcpy(pa.K, KL, sizeof(KL));
  pa.iterations = p.iterations_count;
  pa.reproj_err = (float)p.reprojection_error;
  pa.confidence = p.confidence;
  pa.min_points = p.min_num_3dpoints;  // if (good_prevCam_points.rows > MIN_NUM_3DPOINTS)
  pa.result = s->pnp_result[cur].get();
  pa.inliers = s->inliers[cur].get();
  pa.n_inliers = s->small[cur].get();
  pa.hyps = s->small[cur].get() + 1;
  pa.best = s->small[cur].get() + 2;
  {
    const int iters = std::max(p.iterations_count, 1);
    uint8_t* b = s->pnp_scratch[cur].get();
    auto take = [&](size_t bytes) {
      uint8_t* r = b;
      b += (bytes + 255) & ~(size_t)255;
      return r;
    };
    pa.subsets = (int32_t*)take(sizeof(int32_t) * 5 * iters);
    pa.hyp_model = (double*)take(sizeof(double) * 15 * iters);
    pa.hyp_good = (int*)take(sizeof(int) * iters);
    pa.xs = (float*)take(sizeof(float) * 2 * cap);
    pa.Xf = (float*)take(sizeof(float) * 3 * cap);
  }
  // the gathers (which read buffers the next frame overwrites) and the subset stream stay on the main stream; the
  // long, narrow part of the pose stage -- hypotheses, scoring, bookkeeping, refit, result -- moves to a side stream so
  // that it overlaps the next frame's front end (it only touches its own parity of the pose scratch)
  launch_pnp_prepare(c, pa);
  UVO_CUDA(cudaEventRecord(s->ev_main[cur], c.stream));
  cudaStream_t main_stream = c.stream;
  c.stream = s->side;
  try {
    UVO_CUDA(cudaStreamWaitEvent(c.stream, s->ev_main[cur], 0));
    launch_pnp_solve(c, pa);
    mark(7);
    // 11-13. Rodrigues, t_prevCam_currCam, velocity; one record back to the host
    ResultParams rp{g, dt};
    UVO_KERNEL(c, "k_frame_result");
    k_frame_result<<<1, 1, 0, c.stream>>>(ctrl, s->state.get(), s->pnp_result[cur].get(), s->small[cur].get(),
                                          s->small[cur].get() + 1, s->d_result.get() + slot, rp);
    UVO_LAUNCH_CHECK(c);
    UVO_CUDA(cudaMemcpyAsync(s->h_result.p + slot, s->d_result.get() + slot, sizeof(uvo_stereo_result),
                             cudaMemcpyDeviceToHost, c.stream));
    mark(8);
    UVO_CUDA(cudaEventRecord(s->ev_side[cur], c.stream));
    s->side_used[cur] = true;
  } catch (...) {
    c.stream = main_stream;
    throw;
  }
  c.stream = main_stream;
  // 14. carry curr -> prev (:723-733): swap the after-stereo buffers
  s->parity ^= 1;
  s->frame_no++;
}

extern "C" {

int uvo_stereo_create(uvo_ctx* ctx, int w, int h, const uvo_camera* left, const uvo_camera* right,
                      const double R_right[9], const double t_right[3], const uvo_params* prm, uvo_stereo** out) {
  if (!ctx || !out) return UVO_ERR_INVALID;
  *out = nullptr;
  uvo_stereo* s = new uvo_stereo();
  int rc = guarded(&ctx->c, [&] { stereo_init(s, ctx, w, h, left, right, R_right, t_right, prm); });
  if (rc != UVO_OK) {
    delete s;
    return rc;
  }
  *out = s;
  return UVO_OK;
}

void uvo_stereo_destroy(uvo_stereo* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->c.device);
  cudaStreamSynchronize(s->ctx->c.stream);
  if (s->side) cudaStreamSynchronize(s->side);
  delete s;
}

int uvo_stereo_enqueue_device(uvo_stereo* s, const uint8_t* dL, const uint8_t* dR, size_t pitch, double dt) {
  if (!s) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    UVO_REQUIRE(dL && dR && pitch >= (size_t)3 * s->w && dt != 0.0, "uvo_stereo_enqueue_device: bad argument");
    UVO_REQUIRE((int)s->pending.size() < uvo_stereo::RING, "too many frames in flight: call uvo_stereo_collect");
    Ctx& c = s->ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    const int slot = (int)(s->frame_no % uvo_stereo::RING);
    stereo_enqueue(s, dL, dR, pitch, dt, slot);
    cudaEvent_t e;
    UVO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    UVO_CUDA(cudaEventRecord(e, s->side));
    s->pending.emplace_back(slot, e);
  });
}

int uvo_stereo_collect(uvo_stereo* s, uvo_stereo_result* out) {
  if (!s || !out) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    UVO_REQUIRE(!s->pending.empty(), "uvo_stereo_collect: no frame in flight");
    auto pr = s->pending.front();
    s->pending.pop_front();
    UVO_CUDA(cudaEventSynchronize(pr.second));
    cudaEventDestroy(pr.second);
    *out = s->h_result.p[pr.first];
    if (out->gate == -1)
      throw InvalidArg{"more SURF keypoints than max_features: raise uvo_params.max_features", UVO_ERR_CAPACITY};
  });
}

int uvo_stereo_frame_device(uvo_stereo* s, const uint8_t* dL, const uint8_t* dR, size_t pitch, double dt,
                            uvo_stereo_result* out) {
  if (!s || !out) return UVO_ERR_INVALID;
  s->timing = true;
  int rc = uvo_stereo_enqueue_device(s, dL, dR, pitch, dt);
  s->timing = false;
  if (rc != UVO_OK) return rc;
  rc = uvo_stereo_collect(s, out);
  if (rc == UVO_OK || rc == UVO_ERR_CAPACITY) {
    s->stage_ms[0] = 0.f;
    for (int i = 1; i < UVO_N_STAGES; i++) cudaEventElapsedTime(&s->stage_ms[i], s->ev[i], s->ev[i + 1]);
    s->has_timing = true;
  }
  return rc;
}

int uvo_stereo_frame(uvo_stereo* s, const uint8_t* left3, const uint8_t* right3, size_t pitch, double dt,
                     uvo_stereo_result* out) {
  if (!s || !out) return UVO_ERR_INVALID;
  int rc = guarded(&s->ctx->c, [&] {
    UVO_REQUIRE(left3 && right3 && pitch >= (size_t)3 * s->w, "uvo_stereo_frame: bad argument");
    Ctx& c = s->ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    UVO_CUDA(cudaEventRecord(s->ev[0], c.stream));
    UVO_CUDA(cudaMemcpy2DAsync(s->src[0].get(), s->src_pitch, left3, pitch, (size_t)3 * s->w, s->h,
                               cudaMemcpyHostToDevice, c.stream));
    UVO_CUDA(cudaMemcpy2DAsync(s->src[1].get(), s->src_pitch, right3, pitch, (size_t)3 * s->w, s->h,
                               cudaMemcpyHostToDevice, c.stream));
  });
  if (rc != UVO_OK) return rc;
  rc = uvo_stereo_frame_device(s, s->src[0].get(), s->src[1].get(), s->src_pitch, dt, out);
  if (rc == UVO_OK) cudaEventElapsedTime(&s->stage_ms[0], s->ev[0], s->ev[1]);
  return rc;
}

int uvo_stereo_last_keypoints(uvo_stereo* s, int right, uvo_keypoint* kps, float* desc, int capacity, int* count) {
  if (!s || !count) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    Ctx& c = s->ctx->c;
    const int idx = right ? 1 : 0;
    int cnt[4];
    UVO_CUDA(cudaMemcpyAsync(cnt, s->fe.counters.get() + 4 * idx, sizeof(cnt), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    const int n = std::min(cnt[1], s->cap);
    UVO_REQUIRE(n <= capacity, "uvo_stereo_last_keypoints: capacity too small");
    if (n > 0 && kps) UVO_CUDA(cudaMemcpy(kps, s->fe.kps[idx].get(), sizeof(uvo_keypoint) * n, cudaMemcpyDeviceToHost));
    if (n > 0 && desc) UVO_CUDA(cudaMemcpy(desc, s->fe.desc[idx].get(), sizeof(float) * 64 * n, cudaMemcpyDeviceToHost));
    *count = n;
  });
}

int uvo_stereo_last_matches(uvo_stereo* s, int temporal, uvo_dmatch* m, int capacity, int* count) {
  if (!s || !count) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    Ctx& c = s->ctx->c;
    FrameCtrl fc;
    UVO_CUDA(cudaMemcpyAsync(&fc, s->ctrl.get() + (s->parity ^ 1), sizeof(fc), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    const int n = temporal ? (fc.nq_temporal > 0 ? fc.n_temporal : 0) : fc.n_stereo;
    UVO_REQUIRE(n <= capacity, "uvo_stereo_last_matches: capacity too small");
    if (n > 0 && m)
      UVO_CUDA(cudaMemcpy(m, temporal ? s->m_temporal.get() : s->m_stereo.get(), sizeof(uvo_dmatch) * n,
                          cudaMemcpyDeviceToHost));
    *count = n;
  });
}

int uvo_stereo_last_inliers(uvo_stereo* s, int32_t* inl, int capacity, int* count) {
  if (!s || !count) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    Ctx& c = s->ctx->c;
    int n = 0;
    const int last = s->parity ^ 1;
    UVO_CUDA(cudaStreamSynchronize(s->side));
    UVO_CUDA(cudaMemcpyAsync(&n, s->small[last].get(), sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    UVO_REQUIRE(n <= capacity, "uvo_stereo_last_inliers: capacity too small");
    if (n > 0 && inl) UVO_CUDA(cudaMemcpy(inl, s->inliers[last].get(), sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
    *count = n;
  });
}

int uvo_stereo_stage_ms(uvo_stereo* s, float ms[UVO_N_STAGES]) {
  if (!s || !ms) return UVO_ERR_INVALID;
  if (!s->has_timing) return UVO_ERR_INVALID;
  for (int i = 0; i < UVO_N_STAGES; i++) ms[i] = s->stage_ms[i];
  return UVO_OK;
}

const char* uvo_stage_name(int i) { return (i >= 0 && i < UVO_N_STAGES) ? kStageNames[i] : ""; }

}  // extern "C"
