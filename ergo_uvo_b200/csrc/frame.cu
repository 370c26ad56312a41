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
#include <thread>

#include "capi_internal.cuh"
#include "jpeg.cuh"
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
                                                             uvo_keypoint* kR_as, float* dL_as, GateParams g,
                                                             const int c4_shift) {
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
  // float4 chunks of the descriptors: 1 << c4_shift per row (16 for 64-float rows, 32 for extended 128-float rows)
  const int total = n_as << c4_shift;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int i = e >> c4_shift, q = e & ((1 << c4_shift) - 1);
    const uvo_dmatch mm = m[i];
    reinterpret_cast<float4*>(dL_as)[e] =
        reinterpret_cast<const float4*>(dL)[((size_t)mm.queryIdx << c4_shift) + q];
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
                               const int* hyps_dev, uvo_stereo_result* out, ResultParams p, const int* jpeg_status) {
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
  // a compressed image whose entropy-coded data was corrupt decodes to an empty image: reported, not published
  const bool bad_input = jpeg_status && (jpeg_status[0] | jpeg_status[2]);
  if (bad_input) gate = -2;
  if (was_init) {
    r.gate = gate;
    r.valid = gate == 0 ? 1 : 0;
    if (gate == 0) {
      double R[9];
      rodrigues_vec2mat(r.rvec, R);
      for (int i = 0; i < 3; i++) st->t_prev_curr[i] = -(R[i] * r.tvec[0] + R[3 + i] * r.tvec[1] + R[6 + i] * r.tvec[2]);
    }
  } else {
    r.gate = gate < 0 ? gate : 0;
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

// One lane = every buffer one frame needs + its own stream.  Frame t runs entirely on lane t % N_LANES, so the kernels
// of up to N_LANES consecutive frames are in flight on the GPU at once: the front end of frame t+1 (which does not
// depend on frame t at all) fills the SMs the narrow, latency-bound stages of frame t (rank sort, merges, RANSAC
// hypotheses, refit) leave idle.  Cross-lane order is restored with events exactly where the node's recurrence needs it:
//   gather(t)        after gather(t-1)        (vo_initialized, previous n_as)
//   temporal match/triangulate(t) read the after-stereo sets of lane (t-1)
//   gather(t)        after triangulate(t-N_LANES+1)   (lane t's after-stereo sets are still being read by the frame after
//                                                      the one that last used this lane)
//   result(t)        after result(t-1)        (t_prevCam_currCam carried across gate failures)
// A fixed run of kernels of one lane, captured once into a CUDA graph and replayed: every argument in it is a
// lane-owned buffer or a device-resident count, so nothing changes from frame to frame.  The cross-lane event waits
// and records stay ordinary stream operations BETWEEN segments (an event node inside a graph takes effect when the
// node executes, not when the graph is launched, which is not the order the lanes need).
struct Segment {
  cudaGraphExec_t exec = nullptr;
  int kernels = 0;  // kernel launches the segment stands for (keeps uvo_ctx_launch_count meaningful)
  ~Segment() {
    if (exec) cudaGraphExecDestroy(exec);
  }
};

struct Lane {
  cudaStream_t stream = nullptr;
  Segment seg[3];
  int frames = 0;  // frames this lane has run (the first one runs eagerly: lazy one-time set-up must not be captured)
  FrontEnd fe;
  DevBuf<uvo_keypoint> kL_as, kR_as;
  DevBuf<float> dL_as;
  DevBuf<FrameCtrl> ctrl;
  DevBuf<uvo_dmatch> m_stereo, m_temporal;
  DevBuf<uint8_t> knn_scratch;
  DevBuf<float> pts1, pts2, X4;
  DevBuf<double> good_pts, tmp_pts;
  DevBuf<int32_t> good_idx, tmp_idx;
  DevBuf<double> pnp_result;
  DevBuf<int32_t> inliers;
  DevBuf<int> small;  // [0] n_inliers [1] hyps [2..3] best
  DevBuf<int> jpeg_status;  // GPU entropy decode of the lane's frame: [0] / [2] error of the left / right image
  DevBuf<uint8_t> pnp_scratch;
  cudaEvent_t ev_gather = nullptr, ev_consumed = nullptr, ev_result = nullptr;
  bool used = false;
  ~Lane() {
    if (ev_gather) cudaEventDestroy(ev_gather);
    if (ev_consumed) cudaEventDestroy(ev_consumed);
    if (ev_result) cudaEventDestroy(ev_result);
    if (stream) cudaStreamDestroy(stream);
  }
};

#ifndef UVO_STEREO_LANES
#define UVO_STEREO_LANES 12  // frames whose kernels may be on the GPU at once (one stream + one set of buffers each)
#endif
#ifndef UVO_STEREO_RING
#define UVO_STEREO_RING (2 * UVO_STEREO_LANES)  // result slots = frames that may be enqueued and not yet collected
#endif
struct uvo_stereo {
  static constexpr int N_LANES = UVO_STEREO_LANES;
  static constexpr int RING = UVO_STEREO_RING;
  uvo_ctx* ctx = nullptr;
  int w = 0, h = 0, cap = 0;
  uvo_camera cam[2];
  uvo_params prm;
  double R_right[9], t_right[3];
  double P_left[12], P_right[12];  // P_eye_using_left_as_world, P_using_left_as_world (visual_odometry.h:460-462)
  Lane lane[N_LANES];
  size_t src_pitch = 0;
  DevBuf<FrameState> state;
  DevBuf<uvo_stereo_result> d_result;  // ring
  PinnedBuf<uvo_stereo_result> h_result;
  cudaEvent_t ev_in = nullptr;  // caller's work on the ctx stream -> lanes
  // host images: copied on a dedicated stream into a ring of staging pairs (one per result slot) so that the H2D of a
  // new frame does not queue behind the previous frame of its lane; the lane waits on the slot's event
  cudaStream_t copy_stream = nullptr;
  DevBuf<uint8_t> stage[RING][2];
  DevBuf<uint8_t> stage_bayer[RING][2];  // 1-channel staging of uvo_stereo_enqueue_host_bayer (allocated on first use)
  size_t bayer_pitch = 0;
  // compressed input (uvo_stereo_enqueue_host_sparse / _jpeg): per result slot and eye, the device copy of the sparse
  // coefficients and the component planes the IDCT writes; for _jpeg also the pinned host side of the entropy decode
  DevBuf<uint8_t> jpeg_sparse[RING][2], jpeg_planes[RING][2];
  PinnedBuf<uint32_t> jpeg_host[RING][2];
  // the decode of a compressed frame runs on its result slot's own stream, not on the lane: with up to RING frames
  // enqueued, the (long, narrow) Huffman decode of frame t + N_LANES overlaps the lane work of frame t
  DevBuf<int> jpeg_slot_status[RING];
  cudaStream_t ingest_stream[RING] = {};
  cudaEvent_t ev_ingest[RING] = {};
  cudaEvent_t ev_copied[RING] = {};
  long frame_no = 0;
  std::deque<int> pending;  // result slots in flight, oldest first
  cudaEvent_t ev[UVO_N_STAGES + 1] = {};
  cudaEvent_t ev_done[RING] = {};  // one per result slot, re-recorded
  bool use_graphs = true;          // uvo_stereo_set_graphs
  bool gpu_entropy = true;         // uvo_stereo_set_gpu_entropy: Huffman decoding of compressed input on the GPU
  long gpu_entropy_frames = 0;
  long graph_launches = 0;
  bool timing = false;
  float stage_ms[UVO_N_STAGES] = {};
  bool has_timing = false;

  Lane& last_lane() { return lane[(int)((frame_no + N_LANES - 1) % N_LANES)]; }

  ~uvo_stereo() {
    for (auto& e : ev_done)
      if (e) cudaEventDestroy(e);
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
    if (ev_in) cudaEventDestroy(ev_in);
    for (auto& e : ev_copied)
      if (e) cudaEventDestroy(e);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (auto& e : ev_ingest)
      if (e) cudaEventDestroy(e);
    for (auto& st : ingest_stream)
      if (st) cudaStreamDestroy(st);
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
  s->src_pitch = ((size_t)3 * w + 15) & ~(size_t)15;
  for (Lane& l : s->lane) {
    UVO_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
    l.fe.init(w, h, 2, cap);
    l.kL_as.ensure(cap);
    l.kR_as.ensure(cap);
    const size_t dd = prm->surf_extended ? 128 : 64;  // floats per descriptor row (SURF_EXTENDED, VO_utility.h:86)
    l.dL_as.ensure((size_t)cap * dd);
    UVO_CUDA(cudaMemsetAsync(l.dL_as.get(), 0, (size_t)cap * dd * sizeof(float), c.stream));
    l.ctrl.ensure(1);
    UVO_CUDA(cudaMemsetAsync(l.ctrl.get(), 0, sizeof(FrameCtrl), c.stream));
    l.m_stereo.ensure(cap);
    l.m_temporal.ensure(cap);
    l.knn_scratch.ensure(match_scratch_bytes(cap, cap));
    UVO_CUDA(cudaMemsetAsync(l.knn_scratch.get(), 0, match_scratch_bytes(cap, cap), c.stream));
    l.pts1.ensure(2 * (size_t)cap);
    l.pts2.ensure(2 * (size_t)cap);
    l.X4.ensure(4 * (size_t)cap);
    l.good_pts.ensure(3 * (size_t)cap);
    l.tmp_pts.ensure(3 * (size_t)cap);
    l.good_idx.ensure(cap);
    l.tmp_idx.ensure(cap);
    l.inliers.ensure(cap);
    l.pnp_result.ensure(8);
    l.small.ensure(8);
    UVO_CUDA(cudaMemsetAsync(l.small.get(), 0, 8 * sizeof(int), c.stream));
    l.jpeg_status.ensure(80);
    UVO_CUDA(cudaMemsetAsync(l.jpeg_status.get(), 0, 80 * sizeof(int), c.stream));
    UVO_CUDA(cudaMemsetAsync(l.pnp_result.get(), 0, 8 * sizeof(double), c.stream));
    l.pnp_scratch.ensure(pnp_scratch_bytes(cap, prm->iterations_count) + 4096);
    UVO_CUDA(cudaEventCreateWithFlags(&l.ev_gather, cudaEventDisableTiming));
    UVO_CUDA(cudaEventCreateWithFlags(&l.ev_consumed, cudaEventDisableTiming));
    UVO_CUDA(cudaEventCreateWithFlags(&l.ev_result, cudaEventDisableTiming));
  }
  s->state.ensure(1);
  UVO_CUDA(cudaMemsetAsync(s->state.get(), 0, sizeof(FrameState), c.stream));
  s->d_result.ensure(uvo_stereo::RING);
  s->h_result.ensure(uvo_stereo::RING);
  UVO_CUDA(cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming));
  for (auto& e : s->ev) UVO_CUDA(cudaEventCreate(&e));
  for (auto& e : s->ev_done) UVO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  rng_table_device(c);
  UVO_CUDA(cudaStreamSynchronize(c.stream));
}

// enqueue every kernel of one frame on its lane's stream.  Images: device pointers (host == nullptr) or pinned/pageable
// host pointers that are first copied into the lane's staging buffers on the same stream.
enum { SRC_DEVICE = 0, SRC_HOST_BGR = 1, SRC_HOST_BAYER = 2, SRC_HOST_SPARSE = 3, SRC_HOST_JPEG_GPU = 4 };

static void stereo_enqueue(uvo_stereo* s, const uint8_t* dL, const uint8_t* dR, size_t pitch, double dt, int slot,
                           int src_mode, const uvo_jpeg_sparse* const* sparse = nullptr, int bayer_bggr = 0,
                           const JpegGpuJob* gpu_jobs = nullptr) {
  const bool from_host = src_mode != SRC_DEVICE;
  Ctx& c = s->ctx->c;
  const uvo_params& p = s->prm;
  const int cap = s->cap;
  constexpr int NL = uvo_stereo::N_LANES;
  const int li = (int)(s->frame_no % NL);
  Lane& L = s->lane[li];
  Lane& PL = s->lane[(li + NL - 1) % NL];   // lane of frame t-1
  Lane& NX = s->lane[(li + 1) % NL];        // lane of frame t-NL+1: the last reader of this lane's after-stereo sets
  FrameCtrl* ctrl = L.ctrl.get();
  FrameCtrl* pctrl = PL.ctrl.get();
  const GateParams g{p.min_num_features, p.min_num_3dpoints, p.min_num_inliers, cap};
  const double KL[4] = {s->cam[0].nfx, s->cam[0].nfy, s->cam[0].ncx, s->cam[0].ncy};
  const double KR[4] = {s->cam[1].nfx, s->cam[1].nfy, s->cam[1].ncx, s->cam[1].ncy};
  cudaStream_t caller_stream = c.stream;
  // whatever the caller enqueued on the context stream (e.g. the production of the device images) comes first
  UVO_CUDA(cudaEventRecord(s->ev_in, caller_stream));
  UVO_CUDA(cudaStreamWaitEvent(L.stream, s->ev_in, 0));
  c.stream = L.stream;
  struct Restore {
    Ctx& c;
    cudaStream_t st;
    ~Restore() { c.stream = st; }
  } restore{c, caller_stream};
  auto mark = [&](int i) {
    if (s->timing) UVO_CUDA(cudaEventRecord(s->ev[i], c.stream));
  };
  mark(0);
  bool jpeg_on_gpu = false;
  if (src_mode == SRC_HOST_JPEG_GPU) {
    // compressed frames, Huffman decoding on the GPU: the scan bytes and the table plan go up on the copy stream
    // (one copy per image), then k_jpeg_huff (both images in one launch) + IDCT + colour on an ingest stream
    UVO_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_in, 0));
    uint8_t* d_buf[2];
    uint8_t* d_planes[2];
    uint8_t* d_bgr[2];
    for (int i = 0; i < 2; i++) {
      const size_t need = jpeg_gpu_device_bytes(gpu_jobs[i].L, gpu_jobs[i].upload_bytes);
      if (s->jpeg_sparse[slot][i].n < need) s->jpeg_sparse[slot][i].ensure(need + need / 4);  // head-room: scans vary
      s->jpeg_planes[slot][i].ensure(jpeg_plane_bytes(gpu_jobs[i].L));
      d_buf[i] = s->jpeg_sparse[slot][i].get();
      d_planes[i] = s->jpeg_planes[slot][i].get();
      d_bgr[i] = s->stage[slot][i].get();
      UVO_CUDA(cudaMemcpyAsync(d_buf[i], s->jpeg_host[slot][i].p, gpu_jobs[i].upload_bytes, cudaMemcpyHostToDevice,
                               s->copy_stream));
    }
    UVO_CUDA(cudaEventRecord(s->ev_copied[slot], s->copy_stream));
    // the decode runs on an ingest stream, not on the lane, and depends on nothing but its own upload: every buffer
    // it writes belongs to the result slot (whose previous user, frame_no - RING, has been collected), including
    // the status words -- the lane gets a copy.  Frame t + k therefore decodes while the lanes work on frames t ...
    // The number of ingest streams in use bounds how many decoder launches are in flight (jpeg_gpu_max_concurrent).
    const int n_ing = std::min(uvo_stereo::RING, jpeg_gpu_max_concurrent(c, 2, gpu_jobs));
    const int is = slot % n_ing;
    if (!s->ingest_stream[is]) UVO_CUDA(cudaStreamCreateWithFlags(&s->ingest_stream[is], cudaStreamNonBlocking));
    if (!s->ev_ingest[slot]) UVO_CUDA(cudaEventCreateWithFlags(&s->ev_ingest[slot], cudaEventDisableTiming));
    s->jpeg_slot_status[slot].ensure(80);
    {
      cudaStream_t ing = s->ingest_stream[is];
      UVO_CUDA(cudaStreamWaitEvent(ing, s->ev_copied[slot], 0));
      c.stream = ing;
      jpeg_gpu_launch(c, 2, gpu_jobs, d_buf, d_planes, bayer_bggr, d_bgr, s->src_pitch, s->jpeg_slot_status[slot].get());
      c.stream = L.stream;
      UVO_CUDA(cudaEventRecord(s->ev_ingest[slot], ing));
      UVO_CUDA(cudaStreamWaitEvent(c.stream, s->ev_ingest[slot], 0));
      UVO_CUDA(cudaMemcpyAsync(L.jpeg_status.get(), s->jpeg_slot_status[slot].get(), 80 * sizeof(int),
                               cudaMemcpyDeviceToDevice, c.stream));
    }
    jpeg_on_gpu = true;
    s->gpu_entropy_frames++;
    dL = d_bgr[0];
    dR = d_bgr[1];
    pitch = s->src_pitch;
  } else if (src_mode == SRC_HOST_SPARSE) {
    // compressed frames: the sparse coefficients go up on the copy stream, the transform runs on the lane
    UVO_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_in, 0));
    for (int i = 0; i < 2; i++) {
      s->jpeg_sparse[slot][i].ensure(jpeg_sparse_device_bytes(sparse[i]->layout, sparse[i]->n_entries) + 256);
      s->jpeg_planes[slot][i].ensure(jpeg_plane_bytes(sparse[i]->layout));
      jpeg_upload_sparse(s->copy_stream, *sparse[i], s->jpeg_sparse[slot][i].get());
    }
    UVO_CUDA(cudaEventRecord(s->ev_copied[slot], s->copy_stream));
    UVO_CUDA(cudaStreamWaitEvent(c.stream, s->ev_copied[slot], 0));
    for (int i = 0; i < 2; i++)
      jpeg_launch_transform(c, sparse[i]->layout, sparse[i]->n_entries, s->jpeg_sparse[slot][i].get(),
                            s->jpeg_planes[slot][i].get(), bayer_bggr, s->stage[slot][i].get(), s->src_pitch);
    dL = s->stage[slot][0].get();
    dR = s->stage[slot][1].get();
    pitch = s->src_pitch;
  } else if (src_mode == SRC_HOST_BAYER) {
    // 1-channel bayer images: a third of the bytes over PCIe, demosaiced on the lane's stream into the slot's staging pair
    UVO_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_in, 0));
    for (int i = 0; i < 2; i++)
      UVO_CUDA(cudaMemcpy2DAsync(s->stage_bayer[slot][i].get(), s->bayer_pitch, i == 0 ? dL : dR, pitch, s->w, s->h,
                                 cudaMemcpyHostToDevice, s->copy_stream));
    UVO_CUDA(cudaEventRecord(s->ev_copied[slot], s->copy_stream));
    UVO_CUDA(cudaStreamWaitEvent(c.stream, s->ev_copied[slot], 0));
    for (int i = 0; i < 2; i++)
      launch_demosaic_bggr(c, s->stage_bayer[slot][i].get(), s->bayer_pitch, s->w, s->h, s->stage[slot][i].get(),
                           s->src_pitch);
    dL = s->stage[slot][0].get();
    dR = s->stage[slot][1].get();
    pitch = s->src_pitch;
  } else if (from_host) {
    uint8_t* stL = s->stage[slot][0].get();
    uint8_t* stR = s->stage[slot][1].get();
    // the caller's stream order still applies to the host buffers (ev_in), the lane's previous frame does not
    UVO_CUDA(cudaStreamWaitEvent(s->copy_stream, s->ev_in, 0));
    if (pitch == s->src_pitch) {  // contiguous on both sides: one flat copy per image
      UVO_CUDA(cudaMemcpyAsync(stL, dL, pitch * s->h, cudaMemcpyHostToDevice, s->copy_stream));
      UVO_CUDA(cudaMemcpyAsync(stR, dR, pitch * s->h, cudaMemcpyHostToDevice, s->copy_stream));
    } else {
      UVO_CUDA(cudaMemcpy2DAsync(stL, s->src_pitch, dL, pitch, (size_t)3 * s->w, s->h, cudaMemcpyHostToDevice,
                                 s->copy_stream));
      UVO_CUDA(cudaMemcpy2DAsync(stR, s->src_pitch, dR, pitch, (size_t)3 * s->w, s->h, cudaMemcpyHostToDevice,
                                 s->copy_stream));
    }
    UVO_CUDA(cudaEventRecord(s->ev_copied[slot], s->copy_stream));
    UVO_CUDA(cudaStreamWaitEvent(c.stream, s->ev_copied[slot], 0));
    dL = stL;
    dR = stR;
    pitch = s->src_pitch;
  }
  mark(1);
  // Graph replay: off while per-stage or per-kernel timing is on (the marks and the kernel timer are stream
  // operations between kernels) and for a lane's first frame (one-time set-up inside the launchers -- function
  // attributes, constant tables -- must run for real, not be captured).
  const bool graphs = s->use_graphs && !s->timing && !c.kt.enabled && L.frames > 0;
  auto segment = [&](int k, auto&& body) {
    Segment& sg = L.seg[k];
    if (!graphs) {
      body();
      return;
    }
    if (!sg.exec) {
      const int64_t before = c.launches;
      cudaGraph_t graph = nullptr;
      UVO_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
      try {
        body();
      } catch (...) {
        cudaStreamEndCapture(c.stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      UVO_CUDA(cudaStreamEndCapture(c.stream, &graph));
      sg.kernels = (int)(c.launches - before);
      c.launches = before;
      const cudaError_t e = cudaGraphInstantiate(&sg.exec, graph, 0);
      cudaGraphDestroy(graph);
      UVO_CUDA(e);
    }
    UVO_CUDA(cudaGraphLaunch(sg.exec, c.stream));
    c.launches += sg.kernels;
    s->graph_launches++;
  };
  // 1. get_image x2 (visual_odometry.h:542-543)
  // both images through each preparation kernel at once (blockIdx.z = image).  (Running the right image on a second
  // stream instead was measured: -48 us of frame latency but -9 % end-to-end throughput with 8 frames in flight.)
  // The kernel that reads the source images takes this frame's pointers and is launched directly.
  L.fe.prep_pair(c, dL, dR, pitch, s->cam[0], s->cam[1], p.clahe, p.clip_limit, PREP_PART_SOURCE);
  const int* cL = L.fe.counters.get();
  const int* cR = L.fe.counters.get() + 4;
  MatchArgs ms{};
  // arguments of the stereo match (:558); the temporal match below copies and edits them
  ms.q = L.fe.desc[0].get();
  ms.t = L.fe.desc[1].get();
  ms.nq_dev = &ctrl->nq_stereo;
  ms.nt_dev = cR + 1;
  ms.nq = cap;
  ms.nt = cap;
  ms.dim = p.surf_extended ? 128 : 64;
  ms.ratio = (float)p.lowe_ratio;
  if (p.stereo_gate) {  // off in the reference's configuration (uvo_params.stereo_gate)
    ms.gate_kq = L.fe.kps[0].get();
    ms.gate_kt = L.fe.kps[1].get();
    ms.gate_dy = (float)p.stereo_max_epipolar_dy;
    ms.gate_dmin = (float)p.stereo_min_disparity;
    ms.gate_dmax = (float)p.stereo_max_disparity;
  }
  match_bind_scratch(ms, L.knn_scratch.get(), cap, cap);
  ms.matches = L.m_stereo.get();
  ms.n_matches = &ctrl->n_stereo;
  segment(0, [&] {
  L.fe.prep_pair(c, dL, dR, pitch, s->cam[0], s->cam[1], p.clahe, p.clip_limit, PREP_PART_REST);
  mark(2);
  // 2. detect_features x2 (:548-549), both images batched through each kernel
  L.fe.surf(c, 0, 2, p, /*with_integral=*/false);
  UVO_KERNEL(c, "k_gate_features");
  k_gate_features<<<1, 1, 0, c.stream>>>(cL, cR, ctrl, g);
  UVO_LAUNCH_CHECK(c);
  mark(3);
  // 3. match_features(curr_left, curr_right) (:558)
  launch_match(c, ms);
  });
  // 4. gathers (:569-579).  From here on the frame depends on its predecessor.
  if (PL.used) UVO_CUDA(cudaStreamWaitEvent(c.stream, PL.ev_gather, 0));
  if (NX.used && &NX != &PL) UVO_CUDA(cudaStreamWaitEvent(c.stream, NX.ev_consumed, 0));
  UVO_KERNEL(c, "k_gather_after_stereo");
  k_gather_after_stereo<<<2 * c.sm_count, 256, 0, c.stream>>>(L.m_stereo.get(), ctrl, pctrl, s->state.get(),
                                                            L.fe.kps[0].get(), L.fe.kps[1].get(), L.fe.desc[0].get(),
                                                            L.kL_as.get(), L.kR_as.get(), L.dL_as.get(), g,
                                                            p.surf_extended ? 5 : 4);
  UVO_LAUNCH_CHECK(c);
  UVO_CUDA(cudaEventRecord(L.ev_gather, c.stream));
  mark(4);
  segment(1, [&] {
  // 5. triangular match: prev-left-after-stereo (query) vs all current left features (train) (:592)
  MatchArgs mt = ms;
  mt.gate_kq = mt.gate_kt = nullptr;  // the gate is a stereo (rectified pair) constraint only
  mt.q = PL.dL_as.get();
  mt.t = L.fe.desc[0].get();
  mt.nq_dev = &ctrl->nq_temporal;
  mt.nt_dev = cL + 1;
  mt.matches = L.m_temporal.get();
  mt.n_matches = &ctrl->n_temporal;
  launch_match(c, mt);
  mark(5);
  // 6-8. triangulatePoints(prev left, prev right) (:631) + extract_3Dpoints (:632)
  TriangulateArgs ta{};
  memcpy(ta.P1, s->P_left, sizeof(ta.P1));
  memcpy(ta.P2, s->P_right, sizeof(ta.P2));
  ta.matches = L.m_temporal.get();
  ta.kps1 = PL.kL_as.get();
  ta.kps2 = PL.kR_as.get();
  ta.n_dev = &ctrl->n_temporal;
  ta.n = cap;
  ta.min_points = p.min_num_features;  // if (results_match_prev_curr.size() > MIN_NUM_FEATURES)
  ta.gate_dev = &ctrl->nq_temporal;
  ta.out4 = L.X4.get();
  ta.stride = cap;
  ta.pts1_out = L.pts1.get();
  ta.pts2_out = L.pts2.get();
  ta.n_out = &ctrl->n_tri;
  launch_triangulate(c, ta);
  });
  UVO_CUDA(cudaEventRecord(L.ev_consumed, c.stream));  // the previous frame's after-stereo sets are no longer needed
  segment(2, [&] {
  Extract3dArgs ea{};
  ea.kp1 = L.pts1.get();
  ea.kp2 = L.pts2.get();
  ea.p4 = L.X4.get();
  ea.stride = cap;
  ea.n_dev = &ctrl->n_tri;
  ea.n = cap;
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z3[3] = {0, 0, 0};
  memcpy(ea.R1, I, sizeof(I));
  memcpy(ea.t1, z3, sizeof(z3));
  memcpy(ea.R2, s->R_right, sizeof(I));
  memcpy(ea.t2, s->t_right, sizeof(z3));
  memcpy(ea.K1, KL, sizeof(KL));
  memcpy(ea.K2, KR, sizeof(KR));
  ea.tol = p.reprojection_tolerance;
  ea.min3d = p.min_num_3dpoints;
  ea.out_pts = L.good_pts.get();
  ea.out_idx = L.good_idx.get();
  ea.out_count = &ctrl->n_3d;
  ea.tmp_pts = L.tmp_pts.get();
  ea.tmp_idx = L.tmp_idx.get();
  launch_extract3d(c, ea);
  mark(6);
  // 9-10. solvePnPRansac(good_prevCam_points, curr left keypoints of the surviving matches) (:638-648)
  PnpArgs pa{};
  pa.X = L.good_pts.get();
  pa.x_idx = L.good_idx.get();
  pa.matches = L.m_temporal.get();
  pa.kps = L.fe.kps[0].get();
  pa.n_dev = &ctrl->n_3d;
  pa.n = cap;
  memcpy(pa.K, KL, sizeof(KL));
  pa.iterations = p.iterations_count;
  pa.reproj_err = (float)p.reprojection_error;
  pa.confidence = p.confidence;
  pa.min_points = p.min_num_3dpoints;  // if (good_prevCam_points.rows > MIN_NUM_3DPOINTS)
  pa.result = L.pnp_result.get();
  pa.inliers = L.inliers.get();
  pa.n_inliers = L.small.get();
  pa.hyps = L.small.get() + 1;
  pnp_bind_scratch(pa, L.pnp_scratch.get(), cap, p.iterations_count);
  launch_pnp_ransac(c, pa);
  });
  mark(7);
  // 11-13. Rodrigues, t_prevCam_currCam, velocity; one record back to the host
  if (PL.used) UVO_CUDA(cudaStreamWaitEvent(c.stream, PL.ev_result, 0));
  ResultParams rp{g, dt};
  UVO_KERNEL(c, "k_frame_result");
  k_frame_result<<<1, 1, 0, c.stream>>>(ctrl, s->state.get(), L.pnp_result.get(), L.small.get(), L.small.get() + 1,
                                        s->d_result.get() + slot, rp, jpeg_on_gpu ? L.jpeg_status.get() : nullptr);
  UVO_LAUNCH_CHECK(c);
  UVO_CUDA(cudaEventRecord(L.ev_result, c.stream));
  UVO_CUDA(cudaMemcpyAsync(s->h_result.p + slot, s->d_result.get() + slot, sizeof(uvo_stereo_result),
                           cudaMemcpyDeviceToHost, c.stream));
  mark(8);
  UVO_CUDA(cudaEventRecord(s->ev_done[slot], c.stream));
  s->pending.push_back(slot);
  L.used = true;
  L.frames++;
  // 14. carry curr -> prev (:723-733): the next frame reads this lane's after-stereo sets
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
  for (Lane& l : s->lane)
    if (l.stream) cudaStreamSynchronize(l.stream);
  delete s;
}

static void stereo_prepare_host_ring(uvo_stereo* s) {
  if (s->copy_stream) return;
  // first host frame: the copy stream and the whole staging ring at once (one allocation hiccup, not RING of
  // them).  Slot k's previous user (frame_no - RING) has always been collected before it is reused.
  UVO_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < uvo_stereo::RING; k++) {
    for (int i = 0; i < 2; i++) s->stage[k][i].ensure(s->src_pitch * s->h);
    UVO_CUDA(cudaEventCreateWithFlags(&s->ev_copied[k], cudaEventDisableTiming));
  }
}

static int stereo_enqueue_checked(uvo_stereo* s, const uint8_t* L, const uint8_t* R, size_t pitch, double dt,
                                  int src_mode) {
  if (!s) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    const bool from_host = src_mode != SRC_DEVICE;
    UVO_REQUIRE(L && R && pitch >= (size_t)(src_mode == SRC_HOST_BAYER ? 1 : 3) * s->w && dt != 0.0,
                "uvo_stereo_enqueue: bad argument");
    UVO_REQUIRE(src_mode != SRC_HOST_BAYER || (s->w >= 3 && s->h >= 3), "bayer input needs w, h >= 3");
    UVO_REQUIRE((int)s->pending.size() < uvo_stereo::RING,
                "too many frames in flight (uvo_stereo_max_in_flight): call uvo_stereo_collect");
    Ctx& c = s->ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    const int slot = (int)(s->frame_no % uvo_stereo::RING);
    if (from_host) stereo_prepare_host_ring(s);
    if (src_mode == SRC_HOST_BAYER && !s->bayer_pitch) {
      s->bayer_pitch = ((size_t)s->w + 15) & ~(size_t)15;
      for (int k = 0; k < uvo_stereo::RING; k++)
        for (int i = 0; i < 2; i++) s->stage_bayer[k][i].ensure(s->bayer_pitch * s->h);
    }
    stereo_enqueue(s, L, R, pitch, dt, slot, src_mode);
  });
}

static void check_sparse(const uvo_stereo* s, const uvo_jpeg_sparse* sp, int bayer_bggr) {
  UVO_REQUIRE(sp && sp->block_first && sp->block_count && (sp->n_entries == 0 || sp->entries),
              "uvo_stereo_enqueue_host_sparse: null buffer");
  UVO_REQUIRE(sp->layout.width == s->w && sp->layout.height == s->h,
              "compressed image size differs from the handle's (the reference would resize: use uvo_get_image_resized)");
  const int nc = sp->layout.components;
  if (!(nc == 3 || (nc == 1 && bayer_bggr)))
    throw InvalidArg{"compressed input: a 3-component stream, or a 1-component stream with bayer_bggr, is required",
                     UVO_ERR_UNSUPPORTED};
  UVO_REQUIRE(sp->layout.coeff_total > 0 && sp->n_entries <= (size_t)sp->layout.coeff_total, "bad sparse image");
}

int uvo_stereo_enqueue_host_sparse(uvo_stereo* s, const uvo_jpeg_sparse* left, const uvo_jpeg_sparse* right,
                                   int bayer_bggr, double dt) {
  if (!s) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    UVO_REQUIRE(dt != 0.0, "uvo_stereo_enqueue: bad argument");
    check_sparse(s, left, bayer_bggr);
    check_sparse(s, right, bayer_bggr);
    UVO_REQUIRE(!bayer_bggr || (s->w >= 3 && s->h >= 3), "bayer input needs w, h >= 3");
    UVO_REQUIRE((int)s->pending.size() < uvo_stereo::RING,
                "too many frames in flight (uvo_stereo_max_in_flight): call uvo_stereo_collect");
    Ctx& c = s->ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    stereo_prepare_host_ring(s);
    const int slot = (int)(s->frame_no % uvo_stereo::RING);
    const uvo_jpeg_sparse* both[2] = {left, right};
    stereo_enqueue(s, nullptr, nullptr, 0, dt, slot, SRC_HOST_SPARSE, both, bayer_bggr);
  });
}

int uvo_stereo_enqueue_host_jpeg(uvo_stereo* s, const uint8_t* left_jpeg, size_t left_len, const uint8_t* right_jpeg,
                                 size_t right_len, int bayer_bggr, double dt) {
  if (!s) return UVO_ERR_INVALID;
  uvo_jpeg_sparse sp[2];
  JpegGpuJob gpu_jobs[2];
  bool use_gpu = false;
  int rc = guarded(&s->ctx->c, [&] {
    UVO_REQUIRE(left_jpeg && right_jpeg && left_len && right_len && dt != 0.0, "uvo_stereo_enqueue_host_jpeg: bad argument");
    UVO_REQUIRE((int)s->pending.size() < uvo_stereo::RING,
                "too many frames in flight (uvo_stereo_max_in_flight): call uvo_stereo_collect");
    UVO_CUDA(cudaSetDevice(s->ctx->c.device));
    const int slot = (int)(s->frame_no % uvo_stereo::RING);
    const uint8_t* src[2] = {left_jpeg, right_jpeg};
    const size_t len[2] = {left_len, right_len};
    if (s->gpu_entropy) {
      // Huffman decoding on the GPU: the host only walks the markers and copies the scan with its stuffed zeros
      // removed into the slot's pinned staging.  Streams the GPU decoder does not take fall through to the host one.
      bool ok = true;
      for (int i = 0; i < 2 && ok; i++) {
        const size_t hb = jpeg_gpu_host_bytes(len[i]);
        s->jpeg_host[slot][i].ensure((hb + 3) / 4 + 1024);
        ok = jpeg_gpu_prepare(src[i], len[i], (uint8_t*)s->jpeg_host[slot][i].p, s->jpeg_host[slot][i].n * 4, &gpu_jobs[i]);
        if (ok) {
          const uvo_jpeg_layout& L = gpu_jobs[i].L;
          UVO_REQUIRE(L.width == s->w && L.height == s->h,
                      "compressed image size differs from the handle's (the reference would resize: use uvo_get_image_resized)");
          if (!(L.components == 3 || (L.components == 1 && bayer_bggr)))
            throw InvalidArg{"compressed input: a 3-component stream, or a 1-component stream with bayer_bggr, is required",
                             UVO_ERR_UNSUPPORTED};
        }
      }
      if (ok) {
        use_gpu = true;
        return;
      }
    }
    // pinned host side of the slot, sized from the headers: [first: nb][entries: <= total][count: nb bytes]
    for (int i = 0; i < 2; i++) {
      uvo_jpeg_layout L;
      const int e = uvo_jpeg_info(src[i], len[i], &L);
      if (e != UVO_OK) throw InvalidArg{"uvo_stereo_enqueue_host_jpeg: not a decodable baseline JPEG stream", e};
      // before anything is sized from the header: a stream of another size is refused, it does not get a buffer
      UVO_REQUIRE(L.width == s->w && L.height == s->h,
                  "compressed image size differs from the handle's (the reference would resize: use uvo_get_image_resized)");
      const size_t nb = (size_t)L.coeff_total / 64;
      s->jpeg_host[slot][i].ensure(nb + (size_t)L.coeff_total + (nb + 3) / 4);
    }
    // Huffman decoding, the two images of the pair on two host threads (the slot's previous user was collected, so
    // its upload is long finished)
    std::string err[2];
    int code[2] = {UVO_OK, UVO_OK};
    auto job = [&](int i) {
      try {
        uvo_jpeg_layout L;
        if (uvo_jpeg_info(src[i], len[i], &L) != UVO_OK) throw InvalidArg{"bad stream", UVO_ERR_INVALID};
        const size_t nb = (size_t)L.coeff_total / 64, total = (size_t)L.coeff_total;
        uint32_t* first = s->jpeg_host[slot][i].p;
        uint32_t* entries = first + nb;
        uint8_t* count = (uint8_t*)(entries + total);
        size_t ne = 0;
        jpeg_host_decode_sparse(src[i], len[i], entries, total, first, count, &ne, &sp[i].layout);
        sp[i].entries = entries;
        sp[i].n_entries = ne;
        sp[i].block_first = first;
        sp[i].block_count = count;
      } catch (const InvalidArg& e) {
        err[i] = e.msg;
        code[i] = e.code;
      } catch (const std::exception& e) {
        err[i] = e.what();
        code[i] = UVO_ERR_INVALID;
      }
    };
    std::thread right_thread(job, 1);
    job(0);
    right_thread.join();
    for (int i = 0; i < 2; i++)
      if (code[i] != UVO_OK) throw InvalidArg{(i ? "right image: " : "left image: ") + err[i], code[i]};
  });
  if (rc != UVO_OK) return rc;
  if (use_gpu)
    return guarded(&s->ctx->c, [&] {
      UVO_REQUIRE(!bayer_bggr || (s->w >= 3 && s->h >= 3), "bayer input needs w, h >= 3");
      stereo_prepare_host_ring(s);
      const int slot = (int)(s->frame_no % uvo_stereo::RING);
      stereo_enqueue(s, nullptr, nullptr, 0, dt, slot, SRC_HOST_JPEG_GPU, nullptr, bayer_bggr, gpu_jobs);
    });
  return uvo_stereo_enqueue_host_sparse(s, &sp[0], &sp[1], bayer_bggr, dt);
}

int uvo_stereo_set_gpu_entropy(uvo_stereo* s, int enable) {
  if (!s) return UVO_ERR_INVALID;
  s->gpu_entropy = enable != 0;
  return UVO_OK;
}

int64_t uvo_stereo_gpu_entropy_frames(const uvo_stereo* s) { return s ? (int64_t)s->gpu_entropy_frames : 0; }

int uvo_stereo_enqueue_device(uvo_stereo* s, const uint8_t* dL, const uint8_t* dR, size_t pitch, double dt) {
  return stereo_enqueue_checked(s, dL, dR, pitch, dt, SRC_DEVICE);
}

int uvo_stereo_enqueue_host(uvo_stereo* s, const uint8_t* left3, const uint8_t* right3, size_t pitch, double dt) {
  return stereo_enqueue_checked(s, left3, right3, pitch, dt, SRC_HOST_BGR);
}

int uvo_stereo_enqueue_host_bayer(uvo_stereo* s, const uint8_t* left1, const uint8_t* right1, size_t pitch, double dt) {
  return stereo_enqueue_checked(s, left1, right1, pitch, dt, SRC_HOST_BAYER);
}

int uvo_stereo_collect(uvo_stereo* s, uvo_stereo_result* out) {
  if (!s || !out) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    UVO_REQUIRE(!s->pending.empty(), "uvo_stereo_collect: no frame in flight");
    const int slot = s->pending.front();
    s->pending.pop_front();
    UVO_CUDA(cudaEventSynchronize(s->ev_done[slot]));
    *out = s->h_result.p[slot];
    if (out->gate == -1)
      throw InvalidArg{"more SURF keypoints than max_features: raise uvo_params.max_features", UVO_ERR_CAPACITY};
    if (out->gate == -2)
      throw InvalidArg{"corrupt or truncated entropy-coded data in a compressed input image", UVO_ERR_INVALID};
  });
}

static int stereo_frame_sync(uvo_stereo* s, const uint8_t* L, const uint8_t* R, size_t pitch, double dt,
                             uvo_stereo_result* out, bool from_host) {
  if (!s || !out) return UVO_ERR_INVALID;
  if (!s->pending.empty()) {
    s->ctx->c.err = "uvo_stereo_frame: frames are still in flight (collect them first)";
    return UVO_ERR_INVALID;
  }
  s->timing = true;
  int rc = stereo_enqueue_checked(s, L, R, pitch, dt, from_host ? SRC_HOST_BGR : SRC_DEVICE);
  s->timing = false;
  if (rc != UVO_OK) return rc;
  rc = uvo_stereo_collect(s, out);
  if (rc == UVO_OK || rc == UVO_ERR_CAPACITY) {
    for (int i = 0; i < UVO_N_STAGES; i++) cudaEventElapsedTime(&s->stage_ms[i], s->ev[i], s->ev[i + 1]);
    s->has_timing = true;
  }
  return rc;
}

int uvo_stereo_frame_device(uvo_stereo* s, const uint8_t* dL, const uint8_t* dR, size_t pitch, double dt,
                            uvo_stereo_result* out) {
  return stereo_frame_sync(s, dL, dR, pitch, dt, out, false);
}

int uvo_stereo_frame(uvo_stereo* s, const uint8_t* left3, const uint8_t* right3, size_t pitch, double dt,
                     uvo_stereo_result* out) {
  return stereo_frame_sync(s, left3, right3, pitch, dt, out, true);
}

int uvo_stereo_max_in_flight(void) { return uvo_stereo::RING; }
int uvo_stereo_lanes(void) { return uvo_stereo::N_LANES; }

int uvo_stereo_set_graphs(uvo_stereo* s, int enable) {
  if (!s) return UVO_ERR_INVALID;
  s->use_graphs = enable != 0;
  return UVO_OK;
}

int64_t uvo_stereo_graph_launches(const uvo_stereo* s) { return s ? (int64_t)s->graph_launches : 0; }

// the debug taps read the buffers of the most recent frame; they wait for every lane first
static void stereo_quiesce(uvo_stereo* s) {
  for (Lane& l : s->lane) UVO_CUDA(cudaStreamSynchronize(l.stream));
}

int uvo_stereo_last_keypoints(uvo_stereo* s, int right, uvo_keypoint* kps, float* desc, int capacity, int* count) {
  if (!s || !count) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    UVO_REQUIRE(s->frame_no > 0, "no frame has been processed yet");
    stereo_quiesce(s);
    Lane& L = s->last_lane();
    const int idx = right ? 1 : 0;
    int cnt[4];
    UVO_CUDA(cudaMemcpy(cnt, L.fe.counters.get() + 4 * idx, sizeof(cnt), cudaMemcpyDeviceToHost));
    const int n = std::min(cnt[1], s->cap);
    UVO_REQUIRE(n <= capacity, "uvo_stereo_last_keypoints: capacity too small");
    if (n > 0 && kps) UVO_CUDA(cudaMemcpy(kps, L.fe.kps[idx].get(), sizeof(uvo_keypoint) * n, cudaMemcpyDeviceToHost));
    const size_t dd = s->prm.surf_extended ? 128 : 64;
    if (n > 0 && desc) UVO_CUDA(cudaMemcpy(desc, L.fe.desc[idx].get(), sizeof(float) * dd * n, cudaMemcpyDeviceToHost));
    *count = n;
  });
}

int uvo_stereo_last_matches(uvo_stereo* s, int temporal, uvo_dmatch* m, int capacity, int* count) {
  if (!s || !count) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    UVO_REQUIRE(s->frame_no > 0, "no frame has been processed yet");
    stereo_quiesce(s);
    Lane& L = s->last_lane();
    FrameCtrl fc;
    UVO_CUDA(cudaMemcpy(&fc, L.ctrl.get(), sizeof(fc), cudaMemcpyDeviceToHost));
    const int n = temporal ? (fc.nq_temporal > 0 ? fc.n_temporal : 0) : fc.n_stereo;
    UVO_REQUIRE(n <= capacity, "uvo_stereo_last_matches: capacity too small");
    if (n > 0 && m)
      UVO_CUDA(cudaMemcpy(m, temporal ? L.m_temporal.get() : L.m_stereo.get(), sizeof(uvo_dmatch) * n,
                          cudaMemcpyDeviceToHost));
    *count = n;
  });
}

int uvo_stereo_last_inliers(uvo_stereo* s, int32_t* inl, int capacity, int* count) {
  if (!s || !count) return UVO_ERR_INVALID;
  return guarded(&s->ctx->c, [&] {
    UVO_REQUIRE(s->frame_no > 0, "no frame has been processed yet");
    stereo_quiesce(s);
    Lane& L = s->last_lane();
    int n = 0;
    UVO_CUDA(cudaMemcpy(&n, L.small.get(), sizeof(int), cudaMemcpyDeviceToHost));
    UVO_REQUIRE(n <= capacity, "uvo_stereo_last_inliers: capacity too small");
    if (n > 0 && inl) UVO_CUDA(cudaMemcpy(inl, L.inliers.get(), sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
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
