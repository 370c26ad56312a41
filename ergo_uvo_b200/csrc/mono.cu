// mono.cu -- visual_odometry_node::mono_VO's per-frame body (reference visual_odometry.h:247-397) behind one handle.
// Images, integral images, keypoints and descriptors of the previous and current frame stay on the device (front
// end: imgprep.cu / surf.cu, matcher: match.cu); the matched point lists (a few thousand Point2f) go through the host
// because estimate_relative_pose's control flow -- method switch, gates, "continue" -- is data dependent exactly as in
// the reference.  Every arithmetic step runs in the CUDA kernels of the stage-level entry points.
#include <cstring>
#include <vector>

#include "capi_internal.cuh"
#include "match.cuh"

using namespace uvo;

namespace uvo {
__global__ void __launch_bounds__(256) k_mono_gather(const uvo_dmatch* __restrict__ m, const int* __restrict__ n_dev,
                                                     int cap, const uvo_keypoint* __restrict__ k1,
                                                     const uvo_keypoint* __restrict__ k2, float* __restrict__ p1,
                                                     float* __restrict__ p2) {
  // 7-argument match_features (VO_utility.cpp:567-568): keypoints1_conv / keypoints2_conv of the surviving matches
  const int n = min(*n_dev, cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uvo_keypoint a = k1[m[i].queryIdx], b = k2[m[i].trainIdx];
    p1[2 * i] = a.x;
    p1[2 * i + 1] = a.y;
    p2[2 * i] = b.x;
    p2[2 * i + 1] = b.y;
  }
}
}  // namespace uvo

struct uvo_mono {
  uvo_ctx* ctx = nullptr;
  int w = 0, h = 0, cap = 0;
  uvo_camera cam{};
  uvo_params prm{};
  FrontEnd fe;                       // image slot 0 = current frame
  DevBuf<uint8_t> src;               // staging for host images
  size_t src_pitch = 0;
  DevBuf<uvo_keypoint> prev_kps;
  DevBuf<float> prev_desc;
  int n_prev = 0;
  DevBuf<uint8_t> match_scratch;
  DevBuf<uvo_dmatch> matches;
  DevBuf<int> n_matches;
  DevBuf<float> p1, p2;
  PinnedBuf<int> h_counts;
  bool initialised = false;
  int use_essential = 1;             // VO_utility.h:89
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0}, SF = 1.0;  // visual_odometry.h:195-206
};

static void mono_frame(uvo_mono* m, const uint8_t* img, size_t pitch, bool from_host, double dt, double range,
                       uvo_mono_result* out) {
  uvo_ctx* ctx = m->ctx;
  Ctx& c = ctx->c;
  const uvo_params& p = m->prm;
  const size_t dd = p.surf_extended ? 128 : 64;  // floats per descriptor row
  UVO_CUDA(cudaSetDevice(c.device));
  memset(out, 0, sizeof(*out));
  const uint8_t* d_img = img;
  if (from_host) {
    UVO_CUDA(cudaMemcpy2DAsync(m->src.get(), m->src_pitch, img, pitch, (size_t)3 * m->w, m->h, cudaMemcpyHostToDevice,
                               c.stream));
    d_img = m->src.get();
    pitch = m->src_pitch;
  }
  // get_image + detect_features (visual_odometry.h:232-236 / :262-281)
  m->fe.prep(c, 0, d_img, pitch, m->cam, p.clahe, p.clip_limit);
  m->fe.surf(c, 0, 1, p);
  m->h_counts.ensure(8);
  UVO_CUDA(cudaMemcpyAsync(m->h_counts.p, m->fe.counters.get(), 4 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  UVO_CUDA(cudaStreamSynchronize(c.stream));
  if (m->h_counts.p[0] > m->cap)
    throw InvalidArg{"keypoint capacity exceeded (raise uvo_params.max_features)", UVO_ERR_CAPACITY};
  const int n_curr = m->h_counts.p[1];
  out->n_keypoints = n_curr;
  auto roll = [&]() {  // prev_* = curr_* (visual_odometry.h:392-395)
    if (n_curr > 0) {
      UVO_CUDA(cudaMemcpyAsync(m->prev_kps.get(), m->fe.kps[0].get(), sizeof(uvo_keypoint) * n_curr,
                               cudaMemcpyDeviceToDevice, c.stream));
      UVO_CUDA(cudaMemcpyAsync(m->prev_desc.get(), m->fe.desc[0].get(), sizeof(float) * dd * n_curr,
                               cudaMemcpyDeviceToDevice, c.stream));
    }
    m->n_prev = n_curr;
  };
  auto publish = [&]() {  // mono_output_computation (visual_odometry.h:127-141): -SF * R^T t / dt
    for (int i = 0; i < 3; i++) {
      double v = 0;
      for (int k = 0; k < 3; k++) v += m->R[k * 3 + i] * m->t[k];
      out->velocity[i] = -m->SF * v / dt;
    }
    memcpy(out->R, m->R, sizeof(m->R));
    memcpy(out->t, m->t, sizeof(m->t));
    out->scale_factor = m->SF;
    out->published = 1;
  };
  if (!m->initialised) {
    roll();
    if (n_curr >= p.min_num_features) m->initialised = true;
    out->initialised = m->initialised;
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    return;
  }
  out->initialised = 1;
  if (n_curr < p.min_num_features) {  // "NUMBER OF DETECTED FEATURES IS TOO LOW. SKIP IMAGE!"
    roll();
    out->skipped = 1;
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    return;
  }
  // match_features(prev, curr, ..., prev_keypoints_conv, curr_keypoints_conv) (:285)
  int n_match = 0;
  if (m->n_prev > 0) {
    MatchArgs a{};
    a.q = m->prev_desc.get();
    a.t = m->fe.desc[0].get();
    a.nq = m->n_prev;
    a.nt = n_curr;
    a.dim = (int)dd;
    a.ratio = (float)p.lowe_ratio;
    match_bind_scratch(a, m->match_scratch.get(), m->cap, m->cap);
    a.matches = m->matches.get();
    a.n_matches = m->n_matches.get();
    launch_match(c, a);
    k_mono_gather<<<2 * c.sm_count, 256, 0, c.stream>>>(m->matches.get(), m->n_matches.get(), m->cap,
                                                       m->prev_kps.get(), m->fe.kps[0].get(), m->p1.get(), m->p2.get());
    c.launches++;
    UVO_CUDA(cudaGetLastError());
    UVO_CUDA(cudaMemcpyAsync(m->h_counts.p + 4, m->n_matches.get(), sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    n_match = m->h_counts.p[4];
  }
  out->n_matches = n_match;
  if (n_match < p.min_num_features) {  // "NUMBER OF FEATURES IS TOO LOW. SKIP IMAGE!" (:297-305)
    roll();
    out->skipped = 1;
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    return;
  }
  std::vector<float> hp1(2 * (size_t)n_match), hp2(2 * (size_t)n_match);
  UVO_CUDA(cudaMemcpyAsync(hp1.data(), m->p1.get(), sizeof(float) * 2 * n_match, cudaMemcpyDeviceToHost, c.stream));
  UVO_CUDA(cudaMemcpyAsync(hp2.data(), m->p2.get(), sizeof(float) * 2 * n_match, cudaMemcpyDeviceToHost, c.stream));
  roll();  // the device copies of curr_* are free to become prev_* from here on (same stream order)
  UVO_CUDA(cudaStreamSynchronize(c.stream));
  auto ck = [&](int rc) {
    if (rc != UVO_OK) throw InvalidArg{std::string("mono frame: ") + uvo_last_error(ctx), rc};
  };
  // CHECK CONSECUTIVE IMAGES BASELINE (:308-316)
  int ue = 0;
  ck(uvo_select_estimation_method(ctx, hp1.data(), hp2.data(), n_match, p.distance, &ue));
  m->use_essential = ue;
  const double K[4] = {m->cam.nfx, m->cam.nfy, m->cam.ncx, m->cam.ncy};
  std::vector<uint8_t> mask((size_t)n_match);
  int n_inl = 0, success = 0;
  ck(uvo_estimate_relative_pose(ctx, hp1.data(), hp2.data(), n_match, K, &p, &m->use_essential, m->R, m->t,
                                mask.data(), &n_inl, &success));
  out->used_essential = m->use_essential;
  out->n_inliers = n_inl;
  int valid = success;
  if (success) {
    // extract_inliers (VO_utility.cpp:306-329), then triangulatePoints + extract_3Dpoints (:350-357)
    std::vector<float> i1, i2;
    i1.reserve(2 * (size_t)n_inl);
    i2.reserve(2 * (size_t)n_inl);
    for (int i = 0; i < n_match; i++)
      if (mask[i]) {
        i1.push_back(hp1[2 * i]);
        i1.push_back(hp1[2 * i + 1]);
        i2.push_back(hp2[2 * i]);
        i2.push_back(hp2[2 * i + 1]);
      }
    const int ni = (int)i1.size() / 2;
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z3[3] = {0, 0, 0};
    double P0[12], P1[12];
    auto proj = [&](const double* Rm, const double* tv, double* P) {  // compute_projection_matrix: K [R|t]
      const double Km[9] = {K[0], 0, K[2], 0, K[1], K[3], 0, 0, 1};
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) {
          double acc = 0;
          for (int k = 0; k < 3; k++) acc += Km[i * 3 + k] * (j < 3 ? Rm[k * 3 + j] : tv[k]);
          P[i * 4 + j] = acc;
        }
    };
    proj(I, z3, P0);
    proj(m->R, m->t, P1);
    std::vector<float> X4(4 * (size_t)std::max(ni, 1));
    std::vector<double> good(3 * (size_t)std::max(ni, 1));
    std::vector<int32_t> gidx((size_t)std::max(ni, 1));
    int n3d = 0;
    if (ni > 0) {
      ck(uvo_triangulate_points(ctx, P0, P1, i1.data(), i2.data(), ni, X4.data()));
      ck(uvo_extract_3dpoints(ctx, i1.data(), i2.data(), ni, I, z3, m->R, m->t, K, K, X4.data(),
                              p.reprojection_tolerance, p.min_num_3dpoints, good.data(), gidx.data(), &n3d));
    }
    out->n_3d = n3d;
    if (n3d < p.min_num_3dpoints) {
      valid = 0;  // "NOT ENOUGH TRIANGULATED POINTS - ASSUMING CONSTANT MOTION"
    } else {
      double sf = 0;
      int n_front = 0;
      ck(uvo_scale_factor_front(ctx, good.data(), n3d, m->R, m->t, (float)range, &sf, &n_front));
      // the node assigns SF whenever good_currCam_points is non-empty (visual_odometry.h:365-374) -- also when
      // range == 0 (no altimeter message yet) makes it 0 -- and keeps successful_estimate = 1 then
      if (n_front > 0)
        m->SF = sf;
      else
        valid = 0;
    }
  }
  out->valid = valid;
  publish();
}

extern "C" {

int uvo_mono_create(uvo_ctx* ctx, int width, int height, const uvo_camera* cam, const uvo_params* prm, uvo_mono** out) {
  if (!ctx || !out) return UVO_ERR_INVALID;
  *out = nullptr;
  uvo_mono* m = new uvo_mono();
  const int rc = guarded(&ctx->c, [&] {
    UVO_REQUIRE(width > 0 && height > 0 && cam && prm, "uvo_mono_create: bad argument");
    UVO_REQUIRE(prm->max_features >= 64, "max_features too small");
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    m->ctx = ctx;
    m->w = width;
    m->h = height;
    m->cap = prm->max_features;
    m->cam = *cam;
    m->prm = *prm;
    m->fe.init(width, height, 1, m->cap);
    m->src_pitch = ((size_t)3 * width + 15) & ~(size_t)15;
    m->src.ensure(m->src_pitch * height);
    m->prev_kps.ensure(m->cap);
    const size_t dd = prm->surf_extended ? 128 : 64;  // floats per descriptor row (SURF_EXTENDED, VO_utility.h:86)
    m->prev_desc.ensure((size_t)m->cap * dd);
    UVO_CUDA(cudaMemsetAsync(m->prev_desc.get(), 0, (size_t)m->cap * dd * sizeof(float), c.stream));
    m->match_scratch.ensure(match_scratch_bytes(m->cap, m->cap));
    UVO_CUDA(cudaMemsetAsync(m->match_scratch.get(), 0, match_scratch_bytes(m->cap, m->cap), c.stream));
    m->matches.ensure(m->cap);
    m->n_matches.ensure(4);
    m->p1.ensure(2 * (size_t)m->cap);
    m->p2.ensure(2 * (size_t)m->cap);
    UVO_CUDA(cudaStreamSynchronize(c.stream));
  });
  if (rc != UVO_OK) {
    delete m;
    return rc;
  }
  *out = m;
  return UVO_OK;
}

void uvo_mono_destroy(uvo_mono* m) { delete m; }

int uvo_mono_frame(uvo_mono* m, const uint8_t* img3_host, size_t pitch, double dt, double range, uvo_mono_result* out) {
  if (!m || !out || !img3_host) return UVO_ERR_INVALID;
  return guarded(&m->ctx->c, [&] { mono_frame(m, img3_host, pitch, true, dt, range, out); });
}

int uvo_mono_frame_device(uvo_mono* m, const uint8_t* img3_dev, size_t pitch, double dt, double range,
                          uvo_mono_result* out) {
  if (!m || !out || !img3_dev) return UVO_ERR_INVALID;
  return guarded(&m->ctx->c, [&] { mono_frame(m, img3_dev, pitch, false, dt, range, out); });
}

}  // extern "C"
