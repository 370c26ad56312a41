// capi.cu -- extern "C" entry points of libuvo_b200.so (include/uvo_c.h): context management and the
// stage-level calls that take HOST buffers (the drop-in layer under the reference's VO_utility functions).
// Device-resident whole-frame pipelines live in frame.cu.
#include <cstdlib>
#include <cstring>

#include "capi_internal.cuh"
#include "twoview.cuh"
#include "match.cuh"
#include "pose.cuh"

using namespace uvo;

// bump allocator over one grow-only device buffer, for the stage-level calls below
struct Arena {
  uint8_t* base;
  size_t off = 0, cap;
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = (T*)(base + off);
    off += n * sizeof(T);
    if (off > cap) throw InvalidArg{"internal: stage arena too small", UVO_ERR_INVALID};
    return p;
  }
};


extern "C" {

const char* uvo_version(void) { return "uvo-b200 0.1 (sm_100a)"; }

int uvo_ctx_create(int device, void* cuda_stream, uvo_ctx** out) {
  if (!out) return UVO_ERR_INVALID;
  *out = nullptr;
  // A stereo handle runs 8 lane streams, a copy stream and up to 14 ingest streams; with the default of 8 hardware
  // work queues they alias and a long narrow kernel (the Huffman decode) holds up the lane that shares its queue
  // (measured: 2 116 -> 2 410 frames/s from JPEG bytes).  Only effective when this is the process' first CUDA call;
  // a host that initialises CUDA earlier sets the variable itself.  Never overrides the user's value.
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    cudaGetLastError();
    return UVO_ERR_NO_DEVICE;  // no CPU fallback exists
  }
  uvo_ctx* c = new uvo_ctx();
  int rc = guarded(&c->c, [&] {
    UVO_CUDA(cudaSetDevice(device));
    c->c.device = device;
    if (cuda_stream) {
      c->c.stream = (cudaStream_t)cuda_stream;
      c->c.own_stream = false;
    } else {
      UVO_CUDA(cudaStreamCreateWithFlags(&c->c.stream, cudaStreamNonBlocking));
      c->c.own_stream = true;
    }
    cudaDeviceProp prop;
    UVO_CUDA(cudaGetDeviceProperties(&prop, device));
    c->c.sm_count = prop.multiProcessorCount;
    // the library carries sm_100a machine code only (no PTX): any other part -- older, or sm_103 / sm_12x -- would
    // create a context and then fail every launch with "no kernel image"
    UVO_REQUIRE(prop.major == 10 && prop.minor == 0,
                "libuvo_b200 is built for sm_100a (B200) only; this device has another compute capability");
  });
  if (rc != UVO_OK) {
    fprintf(stderr, "uvo_ctx_create: %s\n", c->c.err.c_str());
    delete c;
    return rc;
  }
  *out = c;
  return UVO_OK;
}

void uvo_ctx_destroy(uvo_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->c.device);
  cudaStreamSynchronize(ctx->c.stream);
  if (ctx->c.own_stream) cudaStreamDestroy(ctx->c.stream);
  delete ctx;
}

const char* uvo_last_error(const uvo_ctx* ctx) { return ctx ? ctx->c.err.c_str() : "null context"; }
void* uvo_ctx_stream(uvo_ctx* ctx) { return ctx ? (void*)ctx->c.stream : nullptr; }
int64_t uvo_ctx_launch_count(const uvo_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

int uvo_ctx_kernel_timing(uvo_ctx* ctx, int enable) {
  if (!ctx) return UVO_ERR_INVALID;
  cudaStreamSynchronize(ctx->c.stream);
  for (auto& r : ctx->c.kt.recs) {
    ctx->c.kt.pool.push_back(r.a);
    ctx->c.kt.pool.push_back(r.b);
  }
  ctx->c.kt.recs.clear();
  ctx->c.kt.enabled = enable != 0;
  return UVO_OK;
}

int uvo_ctx_kernel_report(uvo_ctx* ctx, char* buf, size_t buflen) {
  if (!ctx || !buf || buflen == 0) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_CUDA(cudaStreamSynchronize(ctx->c.stream));
    struct Acc {
      const char* name;
      int count;
      double ms;
    };
    std::vector<Acc> acc;
    for (auto& r : ctx->c.kt.recs) {
      float ms = 0.f;
      UVO_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
      bool found = false;
      for (auto& a : acc)
        if (strcmp(a.name, r.name) == 0) {
          a.count++;
          a.ms += ms;
          found = true;
          break;
        }
      if (!found) acc.push_back({r.name, 1, (double)ms});
      ctx->c.kt.pool.push_back(r.a);
      ctx->c.kt.pool.push_back(r.b);
    }
    ctx->c.kt.recs.clear();
    size_t off = 0;
    buf[0] = 0;
    for (auto& a : acc) {
      int n = snprintf(buf + off, buflen - off, "%s %d %.6f\n", a.name, a.count, a.ms);
      if (n < 0 || (size_t)n >= buflen - off) break;
      off += n;
    }
  });
}

int uvo_ctx_synchronize(uvo_ctx* ctx) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] { UVO_CUDA(cudaStreamSynchronize(ctx->c.stream)); });
}

void* uvo_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void uvo_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

void uvo_default_params(int stereo, uvo_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  // uvo/config/stereo_VO_parameters.yaml:8-47 and mono_VO_parameters.yaml:2-49
  p->clahe = 1;
  p->clip_limit = stereo ? 8 : 3;
  p->distance = 10;
  p->lowe_ratio = stereo ? 0.8 : 0.7;
  p->essential_method = 4;
  p->essential_max_iters = 2000;
  p->essential_confidence = 0.99;
  p->essential_threshold = 0.1;
  p->homography_method = 4;
  p->homography_max_iters = 2000;
  p->homography_confidence = 0.99;
  p->homography_threshold = 0.1;
  p->homography_distance = 50.0;
  p->vpf_threshold = 0.4;
  p->reprojection_tolerance = stereo ? 3.0 : 0.1;
  p->min_num_features = stereo ? 5 : 20;
  p->min_num_3dpoints = 5;
  p->min_num_inliers = stereo ? 5 : 10;
  p->iterations_count = 1000;
  p->reprojection_error = 1.0;
  p->confidence = 0.99;
  p->pnp_method_flag = 1;
  p->surf_min_hessian = stereo ? 1500 : 50;
  p->surf_octaves = 4;
  p->surf_octave_layers = 3;
  p->surf_extended = 0;
  p->surf_upright = 1;
  p->max_features = 16384;
  p->stereo_gate = 0;  // the reference has no epipolar / disparity gate (VO_utility.cpp:515-573): off by default
  p->stereo_max_epipolar_dy = 2.0;
  p->stereo_min_disparity = 0.0;
  p->stereo_max_disparity = 1e9;
}

// ------------------------------------------------------------------------------------------------ K1-K3
int uvo_get_image(uvo_ctx* ctx, const uint8_t* src3, int w, int h, size_t spitch, const uvo_camera* cam, int clahe,
                  int clip_limit, uint8_t* dst, size_t dpitch) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(src3 && dst && cam && w > 0 && h > 0, "uvo_get_image: null argument or empty image");
    UVO_REQUIRE(spitch >= (size_t)3 * w && dpitch >= (size_t)w, "uvo_get_image: pitch smaller than a row");
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    StageScratch& s = ctx->scratch;
    const size_t gp = ((size_t)w + 3) & ~(size_t)3;
    s.src3.ensure(spitch * h);
    s.gray.ensure(gp * h);
    s.lut.ensure(64 * 256);
    UVO_CUDA(cudaMemcpyAsync(s.src3.get(), src3, spitch * h, cudaMemcpyHostToDevice, c.stream));
    launch_gray_undistort(c, s.src3.get(), spitch, w, h, make_undistort_params(*cam), s.gray.get(), gp);
    if (clahe) {
      ClaheGeom g = make_clahe_geom(w, h, (double)clip_limit, 8, 8);
      launch_clahe(c, s.gray.get(), gp, w, h, g, s.lut.get(), s.gray.get(), gp);
    }
    UVO_CUDA(cudaMemcpy2DAsync(dst, dpitch, s.gray.get(), gp, w, h, cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int uvo_demosaic_bggr2bgr(uvo_ctx* ctx, const uint8_t* bayer, int w, int h, size_t spitch, uint8_t* bgr, size_t dpitch) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(bayer && bgr && w >= 3 && h >= 3 && spitch >= (size_t)w && dpitch >= (size_t)3 * w,
                "uvo_demosaic_bggr2bgr: bad argument (needs w, h >= 3)");
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    StageScratch& s = ctx->scratch;
    const size_t sp = ((size_t)w + 15) & ~(size_t)15, dp = ((size_t)3 * w + 15) & ~(size_t)15;
    s.bytes_a.ensure(sp * h);
    s.src3.ensure(dp * h);
    UVO_CUDA(cudaMemcpy2DAsync(s.bytes_a.get(), sp, bayer, spitch, w, h, cudaMemcpyHostToDevice, c.stream));
    launch_demosaic_bggr(c, s.bytes_a.get(), sp, w, h, s.src3.get(), dp);
    UVO_CUDA(cudaMemcpy2DAsync(bgr, dpitch, s.src3.get(), dp, (size_t)3 * w, h, cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int uvo_resize_area(uvo_ctx* ctx, const uint8_t* src, int sw, int sh, size_t spitch, int channels, uint8_t* dst, int dw,
                    int dh, size_t dpitch) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(src && dst && sw > 0 && sh > 0 && dw > 0 && dh > 0 && (channels == 1 || channels == 3) &&
                    spitch >= (size_t)sw * channels && dpitch >= (size_t)dw * channels,
                "uvo_resize_area: bad argument");
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    StageScratch& s = ctx->scratch;
    const size_t dp = ((size_t)dw * channels + 15) & ~(size_t)15;
    s.src3.ensure(spitch * sh);
    s.bytes_b.ensure(dp * dh);
    s.bytes_d.ensure(sizeof(AreaCell) * (size_t)(dw + dh));
    UVO_CUDA(cudaMemcpyAsync(s.src3.get(), src, spitch * sh, cudaMemcpyHostToDevice, c.stream));
    launch_resize_area(c, s.src3.get(), spitch, sw, sh, channels, s.bytes_b.get(), dp, dw, dh, (AreaCell*)s.bytes_d.get());
    UVO_CUDA(cudaMemcpy2DAsync(dst, dpitch, s.bytes_b.get(), dp, (size_t)dw * channels, dh, cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int uvo_get_image_resized(uvo_ctx* ctx, const uint8_t* src3, int w, int h, size_t spitch, int desired_width,
                          const uvo_camera* cam, int clahe, int clip_limit, uint8_t* dst, size_t dpitch, int* out_w,
                          int* out_h) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(src3 && dst && cam && w > 0 && h > 0 && desired_width > 0 && desired_width <= w,
                "uvo_get_image_resized: bad argument");
    // VO_utility.cpp:339-342
    const double ratio = (double)w / (double)desired_width;
    const int dw = desired_width, dh = (int)(h / ratio);
    UVO_REQUIRE(dh > 0 && spitch >= (size_t)3 * w && dpitch >= (size_t)dw, "uvo_get_image_resized: bad pitch / size");
    if (out_w) *out_w = dw;
    if (out_h) *out_h = dh;
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    StageScratch& s = ctx->scratch;
    const size_t rp = ((size_t)3 * dw + 15) & ~(size_t)15, gp = ((size_t)dw + 3) & ~(size_t)3;
    s.src3.ensure(spitch * h);
    s.bytes_b.ensure(rp * dh);
    s.bytes_d.ensure(sizeof(AreaCell) * (size_t)(dw + dh));
    s.gray.ensure(gp * dh);
    s.lut.ensure(64 * 256);
    UVO_CUDA(cudaMemcpyAsync(s.src3.get(), src3, spitch * h, cudaMemcpyHostToDevice, c.stream));
    launch_resize_area(c, s.src3.get(), spitch, w, h, 3, s.bytes_b.get(), rp, dw, dh, (AreaCell*)s.bytes_d.get());
    launch_gray_undistort(c, s.bytes_b.get(), rp, dw, dh, make_undistort_params(*cam), s.gray.get(), gp);
    if (clahe) {
      ClaheGeom g = make_clahe_geom(dw, dh, (double)clip_limit, 8, 8);
      launch_clahe(c, s.gray.get(), gp, dw, dh, g, s.lut.get(), s.gray.get(), gp);
    }
    UVO_CUDA(cudaMemcpy2DAsync(dst, dpitch, s.gray.get(), gp, dw, dh, cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int uvo_integral(uvo_ctx* ctx, const uint8_t* gray, int w, int h, size_t pitch, int32_t* sum) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(gray && sum && w > 0 && h > 0 && pitch >= (size_t)w, "uvo_integral: bad argument");
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    StageScratch& s = ctx->scratch;
    const size_t gp = ((size_t)w + 3) & ~(size_t)3;
    s.gray.ensure(gp * h);
    s.integral.ensure((size_t)(w + 1) * (h + 1));
    UVO_CUDA(cudaMemcpy2DAsync(s.gray.get(), gp, gray, pitch, w, h, cudaMemcpyHostToDevice, c.stream));
    launch_integral(c, s.gray.get(), gp, w, h, s.integral.get());  // dense: pitch w + 1
    UVO_CUDA(cudaMemcpyAsync(sum, s.integral.get(), sizeof(int32_t) * (size_t)(w + 1) * (h + 1),
                             cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
  });
}

// ------------------------------------------------------------------------------------------------ K4-K7
int uvo_detect_features(uvo_ctx* ctx, const uint8_t* gray, int w, int h, size_t pitch, const uvo_params* prm,
                        uvo_keypoint* kps, float* desc, int capacity, int* count) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(gray && prm && kps && desc && count && w > 0 && h > 0 && pitch >= (size_t)w && capacity > 0,
                "uvo_detect_features: bad argument");
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    FrontEnd& fe = ctx->fe;
    fe.init(w, h, 1, std::max(capacity, prm->max_features));
    UVO_CUDA(cudaMemcpy2DAsync(fe.gray[0].get(), fe.gpitch, gray, pitch, w, h, cudaMemcpyHostToDevice, c.stream));
    fe.surf(c, 0, 1, *prm);
    ctx->pinned_counts.ensure(8);
    int* pc = ctx->pinned_counts.p;
    UVO_CUDA(cudaMemcpyAsync(pc, fe.counters.get(), 4 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    *count = 0;
    if (pc[0] > fe.capacity || pc[1] > capacity)
      throw InvalidArg{"uvo_detect_features: more keypoints than capacity (raise capacity / max_features)",
                       UVO_ERR_CAPACITY};
    const int n = pc[1];
    if (n > 0) {
      UVO_CUDA(cudaMemcpyAsync(kps, fe.kps[0].get(), sizeof(uvo_keypoint) * n, cudaMemcpyDeviceToHost, c.stream));
      const size_t dd = prm->surf_extended ? 128 : 64;
      UVO_CUDA(cudaMemcpyAsync(desc, fe.desc[0].get(), sizeof(float) * dd * n, cudaMemcpyDeviceToHost, c.stream));
      UVO_CUDA(cudaStreamSynchronize(c.stream));
    }
    *count = n;
  });
}

// K6 alone (parity tap): SURF's final keypoint order on caller-provided keypoints
int uvo_sort_keypoints(uvo_ctx* ctx, const uvo_keypoint* kps_in, int n, int capacity, uvo_keypoint* kps_out) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(n >= 0 && capacity >= std::max(n, 1) && (n == 0 || (kps_in && kps_out)), "uvo_sort_keypoints: bad argument");
    if (n == 0) return;
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    FrontEnd& fe = ctx->fe;
    fe.init(std::max(fe.w, 16), std::max(fe.h, 16), 1, capacity);
    const int cap = fe.capacity;  // buffers (and the sort variant chosen) follow the front end's capacity
    SurfBatch b = fe.batch(0, 1);
    UVO_CUDA(cudaMemcpyAsync(b.im[0].raw, kps_in, sizeof(uvo_keypoint) * n, cudaMemcpyHostToDevice, c.stream));
    const int counters[4] = {n, 0, 0, 0};
    UVO_CUDA(cudaMemcpyAsync(b.im[0].counters, counters, sizeof(counters), cudaMemcpyHostToDevice, c.stream));
    launch_surf_sort(c, b, cap);
    UVO_CUDA(cudaMemcpyAsync(kps_out, b.im[0].kps, sizeof(uvo_keypoint) * n, cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
  });
}

// ------------------------------------------------------------------------------------------------ K8
struct HostGate {  // optional stereo gate of uvo_match_features_gated
  const uvo_keypoint* k1;
  const uvo_keypoint* k2;
  float dy, dmin, dmax;
};

static void match_host(uvo_ctx* ctx, const float* d1, int n1, const float* d2, int n2, int dim, float ratio,
                       uvo_dmatch* matches, int* count, uvo_dmatch* knn_out, const HostGate* gate = nullptr) {
  if (count) *count = 0;
  UVO_REQUIRE(n1 >= 0 && n2 >= 0, "matcher: bad argument");
  // an empty query set is routine (the frame after a failed gate; knnMatch returns no rows): no match, whatever the
  // row length of the empty matrix says
  if (n1 == 0) return;
  if (dim != 64 && dim != 128)
    throw InvalidArg{"matcher: descriptor rows must be 64 (SURF) or 128 (extended SURF) floats", UVO_ERR_UNSUPPORTED};
  UVO_REQUIRE(d1 && (n2 == 0 || d2), "matcher: bad argument");
  Ctx& c = ctx->c;
  UVO_CUDA(cudaSetDevice(c.device));
  StageScratch& s = ctx->scratch;
  s.bytes_a.ensure(sizeof(float) * dim * (size_t)n1);
  s.bytes_b.ensure(sizeof(float) * dim * (size_t)std::max(n2, 1));
  s.bytes_c.ensure(match_scratch_bytes(n1, std::max(n2, 1)));
  UVO_CUDA(cudaMemsetAsync(s.bytes_c.get(), 0, match_scratch_bytes(n1, std::max(n2, 1)), c.stream));
  s.bytes_d.ensure(sizeof(uvo_dmatch) * (size_t)n1 + 16);
  UVO_CUDA(cudaMemcpyAsync(s.bytes_a.get(), d1, sizeof(float) * dim * (size_t)n1, cudaMemcpyHostToDevice, c.stream));
  if (n2 > 0)
    UVO_CUDA(cudaMemcpyAsync(s.bytes_b.get(), d2, sizeof(float) * dim * (size_t)n2, cudaMemcpyHostToDevice, c.stream));
  MatchArgs a{};
  a.q = (const float*)s.bytes_a.get();
  a.t = (const float*)s.bytes_b.get();
  a.nq = n1;
  a.nt = n2;
  a.dim = dim;
  a.exact_only = ctx->match_exact_only;
  a.ratio = ratio;
  match_bind_scratch(a, s.bytes_c.get(), n1, std::max(n2, 1));
  if (gate && n2 > 0) {
    const size_t b1 = sizeof(uvo_keypoint) * (size_t)n1, b2 = sizeof(uvo_keypoint) * (size_t)n2;
    s.bytes_e.ensure(b1 + b2 + 32);
    uint8_t* base = s.bytes_e.get();
    const size_t off2 = (b1 + 15) & ~(size_t)15;
    UVO_CUDA(cudaMemcpyAsync(base, gate->k1, b1, cudaMemcpyHostToDevice, c.stream));
    UVO_CUDA(cudaMemcpyAsync(base + off2, gate->k2, b2, cudaMemcpyHostToDevice, c.stream));
    a.gate_kq = (const uvo_keypoint*)base;
    a.gate_kt = (const uvo_keypoint*)(base + off2);
    a.gate_dy = gate->dy;
    a.gate_dmin = gate->dmin;
    a.gate_dmax = gate->dmax;
  }
  a.n_matches = (int*)s.bytes_d.get();
  a.matches = (uvo_dmatch*)(s.bytes_d.get() + 16);
  launch_match(c, a);
  ctx->pinned_counts.ensure(8);
  UVO_CUDA(cudaMemcpyAsync(ctx->pinned_counts.p, a.n_matches, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  UVO_CUDA(cudaMemcpyAsync(ctx->pinned_counts.p + 1, a.n_fallback, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  std::vector<Knn2> knn;
  if (knn_out) {
    knn.resize(n1);
    UVO_CUDA(cudaMemcpyAsync(knn.data(), a.knn, sizeof(Knn2) * n1, cudaMemcpyDeviceToHost, c.stream));
  }
  UVO_CUDA(cudaStreamSynchronize(c.stream));
  const int n = ctx->pinned_counts.p[0];
  ctx->last_match_fallbacks = ctx->pinned_counts.p[1];
  if (matches && n > 0) {
    UVO_CUDA(cudaMemcpyAsync(matches, a.matches, sizeof(uvo_dmatch) * n, cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
  }
  if (count) *count = n;
  if (knn_out)
    for (int i = 0; i < n1; i++) {
      knn_out[2 * i] = uvo_dmatch{i, knn[i].i0, 0, knn[i].i0 >= 0 ? knn[i].d0 : 0.f};
      knn_out[2 * i + 1] = uvo_dmatch{i, knn[i].i1, 0, knn[i].i1 >= 0 ? knn[i].d1 : 0.f};
    }
}

int uvo_match_features(uvo_ctx* ctx, const float* d1, int n1, const float* d2, int n2, int dim, float ratio,
                       uvo_dmatch* matches, int* count) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(count && (n1 <= 0 || matches), "uvo_match_features: null output");
    match_host(ctx, d1, n1, d2, n2, dim, ratio, matches, count, nullptr);
  });
}

int uvo_match_features_gated(uvo_ctx* ctx, const uvo_keypoint* k1, const float* d1, int n1, const uvo_keypoint* k2,
                             const float* d2, int n2, int dim, float ratio, float max_dy, float min_disp,
                             float max_disp, uvo_dmatch* matches, int* count) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(count && (n1 <= 0 || matches), "uvo_match_features_gated: null output");
    UVO_REQUIRE((n1 == 0 || k1) && (n2 == 0 || k2), "uvo_match_features_gated: null keypoints");
    const HostGate g{k1, k2, max_dy, min_disp, max_disp};
    match_host(ctx, d1, n1, d2, n2, dim, ratio, matches, count, nullptr, &g);
  });
}

int uvo_knn_match2(uvo_ctx* ctx, const float* d1, int n1, const float* d2, int n2, int dim, uvo_dmatch* knn) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(knn, "uvo_knn_match2: null output");
    match_host(ctx, d1, n1, d2, n2, dim, 0.f, nullptr, nullptr, knn);
  });
}

int uvo_match_last_fallbacks(uvo_ctx* ctx, int* count) {
  if (!ctx || !count) return UVO_ERR_INVALID;
  *count = ctx->last_match_fallbacks;
  return UVO_OK;
}

int uvo_pnp_profile(uvo_ctx* ctx, int enable, int64_t stamps[32]) {
  if (!ctx) return UVO_ERR_INVALID;
  ctx->pnp_profile = enable != 0;
  if (stamps)
    for (int i = 0; i < 32; i++) stamps[i] = (int64_t)ctx->pnp_stamps[i];
  return UVO_OK;
}

int uvo_match_exact_only(uvo_ctx* ctx, int enable) {
  if (!ctx) return UVO_ERR_INVALID;
  ctx->match_exact_only = enable != 0;
  return UVO_OK;
}

// ------------------------------------------------------------------------------------------------ K10a / K10b
namespace {
struct TwoViewDev {  // device copies of the correspondences + result slots, carved from the stage scratch
  float* p1;
  float* p2;
  uint8_t* mask;
  double* model;   // 9
  int* n_inl;
  double* Rt;      // 14
  uint8_t* mask2;  // recoverPose output mask
  uint8_t* flags;
  double* cand;    // 49
  int* counts;     // 4
  void* robust;    // robust scratch
};

TwoViewDev twoview_stage(uvo_ctx* ctx, const float* p1, const float* p2, int n, int iters, int kind) {
  Ctx& c = ctx->c;
  UVO_CUDA(cudaSetDevice(c.device));
  const size_t need = (size_t)n * 32 + robust_scratch_bytes(n, std::max(iters, 1), kind) + 16384;
  ctx->scratch.bytes_a.ensure(need);
  Arena ar{ctx->scratch.bytes_a.get(), 0, ctx->scratch.bytes_a.n};
  TwoViewDev d{};
  d.p1 = ar.take<float>(2 * (size_t)std::max(n, 1));
  d.p2 = ar.take<float>(2 * (size_t)std::max(n, 1));
  d.mask = ar.take<uint8_t>(std::max(n, 1));
  d.mask2 = ar.take<uint8_t>(std::max(n, 1));
  d.flags = ar.take<uint8_t>(std::max(n, 1));
  d.model = ar.take<double>(16);
  d.n_inl = ar.take<int>(4);
  d.Rt = ar.take<double>(16);
  d.cand = ar.take<double>(64);
  d.counts = ar.take<int>(4);
  d.robust = ar.take<uint8_t>(robust_scratch_bytes(n, std::max(iters, 1), kind));
  if (n > 0) {
    UVO_CUDA(cudaMemcpyAsync(d.p1, p1, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c.stream));
    UVO_CUDA(cudaMemcpyAsync(d.p2, p2, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c.stream));
  }
  return d;
}

// findHomography / findEssentialMat on staged device points; results copied back
void robust_host(uvo_ctx* ctx, const float* p1, const float* p2, int n, int kind, const double* K, int method,
                 double threshold, int max_iters, double confidence, double model[9], uint8_t* mask, int* n_inliers,
                 int* hyps, int* ok) {
  UVO_REQUIRE(n >= 0 && (n == 0 || (p1 && p2)) && model, "two-view estimation: bad argument");
  if (!(method == TV_RANSAC || method == TV_LMEDS || (method == TV_LSQ && kind == TV_HOMOGRAPHY)))
    throw InvalidArg{"two-view estimation: method must be 8 (RANSAC), 4 (LMEDS) or, for findHomography, 0 (all points); "
                     "RHO (16) and the USAC family (32-38) are not implemented", UVO_ERR_UNSUPPORTED};
  UVO_REQUIRE(confidence > 0 && confidence < 1, "two-view estimation: confidence must be in (0, 1)");
  const int mp = kind == TV_ESSENTIAL ? 5 : 4;
  for (int k = 0; k < 9; k++) model[k] = 0;
  if (n_inliers) *n_inliers = 0;
  if (hyps) *hyps = 0;
  if (ok) *ok = 0;
  if (n < mp) {
    if (mask)
      for (int i = 0; i < n; i++) mask[i] = 0;
    return;
  }
  Ctx& c = ctx->c;
  RobustArgs a{};
  a.kind = kind;
  a.method = method;
  a.n = n;
  a.iters = robust_iterations(method, mp, confidence, max_iters);
  a.confidence = confidence;
  if (kind == TV_ESSENTIAL) {
    for (int k = 0; k < 4; k++) a.K[k] = K[k];
    a.threshold = threshold / ((K[0] + K[1]) / 2);
  } else {
    a.threshold = threshold <= 0 ? 3.0 : threshold;
  }
  TwoViewDev d = twoview_stage(ctx, p1, p2, n, a.iters, kind);
  a.p1 = d.p1;
  a.p2 = d.p2;
  robust_bind_scratch(a, d.robust);
  a.model_out = d.model;
  a.mask = d.mask;
  a.n_inliers = d.n_inl;
  launch_robust(c, a);
  ctx->pinned_counts.ensure(8);
  int* pc = ctx->pinned_counts.p;
  UVO_CUDA(cudaMemcpyAsync(pc, a.ctl, 6 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  UVO_CUDA(cudaMemcpyAsync(pc + 6, d.n_inl, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  UVO_CUDA(cudaMemcpyAsync(model, d.model, 9 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  if (mask) UVO_CUDA(cudaMemcpyAsync(mask, d.mask, n, cudaMemcpyDeviceToHost, c.stream));
  UVO_CUDA(cudaStreamSynchronize(c.stream));
  if (n_inliers) *n_inliers = pc[6];
  if (hyps) *hyps = pc[4];
  if (ok) *ok = pc[5];
}

void recover_pose_host(uvo_ctx* ctx, const double model[9], const float* p1, const float* p2, int n, const double K[4],
                       int from_h, double dist, uint8_t* mask_inout, double R[9], double t[3], int* good, int* found) {
  UVO_REQUIRE(model && K && R && t && n >= 0 && (n == 0 || (p1 && p2)), "recover pose: bad argument");
  Ctx& c = ctx->c;
  TwoViewDev d = twoview_stage(ctx, p1, p2, n, 1, TV_HOMOGRAPHY);
  UVO_CUDA(cudaMemcpyAsync(d.model, model, 9 * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  if (mask_inout && n > 0) UVO_CUDA(cudaMemcpyAsync(d.mask, mask_inout, n, cudaMemcpyHostToDevice, c.stream));
  RecoverPoseArgs a{};
  a.p1 = d.p1;
  a.p2 = d.p2;
  a.n = n;
  for (int k = 0; k < 4; k++) a.K[k] = K[k];
  a.from_homography = from_h;
  a.model = d.model;
  a.mask_in = (mask_inout && !from_h) ? d.mask : nullptr;
  a.distance_thresh = dist;
  a.cand = d.cand;
  a.flags = d.flags;
  a.counts = d.counts;
  a.Rt = d.Rt;
  a.mask_out = from_h ? nullptr : d.mask2;
  launch_recover_pose(c, a);
  double h[14];
  UVO_CUDA(cudaMemcpyAsync(h, d.Rt, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
  if (mask_inout && !from_h && n > 0) UVO_CUDA(cudaMemcpyAsync(mask_inout, d.mask2, n, cudaMemcpyDeviceToHost, c.stream));
  UVO_CUDA(cudaStreamSynchronize(c.stream));
  if (h[13] != 0) {
    for (int k = 0; k < 9; k++) R[k] = h[k];
    for (int k = 0; k < 3; k++) t[k] = h[9 + k];
  }
  if (good) *good = (int)h[12];
  if (found) *found = h[13] != 0;
}
}  // namespace

int uvo_find_homography(uvo_ctx* ctx, const float* p1, const float* p2, int n, int method, double threshold,
                        int max_iters, double confidence, double H[9], uint8_t* mask, int* n_inliers, int* hyps,
                        int* ok) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    robust_host(ctx, p1, p2, n, TV_HOMOGRAPHY, nullptr, method, threshold, max_iters, confidence, H, mask, n_inliers,
                hyps, ok);
  });
}

int uvo_find_essential_mat(uvo_ctx* ctx, const float* p1, const float* p2, int n, const double K[4], int method,
                           double prob, double threshold, int max_iters, double E[9], uint8_t* mask, int* n_inliers,
                           int* hyps, int* ok) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(K, "uvo_find_essential_mat: null camera");
    robust_host(ctx, p1, p2, n, TV_ESSENTIAL, K, method, threshold, max_iters, prob, E, mask, n_inliers, hyps, ok);
  });
}

int uvo_recover_pose(uvo_ctx* ctx, const double E[9], const float* p1, const float* p2, int n, const double K[4],
                     uint8_t* mask_inout, double R[9], double t[3], int* good) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c,
                 [&] { recover_pose_host(ctx, E, p1, p2, n, K, 0, 50.0, mask_inout, R, t, good, nullptr); });
}

int uvo_recover_pose_homography(uvo_ctx* ctx, const double H[9], const float* p1, const float* p2, int n,
                                const double K[4], double homography_distance, double R[9], double t[3], int* good,
                                int* found) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    recover_pose_host(ctx, H, p1, p2, n, K, 1, homography_distance, nullptr, R, t, good, found);
  });
}

int uvo_estimate_relative_pose(uvo_ctx* ctx, const float* p1, const float* p2, int n, const double K[4],
                               const uvo_params* prm, int* use_essential, double R[9], double t[3],
                               uint8_t* inlier_mask, int* n_inliers, int* success) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(prm && use_essential && K && R && t && success && n >= 0, "uvo_estimate_relative_pose: bad argument");
    *success = 0;
    if (n_inliers) *n_inliers = 0;
    std::vector<uint8_t> mask((size_t)std::max(n, 1)), shrunk;
    bool switched = false;
    for (;;) {  // while (!estimate_completed), VO_utility.cpp:140
      int valid = 0, cnt = 0, hyps = 0, ok = 0;
      double M[9];
      if (*use_essential) {
        robust_host(ctx, p1, p2, n, TV_ESSENTIAL, K, prm->essential_method, prm->essential_threshold,
                    (int)prm->essential_max_iters, prm->essential_confidence, M, mask.data(), &cnt, &hyps, &ok);
        shrunk = mask;  // recoverPose shrinks its own copy; extract_inliers already used the findEssentialMat mask
        int good = 0;
        if (ok) recover_pose_host(ctx, M, p1, p2, n, K, 0, 50.0, shrunk.data(), R, t, &good, nullptr);
        for (int i = 0; i < n; i++) valid += shrunk[i] != 0;
        if (!ok) valid = 0;
      } else {
        robust_host(ctx, p1, p2, n, TV_HOMOGRAPHY, nullptr, prm->homography_method, prm->homography_threshold,
                    (int)prm->homography_max_iters, prm->homography_confidence, M, mask.data(), &cnt, &hyps, &ok);
        int good = 0, found = 0;
        // the reference hands ALL matches to recover_pose_homography, not the inliers (VO_utility.cpp:154, App. D-2)
        if (ok) recover_pose_host(ctx, M, p1, p2, n, K, 1, prm->homography_distance, nullptr, R, t, &good, &found);
        valid = cnt;
      }
      if (inlier_mask)
        for (int i = 0; i < n; i++) inlier_mask[i] = mask[i];
      if (n_inliers) *n_inliers = cnt;
      const double vpf = n > 0 ? (double)valid / n : 0.0;
      if (vpf >= prm->vpf_threshold && valid >= prm->min_num_inliers) {
        *success = 1;
        break;
      }
      if (switched) break;  // both methods failed
      switched = true;
      *use_essential = !*use_essential;
    }
  });
}

// ------------------------------------------------------------------------------------------------ K11, K12, K10c
int uvo_triangulate_points(uvo_ctx* ctx, const double P1[12], const double P2[12], const float* pts1,
                           const float* pts2, int n, float* out4) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(P1 && P2 && n >= 0 && (n == 0 || (pts1 && pts2 && out4)), "uvo_triangulate_points: bad argument");
    if (n == 0) return;
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    ctx->scratch.bytes_a.ensure((size_t)n * 8 * sizeof(float) + 1024);
    Arena ar{ctx->scratch.bytes_a.get(), 0, ctx->scratch.bytes_a.n};
    float* d1 = ar.take<float>(2 * (size_t)n);
    float* d2 = ar.take<float>(2 * (size_t)n);
    float* d4 = ar.take<float>(4 * (size_t)n);
    UVO_CUDA(cudaMemcpyAsync(d1, pts1, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c.stream));
    UVO_CUDA(cudaMemcpyAsync(d2, pts2, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c.stream));
    TriangulateArgs a{};
    memcpy(a.P1, P1, sizeof(a.P1));
    memcpy(a.P2, P2, sizeof(a.P2));
    a.pts1 = d1;
    a.pts2 = d2;
    a.n = n;
    a.min_points = -1;
    a.out4 = d4;
    a.stride = n;
    launch_triangulate(c, a);
    UVO_CUDA(cudaMemcpyAsync(out4, d4, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
  });
}

int uvo_extract_3dpoints(uvo_ctx* ctx, const float* kp1, const float* kp2, int n, const double R1[9],
                         const double t1[3], const double R2[9], const double t2[3], const double K1[4],
                         const double K2[4], const float* p4, double tol, int min3d, double* out_pts,
                         int32_t* out_idx, int* count) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(count && n >= 0 && R1 && t1 && R2 && t2 && K1 && K2, "uvo_extract_3dpoints: bad argument");
    *count = 0;
    if (n == 0) return;
    UVO_REQUIRE(kp1 && kp2 && p4 && out_pts && out_idx, "uvo_extract_3dpoints: null buffer");
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    ctx->scratch.bytes_a.ensure((size_t)n * 128 + 4096);
    Arena ar{ctx->scratch.bytes_a.get(), 0, ctx->scratch.bytes_a.n};
    float* dk1 = ar.take<float>(2 * (size_t)n);
    float* dk2 = ar.take<float>(2 * (size_t)n);
    float* d4 = ar.take<float>(4 * (size_t)n);
    Extract3dArgs a{};
    a.kp1 = dk1;
    a.kp2 = dk2;
    a.p4 = d4;
    a.stride = n;
    a.n = n;
    memcpy(a.R1, R1, sizeof(a.R1));
    memcpy(a.t1, t1, sizeof(a.t1));
    memcpy(a.R2, R2, sizeof(a.R2));
    memcpy(a.t2, t2, sizeof(a.t2));
    memcpy(a.K1, K1, sizeof(a.K1));
    memcpy(a.K2, K2, sizeof(a.K2));
    a.tol = tol;
    a.min3d = min3d;
    a.out_pts = ar.take<double>(3 * (size_t)n);
    a.out_idx = ar.take<int32_t>(n);
    a.tmp_pts = ar.take<double>(3 * (size_t)n);
    a.tmp_idx = ar.take<int32_t>(n);
    a.out_count = ar.take<int>(1);
    UVO_CUDA(cudaMemcpyAsync(dk1, kp1, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c.stream));
    UVO_CUDA(cudaMemcpyAsync(dk2, kp2, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c.stream));
    UVO_CUDA(cudaMemcpyAsync(d4, p4, sizeof(float) * 4 * n, cudaMemcpyHostToDevice, c.stream));
    launch_extract3d(c, a);
    ctx->pinned_counts.ensure(8);
    UVO_CUDA(cudaMemcpyAsync(ctx->pinned_counts.p, a.out_count, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    const int m = ctx->pinned_counts.p[0];
    if (m > 0) {
      UVO_CUDA(cudaMemcpyAsync(out_pts, a.out_pts, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, c.stream));
      UVO_CUDA(cudaMemcpyAsync(out_idx, a.out_idx, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, c.stream));
      UVO_CUDA(cudaStreamSynchronize(c.stream));
    }
    *count = m;
  });
}

int uvo_select_estimation_method(uvo_ctx* ctx, const float* p1, const float* p2, int n, int distance,
                                 int* use_essential) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(use_essential && n >= 0 && (n == 0 || (p1 && p2)), "uvo_select_estimation_method: bad argument");
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    ctx->scratch.bytes_a.ensure((size_t)std::max(n, 1) * 32 + 1024);
    Arena ar{ctx->scratch.bytes_a.get(), 0, ctx->scratch.bytes_a.n};
    float* d1 = ar.take<float>(2 * (size_t)std::max(n, 1));
    float* d2 = ar.take<float>(2 * (size_t)std::max(n, 1));
    double* disp = ar.take<double>(std::max(n, 1));
    double* out = ar.take<double>(4);
    if (n > 0) {
      UVO_CUDA(cudaMemcpyAsync(d1, p1, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c.stream));
      UVO_CUDA(cudaMemcpyAsync(d2, p2, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c.stream));
    }
    launch_displacements(c, d1, d2, n, disp);
    launch_median(c, disp, nullptr, n, out);
    double med = 0;
    UVO_CUDA(cudaMemcpyAsync(&med, out, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    *use_essential = med < (double)distance ? 0 : 1;  // "BASELINE IS TOO LOW. USING HOMOGRAPHY!" when below
  });
}

int uvo_scale_factor(uvo_ctx* ctx, const double* pts, int n, const double R[9], const double t[3], float range,
                     double* scale_factor) {
  return uvo_scale_factor_front(ctx, pts, n, R, t, range, scale_factor, nullptr);
}

int uvo_scale_factor_front(uvo_ctx* ctx, const double* pts, int n, const double R[9], const double t[3], float range,
                           double* scale_factor, int* n_front) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(scale_factor && R && t && n >= 0 && (n == 0 || pts), "uvo_scale_factor: bad argument");
    *scale_factor = 0.0;
    if (n_front) *n_front = 0;
    if (n == 0) return;
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    ctx->scratch.bytes_a.ensure((size_t)n * 40 + 1024);
    Arena ar{ctx->scratch.bytes_a.get(), 0, ctx->scratch.bytes_a.n};
    double* dp = ar.take<double>(3 * (size_t)n);
    double* dz = ar.take<double>(n);
    double* out = ar.take<double>(4);
    int* cnt = ar.take<int>(1);
    UVO_CUDA(cudaMemcpyAsync(dp, pts, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c.stream));
    launch_front_z(c, dp, n, R, t, dz, cnt);
    launch_median(c, dz, cnt, n, out);
    double med = 0;
    int m = 0;
    UVO_CUDA(cudaMemcpyAsync(&med, out, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaMemcpyAsync(&m, cnt, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    // compute_scale_factor(float distance, ...): distance / Zmedian, 0.0 when no point is in front of the camera
    *scale_factor = m > 0 ? (double)range / med : 0.0;
    if (n_front) *n_front = m;
  });
}

int uvo_solve_pnp_ransac(uvo_ctx* ctx, const double* X, const float* x, int n, const double K[4], int iterations,
                         float reproj_err, double confidence, double rvec[3], double tvec[3], int32_t* inliers,
                         int* n_inliers, int* hyps_evaluated) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    UVO_REQUIRE(K && rvec && tvec && n_inliers && n >= 0, "uvo_solve_pnp_ransac: bad argument");
    UVO_REQUIRE(confidence > 0 && confidence < 1, "uvo_solve_pnp_ransac: confidence must be in (0,1)");  // CV_Assert
    *n_inliers = 0;
    if (hyps_evaluated) *hyps_evaluated = 0;
    if (n < 5) return;  // OpenCV asserts npoints >= 4 and uses P3P for 4: outside the reference configuration
    UVO_REQUIRE(X && x && inliers, "uvo_solve_pnp_ransac: null buffer");
    Ctx& c = ctx->c;
    UVO_CUDA(cudaSetDevice(c.device));
    const int iters = std::max(iterations, 1);
    ctx->scratch.bytes_a.ensure((size_t)n * 48 + pnp_scratch_bytes(n, iters) + 16384);
    Arena ar{ctx->scratch.bytes_a.get(), 0, ctx->scratch.bytes_a.n};
    PnpArgs a{};
    double* dX = ar.take<double>(3 * (size_t)n);
    float* dx = ar.take<float>(2 * (size_t)n);
    a.X = dX;
    a.x = dx;
    a.n = n;
    memcpy(a.K, K, sizeof(a.K));
    a.iterations = iterations;
    a.reproj_err = reproj_err;
    a.confidence = confidence;
    a.min_points = -1;
    a.result = ar.take<double>(8);
    a.inliers = ar.take<int32_t>(n);
    a.n_inliers = ar.take<int>(1);
    a.hyps = ar.take<int>(1);
    pnp_bind_scratch(a, ar.take<uint8_t>(pnp_scratch_bytes(n, iters)), n, iters);
    UVO_CUDA(cudaMemcpyAsync(dX, X, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c.stream));
    UVO_CUDA(cudaMemcpyAsync(dx, x, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c.stream));
    if (ctx->pnp_profile) {
      a.prof = ar.take<long long>(32);
      UVO_CUDA(cudaMemsetAsync(a.prof, 0, sizeof(long long) * 32, c.stream));
    }
    launch_pnp_ransac(c, a);
    if (a.prof)
      UVO_CUDA(cudaMemcpyAsync(ctx->pnp_stamps, a.prof, sizeof(long long) * 32, cudaMemcpyDeviceToHost, c.stream));
    double res[8];
    int cnt[2];
    UVO_CUDA(cudaMemcpyAsync(res, a.result, sizeof(double) * 7, cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaMemcpyAsync(&cnt[0], a.n_inliers, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaMemcpyAsync(&cnt[1], a.hyps, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < 3; i++) {
      rvec[i] = res[i];
      tvec[i] = res[3 + i];
    }
    if (hyps_evaluated) *hyps_evaluated = cnt[1];
    if (res[6] != 0.0 && cnt[0] > 0) {
      UVO_CUDA(cudaMemcpyAsync(inliers, a.inliers, sizeof(int32_t) * cnt[0], cudaMemcpyDeviceToHost, c.stream));
      UVO_CUDA(cudaStreamSynchronize(c.stream));
      *n_inliers = cnt[0];
    }
  });
}

}  // extern "C"
