// capi.cu -- extern "C" entry points of libuvo_b200.so (include/uvo_c.h): context management and the
// stage-level calls that take HOST buffers (the drop-in layer under the reference's VO_utility functions).
// Device-resident whole-frame pipelines live in frame.cu.
#include <cstring>

#include "capi_internal.cuh"

using namespace uvo;

extern "C" {

const char* uvo_version(void) { return "uvo-b200 0.1 (sm_100a)"; }

int uvo_ctx_create(int device, void* cuda_stream, uvo_ctx** out) {
  if (!out) return UVO_ERR_INVALID;
  *out = nullptr;
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
    UVO_REQUIRE(prop.major >= 10, "libuvo_b200 is built for sm_100a only; this device is older");
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
    s.hist.ensure(64 * 256);
    s.lut.ensure(64 * 256);
    UVO_CUDA(cudaMemcpyAsync(s.src3.get(), src3, spitch * h, cudaMemcpyHostToDevice, c.stream));
    launch_gray_undistort(c, s.src3.get(), spitch, w, h, make_undistort_params(*cam), s.gray.get(), gp);
    if (clahe) {
      ClaheGeom g = make_clahe_geom(w, h, (double)clip_limit, 8, 8);
      launch_clahe(c, s.gray.get(), gp, w, h, g, s.hist.get(), s.lut.get(), s.gray.get(), gp);
    }
    UVO_CUDA(cudaMemcpy2DAsync(dst, dpitch, s.gray.get(), gp, w, h, cudaMemcpyDeviceToHost, c.stream));
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
    launch_integral(c, s.gray.get(), gp, w, h, s.integral.get());
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
      UVO_CUDA(cudaMemcpyAsync(desc, fe.desc[0].get(), sizeof(float) * 64 * n, cudaMemcpyDeviceToHost, c.stream));
      UVO_CUDA(cudaStreamSynchronize(c.stream));
    }
    *count = n;
  });
}

}  // extern "C"
