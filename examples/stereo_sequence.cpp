// examples/stereo_sequence.cpp -- a C++ consumer of the C ABI (include/uvo_c.h) with no ROS and no OpenCV: the frame
// loop of visual_odometry_node::stereo_VO (reference uvo/include/visual_odometry.h:526-740) over raw image files.
//
//   stereo_sequence <width> <height> <dt> <left_%04d.bgr pattern> <right_%04d.bgr pattern> <first> <count>
//
// Each file is one interleaved 8-bit BGR frame, width * height * 3 bytes.  Camera constants are the shipped stereo
// calibration (uvo/config/stereo_VO_intrinsics.yaml:7-53) scaled to <width> the way resize_camera_matrix does
// (VO_utility.cpp:658-675).  Frames are kept in flight (uvo_stereo_enqueue_host / uvo_stereo_collect), one result line
// per frame: validity, counts, velocity.  Build:
//   g++ -std=c++14 -O2 -Iinclude examples/stereo_sequence.cpp -Lergo_uvo_b200 -luvo_b200 -Wl,-rpath,$PWD/ergo_uvo_b200
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "uvo_c.h"

static void die(uvo_ctx* ctx, const char* what, int rc) {
  std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, ctx ? uvo_last_error(ctx) : "");
  std::exit(1);
}

static bool read_frame(const char* pattern, int index, uint8_t* dst, size_t bytes) {
  char path[1024];
  std::snprintf(path, sizeof(path), pattern, index);
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  const size_t n = std::fread(dst, 1, bytes, f);
  std::fclose(f);
  return n == bytes;
}

int main(int argc, char** argv) {
  if (argc != 8) {
    std::fprintf(stderr,
                 "usage: %s <width> <height> <dt> <left pattern, e.g. left_%%04d.bgr> <right pattern> <first> <count>\n"
                 "library: %s\n",
                 argv[0], uvo_version());
    return argc == 1 ? 0 : 2;
  }
  const int w = std::atoi(argv[1]), h = std::atoi(argv[2]);
  const double dt = std::atof(argv[3]);
  const int first = std::atoi(argv[6]), count = std::atoi(argv[7]);
  if (w <= 0 || h <= 0 || dt <= 0 || count <= 0) return 2;

  // the shipped calibration (uvo/config/stereo_VO_intrinsics.yaml:7-53; a ~1280-pixel-wide sensor), scaled to the working
  // width on the host; pass your own numbers for another camera
  const int calib_w = 1280, calib_h = (int)((long)h * 1280 / w);
  double KL[9] = {1.335036735254999e+03, 0, 0.644564474737301e+03, 0, 1.332419247540885e+03, 0.357685235527149e+03, 0, 0, 1};
  double KR[9] = {1.330461901943011e+03, 0, 0.684598875987595e+03, 0, 1.328225165048530e+03, 0.382841174819059e+03, 0, 0, 1};
  const double DL[4] = {0.475667186716851, 0.126480045385593, 0.0, 0.0};
  const double DR[4] = {0.493006394402676, 0.037112494470407, 0.0, 0.0};
  const double R_right[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t_right[3] = {-0.33, 0.0, 0.0};
  double newKL[9], newKR[9];
  int ow = 0, oh = 0;
  int rc = uvo_resize_camera_matrix(calib_w, calib_h, w, KL, DL, newKL, &ow, &oh);
  if (rc == UVO_OK) rc = uvo_resize_camera_matrix(calib_w, calib_h, w, KR, DR, newKR, &ow, &oh);
  if (rc != UVO_OK) die(nullptr, "uvo_resize_camera_matrix", rc);
  const uvo_camera camL = {KL[0], KL[4], KL[2], KL[5], DL[0], DL[1], DL[2], DL[3], newKL[0], newKL[4], newKL[2], newKL[5]};
  const uvo_camera camR = {KR[0], KR[4], KR[2], KR[5], DR[0], DR[1], DR[2], DR[3], newKR[0], newKR[4], newKR[2], newKR[5]};

  uvo_ctx* ctx = nullptr;
  rc = uvo_ctx_create(0, nullptr, &ctx);
  if (rc != UVO_OK) die(nullptr, "uvo_ctx_create (a B200 is required: there is no CPU fallback)", rc);
  uvo_params prm;
  uvo_default_params(/*stereo=*/1, &prm);
  uvo_stereo* vo = nullptr;
  rc = uvo_stereo_create(ctx, w, h, &camL, &camR, R_right, t_right, &prm, &vo);
  if (rc != UVO_OK) die(ctx, "uvo_stereo_create", rc);

  // one pinned image pair per frame in flight: a buffer is reused only after its frame has been collected
  const size_t bytes = (size_t)w * h * 3;
  const int lanes = uvo_stereo_max_in_flight() < 8 ? uvo_stereo_max_in_flight() : 8;
  std::vector<uint8_t*> bufL(lanes), bufR(lanes);
  for (int i = 0; i < lanes; i++) {
    bufL[i] = (uint8_t*)uvo_host_alloc(bytes);
    bufR[i] = (uint8_t*)uvo_host_alloc(bytes);
    if (!bufL[i] || !bufR[i]) die(ctx, "uvo_host_alloc", UVO_ERR_CUDA);
  }
  int enqueued = 0, collected = 0;
  auto collect_one = [&]() {
    uvo_stereo_result r;
    const int rc2 = uvo_stereo_collect(vo, &r);
    if (rc2 != UVO_OK) die(ctx, "uvo_stereo_collect", rc2);
    std::printf("frame %d valid %d gate %d keypoints %d/%d matches %d/%d points3d %d inliers %d velocity %.6f %.6f %.6f\n",
                first + collected, r.valid, r.gate, r.n_left, r.n_right, r.n_stereo_matches, r.n_temporal_matches, r.n_3d,
                r.n_inliers, r.velocity[0], r.velocity[1], r.velocity[2]);
    collected++;
  };
  for (int k = 0; k < count; k++) {
    if (enqueued - collected == lanes) collect_one();
    const int slot = k % lanes;
    if (!read_frame(argv[4], first + k, bufL[slot], bytes) || !read_frame(argv[5], first + k, bufR[slot], bytes)) {
      std::fprintf(stderr, "frame %d: cannot read %zu bytes per image\n", first + k, bytes);
      break;
    }
    rc = uvo_stereo_enqueue_host(vo, bufL[slot], bufR[slot], (size_t)w * 3, dt);
    if (rc != UVO_OK) die(ctx, "uvo_stereo_enqueue_host", rc);
    enqueued++;
  }
  while (collected < enqueued) collect_one();
  uvo_stereo_destroy(vo);
  for (int i = 0; i < lanes; i++) {
    uvo_host_free(bufL[i]);
    uvo_host_free(bufR[i]);
  }
  uvo_ctx_destroy(ctx);
  return 0;
}
