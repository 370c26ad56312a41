// frontend.cuh -- device-resident per-image front end: K1 gray+undistort -> K2 CLAHE -> K3 integral -> K4-K7 SURF
// for one or two images of identical size (the stereo pair is batched through every kernel: blockIdx.y = image).
#pragma once
#include "common.cuh"
#include "imgprep.cuh"
#include "surf.cuh"

namespace uvo {

struct FrontEnd {
  int w = 0, h = 0, n_img = 0, capacity = 0;
  size_t gpitch = 0;
  int sum_pitch = 0;  // integral row pitch (elements)
  SurfMaps maps{};
  const int32_t* maps_for[2] = {nullptr, nullptr};
  int maps_w = 0, maps_h = 0;
  SurfGeom geom{};
  bool geom_valid = false;
  double geom_thr = -1;
  int geom_oct = 0, geom_lay = 0;
  DevBuf<uint8_t> gray[2];
  DevBuf<int32_t> sum[2];
  DevBuf<uvo_keypoint> raw[2], kps[2];
  DevBuf<float> desc[2];
  DevBuf<uint8_t> patch[2];
  DevBuf<int> counters;  // 4 ints per image
  DevBuf<int> rank[2];
  DevBuf<uint8_t> lut;
  DevBuf<uvo_keypoint> tmp_kps;
  DevBuf<float> tmp_desc;

  void init(int w_, int h_, int n_img_, int capacity_) {
    UVO_REQUIRE(w_ > 0 && h_ > 0 && (n_img_ == 1 || n_img_ == 2) && capacity_ > 0, "FrontEnd::init: bad geometry");
    if (w_ != w || h_ != h) geom_valid = false;
    w = w_;
    h = h_;
    n_img = std::max(n_img, n_img_);
    capacity = std::max(capacity, capacity_);
    gpitch = ((size_t)w + 15) & ~(size_t)15;
    sum_pitch = surf_sum_pitch(w);
    for (int i = 0; i < n_img; i++) {
      gray[i].ensure(gpitch * h);
      sum[i].ensure((size_t)sum_pitch * (h + 1));
      raw[i].ensure(capacity);
      kps[i].ensure(capacity);
      // sized for extended (128-d) rows; the row stride in use is 64 unless surf() is asked for extended descriptors
      const bool fresh = desc[i].n < (size_t)capacity * 128;
      desc[i].ensure((size_t)capacity * 128);
      // rows past the live count are read (and ignored) by the matcher's TMA tiles: keep them finite
      if (fresh) UVO_CUDA(cudaMemset(desc[i].get(), 0, (size_t)capacity * 128 * sizeof(float)));
      rank[i].ensure(capacity);
      patch[i].ensure((size_t)capacity * 448);
    }
    counters.ensure(8);
    lut.ensure(2 * 64 * 256);
    if (maps_for[0] != sum[0].get() || maps_for[1] != sum[n_img > 1 ? 1 : 0].get() || maps_w != w || maps_h != h) {
      maps_w = w;
      maps_h = h;
      maps_for[0] = sum[0].get();
      maps_for[1] = sum[n_img > 1 ? 1 : 0].get();
      maps = make_surf_maps(maps_for[0], maps_for[1], w, h);
    }
  }

  SurfBatch batch(int first, int count) const {
    SurfBatch b{};
    b.n_img = count;
    for (int i = 0; i < count; i++) {
      SurfImage& im = b.im[i];
      im.img = gray[first + i].get();
      im.pitch = gpitch;
      im.sum = sum[first + i].get();
      im.raw = raw[first + i].get();
      im.kps = kps[first + i].get();
      im.desc = desc[first + i].get();
      im.patch = patch[first + i].get();
      im.rank = rank[first + i].get();
      im.counters = counters.get() + 4 * (first + i);
    }
    return b;
  }

  // get_image for image `idx` from a device-resident 3-channel source
  void prep(Ctx& c, int idx, const uint8_t* d_src3, size_t spitch, const uvo_camera& cam, int clahe, int clip_limit) {
    launch_gray_undistort(c, d_src3, spitch, w, h, make_undistort_params(cam), gray[idx].get(), gpitch);
    if (clahe) {
      ClaheGeom g = make_clahe_geom(w, h, (double)clip_limit, 8, 8);
      launch_clahe(c, gray[idx].get(), gpitch, w, h, g, lut.get() + idx * 64 * 256,
                   gray[idx].get(), gpitch);
    }
  }

  // get_image + integral for slots 0 and 1 together (stereo pair), one launch per kernel
  void prep_pair(Ctx& c, const uint8_t* srcL, const uint8_t* srcR, size_t spitch, const uvo_camera& camL,
                 const uvo_camera& camR, int clahe, int clip_limit, int part = PREP_PART_SOURCE | PREP_PART_REST) {
    const uint8_t* src[2] = {srcL, srcR};
    const UndistortParams P[2] = {make_undistort_params(camL), make_undistort_params(camR)};
    uint8_t* g2[2] = {gray[0].get(), gray[1].get()};
    int32_t* s2[2] = {sum[0].get(), sum[1].get()};
    launch_prep_pair(c, src, spitch, w, h, P, clahe, make_clahe_geom(w, h, (double)clip_limit, 8, 8), lut.get(),
                     g2, gpitch, s2, part, sum_pitch);
  }

  // integral image of slot `idx` (the first step of SURF::detectAndCompute)
  void integral(Ctx& c, int idx) { launch_integral(c, gray[idx].get(), gpitch, w, h, sum[idx].get(), sum_pitch); }

  // SURF::detectAndCompute on images [first, first+count); `with_integral` = false when the caller already ran
  // integral() for those slots (the stereo pipeline does, per image, on two streams)
  void surf(Ctx& c, int first, int count, const uvo_params& p, bool with_integral = true) {
    const int dd = p.surf_extended ? 128 : 64;  // floats per descriptor row (SURF::descriptorSize())
    if (!geom_valid || geom_thr != (double)p.surf_min_hessian || geom_oct != p.surf_octaves ||
        geom_lay != p.surf_octave_layers) {
      geom = make_surf_geom(w, h, (double)p.surf_min_hessian, p.surf_octaves, p.surf_octave_layers);
      geom_valid = true;
      geom_thr = (double)p.surf_min_hessian;
      geom_oct = p.surf_octaves;
      geom_lay = p.surf_octave_layers;
    }
    if (with_integral)
      for (int i = 0; i < count; i++) integral(c, first + i);
    SurfBatch b = batch(first, count);
    SurfMaps m = maps;
    if (first == 1) m.sum[0] = maps.sum[1];
    launch_surf_detect(c, geom, b, m, capacity);
    launch_surf_sort(c, b, capacity);
    launch_surf_describe(c, geom, b, capacity, p.surf_upright, p.surf_extended);
    // describe can delete keypoints only in oriented mode or when the image is smaller than the largest
    // gradient wavelet (2*round(2*264*1.2/9) = 142)
    if (!p.surf_upright || std::min(w, h) + 1 < 142) {
      tmp_kps.ensure((size_t)2 * capacity);
      tmp_desc.ensure((size_t)2 * capacity * dd);
      launch_surf_compact(c, b, capacity, tmp_kps.get(), tmp_desc.get(), dd);
    }
  }
};

}  // namespace uvo
