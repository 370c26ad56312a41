// imgprep.cuh -- internal interface of the image-preparation kernels (K1-K3).
#pragma once
#include "common.cuh"

namespace uvo {

struct UndistortParams {
  double ir[9];            // inverse of the new camera matrix
  double fx, fy, u0, v0;   // original camera matrix
  double k1, k2, p1, p2;   // distortion
};

struct ClaheGeom {
  int tiles_x, tiles_y, tw, th, clip;
  float lut_scale, inv_tw, inv_th;
};

__host__ __device__ static inline int div_up_dev(int a, int b) { return (a + b - 1) / b; }

UndistortParams make_undistort_params(const uvo_camera& cam);
ClaheGeom make_clahe_geom(int w, int h, double clip_limit, int tiles_x, int tiles_y);

void launch_gray_undistort(Ctx& c, const uint8_t* d_src3, size_t spitch, int w, int h, const UndistortParams& P,
                           uint8_t* d_dst, size_t dpitch);
// cvtColor(COLOR_BayerBGGR2BGR): 1-channel bayer (w, h >= 3) -> 3-channel interleaved BGR
void launch_demosaic_bggr(Ctx& c, const uint8_t* d_src, size_t spitch, int w, int h, uint8_t* d_dst3, size_t dpitch);
// d_lut: tiles*256 u8 scratch; src and dst may alias
void launch_clahe(Ctx& c, const uint8_t* d_src, size_t spitch, int w, int h, const ClaheGeom& g, uint8_t* d_lut,
                  uint8_t* d_dst, size_t dpitch);
// K0: cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_AREA) for interleaved u8 images with `cn` channels (the
// pre-scaling branch of get_image, reference VO_utility.cpp:362-363).  d_tab: (dw + dh) AreaCell scratch.
struct AreaCell {  // one destination index of computeResizeAreaTab: optional left partial cell, full cells
  int sx1, sx2;    // [sx1, sx2), optional right partial cell
  int has_l, has_r;
  float a_l, a_f, a_r;
};
void launch_resize_area(Ctx& c, const uint8_t* d_src, size_t spitch, int sw, int sh, int cn, uint8_t* d_dst,
                        size_t dpitch, int dw, int dh, AreaCell* d_tab);
// get_image (gray + undistort + optional CLAHE) and the integral image for both images of a stereo pair, batched
// part: the kernel that reads the source images (the only one whose arguments change from frame to frame) and the
// rest can be launched separately, so that the rest can live in a CUDA graph
enum { PREP_PART_SOURCE = 1, PREP_PART_REST = 2 };
void launch_prep_pair(Ctx& c, const uint8_t* d_src3[2], size_t spitch, int w, int h, const UndistortParams P[2], int clahe,
                      const ClaheGeom& g, uint8_t* d_lut, uint8_t* d_gray[2], size_t gpitch,
                      int32_t* d_sum[2], int part = PREP_PART_SOURCE | PREP_PART_REST, int sum_pitch = 0);
// d_sum: (h+1) rows of sum_pitch int32 (w+1 used; sum_pitch = 0: dense, pitch w+1)
void launch_integral(Ctx& c, const uint8_t* d_img, size_t pitch, int w, int h, int32_t* d_sum, int sum_pitch = 0);

}  // namespace uvo
