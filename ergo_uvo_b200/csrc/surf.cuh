// surf.cuh -- internal interface of the SURF kernels (K4-K7): Hessian box-filter pyramid fused with
// 3x3x3 non-max suppression + interpolation, total-order sort, (orientation) and 64-d descriptors.
// Replaces SURF::create(...)->detectAndCompute (reference VO_utility.cpp:117-118).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace uvo {

constexpr int SURF_MAX_LAYERS = 5;   // nOctaveLayers + 2
constexpr int SURF_MAX_OCTAVES = 4;
constexpr int SURF_TILE_W = 32, SURF_TILE_H = 16;

struct SurfBox {
  int p0, p1, p2, p3;  // offsets into the integral image (row stride w+1), as resizeHaarPattern computes them
  float w;
};

struct SurfLayer {
  int size, margin, samples_i, samples_j;
  SurfBox box[10];  // Dx[0..2], Dy[0..2], Dxy[0..3]
  // The three Dxx (Dyy) boxes share edges, so their 12 corners are 8 distinct integral samples: two rows x four
  // columns (four rows x two columns).  Offsets as resizeHaarPattern rounds them, rows pre-multiplied by w+1.
  int xx_row[2], xx_col[4];
  int yy_row[4], yy_col[2];
};

struct SurfOctave {
  int step, lrows, lcols, tiles_x, tiles_y, tile_begin;  // tile_begin: first block index of this octave
  int nms_margin[SURF_MAX_LAYERS];                         // for middle layers 1..n_layers
  SurfLayer layer[SURF_MAX_LAYERS];
};

// row pitch of the integral image in elements: (w + 1) padded to 16 bytes, which is what a TMA tensor map needs of a
// global row stride (the octave-0 tiles of k_surf_detect arrive by cp.async.bulk.tensor)
static inline int surf_sum_pitch(int w) { return (w + 1 + 3) & ~3; }

struct SurfGeom {
  int w, h, n_octaves, n_layers, total_tiles;
  int spitch;  // surf_sum_pitch(w)
  float thr;
  float thr_skip;  // samples whose screened fl(ax ay) is <= this cannot pass `det > thr` (-inf: screen disabled)
  SurfOctave oct[SURF_MAX_OCTAVES];
};

SurfGeom make_surf_geom(int w, int h, double hessian_threshold, int n_octaves, int n_layers);

// Per-image device state of one detectAndCompute.  All buffers are sized once for `capacity` keypoints.
struct SurfImage {
  const uint8_t* img = nullptr;   // gray image (device)
  size_t pitch = 0;
  const int32_t* sum = nullptr;   // integral, (h+1) rows of surf_sum_pitch(w) ints (w+1 used)
  uvo_keypoint* raw = nullptr;    // unordered detections
  uvo_keypoint* kps = nullptr;    // sorted (OpenCV order), compacted
  float* desc = nullptr;          // capacity x 64 (capacity x 128 when extended)
  uint8_t* patch = nullptr;       // capacity x 448: the 21x21 u8 descriptor patches (k_surf_patch -> k_surf_vector)
  int* rank = nullptr;            // capacity ints, zeroed per frame (rank sort accumulator)
  int* counters = nullptr;        // [0] raw count (may exceed capacity => overflow), [1] final count, [2..3] spare
};

struct SurfBatch {
  SurfImage im[2];
  int n_img;
};

// TMA descriptors of the two integral images (2-D int32, (w+1) x (h+1), boxes of 64 x 45: one octave-0 tile); built
// once per front end, passed to k_surf_detect as a __grid_constant__ parameter
struct alignas(64) SurfMaps {
  CUtensorMap sum[2];
};
SurfMaps make_surf_maps(const int32_t* sum0, const int32_t* sum1, int w, int h);

// detection (pyramid + NMS + interpolation) for n_img images of identical geometry, appends to raw[]/counters[0]
void launch_surf_detect(Ctx& c, const SurfGeom& g, const SurfBatch& b, const SurfMaps& maps, int capacity);
// rank sort raw -> kps in KeypointGreater order, sets counters[1] = min(counters[0], capacity)
void launch_surf_sort(Ctx& c, const SurfBatch& b, int capacity);
// upright/oriented descriptors (64-d, or 128-d rows when `extended`) for kps[0..counters[1]); sets angle; marks
// deleted keypoints with size=-1
void launch_surf_describe(Ctx& c, const SurfGeom& g, const SurfBatch& b, int capacity, int upright, int extended);
// order-preserving removal of keypoints with size <= 0 (only needed when describe can delete: oriented mode or
// images smaller than the largest gradient wavelet)
// (dd = floats per descriptor row, 64 or 128)
void launch_surf_compact(Ctx& c, const SurfBatch& b, int capacity, uvo_keypoint* tmp_kps, float* tmp_desc, int dd);

}  // namespace uvo
