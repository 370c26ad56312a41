// surf.cu -- K4..K7: SURF detectAndCompute on sm_100a (see surf.cuh).
//
// Behaviour follows OpenCV-contrib's CPU SURF as restated in SURVEY.md Appendix A (the reference only calls it:
// VO_utility.cpp:117-118).  Design differences from the CPU code (none changes a result bit):
//   * the 20 det/trace maps are never written to HBM: one block computes the 5 layers of an octave for a 32x16 tile
//     (+1 halo) straight from the L2-resident integral image into shared memory and runs the 3x3x3 non-max
//     suppression + interpolation on it.  `trace` is not computed at all (its only use, KeyPoint::class_id, is
//     reset to -1 by detectAndCompute);
//   * detections are appended with an atomic counter and put into OpenCV's total order by a rank sort, so the
//     output order is deterministic regardless of the append order;
//   * descriptors: one block per keypoint does window extraction + INTER_AREA 21x21 resize + Haar + 4x4x4 sums.
// Every floating-point step keeps the CPU operation order (compiled with -fmad=false).
#include <cfloat>
#include <mutex>
#include <cmath>

#include "surf.cuh"
#include "tma.cuh"

namespace uvo {

// getGaussianKernel(13, 2.5f, CV_32F) and getGaussianKernel(20, (double)3.3f, CV_32F) -- SURF_DESC_SIGMA is a float
// constant in OpenCV, so sigma is 3.2999999523..., not 3.3 (values checked against cv2 and the
// oracle in tests/test_oracle_surf.py::test_gaussian_tables)
static const float kGOri[13] = {0x1.282748p-7f,  0x1.64ff86p-6f, 0x1.6eb6e4p-5f, 0x1.40ff98p-4f, 0x1.dedf96p-4f,
                                0x1.30623ap-3f,  0x1.49bc24p-3f, 0x1.30623ap-3f, 0x1.dedf96p-4f, 0x1.40ff98p-4f,
                                0x1.6eb6e4p-5f,  0x1.64ff86p-6f, 0x1.282748p-7f};
static const float kGDesc[20] = {0x1.f6898ap-10f, 0x1.1f18f4p-8f, 0x1.2b4116p-7f, 0x1.1c8ee8p-6f, 0x1.edafecp-6f, 0x1.86ae60p-5f, 0x1.1a0a9cp-4f, 0x1.737ecap-4f, 0x1.be639ap-4f, 0x1.eef842p-4f, 0x1.eef842p-4f, 0x1.be639ap-4f, 0x1.737ecap-4f, 0x1.1a0a9cp-4f, 0x1.86ae60p-5f, 0x1.edafecp-6f, 0x1.1c8ee8p-6f, 0x1.2b4116p-7f, 0x1.1f18f4p-8f, 0x1.f6898ap-10f};

__constant__ float c_DW[400];      // DW[i*20+j] = G_desc[i]*G_desc[j]
__constant__ float c_aptw[128];    // orientation sample weights (113 used)
__constant__ signed char c_apt[128][2];
static bool g_tables_uploaded[64] = {};

static void upload_tables(Ctx& c) {
  if (c.device < 64 && g_tables_uploaded[c.device]) return;
  float DW[400];
  for (int i = 0; i < 20; i++)
    for (int j = 0; j < 20; j++) DW[i * 20 + j] = kGDesc[i] * kGDesc[j];
  float aptw[128] = {};
  signed char apt[128][2] = {};
  int n = 0;
  for (int i = -6; i <= 6; i++)
    for (int j = -6; j <= 6; j++)
      if (i * i + j * j <= 36) {
        apt[n][0] = (signed char)i;
        apt[n][1] = (signed char)j;
        aptw[n++] = kGOri[i + 6] * kGOri[j + 6];
      }
  UVO_CUDA(cudaMemcpyToSymbolAsync(c_DW, DW, sizeof(DW), 0, cudaMemcpyHostToDevice, c.stream));
  UVO_CUDA(cudaMemcpyToSymbolAsync(c_aptw, aptw, sizeof(aptw), 0, cudaMemcpyHostToDevice, c.stream));
  UVO_CUDA(cudaMemcpyToSymbolAsync(c_apt, apt, sizeof(apt), 0, cudaMemcpyHostToDevice, c.stream));
  UVO_CUDA(cudaStreamSynchronize(c.stream));
  if (c.device < 64) g_tables_uploaded[c.device] = true;
}

// ------------------------------------------------------------------------------------------------ geometry (host)
static inline int cv_roundf_h(float v) { return (int)nearbyintf(v); }

static void resize_haar_h(const int src[][5], SurfBox* dst, int n, int old_size, int new_size, int width_step) {
  float ratio = (float)new_size / old_size;
  for (int k = 0; k < n; k++) {
    int dx1 = cv_roundf_h(ratio * src[k][0]), dy1 = cv_roundf_h(ratio * src[k][1]);
    int dx2 = cv_roundf_h(ratio * src[k][2]), dy2 = cv_roundf_h(ratio * src[k][3]);
    dst[k].p0 = dy1 * width_step + dx1;
    dst[k].p1 = dy2 * width_step + dx1;
    dst[k].p2 = dy1 * width_step + dx2;
    dst[k].p3 = dy2 * width_step + dx2;
    dst[k].w = src[k][4] / ((float)(dx2 - dx1) * (dy2 - dy1));
  }
}

SurfGeom make_surf_geom(int w, int h, double hessian_threshold, int n_octaves, int n_layers) {
  static const int dx_s[3][5] = {{0, 2, 3, 7, 1}, {3, 2, 6, 7, -2}, {6, 2, 9, 7, 1}};
  static const int dy_s[3][5] = {{2, 0, 7, 3, 1}, {2, 3, 7, 6, -2}, {2, 6, 7, 9, 1}};
  static const int dxy_s[4][5] = {{1, 1, 4, 4, 1}, {5, 1, 8, 4, -1}, {1, 5, 4, 8, -1}, {5, 5, 8, 8, 1}};
  if (n_octaves < 1 || n_octaves > SURF_MAX_OCTAVES || n_layers < 1 || n_layers + 2 > SURF_MAX_LAYERS)
    throw InvalidArg{"SURF: supported range is 1..4 octaves and 1..3 octave layers", UVO_ERR_UNSUPPORTED};
  SurfGeom g{};
  g.w = w;
  g.h = h;
  g.n_octaves = n_octaves;
  g.n_layers = n_layers;
  g.thr = (float)hessian_threshold;
  g.spitch = surf_sum_pitch(w);
  const int ws_pitch = g.spitch;
  int tile_begin = 0;
  for (int o = 0; o < n_octaves; o++) {
    SurfOctave& O = g.oct[o];
    O.step = 1 << o;
    O.lrows = h / O.step;
    O.lcols = w / O.step;
    O.tiles_x = div_up(std::max(O.lcols, 1), SURF_TILE_W);
    O.tiles_y = div_up(std::max(O.lrows, 1), SURF_TILE_H);
    O.tile_begin = tile_begin;
    tile_begin += O.tiles_x * O.tiles_y;
    for (int l = 0; l < n_layers + 2; l++) {
      SurfLayer& L = O.layer[l];
      L.size = (9 + 6 * l) << o;
      L.margin = (L.size / 2) / O.step;
      if (L.size > h || L.size > w) {
        L.samples_i = L.samples_j = 0;  // calcLayerDetAndTrace returns early
      } else {
        L.samples_i = 1 + (h - L.size) / O.step;
        L.samples_j = 1 + (w - L.size) / O.step;
      }
      resize_haar_h(dx_s, L.box + 0, 3, 9, L.size, ws_pitch);
      resize_haar_h(dy_s, L.box + 3, 3, 9, L.size, ws_pitch);
      resize_haar_h(dxy_s, L.box + 6, 4, 9, L.size, ws_pitch);
      {
        const float ratio = (float)L.size / 9;
        const int ws = ws_pitch;
        L.xx_row[0] = cv_roundf_h(ratio * 2) * ws;
        L.xx_row[1] = cv_roundf_h(ratio * 7) * ws;
        L.yy_col[0] = cv_roundf_h(ratio * 2);
        L.yy_col[1] = cv_roundf_h(ratio * 7);
        for (int k = 0; k < 4; k++) {
          L.xx_col[k] = cv_roundf_h(ratio * (3 * k));
          L.yy_row[k] = cv_roundf_h(ratio * (3 * k)) * ws;
        }
      }
    }
    for (int l = 1; l <= n_layers; l++) O.nms_margin[l] = (O.layer[l + 1].size / 2) / O.step + 1;
  }
  g.total_tiles = tile_begin;
  // the screen of k_surf_detect (see screen_tile0): valid when the Dxx / Dyy weights of every middle layer are
  // (w, -2w, w) in f32 and the integer combination stays below 2^24; otherwise nothing is screened out
  bool screen_ok = true;
  for (int o = 0; o < n_octaves; o++)
    for (int l = 1; l <= n_layers; l++) {
      const SurfLayer& L = g.oct[o].layer[l];
      for (int a = 0; a < 6; a += 3) {
        const float wgt = L.box[a].w;
        screen_ok = screen_ok && wgt > 0.f && L.box[a + 2].w == wgt && L.box[a + 1].w == -2.f * wgt &&
                    4.0 * 255.0 / (double)wgt < 16777216.0;
      }
    }
  g.thr_skip = screen_ok ? std::nextafterf(g.thr - 1.0f - fabsf(g.thr) * 1e-6f, -INFINITY) : -INFINITY;
  return g;
}

// ------------------------------------------------------------------------------------------------ K4+K5
__device__ __forceinline__ float haar3(const int* __restrict__ o, const SurfBox* __restrict__ f) {
  double d = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int v = (int)((unsigned)__ldg(o + f[k].p0) + (unsigned)__ldg(o + f[k].p3) - (unsigned)__ldg(o + f[k].p1) -
                  (unsigned)__ldg(o + f[k].p2));
    d = __dadd_rn(d, (double)__fmul_rn((float)v, f[k].w));
  }
  return (float)d;
}
__device__ __forceinline__ float haar4(const int* __restrict__ o, const SurfBox* __restrict__ f) {
  double d = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int v = (int)((unsigned)__ldg(o + f[k].p0) + (unsigned)__ldg(o + f[k].p3) - (unsigned)__ldg(o + f[k].p1) -
                  (unsigned)__ldg(o + f[k].p2));
    d = __dadd_rn(d, (double)__fmul_rn((float)v, f[k].w));
  }
  return (float)d;
}

// Lazy evaluation (LAZY = true): det = dx dy - 0.81 dxy^2 <= fl(dx dy), because the subtrahend (0.81f dxy) dxy is >= +0
// and f32 subtraction is monotonic.  A sample with fl(dx dy) <= thr can therefore never pass `det > thr` as a centre,
// it compares as smaller than any centre that does, and its exact value is needed only for the interpolation of a
// local maximum next to it -- rare, and filled in cooperatively later.  The Dxy half of such a sample (16 of its 32
// corner reads, 4 of its 10 fp64 terms) is skipped and DET_SKIPPED stored.  Every value that is used is exact.
#define DET_SKIPPED (-INFINITY)

template <bool LAZY = false>
__device__ __forceinline__ float det_at(const int* __restrict__ sum, int scols, const SurfLayer& L, int step, int i,
                                        int j, float thr = 0.f) {
  const int si = i - L.margin, sj = j - L.margin;
  if (si < 0 || sj < 0 || si >= L.samples_i || sj >= L.samples_j) return 0.f;  // never-written map border
  const int* o = sum + (size_t)(si * step) * scols + sj * step;
  // Dxx / Dyy: 8 shared corners each (same integers, same f32 products, same f64 sums as box-by-box evaluation)
  unsigned A[4], B[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    A[k] = (unsigned)__ldg(o + L.xx_row[0] + L.xx_col[k]);
    B[k] = (unsigned)__ldg(o + L.xx_row[1] + L.xx_col[k]);
  }
  double d = 0;
#pragma unroll
  for (int k = 0; k < 3; k++)
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(A[k] + B[k + 1] - B[k] - A[k + 1]), L.box[k].w));
  const float dx = (float)d;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    A[k] = (unsigned)__ldg(o + L.yy_row[k] + L.yy_col[0]);
    B[k] = (unsigned)__ldg(o + L.yy_row[k] + L.yy_col[1]);
  }
  d = 0;
#pragma unroll
  for (int k = 0; k < 3; k++)
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(A[k] + B[k + 1] - A[k + 1] - B[k]), L.box[3 + k].w));
  const float dy = (float)d;
  if (LAZY && !(__fmul_rn(dx, dy) > thr)) return DET_SKIPPED;
  const float dxy = haar4(o, L.box + 6);
  return __fsub_rn(__fmul_rn(dx, dy), __fmul_rn(__fmul_rn(0.81f, dxy), dxy));
}

// ---- octave 0 from a shared-memory tile of the integral image ----------------------------------------------------
// With the integral tile in shared memory at a compile-time pitch and the layer size a template parameter, every one
// of the 32 corner reads of a det sample is a single LDS with an immediate offset (the global path spends an LDC, an
// add and a 64-bit multiply-add per corner).  Same integers, same f32 products, same f64 sums as det_at.
__host__ __device__ constexpr int cround_c(float v) {  // cvRound for v >= 0 (ties to even), usable in constant expressions
  const int f = (int)v;
  const float d = v - (float)f;
  return d > 0.5f ? f + 1 : (d < 0.5f ? f : f + (f & 1));
}
template <int SIZE>
__host__ __device__ constexpr int haar_off(int k) {  // resizeHaarPattern: cvRound(ratio * k), ratio = (float)SIZE / 9
  return cround_c((float)SIZE / 9 * (float)k);
}
constexpr int T0_MAXM = 13;                                 // margin of the largest middle layer (size 27) at step 1
// 45 x 61 integral samples inside one TMA box of 45 x 64: the copy engine wants the inner extent AND the inner start
// coordinate to be multiples of 16 bytes (measured: tools/probes/tma_tile_probe.cu -- a start column that is not a
// multiple of 4 ints faults), so the box starts T0_XOFF = 2 columns left of the first sample the tile needs
// (tile column 0 = image column tj0 - 16, tj0 a multiple of 32)
constexpr int T0_ROWS = SURF_TILE_H + 2 + 27, T0_COLS = 64, T0_XOFF = 2;
static_assert(SURF_TILE_W + 2 + 27 + T0_XOFF <= T0_COLS, "octave-0 tile does not fit its TMA box");
static_assert((SURF_TILE_W % 4) == 0 && ((1 + T0_MAXM + T0_XOFF) % 4) == 0, "TMA start column must be 16-byte aligned");

template <int SIZE, bool LAZY = false>
__device__ __forceinline__ float det_tile0(const int* __restrict__ T, const SurfLayer& L, int i, int j, int y, int x,
                                           float thr = 0.f) {
  constexpr int M = SIZE / 2;
  const int si = i - M, sj = j - M;
  if (si < 0 || sj < 0 || si >= L.samples_i || sj >= L.samples_j) return 0.f;  // never-written map border
  const int* o = T + (y + T0_MAXM - M) * T0_COLS + (x + T0_MAXM - M + T0_XOFF);
  constexpr int c0 = haar_off<SIZE>(0), c1 = haar_off<SIZE>(1), c2 = haar_off<SIZE>(2), c3 = haar_off<SIZE>(3),
                c4 = haar_off<SIZE>(4), c5 = haar_off<SIZE>(5), c6 = haar_off<SIZE>(6), c7 = haar_off<SIZE>(7),
                c8 = haar_off<SIZE>(8), c9 = haar_off<SIZE>(9);
  auto at = [&](int r, int c) -> unsigned { return (unsigned)o[r * T0_COLS + c]; };
  double d = 0;
  {
    const unsigned A0 = at(c2, c0), A1 = at(c2, c3), A2 = at(c2, c6), A3 = at(c2, c9);
    const unsigned B0 = at(c7, c0), B1 = at(c7, c3), B2 = at(c7, c6), B3 = at(c7, c9);
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(A0 + B1 - B0 - A1), L.box[0].w));
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(A1 + B2 - B1 - A2), L.box[1].w));
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(A2 + B3 - B2 - A3), L.box[2].w));
  }
  const float dx = (float)d;
  d = 0;
  {
    const unsigned A0 = at(c0, c2), A1 = at(c3, c2), A2 = at(c6, c2), A3 = at(c9, c2);
    const unsigned B0 = at(c0, c7), B1 = at(c3, c7), B2 = at(c6, c7), B3 = at(c9, c7);
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(A0 + B1 - A1 - B0), L.box[3].w));
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(A1 + B2 - A2 - B1), L.box[4].w));
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(A2 + B3 - A3 - B2), L.box[5].w));
  }
  const float dy = (float)d;
  if (LAZY && !(__fmul_rn(dx, dy) > thr)) return DET_SKIPPED;
  d = 0;
  {
    // Dxy boxes {1,1,4,4} {5,1,8,4} {1,5,4,8} {5,5,8,8} as (x1, y1, x2, y2): p0 + p3 - p1 - p2
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(at(c1, c1) + at(c4, c4) - at(c4, c1) - at(c1, c4)), L.box[6].w));
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(at(c1, c5) + at(c4, c8) - at(c4, c5) - at(c1, c8)), L.box[7].w));
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(at(c5, c1) + at(c8, c4) - at(c8, c1) - at(c5, c4)), L.box[8].w));
    d = __dadd_rn(d, (double)__fmul_rn((float)(int)(at(c5, c5) + at(c8, c8) - at(c8, c5) - at(c5, c8)), L.box[9].w));
  }
  const float dxy = (float)d;
  return __fsub_rn(__fmul_rn(dx, dy), __fmul_rn(__fmul_rn(0.81f, dxy), dxy));
}

// ---- screening ---------------------------------------------------------------------------------------------------
// Most samples fail the lazy test fl(dx dy) > thr, but a warp pays for the exact evaluation (f64 sums, and the Dxy
// half) as soon as ONE of its lanes passes.  So the tile is evaluated in two passes: a cheap screen of every sample,
// then the exact evaluation of the survivors only, packed densely over the block.
// The screen: the three Dxx boxes have equal areas (3 * size / 9 is an integer), so their weights are w, -2w, w with
// the SAME f32 w (the host checks it, screen_weights_ok) and
//     ax = fl((v0 - 2 v1 + v2) w)      (the integer is exact; |v0 - 2 v1 + v2| < 2^24 up to layer size 195)
// differs from the exact dx = fl(fl(v0 w) - fl(2 v1 w) + fl(v2 w)) by at most 2^-24 (v0 + 2 v1 + v2) w + 2 * 2^-24 * 510
// <= 1.3e-4 for 8-bit images (v w <= 255, |dx| <= 510); likewise ay.  Hence |dx dy - ax ay| <= 2 * 510 * 1.3e-4 = 0.13,
// and with the f32 roundings of the two products (<= 0.016 each) a sample with ax ay <= thr - 1 can never have
// fl(dx dy) > thr: it is DET_SKIPPED exactly as the lazy rule would have made it.  Everything else is evaluated
// exactly.  The stored values are bit for bit those of the one-pass evaluation.
__device__ __forceinline__ int screen_combine(unsigned A0, unsigned A1, unsigned A2, unsigned A3, unsigned B0,
                                              unsigned B1, unsigned B2, unsigned B3) {
  // v0 - 2 v1 + v2 with v_k = A_k + B_{k+1} - B_k - A_{k+1}
  return (int)((A0 - A3 - B0 + B3) + 3u * (A2 - A1 + B1 - B2));
}
enum { SCREEN_OUTSIDE = 0, SCREEN_SKIP = 1, SCREEN_KEEP = 2 };

template <int SIZE>
__device__ __forceinline__ int screen_tile0(const int* __restrict__ T, const SurfLayer& L, int i, int j, int y, int x,
                                            float thr_skip) {
  constexpr int M = SIZE / 2;
  const int si = i - M, sj = j - M;
  if (si < 0 || sj < 0 || si >= L.samples_i || sj >= L.samples_j) return SCREEN_OUTSIDE;
  const int* o = T + (y + T0_MAXM - M) * T0_COLS + (x + T0_MAXM - M + T0_XOFF);
  constexpr int c0 = haar_off<SIZE>(0), c2 = haar_off<SIZE>(2), c3 = haar_off<SIZE>(3), c6 = haar_off<SIZE>(6),
                c7 = haar_off<SIZE>(7), c9 = haar_off<SIZE>(9);
  auto at = [&](int r, int c) -> unsigned { return (unsigned)o[r * T0_COLS + c]; };
  const int sx = screen_combine(at(c2, c0), at(c2, c3), at(c2, c6), at(c2, c9), at(c7, c0), at(c7, c3), at(c7, c6),
                                at(c7, c9));
  const int sy = screen_combine(at(c0, c2), at(c3, c2), at(c6, c2), at(c9, c2), at(c0, c7), at(c3, c7), at(c6, c7),
                                at(c9, c7));
  const float ax = __fmul_rn((float)sx, L.box[0].w), ay = __fmul_rn((float)sy, L.box[3].w);
  return __fmul_rn(ax, ay) > thr_skip ? SCREEN_KEEP : SCREEN_SKIP;
}

__device__ __forceinline__ int screen_at(const int* __restrict__ sum, int scols, const SurfLayer& L, int step, int i,
                                         int j, float thr_skip) {
  const int si = i - L.margin, sj = j - L.margin;
  if (si < 0 || sj < 0 || si >= L.samples_i || sj >= L.samples_j) return SCREEN_OUTSIDE;
  const int* o = sum + (size_t)(si * step) * scols + sj * step;
  unsigned A[4], B[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    A[k] = (unsigned)__ldg(o + L.xx_row[0] + L.xx_col[k]);
    B[k] = (unsigned)__ldg(o + L.xx_row[1] + L.xx_col[k]);
  }
  const int sx = screen_combine(A[0], A[1], A[2], A[3], B[0], B[1], B[2], B[3]);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    A[k] = (unsigned)__ldg(o + L.yy_row[k] + L.yy_col[0]);
    B[k] = (unsigned)__ldg(o + L.yy_row[k] + L.yy_col[1]);
  }
  const int sy = screen_combine(A[0], A[1], A[2], A[3], B[0], B[1], B[2], B[3]);
  const float ax = __fmul_rn((float)sx, L.box[0].w), ay = __fmul_rn((float)sy, L.box[3].w);
  return __fmul_rn(ax, ay) > thr_skip ? SCREEN_KEEP : SCREEN_SKIP;
}

// interpolateKeypoint: 3x3 Cramer solve in f32 (Matx33f::solve(DECOMP_LU))
__device__ __forceinline__ bool interpolate_keypoint(const float N[3][9], int step, int ds, float& px, float& py,
                                                     float& psize) {
  const float b0 = -(N[1][5] - N[1][3]) / 2, b1 = -(N[1][7] - N[1][1]) / 2, b2 = -(N[2][4] - N[0][4]) / 2;
  const float a00 = N[1][3] - 2 * N[1][4] + N[1][5];
  const float a01 = (N[1][8] - N[1][6] - N[1][2] + N[1][0]) / 4;
  const float a02 = (N[2][5] - N[2][3] - N[0][5] + N[0][3]) / 4;
  const float a10 = a01;
  const float a11 = N[1][1] - 2 * N[1][4] + N[1][7];
  const float a12 = (N[2][7] - N[2][1] - N[0][7] + N[0][1]) / 4;
  const float a20 = a02, a21 = a12;
  const float a22 = N[0][4] - 2 * N[1][4] + N[2][4];
  float d = a00 * (a11 * a22 - a21 * a12) - a01 * (a10 * a22 - a20 * a12) + a02 * (a10 * a21 - a20 * a11);
  float x0 = 0, x1 = 0, x2 = 0;
  if (d != 0) {
    d = 1 / d;
    x0 = d * (b0 * (a11 * a22 - a12 * a21) - a01 * (b1 * a22 - a12 * b2) + a02 * (b1 * a21 - a11 * b2));
    x1 = d * (a00 * (b1 * a22 - a12 * b2) - b0 * (a10 * a22 - a12 * a20) + a02 * (a10 * b2 - b1 * a20));
    x2 = d * (a00 * (a11 * b2 - b1 * a21) - a01 * (a10 * b2 - b1 * a20) + b0 * (a10 * a21 - a11 * a20));
  }
  const bool ok = (x0 != 0 || x1 != 0 || x2 != 0) && fabsf(x0) <= 1 && fabsf(x1) <= 1 && fabsf(x2) <= 1;
  if (ok) {
    px += x0 * step;
    py += x1 * step;
    psize = (float)__float2int_rn(psize + x2 * ds);
  }
  return ok;
}

constexpr int TW = SURF_TILE_W, TH = SURF_TILE_H;

// Lazy pyramid: only the middle layers 1..n_layers are evaluated everywhere (a keypoint needs det > threshold in a
// middle layer).  The two outer layers are needed solely as NMS neighbours of the few positions that already beat
// their 8 in-layer neighbours and the adjacent middle layers, so they are evaluated on demand, 9 samples per such
// candidate, spread over the block.  Every det value that is computed is computed exactly as before.
constexpr int DET_LIST = 192;  // candidates needing an outer layer, per block (overflow is handled in-thread)

struct DetCand {
  short m, y, x, alive;
};

__device__ __noinline__ void emit_keypoint(const SurfOctave& O, const SurfImage& im, const float N[3][9], int m,
                                              int i, int j, int o, int capacity) {
  const int step = O.step;
  const int size = O.layer[m].size;
  const int sum_i = step * (i - (size / 2) / step), sum_j = step * (j - (size / 2) / step);
  float py = (float)sum_i + (float)(size - 1) * 0.5f;
  float px = (float)sum_j + (float)(size - 1) * 0.5f;
  float psize = (float)size;
  const int ds = size - O.layer[m - 1].size;
  if (!interpolate_keypoint(N, step, ds, px, py, psize)) return;
  const int slot = atomicAdd(&im.counters[0], 1);
  if (slot < capacity) {
    uvo_keypoint k;
    k.x = px;
    k.y = py;
    k.size = psize;
    k.angle = -1.f;
    k.response = N[1][4];
    k.octave = o;
    k.class_id = -1;
    im.raw[slot] = k;
  }
}

#ifndef UVO_DET_MINB
#define UVO_DET_MINB 6
#endif
#ifndef UVO_DET_SCREEN
#define UVO_DET_SCREEN 0  // two-pass evaluation (screen, then packed survivors): exact, but measured slower -- see DESIGN 4
#endif
__global__ void __launch_bounds__(256, UVO_DET_MINB) k_surf_detect(const __grid_constant__ SurfMaps maps, const __grid_constant__ SurfGeom g,
                                                     const __grid_constant__ SurfBatch b, int capacity) {
  // sdet[l] holds pyramid layer l + 1 (the middle layers)
  __shared__ float sdet[SURF_MAX_LAYERS - 2][TH + 2][TW + 2];
  __shared__ DetCand s_cand[DET_LIST];
  __shared__ float s_outer[DET_LIST][9];
  __shared__ int s_ncand;
  __shared__ int s_dead[DET_LIST];  // candidate beaten by an outer-layer neighbour
  __shared__ unsigned s_claim[((SURF_MAX_LAYERS - 2) * (TH + 2) * (TW + 2) + 31) / 32];
#if UVO_DET_SCREEN
  __shared__ unsigned short s_surv[(SURF_MAX_LAYERS - 2) * (TH + 2) * (TW + 2)];  // samples that pass the screen
  __shared__ int s_nsurv;
#endif
  const SurfImage& im = b.im[blockIdx.y];
  int t = blockIdx.x, o = 0;
  while (o + 1 < g.n_octaves && t >= g.oct[o + 1].tile_begin) o++;
  const SurfOctave& O = g.oct[o];
  t -= O.tile_begin;
  const int ti0 = (t / O.tiles_x) * TH, tj0 = (t % O.tiles_x) * TW;
  const int nmid = g.n_layers;
  const int scols = g.spitch;
  const int step = O.step;
  constexpr int PLANE = (TH + 2) * (TW + 2);
  __shared__ __align__(128) int s_tile[T0_ROWS * T0_COLS];
  __shared__ __align__(8) unsigned long long s_bar;
  // exact value of middle layer lm (0-based) at tile position (y, x): the full evaluation of a skipped sample
  auto mid_exact = [&](int lm, int y, int x) -> float {
    const int i = ti0 + y - 1, j = tj0 + x - 1;
    if (o == 0) {
      if (lm == 0) return det_tile0<15>(s_tile, O.layer[1], i, j, y, x);
      if (lm == 1) return det_tile0<21>(s_tile, O.layer[2], i, j, y, x);
      return det_tile0<27>(s_tile, O.layer[3], i, j, y, x);
    }
    return det_at(im.sum, scols, O.layer[lm + 1], step, i, j);
  };
  if (threadIdx.x == 0) {
    s_ncand = 0;
#if UVO_DET_SCREEN
    s_nsurv = 0;
#endif
  }
  if (threadIdx.x < (int)(sizeof(s_claim) / sizeof(unsigned))) s_claim[threadIdx.x] = 0;
  if (threadIdx.x < DET_LIST) s_dead[threadIdx.x] = 0;
  if (o == 0) {
    // octave 0 (three quarters of all samples): the integral tile arrives as ONE TMA box (cp.async.bulk.tensor.2d,
    // completion on an mbarrier) -- coordinates left of / above the image and past its last row / column are filled
    // with zeros by the copy engine, which is the border rule of the evaluation -- then every sample is evaluated
    // from shared memory
    const int R0 = ti0 - 1 - T0_MAXM, C0 = tj0 - 1 - T0_MAXM - T0_XOFF;
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      mbar_expect_tx(bar, T0_ROWS * T0_COLS * (int)sizeof(int));
      tma_load_2d(smem_u32(s_tile), blockIdx.y ? &maps.sum[1] : &maps.sum[0], C0, R0, bar);
    }
    __syncthreads();  // the barrier is initialised before anyone polls it
    mbar_wait(bar, 0);
  } else {
    __syncthreads();
  }
#if UVO_DET_SCREEN
  // pass 1: screen every sample of the middle layers; survivors go to a list (warp-aggregated append)
  const int total = nmid * PLANE;
  for (int base = 0; base < total; base += blockDim.x) {
    const int idx = base + threadIdx.x;
    int verdict = SCREEN_SKIP, l = 0, y = 0, x = 0;
    if (idx < total) {
      l = idx / PLANE;
      const int r = idx - l * PLANE;
      y = r / (TW + 2);
      x = r - y * (TW + 2);
      const int i = ti0 + y - 1, j = tj0 + x - 1;
      if (o == 0) {
        if (l == 0) verdict = screen_tile0<15>(s_tile, O.layer[1], i, j, y, x, g.thr_skip);
        else if (l == 1) verdict = screen_tile0<21>(s_tile, O.layer[2], i, j, y, x, g.thr_skip);
        else verdict = screen_tile0<27>(s_tile, O.layer[3], i, j, y, x, g.thr_skip);
      } else {
        verdict = screen_at(im.sum, scols, O.layer[l + 1], step, i, j, g.thr_skip);
      }
      if (verdict != SCREEN_KEEP) sdet[l][y][x] = verdict == SCREEN_OUTSIDE ? 0.f : DET_SKIPPED;
    }
    const unsigned keep = __ballot_sync(0xffffffffu, verdict == SCREEN_KEEP);
    if (keep) {
      const int lane = threadIdx.x & 31;
      int at0 = 0;
      if (lane == 0) at0 = atomicAdd(&s_nsurv, __popc(keep));
      at0 = __shfl_sync(0xffffffffu, at0, 0);
      if (verdict == SCREEN_KEEP) s_surv[at0 + __popc(keep & ((1u << lane) - 1))] = (unsigned short)idx;
    }
  }
  __syncthreads();
  // pass 2: the survivors, evaluated exactly (the lazy rule still skips the Dxy half of those with fl(dx dy) <= thr)
  for (int k = threadIdx.x; k < s_nsurv; k += blockDim.x) {
    const int idx = s_surv[k];
    const int l = idx / PLANE, r = idx - l * PLANE, y = r / (TW + 2), x = r - y * (TW + 2);
    const int i = ti0 + y - 1, j = tj0 + x - 1;
    float v;
    if (o == 0) {
      if (l == 0) v = det_tile0<15, true>(s_tile, O.layer[1], i, j, y, x, g.thr);
      else if (l == 1) v = det_tile0<21, true>(s_tile, O.layer[2], i, j, y, x, g.thr);
      else v = det_tile0<27, true>(s_tile, O.layer[3], i, j, y, x, g.thr);
    } else {
      v = det_at<true>(im.sum, scols, O.layer[l + 1], step, i, j, g.thr);
    }
    sdet[l][y][x] = v;
  }
#else
  for (int idx = threadIdx.x; idx < nmid * PLANE; idx += blockDim.x) {
    const int l = idx / PLANE, r = idx - l * PLANE, y = r / (TW + 2), x = r - y * (TW + 2);
    const int i = ti0 + y - 1, j = tj0 + x - 1;
    float v;
    if (o == 0) {
      if (l == 0) v = det_tile0<15, true>(s_tile, O.layer[1], i, j, y, x, g.thr);
      else if (l == 1) v = det_tile0<21, true>(s_tile, O.layer[2], i, j, y, x, g.thr);
      else v = det_tile0<27, true>(s_tile, O.layer[3], i, j, y, x, g.thr);
    } else {
      v = det_at<true>(im.sum, scols, O.layer[l + 1], step, i, j, g.thr);
    }
    sdet[l][y][x] = v;
  }
#endif
  __syncthreads();
  // local maxima among the middle layers.  A skipped neighbour is <= thr < val0, and DET_SKIPPED = -inf compares
  // exactly like that.  Every maximum goes to the list; what it still misses (exact values of skipped neighbours,
  // an outer plane) is computed by the whole block below.
  for (int idx = threadIdx.x; idx < nmid * TH * TW; idx += blockDim.x) {
    const int m = 1 + idx / (TH * TW), r = idx % (TH * TW), y = r / TW, x = r % TW;
    const int i = ti0 + y, j = tj0 + x;
    const int margin = O.nms_margin[m];
    if (i < margin || i >= O.lrows - margin || j < margin || j >= O.lcols - margin) continue;
    const float val0 = sdet[m - 1][y + 1][x + 1];
    if (!(val0 > g.thr)) continue;
    bool is_max = true;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const int l = m - 1 + a;  // pyramid layer of this plane
      if (l < 1 || l > nmid) continue;
#pragma unroll
      for (int q = 0; q < 9; q++)
        if (!(a == 1 && q == 4) && !(val0 > sdet[l - 1][y + q / 3][x + q % 3])) is_max = false;
    }
    if (!is_max) continue;
    const int slot = atomicAdd(&s_ncand, 1);
    if (slot < DET_LIST) {
      s_cand[slot] = DetCand{(short)m, (short)y, (short)x, 1};
    } else {  // list full: finish this candidate here
      float N[3][9];
#pragma unroll 1
      for (int a = 0; a < 3; a++) {
        const int l = m - 1 + a;
#pragma unroll 1
        for (int q = 0; q < 9; q++) {
          float v;
          if (l >= 1 && l <= nmid) {
            v = sdet[l - 1][y + q / 3][x + q % 3];
            if (v == DET_SKIPPED) v = mid_exact(l - 1, y + q / 3, x + q % 3);
          } else {
            v = det_at(im.sum, scols, O.layer[l], step, i - 1 + q / 3, j - 1 + q % 3);
            if (!(val0 > v)) is_max = false;
          }
          N[a][q] = v;
        }
      }
      if (is_max) emit_keypoint(O, im, N, m, i, j, o, capacity);
    }
  }
  __syncthreads();
  const int ncand = min(s_ncand, DET_LIST);
  // exact values of the skipped middle-layer neighbours of the listed maxima (27 slots per maximum)
  // Neighbourhoods of nearby maxima overlap, so a cell can be wanted by several threads: each cell is claimed in a
  // bitmap (atomicOr) and filled by the one thread that wins the claim -- no thread reads a cell another may write
  // (compute-sanitizer racecheck clean).  The claim is taken only for cells that were skipped, which every thread
  // establishes before the first write of the phase.
  for (int base = 0; base < ncand * 27; base += blockDim.x) {
    const int idx = base + threadIdx.x;
    int l = 0, y = 0, x = 0;
    bool need = false;
    if (idx < ncand * 27) {
      const int c = idx / 27, r = idx - c * 27, a = r / 9, q = r - a * 9;
      const DetCand cd = s_cand[c];
      l = cd.m - 1 + a;
      y = cd.y + q / 3;
      x = cd.x + q % 3;
      need = l >= 1 && l <= nmid && sdet[l - 1][y][x] == DET_SKIPPED;
    }
    __syncthreads();  // every read of this round precedes its writes
    if (need) {
      const int cell = ((l - 1) * (TH + 2) + y) * (TW + 2) + x;
      const unsigned bit = 1u << (cell & 31);
      if (!(atomicOr(&s_claim[cell >> 5], bit) & bit)) sdet[l - 1][y][x] = mid_exact(l - 1, y, x);
    }
    __syncthreads();
  }
  // outer-layer samples of the listed candidates: 9 per (candidate, outer plane); with a single middle layer a
  // candidate needs both outer planes, handled as two passes
  for (int pass = 0; pass < 2; pass++) {
    for (int idx = threadIdx.x; idx < ncand * 9; idx += blockDim.x) {
      const int c = idx / 9, q = idx - c * 9;
      const DetCand cd = s_cand[c];
      const int l = pass == 0 ? (cd.m == 1 ? 0 : -1) : (cd.m == nmid ? nmid + 1 : -1);
      if (l < 0) continue;
      const float v = det_at(im.sum, scols, O.layer[l], step, ti0 + cd.y - 1 + q / 3, tj0 + cd.x - 1 + q % 3);
      s_outer[c][q] = v;
      if (!(sdet[cd.m - 1][cd.y + 1][cd.x + 1] > v)) atomicExch(&s_dead[c], 1);  // up to nine writers per candidate
    }
    __syncthreads();
    // finish candidates whose last missing plane was this pass's (pass 0 also finishes those that need none)
    for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
      const DetCand cd = s_cand[c];
      const bool needs1 = cd.m == nmid;
      const bool last = pass == 1 ? needs1 : !needs1;
      if (!last || s_dead[c]) continue;
      float N[3][9];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const int l = cd.m - 1 + a;
#pragma unroll
        for (int q = 0; q < 9; q++) {
          if (l >= 1 && l <= nmid)
            N[a][q] = sdet[l - 1][cd.y + q / 3][cd.x + q % 3];
          else if (l == nmid + 1 || !needs1)
            N[a][q] = s_outer[c][q];  // the plane evaluated in this pass
          else
            N[a][q] = det_at(im.sum, scols, O.layer[0], step, ti0 + cd.y - 1 + q / 3, tj0 + cd.x - 1 + q % 3);
        }
      }
      emit_keypoint(O, im, N, cd.m, ti0 + cd.y, tj0 + cd.x, o, capacity);
    }
    __syncthreads();
  }
}

// the compile-time corner offsets of the shared-memory path must be the ones resizeHaarPattern computes at run time
template <int SIZE>
static bool tile0_offsets_match(const SurfLayer& L, int ws) {
  if (L.size != SIZE) return false;
  const float ratio = (float)SIZE / 9;
  for (int k = 0; k < 10; k++)
    if (cv_roundf_h(ratio * k) != haar_off<SIZE>(k)) return false;
  return L.box[6].p0 == haar_off<SIZE>(1) * ws + haar_off<SIZE>(1) && L.box[9].p3 == haar_off<SIZE>(8) * ws + haar_off<SIZE>(8) &&
         L.xx_col[3] == haar_off<SIZE>(9) && L.yy_row[1] == haar_off<SIZE>(3) * ws;
}

SurfMaps make_surf_maps(const int32_t* sum0, const int32_t* sum1, int w, int h) {
  SurfMaps m;
  const int32_t* base[2] = {sum0, sum1};
  for (int i = 0; i < 2; i++) {
    const cuuint64_t dims[2] = {(cuuint64_t)(w + 1), (cuuint64_t)(h + 1)};
    const cuuint64_t strides[1] = {(cuuint64_t)surf_sum_pitch(w) * sizeof(int32_t)};
    const cuuint32_t box[2] = {(cuuint32_t)T0_COLS, (cuuint32_t)T0_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_tiled_fn()(&m.sum[i], CU_TENSOR_MAP_DATA_TYPE_INT32, 2, (void*)base[i], dims, strides, box,
                                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      throw InvalidArg{"cuTensorMapEncodeTiled (integral image) failed (" + std::to_string((int)r) + ")", UVO_ERR_CUDA};
  }
  return m;
}

void launch_surf_detect(Ctx& c, const SurfGeom& g, const SurfBatch& b, const SurfMaps& maps, int capacity) {
  upload_tables(c);
  {
    const SurfOctave& O = g.oct[0];
    const int ws = g.spitch;
    bool ok = O.step == 1 && tile0_offsets_match<15>(O.layer[1], ws);
    if (g.n_layers >= 2) ok = ok && tile0_offsets_match<21>(O.layer[2], ws);
    if (g.n_layers >= 3) ok = ok && tile0_offsets_match<27>(O.layer[3], ws);
    if (!ok) throw InvalidArg{"SURF: octave-0 tile geometry mismatch (internal)", UVO_ERR_UNSUPPORTED};
  }
  for (int i = 0; i < b.n_img; i++) {
    UVO_CUDA(cudaMemsetAsync(b.im[i].counters, 0, 4 * sizeof(int), c.stream));
  }
  UVO_KERNEL(c, "k_surf_detect");
  // the tensor maps come first: the copy engine reads them from the parameter space, and the geometry (> 4 KB of box
  // tables) would push them past the first 4 KB of it
  k_surf_detect<<<dim3(g.total_tiles, b.n_img), 256, 0, c.stream>>>(maps, g, b, capacity);
  UVO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------------ K6 sort
// KeypointGreater: response desc, size desc, octave desc, y desc, x asc
__device__ __forceinline__ bool kp_precedes(float ra, float sa, int oa, float ya, float xa, float rb, float sb,
                                            int ob, float yb, float xb) {
  if (ra > rb) return true;
  if (ra < rb) return false;
  if (sa > sb) return true;
  if (sa < sb) return false;
  if (oa > ob) return true;
  if (oa < ob) return false;
  if (ya > yb) return true;
  if (ya < yb) return false;
  return xa < xb;
}

// 2-D rank sort: block (bi, bj) adds to rank[i] the number of keys of j-tile bj that precede key i.  Ranks are a
// permutation (full ties broken by raw index), so the scatter needs no atomics and the output order is exactly
// std::sort(KeypointGreater)'s for distinct keys.
__global__ void __launch_bounds__(256) k_surf_rank(const __grid_constant__ SurfBatch b, int capacity) {
  __shared__ float4 s_k[256];
  __shared__ int s_o[256];
  const SurfImage& im = b.im[blockIdx.z];
  const int n = min(im.counters[0], capacity);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    im.counters[1] = n;
    im.counters[2] = 0;  // k_surf_describe's work-queue head
  }
  const int i0 = blockIdx.x * 256, j0 = blockIdx.y * 256;
  if (i0 >= n || j0 >= n) return;
  const int i = i0 + threadIdx.x, j = j0 + threadIdx.x;
  if (j < n) {
    const uvo_keypoint k = im.raw[j];
    s_k[threadIdx.x] = make_float4(k.response, k.size, k.y, k.x);
    s_o[threadIdx.x] = k.octave;
  }
  __syncthreads();
  if (i >= n) return;
  const uvo_keypoint me = im.raw[i];
  const int m = min(256, n - j0);
  int rank = 0;
  for (int q = 0; q < m; q++) {
    const float4 k = s_k[q];
    const int jo = s_o[q];
    const bool before = kp_precedes(k.x, k.y, jo, k.z, k.w, me.response, me.size, me.octave, me.y, me.x);
    const bool same = (k.x == me.response) && (k.y == me.size) && (jo == me.octave) && (k.z == me.y) && (k.w == me.x);
    rank += (before || (same && (j0 + q) < i)) ? 1 : 0;
  }
  if (rank) atomicAdd(&im.rank[i], rank);
}

__global__ void __launch_bounds__(256) k_surf_scatter(const __grid_constant__ SurfBatch b, int capacity) {
  const SurfImage& im = b.im[blockIdx.y];
  const int n = min(im.counters[0], capacity);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) im.kps[im.rank[i]] = im.raw[i];
}

// Block sort (capacity <= SORT_BLOCK_MAX): one 1024-thread block per image sorts (response, raw index) pairs with a
// bitonic network in shared memory -- N log^2 N compare-exchanges on one SM instead of the N^2 compares of the rank
// sort on all of them -- then repairs runs of equal responses with the full KeypointGreater comparator (rare; same
// tie rule as the rank sort: identical keys keep raw-index order) and writes the keypoints out in order.
constexpr int SORT_BLOCK_MAX = 16384;

__device__ __forceinline__ bool kp_precedes(const uvo_keypoint& a, const uvo_keypoint& b) {
  return kp_precedes(a.response, a.size, a.octave, a.y, a.x, b.response, b.size, b.octave, b.y, b.x);
}

__global__ void __launch_bounds__(1024) k_surf_sort_block(const __grid_constant__ SurfBatch b, int capacity) {
  extern __shared__ unsigned long long s_key[];  // (monotone response bits << 32) | ~raw index; 0 = padding
  const SurfImage& im = b.im[blockIdx.x];
  const int n = min(im.counters[0], capacity);
  const int tid = threadIdx.x;
  if (tid == 0) {
    im.counters[1] = n;
    im.counters[2] = 0;  // k_surf_describe's work-queue head
  }
  int P = 2;
  while (P < n) P <<= 1;
  for (int i = tid; i < P; i += 1024) {
    unsigned long long key = 0ull;
    if (i < n) {
      const unsigned r = __float_as_uint(im.raw[i].response);
      const unsigned mono = (r & 0x80000000u) ? ~r : (r | 0x80000000u);
      key = ((unsigned long long)mono << 32) | (unsigned)(~(unsigned)i);
    }
    s_key[i] = key;
  }
  __syncthreads();
  // descending bitonic sort
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += 1024) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
        const unsigned long long a = s_key[lo], c = s_key[hi];
        const bool desc = (lo & k) == 0;
        if ((a < c) == desc) {
          s_key[lo] = c;
          s_key[hi] = a;
        }
      }
      __syncthreads();
    }
  }
  // repair runs of equal response (insertion sort by the full comparator; the run head does the work)
  for (int p0 = tid; p0 < n; p0 += 1024) {
    const unsigned r = (unsigned)(s_key[p0] >> 32);
    const bool head = (p0 == 0 || (unsigned)(s_key[p0 - 1] >> 32) != r) && p0 + 1 < n &&
                      (unsigned)(s_key[p0 + 1] >> 32) == r;
    if (!head) continue;
    int q = p0 + 1;
    while (q < n && (unsigned)(s_key[q] >> 32) == r) q++;
    for (int a = p0 + 1; a < q; a++) {
      const unsigned long long ka = s_key[a];
      const uvo_keypoint A = im.raw[~(unsigned)ka];
      int c = a - 1;
      while (c >= p0 && kp_precedes(A, im.raw[~(unsigned)s_key[c]])) {
        s_key[c + 1] = s_key[c];
        c--;
      }
      s_key[c + 1] = ka;
    }
  }
  __syncthreads();
  // ordered write-out, word-wise (a keypoint is 7 words)
  const int* src = (const int*)im.raw;
  int* dst = (int*)im.kps;
  for (int w = tid; w < n * 7; w += 1024) {
    const int p = w / 7, f = w - p * 7;
    dst[w] = src[(size_t)(~(unsigned)s_key[p]) * 7 + f];
  }
}

void launch_surf_sort(Ctx& c, const SurfBatch& b, int capacity) {
  if (capacity <= SORT_BLOCK_MAX) {
    int P = 2;
    while (P < capacity) P <<= 1;
    const size_t smem = (size_t)P * sizeof(unsigned long long);
    {
      static std::mutex attr_mutex;
      static bool attr_set[64] = {};
      std::lock_guard<std::mutex> lock(attr_mutex);
      if (c.device >= 64 || !attr_set[c.device]) {
        UVO_CUDA(cudaFuncSetAttribute(k_surf_sort_block, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SORT_BLOCK_MAX * (int)sizeof(unsigned long long)));
        if (c.device < 64) attr_set[c.device] = true;
      }
    }
    UVO_KERNEL(c, "k_surf_sort_block");
    k_surf_sort_block<<<b.n_img, 1024, smem, c.stream>>>(b, capacity);
    UVO_LAUNCH_CHECK(c);
    return;
  }
  const int tiles = div_up(capacity, 256);
  for (int i = 0; i < b.n_img; i++)
    UVO_CUDA(cudaMemsetAsync(b.im[i].rank, 0, (size_t)capacity * sizeof(int), c.stream));
  UVO_KERNEL(c, "k_surf_rank");
  k_surf_rank<<<dim3(tiles, tiles, b.n_img), 256, 0, c.stream>>>(b, capacity);
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_surf_scatter");
  k_surf_scatter<<<dim3(tiles, b.n_img), 256, 0, c.stream>>>(b, capacity);
  UVO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------------ K7 descriptors
// INTER_AREA table of one destination index (OpenCV computeResizeAreaTab): optional left partial cell, full cells
// [sx1, sx2), optional right partial cell.
struct AreaSpan {
  int sx1, sx2;
  bool has_l, has_r;
  float a_l, a_f, a_r;
};
__device__ __forceinline__ AreaSpan area_span(int d, int ssize, double scale) {
  AreaSpan s;
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  s.sx1 = sx1;
  s.sx2 = sx2;
  s.has_l = (sx1 - fsx1 > 1e-3);
  s.a_l = (float)((sx1 - fsx1) / cell);
  s.a_f = (float)(1.0 / cell);
  s.has_r = (fsx2 - sx2 > 1e-3);
  s.a_r = (float)(fmin(fmin(fsx2 - sx2, 1.), cell) / cell);
  return s;
}

// exact u8 -> f32 of byte `i` of a word without the quarter-rate I2F.U8: PRMT builds the float 2^23 + byte, one FADD
// removes the 2^23 (both exact)
__device__ __forceinline__ float byte_to_float(unsigned v, int i) {
  return __fsub_rn(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7540u | (unsigned)i)), 8388608.f);
}

struct WinSampler {
  const uint8_t* img;
  size_t pitch;
  int w, h;
  bool upright;
  // upright: WIN[i][j] = img(clamp(start_y - j), clamp(start_x + i))
  int start_x, start_y;
  // oriented: bilinear / nearest sampling along the rotated frame (A.4 "Window")
  float fstart_x, fstart_y, sin_dir, cos_dir;
  int win_size;
  __device__ __forceinline__ int at(int i, int j) const {
    if (upright) {
      const int x = min(max(start_x + i, 0), w - 1), y = min(max(start_y - j, 0), h - 1);
      return __ldg(img + (size_t)y * pitch + x);
    }
    // OpenCV accumulates start_x += sin_dir (f32) per row i and pixel_x += cos_dir (f64) per column j
    float sx = fstart_x, sy = fstart_y;
    for (int q = 0; q < i; q++) {
      sx = __fadd_rn(sx, sin_dir);
      sy = __fadd_rn(sy, cos_dir);
    }
    double pxl = sx, pyl = sy;
    for (int q = 0; q < j; q++) {
      pxl = __dadd_rn(pxl, (double)cos_dir);
      pyl = __dsub_rn(pyl, (double)sin_dir);
    }
    const int ix = (int)floor(pxl), iy = (int)floor(pyl);
    if ((unsigned)ix < (unsigned)(w - 1) && (unsigned)iy < (unsigned)(h - 1)) {
      const float a = (float)(pxl - ix), bb = (float)(pyl - iy);
      const uint8_t* p = img + (size_t)iy * pitch + ix;
      const float p00 = (float)__ldg(p), p01 = (float)__ldg(p + 1), p10 = (float)__ldg(p + pitch),
                  p11 = (float)__ldg(p + pitch + 1);
      const float ia = __fsub_rn(1.f, a), ib = __fsub_rn(1.f, bb);
      float v = __fmul_rn(__fmul_rn(p00, ia), ib);
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(p01, a), ib));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(p10, ia), bb));
      v = __fadd_rn(v, __fmul_rn(__fmul_rn(p11, a), bb));
      return (int)(uint8_t)__float2int_rn(v);
    }
    const int x = min(max(__double2int_rn(pxl), 0), w - 1), y = min(max(__double2int_rn(pyl), 0), h - 1);
    return __ldg(img + (size_t)y * pitch + x);
  }
};

// cv::fastAtan2 polynomial (degrees)
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float p1 = 0.9997878412794807f * (float)(180 / M_PI), p3 = -0.3258083974640975f * (float)(180 / M_PI),
              p5 = 0.1555786518463281f * (float)(180 / M_PI), p7 = -0.04432655554792128f * (float)(180 / M_PI);
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

#ifndef UVO_DESC_THREADS
#define UVO_DESC_THREADS 256
#endif
constexpr int DESC_THREADS = UVO_DESC_THREADS;
constexpr int DESC_BUF_ROWS = 168;  // window rows buffered at once (14 KB); taller windows are streamed in chunks
constexpr int PATCH_STRIDE = 448;   // bytes per keypoint in SurfImage::patch (441 used)
// INTER_AREA tables by window size.  The 21 AreaSpan entries, iscale and is_area_fast are functions of win_size alone
// (fp64 divisions, ceil / floor); they are tabulated once per device by the same device code the kernel would run, so
// a keypoint's prologue is a 600-byte load instead of a chain of fp64 divisions on 21 of the block's 256 threads.
constexpr int SPAN_TAB_MAX = 1024;  // window sizes < SPAN_TAB_MAX are tabulated (the largest SURF window is ~740)
struct SpanTable {
  AreaSpan span[SPAN_TAB_MAX][21];
  int iscale[SPAN_TAB_MAX];
  int area_fast[SPAN_TAB_MAX];
};
__global__ void k_span_table(SpanTable* tab) {
  const int win = blockIdx.x, d = threadIdx.x;
  if (win < 21 || d >= 21) return;
  const double inv_scale = (double)21 / win;
  const double scale = 1. / inv_scale;
  tab->span[win][d] = area_span(d, win, scale);
  if (d == 0) {
    const int isc = __double2int_rn(scale);
    tab->iscale[win] = isc;
    tab->area_fast[win] = fabs(scale - isc) < DBL_EPSILON;
  }
}
static SpanTable* g_span_tab[64] = {};
static const SpanTable* span_table(Ctx& c) {
  if (c.device >= 64) return nullptr;
  if (!g_span_tab[c.device]) {
    SpanTable* t = nullptr;
    UVO_CUDA(cudaMalloc(&t, sizeof(SpanTable)));  // lives as long as the process, like the constant tables
    k_span_table<<<SPAN_TAB_MAX, 32, 0, c.stream>>>(t);
    UVO_CUDA(cudaGetLastError());
    UVO_CUDA(cudaStreamSynchronize(c.stream));
    g_span_tab[c.device] = t;
  }
  return g_span_tab[c.device];
}

// K7 runs as two kernels.  k_surf_patch (one block per keypoint, dynamic queue) does the wide part: (orientation,)
// window extraction and the INTER_AREA resize to the 21x21 u8 patch, which it writes to global memory.
// k_surf_vector (one warp per keypoint) does the narrow part: Haar gradients, 4x4x4 sums, L2 normalisation.  Keeping
// the narrow phases out of the block-per-keypoint kernel is what lets its 256 threads stay busy.
#ifndef UVO_DESC_MINB
#define UVO_DESC_MINB 4
#endif
__global__ void __launch_bounds__(DESC_THREADS, UVO_DESC_MINB) k_surf_patch(const __grid_constant__ SurfGeom g,
                                                             const __grid_constant__ SurfBatch b, int upright,
                                                             const SpanTable* __restrict__ tab) {
  __shared__ int s_patch[441];
  __shared__ AreaSpan s_span[21];
  __shared__ float s_buf[DESC_BUF_ROWS * 21];
  __shared__ float s_acc[441];
  __shared__ float s_X[128], s_Y[128], s_ang[128];
  __shared__ int s_nangle;
  __shared__ float s_dir;
  __shared__ int s_next;
  __shared__ int s_iscale, s_area_fast;
  const SurfImage& im = b.im[blockIdx.y];
  const int n = im.counters[1];
  const int w = g.w, h = g.h, srows = h + 1, scols = w + 1;  // sum.rows / sum.cols of the CPU code
  const int spitch = g.spitch;                                // row pitch of the integral in memory
  const int tid = threadIdx.x;
  // dynamic queue (counters[2], zeroed by the sort kernel): window areas span 21^2 .. 576^2 pixels, so a static
  // assignment leaves most blocks waiting for the one that drew the largest windows
  if (tid == 0) s_next = atomicAdd(&im.counters[2], 1);
  __syncthreads();
  for (;;) {
    const int k = s_next;
    __syncthreads();
    if (k >= n) break;
    // the next queue slot is requested now and published at the end of this keypoint, so the atomic's round trip
    // overlaps the work instead of heading every keypoint
    int k_next = 0;
    if (tid == 0) k_next = atomicAdd(&im.counters[2], 1);
    do {
    const uvo_keypoint kp = im.kps[k];
    const float size = kp.size, cx = kp.x, cy = kp.y;
    const float s = __fdiv_rn(__fmul_rn(size, 1.2f), 9.0f);
    const int gws = 2 * __float2int_rn(__fmul_rn(2.f, s));
    if (srows < gws || scols < gws) {  // gradient wavelet larger than the image: mark for deletion
      if (tid == 0) im.kps[k].size = -1.f;
      break;
    }
    float dir = 270.f;
    if (!upright) {
      // ---- dominant orientation (A.4): Haar responses on a radius-6s disc, 60-degree sliding window ----
      if (tid == 0) s_nangle = 0;
      __syncthreads();
      // resizeHaarPattern(dx_s/dy_s, 4 -> gws): boxes {0,0,2,4,-1},{2,0,4,4,1} and {0,0,4,2,1},{0,2,4,4,-1}
      const float ratio = (float)gws / 4;
      const int c2 = __float2int_rn(__fmul_rn(ratio, 2.f)), c4 = __float2int_rn(__fmul_rn(ratio, 4.f));
      const float wx = 1.f / ((float)(c2 - 0) * (c4 - 0));   // |w| of every box (all are c2 x c4 or c4 x c2)
      // samples are visited in table order; order of X/Y entries must match the CPU (sequential append), so one
      // warp-free pass computes validity, a prefix gives the slot.
      float vX = 0.f, vY = 0.f;
      bool valid = false;
      if (tid < 113) {
        const int x = __float2int_rn(__fsub_rn(__fadd_rn(cx, __fmul_rn((float)c_apt[tid][0], s)), (float)(gws - 1) / 2));
        const int y = __float2int_rn(__fsub_rn(__fadd_rn(cy, __fmul_rn((float)c_apt[tid][1], s)), (float)(gws - 1) / 2));
        if (!(y < 0 || y >= srows - gws || x < 0 || x >= scols - gws)) {
          valid = true;
          const int* o = im.sum + (size_t)y * spitch + x;
          auto box = [&](int x1, int y1, int x2, int y2) -> int {
            return (int)((unsigned)__ldg(o + y1 * spitch + x1) + (unsigned)__ldg(o + y2 * spitch + x2) -
                         (unsigned)__ldg(o + y2 * spitch + x1) - (unsigned)__ldg(o + y1 * spitch + x2));
          };
          double dxv = 0, dyv = 0;
          dxv = __dadd_rn(dxv, (double)__fmul_rn((float)box(0, 0, c2, c4), -wx));
          dxv = __dadd_rn(dxv, (double)__fmul_rn((float)box(c2, 0, c4, c4), wx));
          dyv = __dadd_rn(dyv, (double)__fmul_rn((float)box(0, 0, c4, c2), wx));
          dyv = __dadd_rn(dyv, (double)__fmul_rn((float)box(0, c2, c4, c4), -wx));
          vX = __fmul_rn((float)dxv, c_aptw[tid]);
          vY = __fmul_rn((float)dyv, c_aptw[tid]);
        }
      }
      // ordered compaction over the block's warps
      __shared__ int s_wcnt[DESC_THREADS / 32];
      const unsigned bal = __ballot_sync(0xffffffffu, valid);
      const int lane = tid & 31, wid = tid >> 5;
      if (lane == 0) s_wcnt[wid] = __popc(bal);
      __syncthreads();
      int off = 0;
      for (int q = 0; q < wid; q++) off += s_wcnt[q];
      if (valid) {
        const int slot = off + __popc(bal & ((1u << lane) - 1));
        s_X[slot] = vX;
        s_Y[slot] = vY;
        s_ang[slot] = fast_atan2_deg(vY, vX);  // cv::phase(X, Y, angle, true)
      }
      if (tid == 0) {
        int tot = 0;
        for (int q = 0; q < DESC_THREADS / 32; q++) tot += s_wcnt[q];
        s_nangle = tot;
      }
      __syncthreads();
      const int nangle = s_nangle;
      if (nangle == 0) {
        if (tid == 0) im.kps[k].size = -1.f;
        break;
      }
      // 72 window positions; thread t evaluates window i = 5t sequentially over samples (CPU order), then the
      // first-best-wins arg max is taken in window order.
      __shared__ float s_mod[72], s_sx[72], s_sy[72];
      if (tid < 72) {
        const int i = tid * 5;
        float sumx = 0, sumy = 0;
        for (int j = 0; j < nangle; j++) {
          const int d = abs(__float2int_rn(s_ang[j]) - i);
          if (d < 30 || d > 330) {
            sumx = __fadd_rn(sumx, s_X[j]);
            sumy = __fadd_rn(sumy, s_Y[j]);
          }
        }
        s_sx[tid] = sumx;
        s_sy[tid] = sumy;
        s_mod[tid] = __fadd_rn(__fmul_rn(sumx, sumx), __fmul_rn(sumy, sumy));
      }
      __syncthreads();
      if (tid == 0) {
        float bestx = 0, besty = 0, best = 0;
        for (int q = 0; q < 72; q++)
          if (s_mod[q] > best) {
            best = s_mod[q];
            bestx = s_sx[q];
            besty = s_sy[q];
          }
        s_dir = fast_atan2_deg(-besty, bestx);
      }
      __syncthreads();
      dir = s_dir;
    }


    if (tid == 0) im.kps[k].angle = dir;
    // ---- window geometry ----
    const int win_size = (int)__fmul_rn(21.f, s);
    WinSampler ws;
    ws.img = im.img;
    ws.pitch = im.pitch;
    ws.w = w;
    ws.h = h;
    ws.upright = upright != 0;
    ws.win_size = win_size;
    const float win_offset = -(float)(win_size - 1) / 2;
    if (upright) {
      ws.start_x = __float2int_rn(__fadd_rn(cx, win_offset));
      ws.start_y = __float2int_rn(__fsub_rn(cy, win_offset));
    } else {
      const float ddir = __fmul_rn(dir, (float)(M_PI / 180));
      // std::sin/std::cos(float) on the CPU are (nearly) correctly rounded; go through fp64 to match them
      ws.sin_dir = -(float)sin((double)ddir);
      ws.cos_dir = (float)cos((double)ddir);
      ws.fstart_x = __fadd_rn(__fadd_rn(cx, __fmul_rn(win_offset, ws.cos_dir)), __fmul_rn(win_offset, ws.sin_dir));
      ws.fstart_y = __fadd_rn(__fsub_rn(cy, __fmul_rn(win_offset, ws.sin_dir)), __fmul_rn(win_offset, ws.cos_dir));
    }

    // ---- resize(win -> 21x21, INTER_AREA) ----
    // scale / iscale / is_area_fast exactly as cv::resize derives them (fp64); only 21 threads need the divisions
    if (tid < 21) {
      if (tab && win_size >= 21 && win_size < SPAN_TAB_MAX) {  // rows below 21 are never tabulated
        s_span[tid] = tab->span[win_size][tid];
        if (tid == 0) {
          s_iscale = tab->iscale[win_size];
          s_area_fast = tab->area_fast[win_size];
        }
      } else {
        const double inv_scale = (double)21 / win_size;
        const double scale = 1. / inv_scale;
        s_span[tid] = area_span(tid, win_size, scale);  // same table for rows and columns (square window)
        if (tid == 0) {
          const int isc = __double2int_rn(scale);
          s_iscale = isc;
          s_area_fast = fabs(scale - isc) < DBL_EPSILON;
        }
      }
    }
    __syncthreads();
    const int iscale = s_iscale;
    const bool area_fast = s_area_fast != 0;
    if (win_size == 21) {
      for (int t = tid; t < 441; t += DESC_THREADS) s_patch[t] = ws.at(t / 21, t % 21);
    } else if (upright) {
      // Separable form of OpenCV's ResizeArea_Invoker in exactly its arithmetic order (window row i <-> image x,
      // window column j <-> image y descending).  Pass 1: one work item per (aligned group of 4 image columns = 4
      // window rows, destination cell d); the item walks the cell's pixels down image y with one 32-bit load per
      // step and keeps four independent f32 chains, each accumulated in the CPU order.  Pass 2 accumulates
      // beta * buf over the window rows of each destination row, streaming over chunks of DESC_BUF_ROWS rows with
      // the 441 running sums kept in shared memory (0 + v == v exactly, so starting from zero equals OpenCV's
      // "first row assigns").  Gray images are pitch-aligned to 16 bytes (FrontEnd), so word loads are aligned;
      // columns / rows outside the image replicate the border exactly like the CPU's clamped window extraction.
      const uint8_t* __restrict__ img = im.img;
      const int pitch = (int)im.pitch;
      const int sx0 = ws.start_x, sy0 = ws.start_y;
      const bool interior_y = (sy0 - (win_size - 1) >= 0) && (sy0 <= h - 1);
      for (int t = tid; t < 441; t += DESC_THREADS) s_acc[t] = 0.f;
      for (int c0 = 0; c0 < win_size; c0 += DESC_BUF_ROWS) {
        const int rows = min(win_size - c0, DESC_BUF_ROWS);
        __syncthreads();  // s_acc zeroed / previous chunk's pass 2 done with s_buf
        const int x_first = sx0 + c0, xa = x_first & ~3;
        const int ncg = ((x_first + rows - 1) >> 2) - (xa >> 2) + 1;
        // t / ncg by a float reciprocal: (t + 0.5) / ncg is at least 0.5 / ncg (>= 0.01) away from any integer and
        // t < 2^12, so the f32 rounding (relative 2^-23) cannot move the floor
        const float inv_ncg = 1.0f / (float)ncg;
        // pixels of window column j for the four window rows of column group m (image columns xw .. xw + 3), as one
        // little-endian word; columns / rows outside the image replicate the border
        auto load4g = [&](int j, int xw) -> unsigned {
          int y = sy0 - j;
          if (!interior_y) y = min(max(y, 0), h - 1);
          const uint8_t* row = img + (size_t)y * pitch;
          if (xw >= 0 && xw + 3 <= w - 1) return __ldg((const unsigned*)(row + xw));
          unsigned v = 0;
#pragma unroll
          for (int q = 0; q < 4; q++) v |= (unsigned)__ldg(row + min(max(xw + q, 0), w - 1)) << (8 * q);
          return v;
        };
        for (int t = tid; t < ncg * 21; t += DESC_THREADS) {
          const int d = (int)(((float)t + 0.5f) * inv_ncg), m = t - d * ncg;
          const int xw = xa + 4 * m, il0 = xw - x_first;
          const bool x_in = xw >= 0 && xw + 3 <= w - 1;
          auto load4 = [&](int j) -> unsigned { return load4g(j, xw); };
          float r0, r1, r2, r3;
          if (area_fast) {
            const int j0 = d * iscale;
            int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 4
            for (int j = 0; j < iscale; j++) {
              const unsigned v = load4(j0 + j);
              a0 += v & 0xffu;
              a1 += (v >> 8) & 0xffu;
              a2 += (v >> 16) & 0xffu;
              a3 += v >> 24;
            }
            r0 = __int_as_float(a0);
            r1 = __int_as_float(a1);
            r2 = __int_as_float(a2);
            r3 = __int_as_float(a3);
          } else {
            const AreaSpan xs = s_span[d];
            float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
            auto step = [&](unsigned v, float a) {
              b0 = __fadd_rn(b0, __fmul_rn(byte_to_float(v, 0), a));
              b1 = __fadd_rn(b1, __fmul_rn(byte_to_float(v, 1), a));
              b2 = __fadd_rn(b2, __fmul_rn(byte_to_float(v, 2), a));
              b3 = __fadd_rn(b3, __fmul_rn(byte_to_float(v, 3), a));
            };
            if (xs.has_l) step(load4(xs.sx1 - 1), xs.a_l);
            if (interior_y && x_in) {
              const uint8_t* p = img + (size_t)(sy0 - xs.sx1) * pitch + xw;
              const int nfull = xs.sx2 - xs.sx1;
#pragma unroll 4
              for (int j = 0; j < nfull; j++) {
                step(__ldg((const unsigned*)p), xs.a_f);
                p -= pitch;
              }
            } else {
              for (int j = xs.sx1; j < xs.sx2; j++) step(load4(j), xs.a_f);
            }
            if (xs.has_r) step(load4(xs.sx2), xs.a_r);
            r0 = b0;
            r1 = b1;
            r2 = b2;
            r3 = b3;
          }
          float* dst = s_buf + il0 * 21 + d;
          if ((unsigned)il0 < (unsigned)rows) dst[0] = r0;
          if ((unsigned)(il0 + 1) < (unsigned)rows) dst[21] = r1;
          if ((unsigned)(il0 + 2) < (unsigned)rows) dst[42] = r2;
          if ((unsigned)(il0 + 3) < (unsigned)rows) dst[63] = r3;
        }
        __syncthreads();
        for (int t = tid; t < 441; t += DESC_THREADS) {
          const int dy = t / 21, dx = t - dy * 21;
          const float* col = s_buf + dx - c0 * 21;  // col[i * 21] = buf of window row i
          if (area_fast) {
            const int lo = max(dy * iscale, c0), hi = min(dy * iscale + iscale, c0 + rows);
            int acc = __float_as_int(s_acc[t]);
            for (int i = lo; i < hi; i++) acc += __float_as_int(col[i * 21]);
            s_acc[t] = __int_as_float(acc);
          } else {
            const AreaSpan ys = s_span[dy];
            const int c1 = c0 + rows;
            float sum = s_acc[t];
            if (ys.has_l && ys.sx1 - 1 >= c0 && ys.sx1 - 1 < c1)
              sum = __fadd_rn(sum, __fmul_rn(ys.a_l, col[(ys.sx1 - 1) * 21]));
            const int lo = max(ys.sx1, c0), hi = min(ys.sx2, c1);
#pragma unroll 4
            for (int i = lo; i < hi; i++) sum = __fadd_rn(sum, __fmul_rn(ys.a_f, col[i * 21]));
            if (ys.has_r && ys.sx2 >= c0 && ys.sx2 < c1) sum = __fadd_rn(sum, __fmul_rn(ys.a_r, col[ys.sx2 * 21]));
            s_acc[t] = sum;
          }
        }
      }
      __syncthreads();
      for (int t = tid; t < 441; t += DESC_THREADS) {
        int outv;
        if (area_fast) {
          const int acc = __float_as_int(s_acc[t]);
          if (iscale == 2) outv = (acc + 2) >> 2;
          else outv = min(max(__float2int_rn(__fmul_rn((float)acc, 1.f / (float)(iscale * iscale))), 0), 255);
        } else {
          outv = min(max(__float2int_rn(s_acc[t]), 0), 255);
        }
        s_patch[t] = outv;
      }
    } else {
      for (int t = tid; t < 441; t += DESC_THREADS) {
        const int py = t / 21, px = t - py * 21;
        int out;
        if (area_fast) {
          int acc = 0;
          for (int ky = 0; ky < iscale; ky++)
            for (int kx = 0; kx < iscale; kx++) {
              const int sy = py * iscale + ky, sx = px * iscale + kx;
              if (sy < win_size && sx < win_size) acc += ws.at(sy, sx);
            }
          if (iscale == 2) out = (acc + 2) >> 2;
          else out = min(max(__float2int_rn(__fmul_rn((float)acc, 1.f / (float)(iscale * iscale))), 0), 255);
        } else {
          const AreaSpan ys = s_span[py], xs = s_span[px];
          float sum = 0.f;
          bool first = true;
          auto row = [&](int sy, float beta) {
            float buf = 0.f;
            if (xs.has_l) buf = __fadd_rn(buf, __fmul_rn((float)ws.at(sy, xs.sx1 - 1), xs.a_l));
            for (int sx = xs.sx1; sx < xs.sx2; sx++) buf = __fadd_rn(buf, __fmul_rn((float)ws.at(sy, sx), xs.a_f));
            if (xs.has_r) buf = __fadd_rn(buf, __fmul_rn((float)ws.at(sy, xs.sx2), xs.a_r));
            if (first) {
              sum = __fmul_rn(beta, buf);
              first = false;
            } else {
              sum = __fadd_rn(sum, __fmul_rn(beta, buf));
            }
          };
          if (ys.has_l) row(ys.sx1 - 1, ys.a_l);
          for (int sy = ys.sx1; sy < ys.sx2; sy++) row(sy, ys.a_f);
          if (ys.has_r) row(ys.sx2, ys.a_r);
          out = min(max(__float2int_rn(sum), 0), 255);
        }
        s_patch[py * 21 + px] = out;
      }
    }
    __syncthreads();
    {
      uint8_t* dstp = im.patch + (size_t)k * PATCH_STRIDE;
      for (int t = tid; t < 441; t += DESC_THREADS) dstp[t] = (uint8_t)s_patch[t];
    }
    } while (0);
    if (tid == 0) s_next = k_next;
    __syncthreads();
  }
}

// Descriptor vector from the 21x21 patch: one warp per keypoint.
// EXT = extended (128-d) descriptors: eight sums per 5x5 cell, split by the sign of the other gradient (A.4).
constexpr int VEC_WARPS = 8;
template <bool EXT>
__global__ void __launch_bounds__(VEC_WARPS * 32) k_surf_vector(const __grid_constant__ SurfBatch b) {
  __shared__ __align__(16) uint8_t s_p[VEC_WARPS][PATCH_STRIDE];
  __shared__ __align__(16) float s_dx[VEC_WARPS][400];
  __shared__ float s_dy[VEC_WARPS][400];
  __shared__ float s_dw[400];  // the Gaussian table: lanes index it divergently, which constant memory serialises
  const SurfImage& im = b.im[blockIdx.y];
  const int n = im.counters[1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < 400; t += VEC_WARPS * 32) s_dw[t] = c_DW[t];
  __syncthreads();
  uint8_t* P = s_p[wid];
  float* DX = s_dx[wid];
  float* DY = s_dy[wid];
  for (int k = blockIdx.x * VEC_WARPS + wid; k < n; k += gridDim.x * VEC_WARPS) {
    if (!(im.kps[k].size > 0.f)) continue;  // marked for deletion by k_surf_patch (warp-uniform)
    const unsigned* src = (const unsigned*)(im.patch + (size_t)k * PATCH_STRIDE);
    for (int t = lane; t < PATCH_STRIDE / 4; t += 32) ((unsigned*)P)[t] = __ldg(src + t);
    __syncwarp();
    // ---- Haar gradients with Gaussian weights ----
    for (int t = lane; t < 400; t += 32) {
      const int i = t / 20, j = t - i * 20;
      const int p00 = P[i * 21 + j], p01 = P[i * 21 + j + 1], p10 = P[i * 21 + 21 + j], p11 = P[i * 21 + 21 + j + 1];
      const float dw = s_dw[t];
      DX[t] = __fmul_rn((float)(p01 - p00 + p11 - p10), dw);
      DY[t] = __fmul_rn((float)(p10 - p00 + p11 - p01), dw);
    }
    __syncwarp();
    // ---- 4x4 cells x (sum dx, sum dy, sum |dx|, sum |dy|), y-major over each 5x5 cell ----
    // extended: per cell (sum dx, sum |dx|) over dy >= 0, then over dy < 0, (sum dy, sum |dy|) over dx >= 0, then
    // over dx < 0; an excluded sample leaves the running sum untouched, as the CPU's if / else does
    constexpr int NV = EXT ? 4 : 2;  // outputs per lane
    float v2[NV];
#pragma unroll
    for (int half = 0; half < NV; half++) {
      const int q = lane + 32 * half;
      float v = 0.f;
      if (EXT) {
        const int cell = q >> 3, comp = q & 7, ci = cell >> 2, cj = cell & 3;
        const int o = ci * 100 + cj * 5;
        const float* srcv = (comp & 4 ? DY : DX) + o;  // the summed gradient
        const float* selv = (comp & 4 ? DX : DY) + o;  // the gradient whose sign selects
        const bool use_abs = comp & 1, want_neg = comp & 2;
#pragma unroll
        for (int y = 0; y < 5; y++)
#pragma unroll
          for (int x = 0; x < 5; x++) {
            const float e = srcv[y * 20 + x];
            const bool nonneg = selv[y * 20 + x] >= 0.f;  // false for NaN: the CPU's else branch
            if (nonneg != want_neg) v = __fadd_rn(v, use_abs ? fabsf(e) : e);
          }
      } else {
        const int cell = q >> 2, comp = q & 3, ci = cell >> 2, cj = cell & 3;
        const float* srcv = ((comp & 1) ? DY : DX) + ci * 100 + cj * 5;
        const bool use_abs = comp >= 2;
        float e[25];
#pragma unroll
        for (int y = 0; y < 5; y++)
#pragma unroll
          for (int x = 0; x < 5; x++) e[y * 5 + x] = srcv[y * 20 + x];
#pragma unroll
        for (int q2 = 0; q2 < 25; q2++) v = __fadd_rn(v, use_abs ? fabsf(e[q2]) : e[q2]);
      }
      v2[half] = v;
    }
    __syncwarp();  // every lane is done reading DX / DY: DX is reused for the fp64 squares
    double* SQ = (double*)DX;  // 32 * NV doubles <= 400 floats
#pragma unroll
    for (int half = 0; half < NV; half++) SQ[lane + 32 * half] = (double)__fmul_rn(v2[half], v2[half]);
    __syncwarp();
    // ---- L2 normalisation: the squares are summed sequentially in fp64, as the CPU does ----
    float scale = 0.f;
    if (lane == 0) {
      double sq = 0;
#pragma unroll
      for (int q = 0; q < 32 * NV; q++) sq = __dadd_rn(sq, SQ[q]);
      scale = (float)(1. / (sqrt(sq) + (double)FLT_EPSILON));
    }
    scale = __shfl_sync(0xffffffffu, scale, 0);
#pragma unroll
    for (int half = 0; half < NV; half++)
      im.desc[(size_t)k * (32 * NV) + 32 * half + lane] = __fmul_rn(v2[half], scale);
    __syncwarp();
  }
}

void launch_surf_describe(Ctx& c, const SurfGeom& g, const SurfBatch& b, int capacity, int upright, int extended) {
  upload_tables(c);
  const SpanTable* tab = span_table(c);
  const int blocks = std::min(capacity, 8 * c.sm_count);
  UVO_KERNEL(c, "k_surf_patch");
  k_surf_patch<<<dim3(blocks, b.n_img), DESC_THREADS, 0, c.stream>>>(g, b, upright, tab);
  UVO_LAUNCH_CHECK(c);
  const int vblocks = std::min(div_up(capacity, VEC_WARPS), 4 * c.sm_count);
  UVO_KERNEL(c, "k_surf_vector");
  if (extended)
    k_surf_vector<true><<<dim3(vblocks, b.n_img), VEC_WARPS * 32, 0, c.stream>>>(b);
  else
    k_surf_vector<false><<<dim3(vblocks, b.n_img), VEC_WARPS * 32, 0, c.stream>>>(b);
  UVO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------------ compaction
// single block per image: ordered removal of keypoints marked size <= 0
__global__ void __launch_bounds__(1024) k_surf_compact(const __grid_constant__ SurfBatch b, uvo_keypoint* tmp_kps, float* tmp_desc,
                                                       int capacity, int dd) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const SurfImage& im = b.im[blockIdx.x];
  uvo_keypoint* tk = tmp_kps + (size_t)blockIdx.x * capacity;
  float* td = tmp_desc + (size_t)blockIdx.x * capacity * dd;
  const int n = im.counters[1];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int k = base + tid;
    const bool keep = k < n && im.kps[k].size > 0.f;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int q = 0; q < wid; q++) off += s_warp[q];
    if (keep) {
      const int dst = off + __popc(bal & ((1u << lane) - 1));
      tk[dst] = im.kps[k];
      for (int q = 0; q < dd; q++) td[(size_t)dst * dd + q] = im.desc[(size_t)k * dd + q];
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int q = 0; q < 32; q++) tot += s_warp[q];
      s_base += tot;
    }
    __syncthreads();
  }
  const int m = s_base;
  __syncthreads();
  for (int k = tid; k < m; k += 1024) im.kps[k] = tk[k];
  for (int q = tid; q < m * dd; q += 1024) im.desc[q] = td[q];
  if (tid == 0) im.counters[1] = m;
}

void launch_surf_compact(Ctx& c, const SurfBatch& b, int capacity, uvo_keypoint* tmp_kps, float* tmp_desc, int dd) {
  UVO_KERNEL(c, "k_surf_compact");
  k_surf_compact<<<b.n_img, 1024, 0, c.stream>>>(b, tmp_kps, tmp_desc, capacity, dd);
  UVO_LAUNCH_CHECK(c);
}

}  // namespace uvo
