// imgprep.cu -- K1 gray+undistort, K2 CLAHE, K3 integral image.
// Replaces get_image (reference VO_utility.cpp:337-379) and the integral() inside SURF::detectAndCompute
// (VO_utility.cpp:118).  All three are integer / exactly-specified f32 pipelines: results are bit-identical to the
// OpenCV CPU path (tests/test_gpu_imgprep.py compares against the oracle and committed cv2 fixtures).
#include <cfloat>
#include <cmath>

#include "imgprep.cuh"

namespace uvo {

// ------------------------------------------------------------------------------------------------ K1
// One thread produces 4 consecutive output pixels.  The rectification map is evaluated per pixel in fp64 exactly
// as cv::initUndistortRectifyMap does (closed form), quantised to 1/32 px, and the four bilinear taps
// are converted to gray on the fly with cvtColor's 15-bit weights (C.1) -- the gray image is never materialised.
// Algorithmic HBM bytes: read 3P (source, each byte touched ~once through L1/L2) + write P.
__device__ __forceinline__ int gray_tap(const uint8_t* __restrict__ src, size_t pitch, int w, int h, int x, int y) {
  if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return 0;  // BORDER_CONSTANT(0)
  const uint8_t* p = src + (size_t)y * pitch + 3 * x;
  return (9798 * (int)p[0] + 19235 * (int)p[1] + 3735 * (int)p[2] + 16384) >> 15;
}

// blockIdx.z selects the image of a stereo pair (second pointer set / camera); single-image launches use z = 1
__global__ void __launch_bounds__(256) k_gray_undistort(const uint8_t* src0, const uint8_t* src1,
                                                        size_t spitch, int w, int h, const __grid_constant__ UndistortParams P0,
                                                        const __grid_constant__ UndistortParams P1, uint8_t* dst0,
                                                        uint8_t* dst1, size_t dpitch) {
  const uint8_t* __restrict__ src = blockIdx.z ? src1 : src0;
  uint8_t* __restrict__ dst = blockIdx.z ? dst1 : dst0;
  const UndistortParams& P = blockIdx.z ? P1 : P0;
  const int j0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (j0 >= w || i >= h) return;
  uint8_t out[4];
  const double bx = i * P.ir[1] + P.ir[2], by = i * P.ir[4] + P.ir[5], bw = i * P.ir[7] + P.ir[8];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int j = j0 + k;
    double _x = j * P.ir[0] + bx, _y = j * P.ir[3] + by, _w = j * P.ir[6] + bw;
    double iw = 1. / _w, x = _x * iw, y = _y * iw;
    double x2 = x * x, y2 = y * y;
    double r2 = x2 + y2, _2xy = 2 * x * y;
    double kr = 1 + ((0 * r2 + P.k2) * r2 + P.k1) * r2;
    double xd = x * kr + P.p1 * _2xy + P.p2 * (r2 + 2 * x2);
    double yd = y * kr + P.p1 * (r2 + 2 * y2) + P.p2 * _2xy;
    double u = P.fx * xd + P.u0, v = P.fy * yd + P.v0;
    int iu = __double2int_rn(u * 32), iv = __double2int_rn(v * 32);
    int sx = iu >> 5, sy = iv >> 5;
    sx = min(max(sx, -32768), 32767);
    sy = min(max(sy, -32768), 32767);
    int fx = iu & 31, fy = iv & 31;
    int acc = gray_tap(src, spitch, w, h, sx, sy) * ((32 - fy) * (32 - fx) * 32) +
              gray_tap(src, spitch, w, h, sx + 1, sy) * ((32 - fy) * fx * 32) +
              gray_tap(src, spitch, w, h, sx, sy + 1) * (fy * (32 - fx) * 32) +
              gray_tap(src, spitch, w, h, sx + 1, sy + 1) * (fy * fx * 32);
    int r = (acc + 16384) >> 15;
    out[k] = (uint8_t)min(max(r, 0), 255);
  }
  uint8_t* d = dst + (size_t)i * dpitch + j0;
  if (j0 + 3 < w && (dpitch & 3) == 0) {
    *reinterpret_cast<uchar4*>(d) = make_uchar4(out[0], out[1], out[2], out[3]);
  } else {
    for (int k = 0; k < 4 && j0 + k < w; k++) d[k] = out[k];
  }
}

// plain gray image -> undistorted gray (used by the stage-level API when the caller already holds a gray image)
void launch_gray_undistort(Ctx& c, const uint8_t* d_src3, size_t spitch, int w, int h, const UndistortParams& P,
                           uint8_t* d_dst, size_t dpitch) {
  dim3 block(32, 8);
  dim3 grid(div_up(div_up(w, 4), 32), div_up(h, 8));
  UVO_KERNEL(c, "k_gray_undistort");
  k_gray_undistort<<<grid, block, 0, c.stream>>>(d_src3, d_src3, spitch, w, h, P, P, d_dst, d_dst, dpitch);
  UVO_LAUNCH_CHECK(c);
}

UndistortParams make_undistort_params(const uvo_camera& cam) {
  UndistortParams P;
  // (newK * I).inv(DECOMP_LU): cv::invert's closed form for 3x3 (cofactors times 1/det)
  const double m[9] = {cam.nfx, 0, cam.ncx, 0, cam.nfy, cam.ncy, 0, 0, 1};
  double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
             m[2] * (m[3] * m[7] - m[4] * m[6]);
  d = 1. / d;
  P.ir[0] = (m[4] * m[8] - m[5] * m[7]) * d;
  P.ir[1] = (m[2] * m[7] - m[1] * m[8]) * d;
  P.ir[2] = (m[1] * m[5] - m[2] * m[4]) * d;
  P.ir[3] = (m[5] * m[6] - m[3] * m[8]) * d;
  P.ir[4] = (m[0] * m[8] - m[2] * m[6]) * d;
  P.ir[5] = (m[2] * m[3] - m[0] * m[5]) * d;
  P.ir[6] = (m[3] * m[7] - m[4] * m[6]) * d;
  P.ir[7] = (m[1] * m[6] - m[0] * m[7]) * d;
  P.ir[8] = (m[0] * m[4] - m[1] * m[3]) * d;
  P.fx = cam.fx;
  P.fy = cam.fy;
  P.u0 = cam.cx;
  P.v0 = cam.cy;
  P.k1 = cam.k1;
  P.k2 = cam.k2;
  P.p1 = cam.p1;
  P.p2 = cam.p2;
  return P;
}

// ------------------------------------------------------------------------------------------------ Bayer demosaic
// cvtColor(COLOR_BayerBGGR2BGR) as from_ros_to_cv_image applies it to bayer-format camera messages (reference
// math_utility.cpp:161-164): OpenCV's bilinear demosaic -- rounded means of the 2 or 4 nearest samples of a colour at
// the interior pixels, first / last column and row copied from their neighbours.  One thread per 4 output pixels;
// the 3 x 6 neighbourhood is read through L1.  Bytes: read P, write 3P.
__global__ void __launch_bounds__(256) k_demosaic_bggr(const uint8_t* __restrict__ src, size_t spitch, int w, int h,
                                                       uint8_t* __restrict__ dst, size_t dpitch) {
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x0 >= w || y >= h) return;
  const int yc = min(max(y, 1), h - 2);
  const uint8_t* r0 = src + (size_t)(yc - 1) * spitch;
  const uint8_t* r1 = src + (size_t)yc * spitch;
  const uint8_t* r2 = src + (size_t)(yc + 1) * spitch;
  const bool ey = (yc & 1) == 0;
  uint8_t* o = dst + (size_t)y * dpitch + 3 * x0;
  for (int k = 0; k < 4 && x0 + k < w; k++) {
    const int xc = min(max(x0 + k, 1), w - 2);
    const int a = r0[xc - 1], b = r0[xc], c = r0[xc + 1], d = r1[xc - 1], v = r1[xc], e = r1[xc + 1], f = r2[xc - 1],
              g = r2[xc], i = r2[xc + 1];
    const int cross = (b + g + d + e + 2) >> 2, diag = (a + c + f + i + 2) >> 2, hor = (d + e + 1) >> 1,
              ver = (b + g + 1) >> 1;
    const bool ex = (xc & 1) == 0;
    int B, G, R;
    if (ey == ex) {  // blue (even, even) or red (odd, odd) site
      G = cross;
      B = ey ? v : diag;
      R = ey ? diag : v;
    } else {  // green site: blue row (even y) or red row
      G = v;
      B = ey ? hor : ver;
      R = ey ? ver : hor;
    }
    o[3 * k] = (uint8_t)B;
    o[3 * k + 1] = (uint8_t)G;
    o[3 * k + 2] = (uint8_t)R;
  }
}

void launch_demosaic_bggr(Ctx& c, const uint8_t* d_src, size_t spitch, int w, int h, uint8_t* d_dst3, size_t dpitch) {
  dim3 block(32, 8), grid(div_up(div_up(w, 4), 32), div_up(h, 8));
  UVO_KERNEL(c, "k_demosaic_bggr");
  k_demosaic_bggr<<<grid, block, 0, c.stream>>>(d_src, spitch, w, h, d_dst3, dpitch);
  UVO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------------ K2 CLAHE
__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
  return p;
}

// One block per (tile, image): histogram of the tile in shared memory (four sub-histograms, one aligned 32-bit load
// per four pixels), then clip, redistribute, cumulative sum and the LUT (OpenCV clahe.cpp CLAHE_CalcLut_Body) in the
// same block -- no global histogram, no atomics to global memory, no separate LUT launch.  Pixels right of / below the
// image (tile sizes that do not divide it) are BORDER_REFLECT_101 copies, as copyMakeBorder provides them.
constexpr int CLAHE_HIST_THREADS = 1024, CLAHE_HIST_UN = 5;
__global__ void __launch_bounds__(CLAHE_HIST_THREADS) k_clahe_tile_lut(const uint8_t* img0, const uint8_t* img1,
                                                                       size_t pitch, int w, int h, ClaheGeom g,
                                                                       uint8_t* __restrict__ lut) {
  __shared__ unsigned int sh[4][256];
  __shared__ int warp_sums[8];
  __shared__ int s_total;
  const uint8_t* __restrict__ img = blockIdx.z ? img1 : img0;
  lut += (size_t)blockIdx.z * g.tiles_x * g.tiles_y * 256;  // the pair's LUTs are contiguous
  const int b = threadIdx.x, lane = b & 31, wid = b >> 5;
  sh[b >> 8][b & 255] = 0;
  __syncthreads();
  const int tile = blockIdx.x, txi = tile % g.tiles_x, tyi = tile / g.tiles_x;
  const int x0 = txi * g.tw, y0 = tyi * g.th;
  unsigned int* mine = sh[wid & 3];
  const bool words = x0 + g.tw <= w && y0 + g.th <= h && ((g.tw | x0) & 3) == 0 && (pitch & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(img) & 3) == 0;
  if (words) {
    // batches of CLAHE_HIST_UN words per thread, all loads of a batch in flight before the first atomic
    const int wpr = g.tw >> 2, nw = wpr * g.th;
    for (int t0 = 0; t0 < nw; t0 += CLAHE_HIST_THREADS * CLAHE_HIST_UN) {
      unsigned v[CLAHE_HIST_UN];
      bool in[CLAHE_HIST_UN];
#pragma unroll
      for (int u = 0; u < CLAHE_HIST_UN; u++) {
        const int t = t0 + u * CLAHE_HIST_THREADS + b;
        in[u] = t < nw;
        const int y = t / wpr, xw = t - y * wpr;
        v[u] = in[u] ? __ldg(reinterpret_cast<const unsigned*>(img + (size_t)(y0 + y) * pitch + x0) + xw) : 0u;
      }
#pragma unroll
      for (int u = 0; u < CLAHE_HIST_UN; u++) {
        if (!in[u]) continue;
        atomicAdd(&mine[v[u] & 255u], 1u);
        atomicAdd(&mine[(v[u] >> 8) & 255u], 1u);
        atomicAdd(&mine[(v[u] >> 16) & 255u], 1u);
        atomicAdd(&mine[v[u] >> 24], 1u);
      }
    }
  } else {
    const int npx = g.tw * g.th;
    for (int t = b; t < npx; t += blockDim.x) {
      int y = y0 + t / g.tw, x = x0 + t % g.tw;
      if (y >= h) y = reflect101(y, h);
      if (x >= w) x = reflect101(x, w);
      atomicAdd(&mine[img[(size_t)y * pitch + x]], 1u);
    }
  }
  __syncthreads();
  if (b >= 256) return;
  // the LUT on the first 256 threads (8 warps); their barriers are named and counted, the other warps are gone
  auto sync256 = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
  int v = (int)(sh[0][b] + sh[1][b] + sh[2][b] + sh[3][b]);
  if (g.clip > 0) {
    int excess = v > g.clip ? v - g.clip : 0;
    if (v > g.clip) v = g.clip;
    int e = excess;
#pragma unroll
    for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) warp_sums[wid] = e;
    sync256();
    if (b == 0) {
      int t = 0;
      for (int k = 0; k < 8; k++) t += warp_sums[k];
      s_total = t;
    }
    sync256();
    const int clipped = s_total;
    const int batch = clipped / 256, residual = clipped - batch * 256;
    v += batch;
    if (residual != 0) {
      const int step = max(256 / residual, 1);
      if (b % step == 0 && b / step < residual) v += 1;
    }
    sync256();
  }
  // inclusive scan over the 256 bins
  int s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, s, o);
    if (lane >= o) s += t;
  }
  if (lane == 31) warp_sums[wid] = s;
  sync256();
  int base = 0;
  for (int k = 0; k < wid; k++) base += warp_sums[k];
  s += base;
  float f = __fmul_rn((float)s, g.lut_scale);
  int r = __float2int_rn(f);
  lut[tile * 256 + b] = (uint8_t)min(max(r, 0), 255);
}

// bilinear blend of the four neighbouring tile LUTs (CLAHE_Interpolation_Body); 4 px per thread, in place allowed.
// Optionally also emits the per-row inclusive prefix sums of the result (first half of the integral image, K3):
// not fused here -- see k_integral_rows.
__global__ void __launch_bounds__(256) k_clahe_apply(const uint8_t* src0, const uint8_t* src1,
                                                     size_t spitch, int w, int h, ClaheGeom g,
                                                     const uint8_t* __restrict__ lut, uint8_t* dst0,
                                                     uint8_t* dst1, size_t dpitch) {
  const uint8_t* __restrict__ src = blockIdx.z ? src1 : src0;
  uint8_t* __restrict__ dst = blockIdx.z ? dst1 : dst0;
  lut += (size_t)blockIdx.z * g.tiles_x * g.tiles_y * 256;
  const int j0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (j0 >= w || y >= h) return;
  const float tyf = __fsub_rn(__fmul_rn((float)y, g.inv_th), 0.5f);
  int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
  const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
  ty1 = max(ty1, 0);
  ty2 = min(ty2, g.tiles_y - 1);
  const uint8_t* l1 = lut + (size_t)ty1 * g.tiles_x * 256;
  const uint8_t* l2 = lut + (size_t)ty2 * g.tiles_x * 256;
  uint8_t in[4], out[4];
  const uint8_t* s = src + (size_t)y * spitch + j0;
  const bool vec = (j0 + 3 < w) && ((spitch & 3) == 0) && ((dpitch & 3) == 0);
  if (vec) {
    uchar4 q = *reinterpret_cast<const uchar4*>(s);
    in[0] = q.x; in[1] = q.y; in[2] = q.z; in[3] = q.w;
  } else {
    for (int k = 0; k < 4; k++) in[k] = (j0 + k < w) ? s[k] : 0;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = j0 + k;
    const float txf = __fsub_rn(__fmul_rn((float)x, g.inv_tw), 0.5f);
    int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
    const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
    tx1 = max(tx1, 0);
    tx2 = min(tx2, g.tiles_x - 1);
    const int v = in[k];
    const float a = (float)l1[tx1 * 256 + v], b = (float)l1[tx2 * 256 + v];
    const float c = (float)l2[tx1 * 256 + v], e = (float)l2[tx2 * 256 + v];
    const float top = __fadd_rn(__fmul_rn(a, xa1), __fmul_rn(b, xa));
    const float bot = __fadd_rn(__fmul_rn(c, xa1), __fmul_rn(e, xa));
    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
    const int r = __float2int_rn(res);
    out[k] = (uint8_t)min(max(r, 0), 255);
  }
  uint8_t* d = dst + (size_t)y * dpitch + j0;
  if (vec) *reinterpret_cast<uchar4*>(d) = make_uchar4(out[0], out[1], out[2], out[3]);
  else
    for (int k = 0; k < 4 && j0 + k < w; k++) d[k] = out[k];
}

ClaheGeom make_clahe_geom(int w, int h, double clip_limit, int tiles_x, int tiles_y) {
  ClaheGeom g;
  int pw = w, ph = h;
  if (w % tiles_x != 0 || h % tiles_y != 0) {
    pw = w + (tiles_x - (w % tiles_x));
    ph = h + (tiles_y - (h % tiles_y));
  }
  g.tiles_x = tiles_x;
  g.tiles_y = tiles_y;
  g.tw = pw / tiles_x;
  g.th = ph / tiles_y;
  const int area = g.tw * g.th;
  g.clip = 0;
  if (clip_limit > 0.0) {
    g.clip = (int)(clip_limit * area / 256);
    if (g.clip < 1) g.clip = 1;
  }
  g.lut_scale = (float)255 / area;
  g.inv_tw = 1.0f / g.tw;
  g.inv_th = 1.0f / g.th;
  return g;
}

void launch_clahe(Ctx& c, const uint8_t* d_src, size_t spitch, int w, int h, const ClaheGeom& g, uint8_t* d_lut,
                  uint8_t* d_dst, size_t dpitch) {
  const int tiles = g.tiles_x * g.tiles_y;
  UVO_KERNEL(c, "k_clahe_tile_lut");
  k_clahe_tile_lut<<<tiles, CLAHE_HIST_THREADS, 0, c.stream>>>(d_src, d_src, spitch, w, h, g, d_lut);
  UVO_LAUNCH_CHECK(c);
  dim3 block(32, 8), grid(div_up(div_up(w, 4), 32), div_up(h, 8));
  UVO_KERNEL(c, "k_clahe_apply");
  k_clahe_apply<<<grid, block, 0, c.stream>>>(d_src, d_src, spitch, w, h, g, d_lut, d_dst, d_dst, dpitch);
  UVO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------------ K3 integral
// Pass 1: one warp per image row; inclusive prefix along the row, written to sum[(i+1)][1..w]; also zeroes column 0
// and (warp 0) row 0.  The row is taken in batches of INT_UN x 128 pixels whose loads are all issued before the first
// scan (a warp is alone with its row: what it waits for is the load latency, once per batch instead of once per 128 px).
// Pass 2: column prefix in place, each block owns INT_CW columns x (1024 / INT_CW) row segments.
constexpr int INT_UN = 5;
__global__ void __launch_bounds__(256) k_integral_rows(const uint8_t* img0, const uint8_t* img1,
                                                       size_t pitch, int w, int h, int32_t* sum0,
                                                       int32_t* sum1, int sw) {
  const uint8_t* __restrict__ img = blockIdx.z ? img1 : img0;
  int32_t* __restrict__ sum = blockIdx.z ? sum1 : sum0;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  // sw: row pitch of the integral in elements (>= w + 1; the SURF front end pads it to 16 bytes for TMA)
  if (warp == 0)
    for (int j = lane; j < w + 1; j += 32) sum[j] = 0;
  if (warp >= h) return;
  const uint8_t* row = img + (size_t)warp * pitch;
  int32_t* out = sum + (size_t)(warp + 1) * sw;
  if (lane == 0) out[0] = 0;
  const bool aligned = (pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(img) & 3) == 0;
  int carry = 0;
  for (int base0 = 0; base0 < w; base0 += 128 * INT_UN) {
    unsigned q[INT_UN];
#pragma unroll
    for (int u = 0; u < INT_UN; u++) {
      const int j = base0 + 128 * u + lane * 4;
      unsigned v = 0;
      if (j + 3 < w && aligned) {
        v = __ldg(reinterpret_cast<const unsigned*>(row + j));
      } else {
        if (j < w) v |= row[j];
        if (j + 1 < w) v |= (unsigned)row[j + 1] << 8;
        if (j + 2 < w) v |= (unsigned)row[j + 2] << 16;
        if (j + 3 < w) v |= (unsigned)row[j + 3] << 24;
      }
      q[u] = v;
    }
#pragma unroll
    for (int u = 0; u < INT_UN; u++) {
      const int j = base0 + 128 * u + lane * 4;
      if (base0 + 128 * u >= w) break;  // warp-uniform
      int v0 = q[u] & 255u, v1 = (q[u] >> 8) & 255u, v2 = (q[u] >> 16) & 255u, v3 = q[u] >> 24;
      v1 += v0; v2 += v1; v3 += v2;
      int s = v3;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      const int excl = s - v3 + carry;
      if (j < w) out[j + 1] = v0 + excl;
      if (j + 1 < w) out[j + 2] = v1 + excl;
      if (j + 2 < w) out[j + 3] = v2 + excl;
      if (j + 3 < w) out[j + 4] = v3 + excl;
      carry += __shfl_sync(0xffffffffu, s, 31);
    }
  }
}

// block = INT_CW (columns) x INT_SEG (row segments).  Segment totals -> exclusive scan over the segments of each
// column (one warp per column) -> second sweep adds.  Block shape measured on the stereo pair (tools/pair_probe.py):
// 16 x 64 is the fastest kernel run alone (17 us against 22 for 32 x 32) but 32 x 32 -- 80 blocks, half the SMs left
// to the kernels of the other frames in flight -- gives the highest frame rate, which is what counts.
#ifndef UVO_INT_CW
#define UVO_INT_CW 32
#endif
#ifndef UVO_INT_SEG
#define UVO_INT_SEG 32
#endif
constexpr int INT_CW = UVO_INT_CW, INT_SEG = UVO_INT_SEG;
static_assert(INT_CW * INT_SEG <= 1024 && INT_CW * 32 <= INT_CW * INT_SEG && INT_SEG % 32 == 0,
              "column scan: one warp per column, whole words of segments per lane");
__global__ void __launch_bounds__(INT_CW* INT_SEG) k_integral_cols(int32_t* sum0, int32_t* sum1, int w,
                                                        int h, int sw) {
  __shared__ int32_t tot[INT_SEG][INT_CW + 1];
  int32_t* __restrict__ sum = blockIdx.z ? sum1 : sum0;
  const int col = 1 + blockIdx.x * INT_CW + threadIdx.x;
  const int seg = threadIdx.y;
  const int rows_per = (h + INT_SEG - 1) / INT_SEG;
  const int r0 = min(1 + seg * rows_per, h + 1), r1 = min(r0 + rows_per, h + 1);
  int32_t acc = 0;
  if (col < w + 1)
    for (int r = r0; r < r1; r++) acc += sum[(size_t)r * sw + col];
  tot[seg][threadIdx.x] = acc;
  __syncthreads();
  {  // exclusive scan of tot[.][c] over the segments: warp c takes column c, INT_SEG / 32 consecutive segments per lane
    const int t = threadIdx.y * INT_CW + threadIdx.x, c = t >> 5, lane = t & 31;
    if (c < INT_CW) {
      constexpr int PER = INT_SEG / 32;
      int32_t loc[PER], run = 0;
#pragma unroll
      for (int k = 0; k < PER; k++) {
        loc[k] = tot[lane * PER + k][c];
        run += loc[k];
      }
      int32_t s = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t x = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += x;
      }
      int32_t e = s - run;
#pragma unroll
      for (int k = 0; k < PER; k++) {
        tot[lane * PER + k][c] = e;
        e += loc[k];
      }
    }
  }
  __syncthreads();
  int32_t run = tot[seg][threadIdx.x];
  if (col < w + 1)
    for (int r = r0; r < r1; r++) {
      run += sum[(size_t)r * sw + col];
      sum[(size_t)r * sw + col] = run;
    }
}

void launch_integral(Ctx& c, const uint8_t* d_img, size_t pitch, int w, int h, int32_t* d_sum, int sum_pitch) {
  const int warps_per_block = 8;
  const int sw = sum_pitch > 0 ? sum_pitch : w + 1;
  UVO_KERNEL(c, "k_integral_rows");
  k_integral_rows<<<div_up(h, warps_per_block), 32 * warps_per_block, 0, c.stream>>>(d_img, d_img, pitch, w, h, d_sum,
                                                                                      d_sum, sw);
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_integral_cols");
  k_integral_cols<<<div_up(w, INT_CW), dim3(INT_CW, INT_SEG), 0, c.stream>>>(d_sum, d_sum, w, h, sw);
  UVO_LAUNCH_CHECK(c);
}

// K1-K3 for both images of a stereo pair in one launch per kernel (blockIdx.z = image): half the launches, and the
// narrow kernels (tile histograms + LUTs, the two scan passes) get twice the blocks.  d_lut: 2 * tiles * 256.
void launch_prep_pair(Ctx& c, const uint8_t* d_src3[2], size_t spitch, int w, int h, const UndistortParams P[2], int clahe,
                      const ClaheGeom& g, uint8_t* d_lut, uint8_t* d_gray[2], size_t gpitch,
                      int32_t* d_sum[2], int part, int sum_pitch) {
  if (part & PREP_PART_SOURCE) {
    dim3 block(32, 8), grid(div_up(div_up(w, 4), 32), div_up(h, 8), 2);
    UVO_KERNEL(c, "k_gray_undistort");
    k_gray_undistort<<<grid, block, 0, c.stream>>>(d_src3[0], d_src3[1], spitch, w, h, P[0], P[1], d_gray[0], d_gray[1],
                                                   gpitch);
    UVO_LAUNCH_CHECK(c);
  }
  if (!(part & PREP_PART_REST)) return;
  if (clahe) {
    const int tiles = g.tiles_x * g.tiles_y;
    UVO_KERNEL(c, "k_clahe_tile_lut");
    k_clahe_tile_lut<<<dim3(tiles, 1, 2), CLAHE_HIST_THREADS, 0, c.stream>>>(d_gray[0], d_gray[1], gpitch, w, h, g, d_lut);
    UVO_LAUNCH_CHECK(c);
    dim3 block(32, 8), grid(div_up(div_up(w, 4), 32), div_up(h, 8), 2);
    UVO_KERNEL(c, "k_clahe_apply");
    k_clahe_apply<<<grid, block, 0, c.stream>>>(d_gray[0], d_gray[1], gpitch, w, h, g, d_lut, d_gray[0], d_gray[1],
                                                gpitch);
    UVO_LAUNCH_CHECK(c);
  }
  const int warps_per_block = 8;
  const int sw = sum_pitch > 0 ? sum_pitch : w + 1;
  UVO_KERNEL(c, "k_integral_rows");
  k_integral_rows<<<dim3(div_up(h, warps_per_block), 1, 2), 32 * warps_per_block, 0, c.stream>>>(
      d_gray[0], d_gray[1], gpitch, w, h, d_sum[0], d_sum[1], sw);
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_integral_cols");
  k_integral_cols<<<dim3(div_up(w, INT_CW), 1, 2), dim3(INT_CW, INT_SEG), 0, c.stream>>>(d_sum[0], d_sum[1], w, h, sw);
  UVO_LAUNCH_CHECK(c);
}

// ------------------------------------------------------------------------------------------------ K0 INTER_AREA
// OpenCV's ResizeArea_Invoker / ResizeAreaFast for u8 (imgproc/src/resize.cpp): scale in fp64, weights
// stored as f32, horizontal accumulation per source row in table order, then vertical accumulation, rint + saturate;
// integer scales in both directions take the block-sum path ((s + 2) >> 2 for 2x2, rint(s * (1.f / area)) otherwise).
__global__ void k_area_tab(int ssize, int dsize, double scale, AreaCell* __restrict__ tab) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dsize) return;
  const double fsx1 = d * scale, fsx2 = fsx1 + scale;
  const double cell = fmin(scale, ssize - fsx1);
  int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
  sx2 = min(sx2, ssize - 1);
  sx1 = min(sx1, sx2);
  AreaCell t;
  t.sx1 = sx1;
  t.sx2 = sx2;
  t.has_l = (sx1 - fsx1 > 1e-3) ? 1 : 0;
  t.a_l = (float)((sx1 - fsx1) / cell);
  t.a_f = (float)(1.0 / cell);
  t.has_r = (fsx2 - sx2 > 1e-3) ? 1 : 0;
  t.a_r = (float)(fmin(fmin(fsx2 - sx2, 1.), cell) / cell);
  tab[d] = t;
}

__global__ void __launch_bounds__(256) k_resize_area(const uint8_t* __restrict__ src, size_t spitch, int sw, int sh,
                                                     int cn, uint8_t* __restrict__ dst, size_t dpitch, int dw, int dh,
                                                     const AreaCell* __restrict__ xtab,
                                                     const AreaCell* __restrict__ ytab, int iscale_x, int iscale_y,
                                                     int fast) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;  // element of the destination row: dx * cn + c
  const int dy = blockIdx.y;
  if (e >= dw * cn) return;
  const int dx = e / cn, ch = e - dx * cn;
  int outv;
  if (fast) {
    int sum = 0;
    for (int ky = 0; ky < iscale_y; ky++) {
      const int sy = dy * iscale_y + ky;
      if (sy >= sh) break;
      const uint8_t* S = src + (size_t)sy * spitch + ch;
      for (int kx = 0; kx < iscale_x; kx++) {
        const int sx = dx * iscale_x + kx;
        if (sx < sw) sum += __ldg(S + sx * cn);
      }
    }
    if (iscale_x == 2 && iscale_y == 2) outv = (sum + 2) >> 2;
    else outv = min(max(__float2int_rn(__fmul_rn((float)sum, 1.f / (float)(iscale_x * iscale_y))), 0), 255);
  } else {
    const AreaCell xs = xtab[dx], ys = ytab[dy];
    float sum = 0.f;
    bool first = true;
    auto row = [&](int sy, float beta) {
      const uint8_t* S = src + (size_t)sy * spitch + ch;
      float buf = 0.f;
      if (xs.has_l) buf = __fadd_rn(buf, __fmul_rn((float)__ldg(S + (xs.sx1 - 1) * cn), xs.a_l));
      for (int sx = xs.sx1; sx < xs.sx2; sx++) buf = __fadd_rn(buf, __fmul_rn((float)__ldg(S + sx * cn), xs.a_f));
      if (xs.has_r) buf = __fadd_rn(buf, __fmul_rn((float)__ldg(S + xs.sx2 * cn), xs.a_r));
      sum = first ? __fmul_rn(beta, buf) : __fadd_rn(sum, __fmul_rn(beta, buf));
      first = false;
    };
    if (ys.has_l) row(ys.sx1 - 1, ys.a_l);
    for (int sy = ys.sx1; sy < ys.sx2; sy++) row(sy, ys.a_f);
    if (ys.has_r) row(ys.sx2, ys.a_r);
    outv = min(max(__float2int_rn(sum), 0), 255);
  }
  dst[(size_t)dy * dpitch + e] = (uint8_t)outv;
}

void launch_resize_area(Ctx& c, const uint8_t* d_src, size_t spitch, int sw, int sh, int cn, uint8_t* d_dst,
                        size_t dpitch, int dw, int dh, AreaCell* d_tab) {
  if (sw == dw && sh == dh) {
    UVO_CUDA(cudaMemcpy2DAsync(d_dst, dpitch, d_src, spitch, (size_t)sw * cn, sh, cudaMemcpyDeviceToDevice, c.stream));
    return;
  }
  UVO_REQUIRE(dw <= sw && dh <= sh, "resize: INTER_AREA enlargement is not implemented (the reference only shrinks)");
  // cv::resize derives inv_scale = dsize / ssize and hal::resize inverts it again (scale = 1. / inv_scale)
  const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
  const double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
  const int isx = std::max((int)nearbyint(scale_x), 1), isy = std::max((int)nearbyint(scale_y), 1);
  const int fast = (fabs(scale_x - isx) < DBL_EPSILON && fabs(scale_y - isy) < DBL_EPSILON) ? 1 : 0;
  if (!fast) {
    UVO_KERNEL(c, "k_area_tab");
    k_area_tab<<<div_up(dw, 128), 128, 0, c.stream>>>(sw, dw, scale_x, d_tab);
    UVO_LAUNCH_CHECK(c);
    UVO_KERNEL(c, "k_area_tab");
    k_area_tab<<<div_up(dh, 128), 128, 0, c.stream>>>(sh, dh, scale_y, d_tab + dw);
    UVO_LAUNCH_CHECK(c);
  }
  UVO_KERNEL(c, "k_resize_area");
  k_resize_area<<<dim3(div_up(dw * cn, 256), dh), 256, 0, c.stream>>>(d_src, spitch, sw, sh, cn, d_dst, dpitch, dw, dh,
                                                                     d_tab, d_tab + dw, isx, isy, fast);
  UVO_LAUNCH_CHECK(c);
}

}  // namespace uvo
