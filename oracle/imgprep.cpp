// oracle/imgprep.cpp -- CPU ORACLE (test infrastructure only; see uvo_oracle.h).
// Restates the OpenCV imgproc calls made by get_image (reference VO_utility.cpp:337-379) and the integral image
// that SURF::detectAndCompute builds (VO_utility.cpp:117-118).  OpenCV is not vendored by the reference; the
// specs below are SURVEY.md Appendix C.1-C.5 and are pinned against cv2 4.13 by tests/test_oracle_imgprep.py.
#include "uvo_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

// rows [0, n) split over the host threads, like OpenCV's parallel_for_ over stripes: every output element is an
// independent function of the inputs, so the result does not depend on the split
template <class F>
static void parallel_rows(int n, F f) {
  const unsigned nt = std::max(1u, std::min({16u, std::thread::hardware_concurrency(), (unsigned)std::max(n / 16, 1)}));
  if (nt == 1) {
    f(0, n);
    return;
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++) th.emplace_back([=] { f((int)((long long)n * t / nt), (int)((long long)n * (t + 1) / nt)); });
  for (auto& x : th) x.join();
}

static inline int round_half_even(double v) { return (int)std::nearbyint(v); }  // cvRound (default FE_TONEAREST)
static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// ---- cvtColor(COLOR_RGB2GRAY), u8: VO_utility.cpp:347 ; SURVEY C.1 ----
extern "C" void orc_gray(const uint8_t* s, int w, int h, uint8_t* d) {
  parallel_rows(h, [=](int r0, int r1) {
    for (size_t i = (size_t)r0 * w; i < (size_t)r1 * w; i++)
      d[i] = (uint8_t)((9798 * s[3 * i] + 19235 * s[3 * i + 1] + 3735 * s[3 * i + 2] + 16384) >> 15);
  });
}

// 3x3 inverse as cv::invert(DECOMP_LU) does for n==3 (cofactors * 1/det, in double).
static void inv3x3(const double m[9], double o[9]) {
  double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
             m[2] * (m[3] * m[7] - m[4] * m[6]);
  d = 1. / d;
  o[0] = (m[4] * m[8] - m[5] * m[7]) * d;
  o[1] = (m[2] * m[7] - m[1] * m[8]) * d;
  o[2] = (m[1] * m[5] - m[2] * m[4]) * d;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * d;
  o[4] = (m[0] * m[8] - m[2] * m[6]) * d;
  o[5] = (m[2] * m[3] - m[0] * m[5]) * d;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * d;
  o[7] = (m[1] * m[6] - m[0] * m[7]) * d;
  o[8] = (m[0] * m[4] - m[1] * m[3]) * d;
}

// ---- cv::undistort == initUndistortRectifyMap(CV_16SC2) + remap: VO_utility.cpp:350 ; SURVEY C.2 ----
extern "C" void orc_undistort_map(const double K[4], const double D[4], const double nK[4], int w, int h,
                                  int16_t* mxy, uint16_t* mfr) {
  const double A[9] = {nK[0], 0, nK[2], 0, nK[1], nK[3], 0, 0, 1};
  double ir[9];
  inv3x3(A, ir);
  const double fx = K[0], fy = K[1], u0 = K[2], v0 = K[3];
  const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3];
  parallel_rows(h, [=](int r0, int r1) {
  for (int i = r0; i < r1; i++) {
    for (int j = 0; j < w; j++) {
      double _x = j * ir[0] + (i * ir[1] + ir[2]);
      double _y = j * ir[3] + (i * ir[4] + ir[5]);
      double _w = j * ir[6] + (i * ir[7] + ir[8]);
      double iw = 1. / _w, x = _x * iw, y = _y * iw;
      double x2 = x * x, y2 = y * y;
      double r2 = x2 + y2, _2xy = 2 * x * y;
      double kr = 1 + ((0 * r2 + k2) * r2 + k1) * r2;  // k3 = 0; rational denominator = 1
      double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
      double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
      double u = fx * xd + u0;
      double v = fy * yd + v0;
      int iu = round_half_even(u * 32), iv = round_half_even(v * 32);  // saturate_cast<int>(u*INTER_TAB_SIZE)
      int mx = iu >> 5, my = iv >> 5;
      mx = std::min(std::max(mx, -32768), 32767);  // saturate_cast<short>
      my = std::min(std::max(my, -32768), 32767);
      mxy[2 * ((size_t)i * w + j)] = (int16_t)mx;
      mxy[2 * ((size_t)i * w + j) + 1] = (int16_t)my;
      mfr[(size_t)i * w + j] = (uint16_t)(((iv & 31) << 5) | (iu & 31));
    }
  }
  });
}

// ---- remap INTER_LINEAR fixed point, BORDER_CONSTANT 0: SURVEY C.3 ----
// weights for fractions (fx,fy)/32 are exact: (32-fy)(32-fx)*32 etc. sum to 32768, so the table fix-up never fires.
extern "C" void orc_remap_bilinear(const uint8_t* s, int w, int h, const int16_t* mxy, const uint16_t* mfr,
                                   uint8_t* d) {
  parallel_rows(h, [=](int r0, int r1) {
  for (int i = r0; i < r1; i++)
    for (int j = 0; j < w; j++) {
      size_t o = (size_t)i * w + j;
      int sx = mxy[2 * o], sy = mxy[2 * o + 1];
      int fxi = mfr[o] & 31, fyi = (mfr[o] >> 5) & 31;
      int w00 = (32 - fyi) * (32 - fxi) * 32, w01 = (32 - fyi) * fxi * 32, w10 = fyi * (32 - fxi) * 32,
          w11 = fyi * fxi * 32;
      auto px = [&](int x, int y) -> int {
        return ((unsigned)x < (unsigned)w && (unsigned)y < (unsigned)h) ? s[(size_t)y * w + x] : 0;
      };
      int acc = px(sx, sy) * w00 + px(sx + 1, sy) * w01 + px(sx, sy + 1) * w10 + px(sx + 1, sy + 1) * w11;
      d[o] = sat_u8((acc + 16384) >> 15);
    }
  });
}

extern "C" void orc_undistort(const uint8_t* g, int w, int h, const double K[4], const double D[4],
                              const double nK[4], uint8_t* d) {
  std::vector<int16_t> mxy((size_t)w * h * 2);
  std::vector<uint16_t> mfr((size_t)w * h);
  orc_undistort_map(K, D, nK, w, h, mxy.data(), mfr.data());
  orc_remap_bilinear(g, w, h, mxy.data(), mfr.data(), d);
}

// ---- CLAHE: VO_utility.cpp:352-357 ; SURVEY C.4 (OpenCV imgproc/src/clahe.cpp) ----
static inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

extern "C" void orc_clahe(const uint8_t* src, int w, int h, double clip_limit, int tx, int ty, uint8_t* dst) {
  // non-divisible sizes are padded right/bottom with BORDER_REFLECT_101 for the histogram pass only
  int pw = w, ph = h;
  if (w % tx != 0 || h % ty != 0) {
    pw = w + (tx - (w % tx));  // OpenCV pads by tiles - (size % tiles) on BOTH axes (a full tile count when
    ph = h + (ty - (h % ty));  // that axis was already divisible)
  }
  const int tw = pw / tx, th = ph / ty;
  const int area = tw * th;
  int clip = 0;
  if (clip_limit > 0.0) {
    clip = (int)(clip_limit * area / 256);
    clip = std::max(clip, 1);
  }
  const float lut_scale = (float)255 / area;
  std::vector<uint8_t> lut((size_t)tx * ty * 256);
  uint8_t* lutp = lut.data();
  parallel_rows(tx * ty, [=](int t0, int t1) {
  for (int t = t0; t < t1; t++) {
    int tyi = t / tx, txi = t % tx;
    int hist[256];
    std::memset(hist, 0, sizeof(hist));
    for (int y = tyi * th; y < (tyi + 1) * th; y++)
      for (int x = txi * tw; x < (txi + 1) * tw; x++) {
        int sy = y < h ? y : reflect101(y, h), sx = x < w ? x : reflect101(x, w);
        hist[src[(size_t)sy * w + sx]]++;
      }
    if (clip > 0) {
      int clipped = 0;
      for (int i = 0; i < 256; i++)
        if (hist[i] > clip) {
          clipped += hist[i] - clip;
          hist[i] = clip;
        }
      int batch = clipped / 256, residual = clipped - batch * 256;
      for (int i = 0; i < 256; i++) hist[i] += batch;
      if (residual != 0) {
        int step = std::max(256 / residual, 1);
        for (int i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++;
      }
    }
    int sum = 0;
    for (int i = 0; i < 256; i++) {
      sum += hist[i];
      float v = (float)sum * lut_scale;
      lutp[(size_t)t * 256 + i] = sat_u8(round_half_even(v));
    }
  }
  });
  std::vector<uint8_t> out((size_t)w * h);
  uint8_t* outp = out.data();
  const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
  parallel_rows(h, [=](int y0, int y1) {
  for (int y = y0; y < y1; y++) {
    float tyf = y * inv_th - 0.5f;
    int ty1 = (int)std::floor(tyf), ty2 = ty1 + 1;
    float ya = tyf - ty1, ya1 = 1.0f - ya;
    ty1 = std::max(ty1, 0);
    ty2 = std::min(ty2, ty - 1);
    for (int x = 0; x < w; x++) {
      float txf = x * inv_tw - 0.5f;
      int tx1 = (int)std::floor(txf), tx2 = tx1 + 1;
      float xa = txf - tx1, xa1 = 1.0f - xa;
      tx1 = std::max(tx1, 0);
      tx2 = std::min(tx2, tx - 1);
      int v = src[(size_t)y * w + x];
      float a = lutp[((size_t)ty1 * tx + tx1) * 256 + v], b = lutp[((size_t)ty1 * tx + tx2) * 256 + v];
      float c = lutp[((size_t)ty2 * tx + tx1) * 256 + v], e = lutp[((size_t)ty2 * tx + tx2) * 256 + v];
      float res = (a * xa1 + b * xa) * ya1 + (c * xa1 + e * xa) * ya;
      outp[(size_t)y * w + x] = sat_u8(round_half_even(res));
    }
  }
  });
  std::memcpy(dst, out.data(), out.size());
}

extern "C" void orc_get_image(const uint8_t* src3, int w, int h, const double K[4], const double D[4],
                              const double nK[4], int clahe, double clip_limit, uint8_t* dst) {
  std::vector<uint8_t> g((size_t)w * h);
  orc_gray(src3, w, h, g.data());
  orc_undistort(g.data(), w, h, K, D, nK, dst);
  if (clahe) orc_clahe(dst, w, h, clip_limit, 8, 8, dst);
}

// ---- cvtColor(COLOR_BayerBGGR2BGR), u8 (from_ros_to_cv_image, math_utility.cpp:161-164): OpenCV's bilinear demosaic.
// Interior pixels (1 <= y <= h-2, 1 <= x <= w-2): the missing colours are rounded means of the 2 or 4 nearest samples
// of that colour; the first / last column then copy their neighbour column and the first / last row their neighbour
// row.  Pinned bit-exact against cv2 4.13 (tests/test_oracle_imgprep.py).  Sites: (even, even) blue, (odd, odd) red.
extern "C" void orc_bayer_bggr2bgr(const uint8_t* s, int w, int h, uint8_t* d) {
  auto at = [&](int y, int x) -> int { return s[(size_t)y * w + x]; };
  parallel_rows(h, [=](int r0, int r1) {
    for (int y = r0; y < r1; y++)
      for (int x = 0; x < w; x++) {
        const int yc = std::min(std::max(y, 1), h - 2), xc = std::min(std::max(x, 1), w - 2);
        const int cross = (at(yc - 1, xc) + at(yc + 1, xc) + at(yc, xc - 1) + at(yc, xc + 1) + 2) >> 2;
        const int diag = (at(yc - 1, xc - 1) + at(yc - 1, xc + 1) + at(yc + 1, xc - 1) + at(yc + 1, xc + 1) + 2) >> 2;
        const int hor = (at(yc, xc - 1) + at(yc, xc + 1) + 1) >> 1, ver = (at(yc - 1, xc) + at(yc + 1, xc) + 1) >> 1;
        const int v = at(yc, xc);
        const bool ey = (yc & 1) == 0, ex = (xc & 1) == 0;
        int B, G, R;
        if (ey && ex) { B = v; G = cross; R = diag; }
        else if (!ey && !ex) { B = diag; G = cross; R = v; }
        else if (ey) { B = hor; G = v; R = ver; }
        else { B = ver; G = v; R = hor; }
        uint8_t* o = d + ((size_t)y * w + x) * 3;
        o[0] = (uint8_t)B;
        o[1] = (uint8_t)G;
        o[2] = (uint8_t)R;
      }
  });
}

// ---- integral(CV_32S) ----
extern "C" void orc_integral(const uint8_t* s, int w, int h, int32_t* sum) {
  const int sw = w + 1;
  for (int j = 0; j <= w; j++) sum[j] = 0;
  for (int i = 0; i < h; i++) {
    int32_t row = 0;
    sum[(size_t)(i + 1) * sw] = 0;
    for (int j = 0; j < w; j++) {
      row += s[(size_t)i * w + j];
      sum[(size_t)(i + 1) * sw + j + 1] = sum[(size_t)i * sw + j + 1] + row;
    }
  }
}

// ---- resize INTER_AREA for u8 (OpenCV imgproc/src/resize.cpp) ----
struct DecimateAlpha {
  int si, di;
  float alpha;
};
static int area_tab(int ssize, int dsize, int cn, double scale, std::vector<DecimateAlpha>& tab) {
  tab.clear();
  for (int dx = 0; dx < dsize; dx++) {
    double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    double cell = std::min(scale, ssize - fsx1);
    int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) tab.push_back({(sx1 - 1) * cn, dx * cn, (float)((sx1 - fsx1) / cell)});
    for (int sx = sx1; sx < sx2; sx++) tab.push_back({sx * cn, dx * cn, (float)(1.0 / cell)});
    if (fsx2 - sx2 > 1e-3)
      tab.push_back({sx2 * cn, dx * cn, (float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)});
  }
  return (int)tab.size();
}

extern "C" void orc_resize_area(const uint8_t* src, int sw, int sh, int cn, uint8_t* dst, int dw, int dh) {
  if (sw == dw && sh == dh) {
    std::memcpy(dst, src, (size_t)sw * sh * cn);
    return;
  }
  // cv::resize derives inv_scale = dsize/ssize and hal::resize inverts it again (scale = 1./inv_scale)
  double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
  int iscale_x = (int)std::nearbyint(scale_x) < 1 ? 1 : (int)std::nearbyint(scale_x);  // saturate_cast<int>
  int iscale_y = (int)std::nearbyint(scale_y) < 1 ? 1 : (int)std::nearbyint(scale_y);
  bool is_area_fast = std::abs(scale_x - iscale_x) < 2.220446049250313e-16 &&
                      std::abs(scale_y - iscale_y) < 2.220446049250313e-16;
  if (is_area_fast) {
    // ResizeAreaFast: integer block sums; 2x2 has the (sum+2)>>2 SIMD path, others sum*(1.f/area) rounded
    int area = iscale_x * iscale_y;
    float scale = 1.f / area;
    for (int dy = 0; dy < dh; dy++)
      for (int dx = 0; dx < dw; dx++)
        for (int c = 0; c < cn; c++) {
          int sum = 0;
          for (int ky = 0; ky < iscale_y; ky++)
            for (int kx = 0; kx < iscale_x; kx++) {
              int sy = dy * iscale_y + ky, sx = dx * iscale_x + kx;
              if (sy < sh && sx < sw) sum += src[((size_t)sy * sw + sx) * cn + c];
            }
          uint8_t v;
          if (iscale_x == 2 && iscale_y == 2) v = (uint8_t)((sum + 2) >> 2);
          else v = sat_u8(round_half_even((float)sum * scale));
          dst[((size_t)dy * dw + dx) * cn + c] = v;
        }
    return;
  }
  std::vector<DecimateAlpha> xtab, ytab;
  area_tab(sw, dw, cn, scale_x, xtab);
  area_tab(sh, dh, 1, scale_y, ytab);
  const int dwidth = dw * cn;
  std::vector<float> buf(dwidth), sum(dwidth, 0.f);
  int prev_dy = ytab[0].di;
  for (size_t j = 0; j < ytab.size(); j++) {
    float beta = ytab[j].alpha;
    int dy = ytab[j].di, sy = ytab[j].si;
    const uint8_t* S = src + (size_t)sy * sw * cn;
    std::fill(buf.begin(), buf.end(), 0.f);
    for (size_t k = 0; k < xtab.size(); k++) {
      int dxn = xtab[k].di;
      float alpha = xtab[k].alpha;
      for (int c = 0; c < cn; c++) buf[dxn + c] += S[xtab[k].si + c] * alpha;
    }
    if (dy != prev_dy) {
      uint8_t* Dp = dst + (size_t)prev_dy * dwidth;
      for (int dx = 0; dx < dwidth; dx++) {
        Dp[dx] = sat_u8(round_half_even(sum[dx]));
        sum[dx] = beta * buf[dx];
      }
      prev_dy = dy;
    } else {
      for (int dx = 0; dx < dwidth; dx++) sum[dx] += beta * buf[dx];
    }
  }
  uint8_t* Dp = dst + (size_t)prev_dy * dwidth;
  for (int dx = 0; dx < dwidth; dx++) Dp[dx] = sat_u8(round_half_even(sum[dx]));
}
