// oracle/surf.cpp -- CPU ORACLE (test infrastructure only; see uvo_oracle.h).
// Restates cv::xfeatures2d::SURF::detectAndCompute as called by detect_features (reference VO_utility.cpp:114-119:
// SURF::create(SURF_MIN_HESSIAN, SURF_OCTAVES_NUMBER, SURF_OCTAVES_LAYERS, SURF_EXTENDED, SURF_UPRIGHT)).
// Source of the algorithm: opencv_contrib 4.5 modules/xfeatures2d/src/surf.cpp -- NOT present in this image and not
// vendored by the reference.  The restatement follows SURVEY.md Appendix A (A.1-A.4).
//
//   *** PARITY UNPINNED against real OpenCV for this file ***  (no contrib wheel, no network).  Sub-steps that
//   exist in the importable cv2 4.13 are pinned by tests: integral, getGaussianKernel, resize INTER_AREA,
//   fastAtan2/phase.
//
// Unwritten borders of the det/trace maps are uninitialised memory in OpenCV and never meant to be read; here they
// are zero (documented divergence: cannot matter unless OpenCV itself reads uninitialised memory).
#include "uvo_oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace {

inline int cv_round(double v) { return (int)std::nearbyint(v); }
inline int cv_roundf(float v) { return (int)std::nearbyintf(v); }

const int HAAR_SIZE0 = 9, HAAR_SIZE_INC = 6, SAMPLE_STEP0 = 1;
const int ORI_RADIUS = 6, ORI_WIN = 60, PATCH_SZ = 20, ORI_SEARCH_INC = 5;
const float ORI_SIGMA = 2.5f, DESC_SIGMA = 3.3f;

struct SurfHF {
  int p0, p1, p2, p3;
  float w;
};

// A.2 resizeHaarPattern
void resize_haar(const int src[][5], SurfHF* dst, int n, int old_size, int new_size, int width_step) {
  float ratio = (float)new_size / old_size;
  for (int k = 0; k < n; k++) {
    int dx1 = cv_roundf(ratio * src[k][0]);
    int dy1 = cv_roundf(ratio * src[k][1]);
    int dx2 = cv_roundf(ratio * src[k][2]);
    int dy2 = cv_roundf(ratio * src[k][3]);
    dst[k].p0 = dy1 * width_step + dx1;
    dst[k].p1 = dy2 * width_step + dx1;
    dst[k].p2 = dy1 * width_step + dx2;
    dst[k].p3 = dy2 * width_step + dx2;
    dst[k].w = src[k][4] / ((float)(dx2 - dx1) * (dy2 - dy1));
  }
}

// A.2 calcHaarPattern: int box sum times f32 weight (f32 product), accumulated in f64, returned as f32
inline float calc_haar(const int* origin, const SurfHF* f, int n) {
  double d = 0;
  for (int k = 0; k < n; k++) d += (origin[f[k].p0] + origin[f[k].p3] - origin[f[k].p1] - origin[f[k].p2]) * f[k].w;
  return (float)d;
}

const int dx_s[3][5] = {{0, 2, 3, 7, 1}, {3, 2, 6, 7, -2}, {6, 2, 9, 7, 1}};
const int dy_s[3][5] = {{2, 0, 7, 3, 1}, {2, 3, 7, 6, -2}, {2, 6, 7, 9, 1}};
const int dxy_s[4][5] = {{1, 1, 4, 4, 1}, {5, 1, 8, 4, -1}, {1, 5, 4, 8, -1}, {5, 5, 8, 8, 1}};

// sum: (h+1)x(w+1); det/trace: (h/step)x(w/step), pre-zeroed.  Sample rows [row0, row1) of the layer (clamped to the
// layer's sample rows): every sample is independent, so row ranges can run on different threads.
void calc_layer_det_trace(const int32_t* sum, int w, int h, int size, int step, float* det, float* trace,
                          int row0 = 0, int row1 = 1 << 30) {
  const int srows = h + 1, scols = w + 1;
  if (size > srows - 1 || size > scols - 1) return;
  SurfHF Dx[3], Dy[3], Dxy[4];
  resize_haar(dx_s, Dx, 3, 9, size, scols);
  resize_haar(dy_s, Dy, 3, 9, size, scols);
  resize_haar(dxy_s, Dxy, 4, 9, size, scols);
  const int samples_i = 1 + (srows - 1 - size) / step, samples_j = 1 + (scols - 1 - size) / step;
  const int margin = (size / 2) / step;
  const int lcols = w / step;
  for (int i = std::max(row0, 0); i < std::min(samples_i, row1); i++) {
    const int* sum_ptr = sum + (size_t)(i * step) * scols;
    float* det_ptr = det + (size_t)(i + margin) * lcols + margin;
    float* tr_ptr = trace + (size_t)(i + margin) * lcols + margin;
    for (int j = 0; j < samples_j; j++) {
      float dx = calc_haar(sum_ptr, Dx, 3);
      float dy = calc_haar(sum_ptr, Dy, 3);
      float dxy = calc_haar(sum_ptr, Dxy, 4);
      sum_ptr += step;
      det_ptr[j] = dx * dy - 0.81f * dxy * dxy;
      tr_ptr[j] = dx + dy;
    }
  }
}

// A.3 interpolateKeypoint: Matx33f::solve(DECOMP_LU) == Cramer's rule in f32
bool interpolate_keypoint(const float N9[3][9], int dx, int dy, int ds, orc_keypoint& kpt) {
  float b0 = -(N9[1][5] - N9[1][3]) / 2, b1 = -(N9[1][7] - N9[1][1]) / 2, b2 = -(N9[2][4] - N9[0][4]) / 2;
  float a00 = N9[1][3] - 2 * N9[1][4] + N9[1][5];
  float a01 = (N9[1][8] - N9[1][6] - N9[1][2] + N9[1][0]) / 4;
  float a02 = (N9[2][5] - N9[2][3] - N9[0][5] + N9[0][3]) / 4;
  float a10 = a01;
  float a11 = N9[1][1] - 2 * N9[1][4] + N9[1][7];
  float a12 = (N9[2][7] - N9[2][1] - N9[0][7] + N9[0][1]) / 4;
  float a20 = a02, a21 = a12;
  float a22 = N9[0][4] - 2 * N9[1][4] + N9[2][4];
  float d = a00 * (a11 * a22 - a21 * a12) - a01 * (a10 * a22 - a20 * a12) + a02 * (a10 * a21 - a20 * a11);
  float x0 = 0, x1 = 0, x2 = 0;
  if (d != 0) {
    d = 1 / d;
    x0 = d * (b0 * (a11 * a22 - a12 * a21) - a01 * (b1 * a22 - a12 * b2) + a02 * (b1 * a21 - a11 * b2));
    x1 = d * (a00 * (b1 * a22 - a12 * b2) - b0 * (a10 * a22 - a12 * a20) + a02 * (a10 * b2 - b1 * a20));
    x2 = d * (a00 * (a11 * b2 - b1 * a21) - a01 * (a10 * b2 - b1 * a20) + b0 * (a10 * a21 - a11 * a20));
  }
  bool ok = (x0 != 0 || x1 != 0 || x2 != 0) && std::abs(x0) <= 1 && std::abs(x1) <= 1 && std::abs(x2) <= 1;
  if (ok) {
    kpt.x += x0 * dx;
    kpt.y += x1 * dy;
    kpt.size = (float)cv_roundf(kpt.size + x2 * ds);
  }
  return ok;
}

// a layer map: (h/step) x (w/step) floats, zero-filled by the band workers before any sample is written
struct LayerMap {
  std::unique_ptr<float[]> p;
  size_t n = 0;
  float* data() const { return p.get(); }
};

void find_maxima_in_layer(int w, int h, const std::vector<LayerMap>& dets,
                          const std::vector<LayerMap>& traces, const std::vector<int>& sizes,
                          std::vector<orc_keypoint>& kps, int octave, int layer, float thr, int step) {
  const int size = sizes[layer];
  const int lrows = h / step, lcols = w / step;
  const int margin = (sizes[layer + 1] / 2) / step + 1;
  const int st = lcols;
  for (int i = margin; i < lrows - margin; i++) {
    const float* det_ptr = dets[layer].data() + (size_t)i * st;
    const float* tr_ptr = traces[layer].data() + (size_t)i * st;
    for (int j = margin; j < lcols - margin; j++) {
      float val0 = det_ptr[j];
      if (val0 > thr) {
        int sum_i = step * (i - (size / 2) / step);
        int sum_j = step * (j - (size / 2) / step);
        const float* det1 = dets[layer - 1].data() + (size_t)i * st + j;
        const float* det2 = dets[layer].data() + (size_t)i * st + j;
        const float* det3 = dets[layer + 1].data() + (size_t)i * st + j;
        float N9[3][9] = {{det1[-st - 1], det1[-st], det1[-st + 1], det1[-1], det1[0], det1[1], det1[st - 1],
                           det1[st], det1[st + 1]},
                          {det2[-st - 1], det2[-st], det2[-st + 1], det2[-1], det2[0], det2[1], det2[st - 1],
                           det2[st], det2[st + 1]},
                          {det3[-st - 1], det3[-st], det3[-st + 1], det3[-1], det3[0], det3[1], det3[st - 1],
                           det3[st], det3[st + 1]}};
        bool is_max = true;
        for (int a = 0; a < 3 && is_max; a++)
          for (int b = 0; b < 9; b++) {
            if (a == 1 && b == 4) continue;
            if (!(val0 > N9[a][b])) {
              is_max = false;
              break;
            }
          }
        if (is_max) {
          float center_i = sum_i + (size - 1) * 0.5f;
          float center_j = sum_j + (size - 1) * 0.5f;
          orc_keypoint kpt;
          kpt.x = center_j;
          kpt.y = center_i;
          kpt.size = (float)sizes[layer];
          kpt.angle = -1;
          kpt.response = val0;
          kpt.octave = octave;
          kpt.class_id = (tr_ptr[j] > 0) - (tr_ptr[j] < 0);
          int ds = size - sizes[layer - 1];
          if (interpolate_keypoint(N9, step, step, ds, kpt)) kps.push_back(kpt);
        }
      }
    }
  }
}

// A.3 KeypointGreater
bool keypoint_greater(const orc_keypoint& a, const orc_keypoint& b) {
  if (a.response > b.response) return true;
  if (a.response < b.response) return false;
  if (a.size > b.size) return true;
  if (a.size < b.size) return false;
  if (a.octave > b.octave) return true;
  if (a.octave < b.octave) return false;
  if (a.y > b.y) return true;
  if (a.y < b.y) return false;
  return a.x < b.x;
}

// cv::getGaussianKernel(n, sigma, CV_32F), sigma > 0.  Follows OpenCV's bit-exact construction (imgproc
// smooth.dispatch.cpp getGaussianKernelBitExact): only the first (n-1)/2 taps are evaluated, at x = 1-n+2i with
// exponent x*x*(-0.125/sigma^2); the centre tap -- and for even n BOTH centre taps -- are taken as 1 (an OpenCV
// quirk that matters for the 20-tap descriptor window).  Pinned bit-equal (as f32) to cv2 4.13 by the tests.
void gaussian_kernel(int n, double sigma, float* out) {
  const double scale2x = -0.125 / (sigma * sigma);
  const int n2 = (n - 1) / 2;
  std::vector<double> v(n2 + 1);
  double sum = 0;
  for (int i = 0, x = 1 - n; i < n2; i++, x += 2) {
    v[i] = std::exp((double)(x * x) * scale2x);
    sum += v[i];
  }
  sum *= 2.0;
  sum += 1.0;
  if ((n & 1) == 0) sum += 1.0;
  const double mul1 = 1.0 / sum;
  for (int i = 0; i < n2; i++) {
    out[i] = (float)(v[i] * mul1);
    out[n - 1 - i] = out[i];
  }
  out[n2] = (float)mul1;
  if ((n & 1) == 0) out[n2 + 1] = (float)mul1;
}

// cv::fastAtan2 (degrees), also the kernel of cv::phase(..., angleInDegrees=true)
float fast_atan2(float y, float x) {
  const float p1 = 0.9997878412794807f * (float)(180 / M_PI), p3 = -0.3258083974640975f * (float)(180 / M_PI),
              p5 = 0.1555786518463281f * (float)(180 / M_PI), p7 = -0.04432655554792128f * (float)(180 / M_PI);
  float ax = std::abs(x), ay = std::abs(y), a, c, c2;
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

struct DescTables {
  int n_ori;
  int apt_x[169], apt_y[169];
  float aptw[169];
  float DW[PATCH_SZ * PATCH_SZ];
  DescTables() {
    float G_ori[2 * ORI_RADIUS + 1];
    gaussian_kernel(2 * ORI_RADIUS + 1, ORI_SIGMA, G_ori);
    n_ori = 0;
    for (int i = -ORI_RADIUS; i <= ORI_RADIUS; i++)
      for (int j = -ORI_RADIUS; j <= ORI_RADIUS; j++)
        if (i * i + j * j <= ORI_RADIUS * ORI_RADIUS) {
          apt_x[n_ori] = i;  // Point(i, j): x = i, y = j
          apt_y[n_ori] = j;
          aptw[n_ori++] = G_ori[i + ORI_RADIUS] * G_ori[j + ORI_RADIUS];
        }
    float G_desc[PATCH_SZ];
    gaussian_kernel(PATCH_SZ, DESC_SIGMA, G_desc);
    for (int i = 0; i < PATCH_SZ; i++)
      for (int j = 0; j < PATCH_SZ; j++) DW[i * PATCH_SZ + j] = G_desc[i] * G_desc[j];
  }
};

// window extraction (A.4 "Window"); returns win_size, fills win (win_size*win_size)
int extract_window(const uint8_t* img, int w, int h, float cx, float cy, float size, float dir_deg, bool upright,
                   std::vector<uint8_t>& win) {
  float s = size * 1.2f / 9.0f;
  int win_size = (int)((PATCH_SZ + 1) * s);
  win.resize((size_t)std::max(win_size, 1) * std::max(win_size, 1));
  uint8_t* WIN = win.data();
  if (!upright) {
    float descriptor_dir = dir_deg * (float)(M_PI / 180);
    float sin_dir = -std::sin(descriptor_dir), cos_dir = std::cos(descriptor_dir);
    float win_offset = -(float)(win_size - 1) / 2;
    float start_x = cx + win_offset * cos_dir + win_offset * sin_dir;
    float start_y = cy - win_offset * sin_dir + win_offset * cos_dir;
    int ncols1 = w - 1, nrows1 = h - 1;
    for (int i = 0; i < win_size; i++, start_x += sin_dir, start_y += cos_dir) {
      double pixel_x = start_x, pixel_y = start_y;
      for (int j = 0; j < win_size; j++, pixel_x += cos_dir, pixel_y -= sin_dir) {
        int ix = (int)std::floor(pixel_x), iy = (int)std::floor(pixel_y);
        if ((unsigned)ix < (unsigned)ncols1 && (unsigned)iy < (unsigned)nrows1) {
          float a = (float)(pixel_x - ix), b = (float)(pixel_y - iy);
          const uint8_t* p = img + (size_t)iy * w + ix;
          WIN[i * win_size + j] = (uint8_t)cv_roundf(p[0] * (1.f - a) * (1.f - b) + p[1] * a * (1.f - b) +
                                                     p[w] * (1.f - a) * b + p[w + 1] * a * b);
        } else {
          int x = std::min(std::max(cv_round(pixel_x), 0), ncols1);
          int y = std::min(std::max(cv_round(pixel_y), 0), nrows1);
          WIN[i * win_size + j] = img[(size_t)y * w + x];
        }
      }
    }
  } else {
    float win_offset = -(float)(win_size - 1) / 2;
    int start_x = cv_roundf(cx + win_offset);
    int start_y = cv_roundf(cy - win_offset);
    for (int i = 0; i < win_size; i++, start_x++) {
      int pixel_x = start_x, pixel_y = start_y;
      for (int j = 0; j < win_size; j++, pixel_y--) {
        int x = std::min(std::max(pixel_x, 0), w - 1);
        int y = std::min(std::max(pixel_y, 0), h - 1);
        WIN[i * win_size + j] = img[(size_t)y * w + x];
      }
    }
  }
  return win_size;
}

const int ori_dx_s[2][5] = {{0, 0, 2, 4, -1}, {2, 0, 4, 4, 1}};
const int ori_dy_s[2][5] = {{0, 0, 4, 2, 1}, {0, 2, 4, 4, -1}};

// A.4 SURFInvoker body for one keypoint.  Returns false when the keypoint is marked for deletion (size = -1).
bool describe_keypoint(const uint8_t* img, const int32_t* sum, int w, int h, const DescTables& T, bool extended,
                       bool upright, orc_keypoint& kp, float* vec) {
  const int srows = h + 1, scols = w + 1;
  const float size = kp.size;
  const float cx = kp.x, cy = kp.y;
  const float s = size * 1.2f / 9.0f;
  const int grad_wav_size = 2 * cv_roundf(2 * s);
  if (srows < grad_wav_size || scols < grad_wav_size) {
    kp.size = -1;
    return false;
  }
  float descriptor_dir = 360.f - 90.f;
  if (!upright) {
    SurfHF dx_t[2], dy_t[2];
    resize_haar(ori_dx_s, dx_t, 2, 4, grad_wav_size, scols);
    resize_haar(ori_dy_s, dy_t, 2, 4, grad_wav_size, scols);
    float X[169], Y[169], angle[169];
    int nangle = 0;
    for (int kk = 0; kk < T.n_ori; kk++) {
      int x = cv_roundf(cx + T.apt_x[kk] * s - (float)(grad_wav_size - 1) / 2);
      int y = cv_roundf(cy + T.apt_y[kk] * s - (float)(grad_wav_size - 1) / 2);
      if (y < 0 || y >= srows - grad_wav_size || x < 0 || x >= scols - grad_wav_size) continue;
      const int* ptr = sum + (size_t)y * scols + x;
      float vx = calc_haar(ptr, dx_t, 2), vy = calc_haar(ptr, dy_t, 2);
      X[nangle] = vx * T.aptw[kk];
      Y[nangle] = vy * T.aptw[kk];
      nangle++;
    }
    if (nangle == 0) {
      kp.size = -1;
      return false;
    }
    for (int j = 0; j < nangle; j++) angle[j] = fast_atan2(Y[j], X[j]);  // cv::phase(X, Y, angle, true)
    float bestx = 0, besty = 0, descriptor_mod = 0;
    for (int i = 0; i < 360; i += ORI_SEARCH_INC) {
      float sumx = 0, sumy = 0, temp_mod;
      for (int j = 0; j < nangle; j++) {
        int d = std::abs(cv_roundf(angle[j]) - i);
        if (d < ORI_WIN / 2 || d > 360 - ORI_WIN / 2) {
          sumx += X[j];
          sumy += Y[j];
        }
      }
      temp_mod = sumx * sumx + sumy * sumy;
      if (temp_mod > descriptor_mod) {
        descriptor_mod = temp_mod;
        bestx = sumx;
        besty = sumy;
      }
    }
    descriptor_dir = fast_atan2(-besty, bestx);
  }
  kp.angle = descriptor_dir;
  if (!vec) return true;

  std::vector<uint8_t> win;
  int win_size = extract_window(img, w, h, cx, cy, size, descriptor_dir, upright, win);
  uint8_t PATCH[PATCH_SZ + 1][PATCH_SZ + 1];
  orc_resize_area(win.data(), win_size, win_size, 1, &PATCH[0][0], PATCH_SZ + 1, PATCH_SZ + 1);

  float DX[PATCH_SZ][PATCH_SZ], DY[PATCH_SZ][PATCH_SZ];
  for (int i = 0; i < PATCH_SZ; i++)
    for (int j = 0; j < PATCH_SZ; j++) {
      float dw = T.DW[i * PATCH_SZ + j];
      float vx = (PATCH[i][j + 1] - PATCH[i][j] + PATCH[i + 1][j + 1] - PATCH[i + 1][j]) * dw;
      float vy = (PATCH[i + 1][j] - PATCH[i][j] + PATCH[i + 1][j + 1] - PATCH[i][j + 1]) * dw;
      DX[i][j] = vx;
      DY[i][j] = vy;
    }
  const int dsize = extended ? 128 : 64;
  for (int kk = 0; kk < dsize; kk++) vec[kk] = 0;
  double square_mag = 0;
  float* v = vec;
  if (extended) {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        for (int y = i * 5; y < i * 5 + 5; y++)
          for (int x = j * 5; x < j * 5 + 5; x++) {
            float tx = DX[y][x], ty = DY[y][x];
            if (ty >= 0) {
              v[0] += tx;
              v[1] += (float)fabs(tx);
            } else {
              v[2] += tx;
              v[3] += (float)fabs(tx);
            }
            if (tx >= 0) {
              v[4] += ty;
              v[5] += (float)fabs(ty);
            } else {
              v[6] += ty;
              v[7] += (float)fabs(ty);
            }
          }
        for (int kk = 0; kk < 8; kk++) square_mag += v[kk] * v[kk];
        v += 8;
      }
  } else {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        for (int y = i * 5; y < i * 5 + 5; y++)
          for (int x = j * 5; x < j * 5 + 5; x++) {
            float tx = DX[y][x], ty = DY[y][x];
            v[0] += tx;
            v[1] += ty;
            v[2] += (float)fabs(tx);
            v[3] += (float)fabs(ty);
          }
        for (int kk = 0; kk < 4; kk++) square_mag += v[kk] * v[kk];
        v += 4;
      }
  }
  float scale = (float)(1. / (std::sqrt(square_mag) + FLT_EPSILON));
  for (int kk = 0; kk < dsize; kk++) vec[kk] *= scale;
  return true;
}

}  // namespace

extern "C" void orc_surf_det_trace_layer(const int32_t* sum, int w, int h, int size, int step, float* det,
                                         float* trace) {
  calc_layer_det_trace(sum, w, h, size, step, det, trace);
}

extern "C" void orc_gaussian_kernel_f32(int n, double sigma, float* out) { gaussian_kernel(n, sigma, out); }
extern "C" float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }

extern "C" void orc_surf_patch(const uint8_t* img, int w, int h, float cx, float cy, float size, float angle_deg,
                               int upright, uint8_t* patch, int* win_size_out) {
  std::vector<uint8_t> win;
  int ws = extract_window(img, w, h, cx, cy, size, angle_deg, upright != 0, win);
  orc_resize_area(win.data(), ws, ws, 1, patch, PATCH_SZ + 1, PATCH_SZ + 1);
  if (win_size_out) *win_size_out = ws;
}

extern "C" int orc_surf_detect_and_compute(const uint8_t* img, int w, int h, double hessian_threshold, int n_octaves,
                                           int n_layers, int extended, int upright, orc_keypoint* out_kps,
                                           float* out_desc, int capacity) {
  std::vector<int32_t> sum((size_t)(w + 1) * (h + 1));
  orc_integral(img, w, h, sum.data());
  const int n_total = (n_layers + 2) * n_octaves, n_middle = n_layers * n_octaves;
  std::vector<LayerMap> dets(n_total), traces(n_total);
  std::vector<int> sizes(n_total), steps(n_total), middle(n_middle);
  int index = 0, mi = 0, step = SAMPLE_STEP0;
  for (int o = 0; o < n_octaves; o++) {
    for (int l = 0; l < n_layers + 2; l++) {
      dets[index].n = traces[index].n = (size_t)(h / step) * (w / step);
      dets[index].p.reset(new float[dets[index].n]);  // uninitialised: zero-filled in parallel below
      traces[index].p.reset(new float[traces[index].n]);
      sizes[index] = (HAAR_SIZE0 + HAAR_SIZE_INC * l) << o;
      steps[index] = step;
      if (0 < l && l <= n_layers) middle[mi++] = index;
      index++;
    }
    step *= 2;
  }
  const unsigned n_thr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  {
    // OpenCV runs one parallel_for_ item per layer (SURFBuildInvoker); the five octave-0 layers hold 3/4 of the
    // samples, so the items here are 64-row bands of a layer, handed out from a shared counter, octave 0 first
    struct Band { int layer, row0; };
    std::vector<Band> bands;
    for (int i = 0; i < n_total; i++)
      for (int r = 0; r < h / steps[i] + 1; r += 64) bands.push_back({i, r});
    {  // phase 1: zero the maps (the unwritten borders read as 0), same bands
      std::atomic<size_t> nz{0};
      std::vector<std::thread> tz;
      for (unsigned t = 0; t < n_thr; t++)
        tz.emplace_back([&] {
          for (size_t b = nz++; b < bands.size(); b = nz++) {
            const int i = bands[b].layer, lcols = w / steps[i], lrows = h / steps[i];
            const int r0 = std::min(bands[b].row0, lrows), r1 = std::min(bands[b].row0 + 64, lrows);
            std::memset(dets[i].data() + (size_t)r0 * lcols, 0, sizeof(float) * (size_t)(r1 - r0) * lcols);
            std::memset(traces[i].data() + (size_t)r0 * lcols, 0, sizeof(float) * (size_t)(r1 - r0) * lcols);
          }
        });
      for (auto& t : tz) t.join();
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < n_thr; t++)
      th.emplace_back([&] {
        for (size_t b = next++; b < bands.size(); b = next++) {
          const int i = bands[b].layer;
          calc_layer_det_trace(sum.data(), w, h, sizes[i], steps[i], dets[i].data(), traces[i].data(), bands[b].row0,
                               bands[b].row0 + 64);
        }
      });
    for (auto& t : th) t.join();
  }
  // one item per middle layer, as OpenCV's SURFFindInvoker; the per-layer lists are concatenated in layer order, which
  // is the order the sequential loop produces (the sort below is a total order anyway)
  std::vector<orc_keypoint> kps;
  const float thr = (float)hessian_threshold;
  {
    std::vector<std::vector<orc_keypoint>> per_layer(n_middle);
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < std::min<unsigned>(n_thr, n_middle); t++)
      th.emplace_back([&] {
        for (int i = next++; i < n_middle; i = next++)
          find_maxima_in_layer(w, h, dets, traces, sizes, per_layer[i], i / n_layers, middle[i], thr, steps[middle[i]]);
      });
    for (auto& t : th) t.join();
    for (auto& v : per_layer) kps.insert(kps.end(), v.begin(), v.end());
  }
  std::sort(kps.begin(), kps.end(), keypoint_greater);
  for (auto& k : kps) k.class_id = -1;  // detectAndCompute resets class_id (A.1)
  const int N = (int)kps.size();
  const int dsize = extended ? 128 : 64;
  std::vector<float> desc((size_t)N * dsize);
  static const DescTables T;
  {
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
      th.emplace_back([&, t] {
        for (int k = t; k < N; k += nt)
          describe_keypoint(img, sum.data(), w, h, T, extended != 0, upright != 0, kps[k], desc.data() + (size_t)k * dsize);
      });
    for (auto& t : th) t.join();
  }
  // remove keypoints marked for deletion, preserving order
  int n_out = 0;
  for (int k = 0; k < N; k++)
    if (kps[k].size > 0) n_out++;
  if (n_out > capacity) return -n_out;
  int j = 0;
  for (int k = 0; k < N; k++)
    if (kps[k].size > 0) {
      out_kps[j] = kps[k];
      if (out_desc) std::memcpy(out_desc + (size_t)j * dsize, desc.data() + (size_t)k * dsize, sizeof(float) * dsize);
      j++;
    }
  return n_out;
}
