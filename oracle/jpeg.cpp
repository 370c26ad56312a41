// oracle/jpeg.cpp -- CPU ORACLE (test infrastructure only; see uvo_oracle.h).
// Baseline / extended-sequential Huffman JPEG decode as the reference obtains it: from_ros_to_cv_image
// (math_utility.cpp:154-173) -> cv_bridge::toCvCopy(CompressedImage) -> cv::imdecode(IMREAD_UNCHANGED), i.e.
// libjpeg-turbo with its defaults (JDCT_ISLOW, fancy upsampling, YCbCr -> BGR).  The algorithm lives in a third-party
// dependency that is absent from /root/reference (libjpeg-turbo inside OpenCV; the wheel here bundles 3.1.2); this file
// restates its published algorithm: ITU-T T.81 entropy decoding, the "islow" integer IDCT (jidctint: 13-bit constants,
// two passes), the triangle-filter chroma upsampling (jdsample: h2v1 / h2v2 / h1v2 "fancy") and the 16-bit fixed-point
// YCbCr -> RGB tables (jdcolor).  PINNED bit-exact against cv2.imdecode (tests/test_oracle_jpeg.py: live over
// sub-samplings, qualities, odd sizes, restart intervals, optimised tables, grayscale; fixture tests/golden/jpeg_*.npz).
// Not handled (returns a negative code): progressive / arithmetic / lossless / 12-bit streams, CMYK, Adobe RGB.
#include "uvo_oracle.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

const int ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                        41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                        30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
  bool present = false;
  int mincode[17], maxcode[18], valptr[17];
  uint8_t vals[256];
  void build(const uint8_t counts[16], const uint8_t* v, int nv) {
    std::memcpy(vals, v, nv);
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
      valptr[l] = k;
      mincode[l] = code;
      code += counts[l - 1];
      k += counts[l - 1];
      maxcode[l] = counts[l - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    present = true;
  }
};

struct Comp {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int bw = 0, bh = 0;  // blocks per row / column, padded to whole MCUs
  int dw = 0, dh = 0;  // downsampled_width / downsampled_height: the real sample counts
  int pred = 0;
  std::vector<int16_t> coef;   // bh * bw blocks of 64, natural order, not yet dequantised
  std::vector<uint8_t> plane;  // (bh*8) x (bw*8) samples after the IDCT
};

struct Bits {
  const uint8_t* p;
  const uint8_t* end;
  uint32_t acc = 0;
  int n = 0;
  bool hit_marker = false;
  int bit() {
    if (n == 0) {
      int b = 0;
      if (!hit_marker && p < end) {
        b = *p++;
        if (b == 0xFF) {
          if (p < end && *p == 0x00) {
            p++;
          } else {  // a marker: feed zeros from here on (libjpeg does the same past the end of a segment)
            p--;
            hit_marker = true;
            b = 0;
          }
        }
      }
      acc = (uint32_t)b;
      n = 8;
    }
    n--;
    return (acc >> n) & 1;
  }
  int get(int len) {
    int v = 0;
    for (int i = 0; i < len; i++) v = (v << 1) | bit();
    return v;
  }
  void reset() {
    n = 0;
    acc = 0;
    hit_marker = false;
  }
};

int decode_symbol(Bits& b, const Huff& h) {
  int code = 0;
  for (int l = 1; l <= 16; l++) {
    code = (code << 1) | b.bit();
    if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
  }
  return 0;  // corrupt stream: libjpeg warns and returns 0
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

bool decode_block(Bits& b, const Huff& dc, const Huff& ac, Comp& c, int16_t* blk) {
  int s = decode_symbol(b, dc);
  int diff = s ? extend(b.get(s), s) : 0;
  c.pred += diff;
  blk[0] = (int16_t)c.pred;
  for (int k = 1; k < 64;) {
    int rs = decode_symbol(b, ac);
    int r = rs >> 4;
    s = rs & 15;
    if (s == 0) {
      if (r != 15) break;  // EOB
      k += 16;
      continue;
    }
    k += r;
    if (k > 63) return false;
    blk[ZIGZAG[k]] = (int16_t)extend(b.get(s), s);
    k++;
  }
  return true;
}

// jidctint.c (jpeg_idct_islow): CONST_BITS 13, PASS1_BITS 2
inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
inline uint8_t range_limit(int v) {  // the post-IDCT part of libjpeg's sample_range_limit table, index masked to 10 bits
  const int idx = v & 1023;
  return (uint8_t)(idx < 128 ? idx + 128 : idx < 512 ? 255 : idx < 896 ? 0 : idx - 896);
}
void idct_islow(const int16_t* coef, const uint16_t* q, uint8_t* out, int stride) {
  const int F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633,
            F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
  int ws[64];
  for (int c = 0; c < 8; c++) {
    const int16_t* in = coef + c;
    const uint16_t* qq = q + c;
    int* w = ws + c;
    if (in[8] == 0 && in[16] == 0 && in[24] == 0 && in[32] == 0 && in[40] == 0 && in[48] == 0 && in[56] == 0) {
      const int dcval = (in[0] * qq[0]) * 4;  // << PASS1_BITS
      for (int r = 0; r < 8; r++) w[8 * r] = dcval;
      continue;
    }
    int z2 = in[16] * qq[16], z3 = in[48] * qq[48];
    int z1 = (z2 + z3) * F_0_541;
    int tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
    z2 = in[0] * qq[0];
    z3 = in[32] * qq[32];
    int tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[56] * qq[56];
    tmp1 = in[40] * qq[40];
    tmp2 = in[24] * qq[24];
    tmp3 = in[8] * qq[8];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * F_1_175;
    tmp0 *= F_0_298;
    tmp1 *= F_2_053;
    tmp2 *= F_3_072;
    tmp3 *= F_1_501;
    z1 *= -F_0_899;
    z2 *= -F_2_562;
    z3 *= -F_1_961;
    z4 *= -F_0_390;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    w[0] = descale(tmp10 + tmp3, 11);
    w[56] = descale(tmp10 - tmp3, 11);
    w[8] = descale(tmp11 + tmp2, 11);
    w[48] = descale(tmp11 - tmp2, 11);
    w[16] = descale(tmp12 + tmp1, 11);
    w[40] = descale(tmp12 - tmp1, 11);
    w[24] = descale(tmp13 + tmp0, 11);
    w[32] = descale(tmp13 - tmp0, 11);
  }
  for (int r = 0; r < 8; r++) {
    const int* w = ws + 8 * r;
    uint8_t* o = out + (size_t)r * stride;
    int z2 = w[2], z3 = w[6];
    int z1 = (z2 + z3) * F_0_541;
    int tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
    int tmp0 = (w[0] + w[4]) * 8192, tmp1 = (w[0] - w[4]) * 8192;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = w[7];
    tmp1 = w[5];
    tmp2 = w[3];
    tmp3 = w[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * F_1_175;
    tmp0 *= F_0_298;
    tmp1 *= F_2_053;
    tmp2 *= F_3_072;
    tmp3 *= F_1_501;
    z1 *= -F_0_899;
    z2 *= -F_2_562;
    z3 *= -F_1_961;
    z4 *= -F_0_390;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    o[0] = range_limit(descale(tmp10 + tmp3, 18));
    o[7] = range_limit(descale(tmp10 - tmp3, 18));
    o[1] = range_limit(descale(tmp11 + tmp2, 18));
    o[6] = range_limit(descale(tmp11 - tmp2, 18));
    o[2] = range_limit(descale(tmp12 + tmp1, 18));
    o[5] = range_limit(descale(tmp12 - tmp1, 18));
    o[3] = range_limit(descale(tmp13 + tmp0, 18));
    o[4] = range_limit(descale(tmp13 - tmp0, 18));
  }
}

// jdsample.c.  in: dh rows x dw samples (stride s).  out: full resolution, at least W x H written.
void upsample(const Comp& c, int hexp, int vexp, int W, int H, std::vector<uint8_t>& out) {
  const int s = c.bw * 8, dw = c.dw, dh = c.dh;
  const uint8_t* in = c.plane.data();
  const int ow = dw * hexp, oh = dh * vexp;  // >= W, H
  out.assign((size_t)ow * oh, 0);
  auto row = [&](int r) { return in + (size_t)(r < 0 ? 0 : r >= dh ? dh - 1 : r) * s; };
  const bool fancy = dw > 2;
  if (hexp == 1 && vexp == 1) {
    for (int y = 0; y < dh; y++) std::memcpy(&out[(size_t)y * ow], row(y), dw);
  } else if (hexp == 2 && vexp == 1 && fancy) {  // h2v1_fancy_upsample
    for (int y = 0; y < dh; y++) {
      const uint8_t* p = row(y);
      uint8_t* o = &out[(size_t)y * ow];
      o[0] = p[0];
      o[1] = (uint8_t)((p[0] * 3 + p[1] + 2) >> 2);
      for (int x = 1; x < dw - 1; x++) {
        const int v = p[x] * 3;
        o[2 * x] = (uint8_t)((v + p[x - 1] + 1) >> 2);
        o[2 * x + 1] = (uint8_t)((v + p[x + 1] + 2) >> 2);
      }
      o[2 * dw - 2] = (uint8_t)((p[dw - 1] * 3 + p[dw - 2] + 1) >> 2);
      o[2 * dw - 1] = p[dw - 1];
    }
  } else if (hexp == 2 && vexp == 2 && fancy) {  // h2v2_fancy_upsample
    for (int y = 0; y < dh; y++)
      for (int v = 0; v < 2; v++) {
        const uint8_t* p0 = row(y);
        const uint8_t* p1 = row(v == 0 ? y - 1 : y + 1);
        uint8_t* o = &out[(size_t)(2 * y + v) * ow];
        int thiscol = p0[0] * 3 + p1[0], nextcol = p0[1] * 3 + p1[1], lastcol;
        o[0] = (uint8_t)((thiscol * 4 + 8) >> 4);
        o[1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
        lastcol = thiscol;
        thiscol = nextcol;
        for (int x = 1; x < dw - 1; x++) {
          nextcol = p0[x + 1] * 3 + p1[x + 1];
          o[2 * x] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
          o[2 * x + 1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
          lastcol = thiscol;
          thiscol = nextcol;
        }
        o[2 * dw - 2] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
        o[2 * dw - 1] = (uint8_t)((thiscol * 4 + 7) >> 4);
      }
  } else if (hexp == 1 && vexp == 2) {  // h1v2_fancy_upsample
    for (int y = 0; y < dh; y++)
      for (int v = 0; v < 2; v++) {
        const uint8_t* p0 = row(y);
        const uint8_t* p1 = row(v == 0 ? y - 1 : y + 1);
        const int bias = v == 0 ? 1 : 2;
        uint8_t* o = &out[(size_t)(2 * y + v) * ow];
        for (int x = 0; x < dw; x++) o[x] = (uint8_t)((p0[x] * 3 + p1[x] + bias) >> 2);
      }
  } else {  // int_upsample / h2v1_upsample / h2v2_upsample: replication
    for (int y = 0; y < oh; y++) {
      const uint8_t* p = row(y / vexp);
      uint8_t* o = &out[(size_t)y * ow];
      for (int x = 0; x < ow; x++) o[x] = p[x / hexp];
    }
  }
  (void)W;
  (void)H;
}

struct Decoder {
  int W = 0, H = 0, nc = 0, hmax = 1, vmax = 1, restart = 0;
  bool have_sof = false;
  uint16_t qt[4][64] = {};
  bool qt_present[4] = {};
  Huff hdc[4], hac[4];
  Comp comp[4];
  int adobe_transform = -1;

  int parse_and_decode(const uint8_t* d, size_t len, bool header_only = false) {
    if (len < 4 || d[0] != 0xFF || d[1] != 0xD8) return -1;
    size_t i = 2;
    while (i + 4 <= len) {
      if (d[i] != 0xFF) {
        i++;
        continue;
      }
      const int m = d[i + 1];
      if (m == 0xFF) {
        i++;
        continue;
      }
      i += 2;
      if (m == 0xD9) break;
      if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
      if (i + 2 > len) return -1;
      const size_t L = ((size_t)d[i] << 8) | d[i + 1];
      if (L < 2 || i + L > len) return -1;
      const uint8_t* s = d + i + 2;
      const size_t n = L - 2;
      if (m == 0xDB) {  // DQT
        size_t k = 0;
        while (k < n) {
          const int pq = s[k] >> 4, tq = s[k] & 15;
          k++;
          if (tq > 3) return -1;
          for (int j = 0; j < 64; j++) {
            int v;
            if (pq) {
              v = (s[k] << 8) | s[k + 1];
              k += 2;
            } else {
              v = s[k++];
            }
            qt[tq][ZIGZAG[j]] = (uint16_t)v;
          }
          qt_present[tq] = true;
        }
      } else if (m == 0xC0 || m == 0xC1) {  // SOF0 / SOF1
        if (s[0] != 8) return -2;
        H = (s[1] << 8) | s[2];
        W = (s[3] << 8) | s[4];
        nc = s[5];
        if (W <= 0 || H <= 0 || (nc != 1 && nc != 3)) return -2;
        for (int c = 0; c < nc; c++) {
          comp[c].id = s[6 + 3 * c];
          comp[c].h = s[7 + 3 * c] >> 4;
          comp[c].v = s[7 + 3 * c] & 15;
          comp[c].tq = s[8 + 3 * c];
          if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3) return -1;
          hmax = std::max(hmax, comp[c].h);
          vmax = std::max(vmax, comp[c].v);
        }
        if (nc == 1) comp[0].h = comp[0].v = hmax = vmax = 1;  // a single component is never subsampled (T.81 A.2.2)
        const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
        for (int c = 0; c < nc; c++) {
          Comp& k = comp[c];
          if (hmax % k.h || vmax % k.v) return -2;
          k.bw = mcux * k.h;
          k.bh = mcuy * k.v;
          k.dw = (W * k.h + hmax - 1) / hmax;
          k.dh = (H * k.v + vmax - 1) / vmax;
          k.coef.assign((size_t)k.bw * k.bh * 64, 0);
        }
        have_sof = true;
      } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
        return -2;  // progressive, lossless, arithmetic ...
      } else if (m == 0xC4) {  // DHT
        size_t k = 0;
        while (k + 17 <= n) {
          const int tc = s[k] >> 4, th = s[k] & 15;
          if (th > 3 || tc > 1) return -1;
          const uint8_t* counts = s + k + 1;
          int nv = 0;
          for (int j = 0; j < 16; j++) nv += counts[j];
          if (nv > 256 || k + 17 + nv > n) return -1;
          (tc ? hac : hdc)[th].build(counts, s + k + 17, nv);
          k += 17 + nv;
        }
      } else if (m == 0xDD) {  // DRI
        restart = (s[0] << 8) | s[1];
      } else if (m == 0xEE && n >= 12 && std::memcmp(s, "Adobe", 5) == 0) {
        adobe_transform = s[11];
      } else if (m == 0xDA) {  // SOS
        if (!have_sof) return -1;
        if (header_only) return 0;
        const int ns = s[0];
        if (ns < 1 || ns > nc) return -1;
        int idx[4];
        for (int j = 0; j < ns; j++) {
          int c = -1;
          for (int q = 0; q < nc; q++)
            if (comp[q].id == s[1 + 2 * j]) c = q;
          if (c < 0) return -1;
          comp[c].td = s[2 + 2 * j] >> 4;
          comp[c].ta = s[2 + 2 * j] & 15;
          idx[j] = c;
        }
        const uint8_t* e = scan(d + i + L, d + len, idx, ns);
        if (!e) return -1;
        i = (size_t)(e - d);
        continue;
      }
      i += L;
    }
    return have_sof ? 0 : -1;
  }

  // entropy-coded segment; returns the position of the marker that ends it
  const uint8_t* scan(const uint8_t* p, const uint8_t* end, const int* idx, int ns) {
    Bits b{p, end};
    for (int j = 0; j < ns; j++) {
      comp[idx[j]].pred = 0;
      if (!hdc[comp[idx[j]].td].present || !hac[comp[idx[j]].ta].present || !qt_present[comp[idx[j]].tq]) return nullptr;
    }
    int mcus_x, mcus_y;
    if (ns == 1) {  // non-interleaved: one block per MCU, only the blocks that hold real samples
      const Comp& c = comp[idx[0]];
      mcus_x = (c.dw + 7) / 8;
      mcus_y = (c.dh + 7) / 8;
    } else {
      mcus_x = comp[idx[0]].bw / comp[idx[0]].h;
      mcus_y = comp[idx[0]].bh / comp[idx[0]].v;
    }
    int left = restart, rst = 0;
    for (int my = 0; my < mcus_y; my++)
      for (int mx = 0; mx < mcus_x; mx++) {
        if (restart && left == 0) {  // expect RSTn
          b.reset();
          const uint8_t* q = b.p;
          while (q + 1 < end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
          if (q + 1 >= end) return nullptr;
          (void)rst;
          b.p = q + 2;
          for (int j = 0; j < ns; j++) comp[idx[j]].pred = 0;
          left = restart;
        }
        for (int j = 0; j < ns; j++) {
          Comp& c = comp[idx[j]];
          const int bh_ = ns == 1 ? 1 : c.h, bv_ = ns == 1 ? 1 : c.v;
          for (int by = 0; by < bv_; by++)
            for (int bx = 0; bx < bh_; bx++) {
              const int X = mx * bh_ + bx, Y = my * bv_ + by;
              if (!decode_block(b, hdc[c.td], hac[c.ta], c, &c.coef[((size_t)Y * c.bw + X) * 64])) return nullptr;
            }
        }
        if (restart) left--;
      }
    // advance to the next marker
    const uint8_t* q = b.p;
    while (q + 1 < end && !(q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) q++;
    return q;
  }

  void reconstruct() {
    for (int c = 0; c < nc; c++) {
      Comp& k = comp[c];
      const int stride = k.bw * 8;
      k.plane.assign((size_t)stride * k.bh * 8, 0);
      for (int by = 0; by < k.bh; by++)
        for (int bx = 0; bx < k.bw; bx++)
          idct_islow(&k.coef[((size_t)by * k.bw + bx) * 64], qt[k.tq], &k.plane[(size_t)by * 8 * stride + bx * 8], stride);
    }
  }
};

}  // namespace

extern "C" int orc_jpeg_info(const uint8_t* data, size_t len, int* w, int* h, int* channels) {
  Decoder D;
  const int rc = D.parse_and_decode(data, len, /*header_only=*/true);
  if (rc < 0) return rc;
  *w = D.W;
  *h = D.H;
  *channels = D.nc;
  return 0;
}

// parity tap: the quantised coefficients after entropy decoding, component after component, each blocks_y x blocks_x
// x 64 in natural order with the block grid padded to whole MCUs.  Returns the count (negative: error / too small).
extern "C" long orc_jpeg_coefficients(const uint8_t* data, size_t len, int16_t* out, size_t capacity) {
  Decoder D;
  const int rc = D.parse_and_decode(data, len);
  if (rc < 0) return rc;
  size_t total = 0;
  for (int c = 0; c < D.nc; c++) total += D.comp[c].coef.size();
  if (total > capacity) return -3;
  size_t off = 0;
  for (int c = 0; c < D.nc; c++) {
    std::memcpy(out + off, D.comp[c].coef.data(), sizeof(int16_t) * D.comp[c].coef.size());
    off += D.comp[c].coef.size();
  }
  return (long)total;
}

// out: h x w (1 component) or h x w x 3 BGR, as cv::imdecode(IMREAD_UNCHANGED) returns it
extern "C" int orc_jpeg_decode(const uint8_t* data, size_t len, uint8_t* out) {
  Decoder D;
  const int rc = D.parse_and_decode(data, len);
  if (rc < 0) return rc;
  if (D.nc == 3 && D.adobe_transform == 0) return -2;  // Adobe RGB: not a case the camera path produces
  D.reconstruct();
  const int W = D.W, H = D.H;
  if (D.nc == 1) {
    const Comp& y = D.comp[0];
    for (int r = 0; r < H; r++) std::memcpy(out + (size_t)r * W, &y.plane[(size_t)r * y.bw * 8], W);
    return 0;
  }
  std::vector<uint8_t> full[3];
  int fw[3];
  for (int c = 0; c < 3; c++) {
    const Comp& k = D.comp[c];
    upsample(k, D.hmax / k.h, D.vmax / k.v, W, H, full[c]);
    fw[c] = k.dw * (D.hmax / k.h);
  }
  // jdcolor.c build_ycc_rgb_table / ycc_rgb_convert, SCALEBITS 16
  int cr_r[256], cb_b[256], cr_g[256], cb_g[256];
  for (int i = 0; i < 256; i++) {
    const int x = i - 128;
    auto FIX = [](double v) { return (int)(v * 65536.0 + 0.5); };
    cr_r[i] = (FIX(1.40200) * x + 32768) >> 16;
    cb_b[i] = (FIX(1.77200) * x + 32768) >> 16;
    cr_g[i] = -FIX(0.71414) * x;
    cb_g[i] = -FIX(0.34414) * x + 32768;  // + ONE_HALF
  }
  auto clamp = [](int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); };
  for (int r = 0; r < H; r++) {
    const uint8_t* Y = &full[0][(size_t)r * fw[0]];
    const uint8_t* Cb = &full[1][(size_t)r * fw[1]];
    const uint8_t* Cr = &full[2][(size_t)r * fw[2]];
    uint8_t* o = out + (size_t)r * W * 3;
    for (int x = 0; x < W; x++) {
      const int y = Y[x], cb = Cb[x], cr = Cr[x];
      o[3 * x + 2] = clamp(y + cr_r[cr]);
      o[3 * x + 1] = clamp(y + ((cb_g[cb] + cr_g[cr]) >> 16));
      o[3 * x + 0] = clamp(y + cb_b[cb]);
    }
  }
  return 0;
}
