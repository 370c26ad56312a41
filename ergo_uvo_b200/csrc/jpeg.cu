// jpeg.cu -- ingest: baseline JPEG decode, the cv::imdecode inside from_ros_to_cv_image (reference
// math_utility.cpp:154-173: cv_bridge::toCvCopy(CompressedImage) -> cv::imdecode(IMREAD_UNCHANGED) -> libjpeg-turbo with
// JDCT_ISLOW, fancy upsampling, YCbCr -> BGR).  SURVEY.md 8f-2.
//
// Split (the "hybrid" arrangement): the entropy-coded segment is a serial bit stream, so Huffman decoding runs on the
// host (uvo_jpeg_entropy_decode: table-driven, 64-bit bit buffer) into planes of quantised coefficients; everything
// after it is independent per block / per pixel and runs on the GPU:
//   k_jpeg_idct   dequantisation + the 8x8 "islow" integer IDCT (13-bit constants, two passes, descale 11 / 18 bits):
//                 eight threads per block, columns then rows through shared memory, one 8-byte store per sample row
//   k_jpeg_color  chroma upsampling with libjpeg's triangle filters (h2v1, h2v2, h1v2; replication elsewhere and for
//                 planes narrower than three samples) evaluated per output pixel from the component planes, then the
//                 16-bit fixed-point YCbCr -> BGR conversion; interleaved u8 output
// Integer arithmetic throughout, the same operations in the same order as libjpeg-turbo: results are bit-identical.
// Not handled: progressive / arithmetic / lossless / 12-bit streams, CMYK, Adobe RGB -> UVO_ERR_UNSUPPORTED.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "capi_internal.cuh"
#include "jpeg.cuh"
#include "jpeg_huff.cuh"
#include "jpeg_kernels.cuh"

using namespace uvo;

namespace {
constexpr int64_t JPEG_MAX_PIXELS = (int64_t)1 << 28;  // 16384 x 16384; the cameras of the reference are 5 MP

// ------------------------------------------------------------------------------------------------ host: parsing
const uint8_t kNatural[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                              41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                              30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int FAST_BITS = 10;

struct HuffTab {
  bool present = false;
  uint16_t fast[1 << FAST_BITS];  // (code length << 8) | symbol for codes of up to FAST_BITS bits, 0 = longer
  int32_t maxcode[18];            // largest code of each length, -1 if none
  int32_t valoff[17];             // huffval index of the first code of a length minus that code
  uint8_t huffval[256];
  // AC tables only: where code + magnitude bits fit in FAST_BITS, the whole coefficient in one look-up:
  // (value << 16) | (zero run << 8) | bits consumed; 0 = take the two-step path
  int32_t fast_ac[1 << FAST_BITS];
  void build_fast_ac() {
    for (int i = 0; i < (1 << FAST_BITS); i++) {
      fast_ac[i] = 0;
      const unsigned e = fast[i];
      if (!e) continue;
      const int len = e >> 8, run = (e & 255) >> 4, mag = e & 15;
      if (mag == 0 || len + mag > FAST_BITS) continue;
      int v = ((i << len) & ((1 << FAST_BITS) - 1)) >> (FAST_BITS - mag);
      if (v < (1 << (mag - 1))) v += -(1 << mag) + 1;
      fast_ac[i] = v * 65536 + (run << 8) + (len + mag);
    }
  }
  void build(const uint8_t counts[16], const uint8_t* vals, int nv) {
    // the counts must describe a prefix code: at most 2^l codes of length l once the shorter ones are taken out
    for (int l = 1, code = 0; l <= 16; l++) {
      code += counts[l - 1];
      if (code > (1 << l)) throw InvalidArg{"jpeg: bad DHT (not a prefix code)", UVO_ERR_INVALID};
      code <<= 1;
    }
    memset(huffval, 0, sizeof(huffval));  // entries past nv stay defined: a corrupt code can index them
    memcpy(huffval, vals, nv);
    memset(fast, 0, sizeof(fast));
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
      valoff[l] = k - code;
      for (int i = 0; i < counts[l - 1]; i++, k++, code++)
        if (l <= FAST_BITS) {
          const int lo = code << (FAST_BITS - l), n = 1 << (FAST_BITS - l);
          for (int j = 0; j < n; j++) fast[lo + j] = (uint16_t)((l << 8) | vals[k]);
        }
      maxcode[l] = counts[l - 1] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    present = true;
  }
};

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t buf = 0;
  int cnt = 0;
  bool marker = false;  // a marker (or the end of the data) was reached: zeros are fed from there on
  void refill() {
    while (cnt <= 56) {
      unsigned b = 0;
      if (!marker && p < end) {
        b = *p++;
        if (b == 0xFF) {
          if (p < end && *p == 0x00) {
            p++;  // stuffed zero
          } else {
            p--;
            marker = true;
            b = 0;
          }
        }
      }
      buf |= (uint64_t)b << (56 - cnt);
      cnt += 8;
    }
  }
  unsigned peek(int n) const { return (unsigned)(buf >> (64 - n)); }
  void skip(int n) {
    buf <<= n;
    cnt -= n;
  }
  void restart_at(const uint8_t* q) {
    p = q;
    buf = 0;
    cnt = 0;
    marker = false;
  }
};

inline int huff_decode(BitReader& b, const HuffTab& h) {
  if (b.cnt < 16) b.refill();
  const unsigned e = h.fast[b.peek(FAST_BITS)];
  if (e) {
    b.skip(e >> 8);
    return e & 255;
  }
  for (int l = FAST_BITS + 1; l <= 16; l++) {
    const int code = (int)b.peek(l);
    if (code <= h.maxcode[l]) {
      b.skip(l);
      return h.huffval[(code + h.valoff[l]) & 255];
    }
  }
  b.skip(16);
  return 0;  // not a code of this table (corrupt data): libjpeg carries on with a zero as well
}

inline int receive_extend(BitReader& b, int s) {
  if (b.cnt < s) b.refill();
  const int v = (int)b.peek(s);
  b.skip(s);
  return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// Where the decoded coefficients go.  Dense: int16 planes (the parity tap uvo_jpeg_entropy_decode).  Sparse: what
// uvo_jpeg_decode ships to the GPU -- one 32-bit entry per NON-ZERO coefficient, (natural index << 16) | value, in scan
// order, plus per block (numbered component-major, row-major: the k_jpeg_idct launch order) the position of its first
// entry and its entry count.  At q75-q90 11-16 % of the coefficients are non-zero: ~1 MB instead of 3.9 MB per 1280x1024.
struct CoefSink {
  int16_t* dense = nullptr;
  uint32_t* entries = nullptr;
  uint32_t* first = nullptr;
  uint8_t* count = nullptr;
  size_t n = 0, cap = 0;
};

struct CompInfo {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0;
};

struct Parser {
  uvo_jpeg_layout L;
  CompInfo comp[3];
  HuffTab dc[4], ac[4];
  uint16_t qt[4][64];
  bool qt_present[4] = {false, false, false, false};
  int hmax = 1, vmax = 1, restart = 0, adobe_transform = -1;
  bool have_sof = false;

  Parser() {
    memset(&L, 0, sizeof(L));
    memset(qt, 0, sizeof(qt));
  }

  // walks the marker segments; with `coef` decodes every scan into it, without stops at the first SOS
  CoefSink sink;
  void run(const uint8_t* d, size_t len, int16_t* coef) {
    sink.dense = coef;
    walk(d, len, coef != nullptr);
  }
  void run_sparse(const uint8_t* d, size_t len, uint32_t* entries, size_t cap, uint32_t* first, uint8_t* count) {
    sink.entries = entries;
    sink.cap = cap;
    sink.first = first;
    sink.count = count;
    walk(d, len, true);
  }
  void walk(const uint8_t* d, size_t len, bool decode) {
    if (!d || len < 4 || d[0] != 0xFF || d[1] != 0xD8) throw InvalidArg{"jpeg: not a JPEG stream (no SOI)", UVO_ERR_INVALID};
    size_t i = 2;
    while (i + 4 <= len) {
      if (d[i] != 0xFF || d[i + 1] == 0xFF) {
        i++;
        continue;
      }
      const int m = d[i + 1];
      i += 2;
      if (m == 0xD9) break;                                    // EOI
      if (m == 0x01 || m == 0x00 || (m >= 0xD0 && m <= 0xD7)) continue;  // stand-alone
      const size_t seg = ((size_t)d[i] << 8) | d[i + 1];
      if (seg < 2 || i + seg > len) throw InvalidArg{"jpeg: truncated marker segment", UVO_ERR_INVALID};
      const uint8_t* s = d + i + 2;
      const size_t n = seg - 2;
      switch (m) {
        case 0xDB: dqt(s, n); break;
        case 0xC4: dht(s, n); break;
        case 0xC0:
        case 0xC1: sof(s, n); break;
        case 0xDD:
          if (n < 2) throw InvalidArg{"jpeg: bad DRI", UVO_ERR_INVALID};
          restart = (s[0] << 8) | s[1];
          break;
        case 0xEE:
          if (n >= 12 && memcmp(s, "Adobe", 5) == 0) adobe_transform = s[11];
          break;
        case 0xDA: {
          if (!have_sof) throw InvalidArg{"jpeg: SOS before SOF", UVO_ERR_INVALID};
          if (!decode) {  // header only: report the quantisation tables as defined so far
            for (int c = 0; c < L.components; c++)
              if (qt_present[comp[c].tq]) memcpy(L.quant[c], qt[comp[c].tq], sizeof(L.quant[0]));
            return;
          }
          const uint8_t* e = scan(s, n, d + i + seg, d + len);
          i = (size_t)(e - d);
          continue;
        }
        default:
          if (m >= 0xC2 && m <= 0xCF && m != 0xC8)
            throw InvalidArg{"jpeg: only baseline / extended-sequential Huffman streams are supported", UVO_ERR_UNSUPPORTED};
      }
      i += seg;
    }
    if (!have_sof) throw InvalidArg{"jpeg: no frame header", UVO_ERR_INVALID};
    if (L.components == 3 && adobe_transform == 0)
      throw InvalidArg{"jpeg: Adobe RGB streams are not supported", UVO_ERR_UNSUPPORTED};
  }

  void dqt(const uint8_t* s, size_t n) {
    size_t k = 0;
    while (k < n) {
      const int pq = s[k] >> 4, tq = s[k] & 15;
      k++;
      if (tq > 3 || pq > 1 || k + (pq ? 128 : 64) > n) throw InvalidArg{"jpeg: bad DQT", UVO_ERR_INVALID};
      for (int j = 0; j < 64; j++) {
        qt[tq][kNatural[j]] = pq ? (uint16_t)((s[k] << 8) | s[k + 1]) : s[k];
        k += pq ? 2 : 1;
      }
      qt_present[tq] = true;
    }
  }

  void dht(const uint8_t* s, size_t n) {
    size_t k = 0;
    while (k + 17 <= n) {
      const int tc = s[k] >> 4, th = s[k] & 15;
      int nv = 0;
      for (int j = 0; j < 16; j++) nv += s[k + 1 + j];
      if (th > 3 || tc > 1 || nv > 256 || k + 17 + nv > n) throw InvalidArg{"jpeg: bad DHT", UVO_ERR_INVALID};
      (tc ? ac : dc)[th].build(s + k + 1, s + k + 17, nv);
      if (tc) ac[th].build_fast_ac();
      k += 17 + nv;
    }
  }

  void sof(const uint8_t* s, size_t n) {
    if (have_sof) throw InvalidArg{"jpeg: more than one frame header", UVO_ERR_INVALID};  // buffers follow the first
    if (n < 6) throw InvalidArg{"jpeg: bad SOF", UVO_ERR_INVALID};
    if (s[0] != 8) throw InvalidArg{"jpeg: only 8-bit samples are supported", UVO_ERR_UNSUPPORTED};
    L.height = (s[1] << 8) | s[2];
    L.width = (s[3] << 8) | s[4];
    L.components = s[5];
    if (L.width <= 0 || L.height <= 0) throw InvalidArg{"jpeg: empty frame", UVO_ERR_INVALID};
    // the bytes come off the network (a ROS CompressedImage): a 65535 x 65535 frame header in a tiny stream must not
    // size tens of GB of buffers (cv::imdecode has CV_IO_MAX_IMAGE_PIXELS = 2^30 for the same reason)
    if ((int64_t)L.width * L.height > JPEG_MAX_PIXELS)
      throw InvalidArg{"jpeg: frame larger than the supported 2^28 pixels", UVO_ERR_UNSUPPORTED};
    if (L.components != 1 && L.components != 3)
      throw InvalidArg{"jpeg: only 1- and 3-component streams are supported", UVO_ERR_UNSUPPORTED};
    if (n < (size_t)(6 + 3 * L.components)) throw InvalidArg{"jpeg: bad SOF", UVO_ERR_INVALID};
    for (int c = 0; c < L.components; c++) {
      comp[c].id = s[6 + 3 * c];
      comp[c].h = s[7 + 3 * c] >> 4;
      comp[c].v = s[7 + 3 * c] & 15;
      comp[c].tq = s[8 + 3 * c];
      if (comp[c].h < 1 || comp[c].h > 4 || comp[c].v < 1 || comp[c].v > 4 || comp[c].tq > 3)
        throw InvalidArg{"jpeg: bad component specification", UVO_ERR_INVALID};
    }
    if (L.components == 1) comp[0].h = comp[0].v = 1;  // a lone component is not subsampled (T.81 A.2.2)
    for (int c = 0; c < L.components; c++) {
      hmax = std::max(hmax, comp[c].h);
      vmax = std::max(vmax, comp[c].v);
    }
    const int mcux = div_up(L.width, 8 * hmax), mcuy = div_up(L.height, 8 * vmax);
    int64_t off = 0;
    for (int c = 0; c < L.components; c++) {
      if (hmax % comp[c].h || vmax % comp[c].v)
        throw InvalidArg{"jpeg: fractional sampling ratios are not supported", UVO_ERR_UNSUPPORTED};
      L.h_samp[c] = comp[c].h;
      L.v_samp[c] = comp[c].v;
      L.blocks_x[c] = mcux * comp[c].h;
      L.blocks_y[c] = mcuy * comp[c].v;
      L.samples_x[c] = div_up(L.width * comp[c].h, hmax);
      L.samples_y[c] = div_up(L.height * comp[c].v, vmax);
      L.coeff_offset[c] = off;
      off += (int64_t)L.blocks_x[c] * L.blocks_y[c] * 64;
    }
    L.coeff_total = off;
    have_sof = true;
  }

  // one scan: header at s, entropy-coded data from p; returns the position of the marker that ends it
  const uint8_t* scan(const uint8_t* s, size_t n, const uint8_t* p, const uint8_t* end) {
    const int ns = n ? s[0] : 0;
    if (ns < 1 || ns > L.components || n < (size_t)(1 + 2 * ns + 3)) throw InvalidArg{"jpeg: bad SOS", UVO_ERR_INVALID};
    int idx[3];
    for (int j = 0; j < ns; j++) {
      int c = -1;
      for (int q = 0; q < L.components; q++)
        if (comp[q].id == s[1 + 2 * j]) c = q;
      if (c < 0) throw InvalidArg{"jpeg: SOS names an unknown component", UVO_ERR_INVALID};
      for (int k = 0; k < j; k++)  // libjpeg-turbo: JERR_BAD_COMPONENT_ID
        if (idx[k] == c) throw InvalidArg{"jpeg: SOS names a component twice", UVO_ERR_INVALID};
      comp[c].td = s[2 + 2 * j] >> 4;
      comp[c].ta = s[2 + 2 * j] & 15;
      if (comp[c].td > 3 || comp[c].ta > 3 || !dc[comp[c].td].present || !ac[comp[c].ta].present ||
          !qt_present[comp[c].tq])
        throw InvalidArg{"jpeg: scan refers to a table that was not defined", UVO_ERR_INVALID};
      comp[c].pred = 0;
      idx[j] = c;
    }
    // the quantisation tables in force at the first scan of a component are the ones the frame is decoded with
    for (int j = 0; j < ns; j++) memcpy(L.quant[idx[j]], qt[comp[idx[j]].tq], sizeof(L.quant[0]));
    int mcus_x, mcus_y;
    if (ns == 1) {  // non-interleaved: one block per MCU, only the blocks that hold real samples
      mcus_x = div_up(L.samples_x[idx[0]], 8);
      mcus_y = div_up(L.samples_y[idx[0]], 8);
    } else {
      mcus_x = L.blocks_x[idx[0]] / comp[idx[0]].h;
      mcus_y = L.blocks_y[idx[0]] / comp[idx[0]].v;
    }
    BitReader b{p, end};
    int left = restart;
    for (int my = 0; my < mcus_y; my++)
      for (int mx = 0; mx < mcus_x; mx++) {
        if (restart && left == 0) {
          const uint8_t* q = b.p;  // the reader never moves past a marker: the RSTn is at or after b.p
          while (q + 1 < end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
          if (q + 1 >= end) throw InvalidArg{"jpeg: missing restart marker", UVO_ERR_INVALID};
          b.restart_at(q + 2);
          for (int j = 0; j < ns; j++) comp[idx[j]].pred = 0;
          left = restart;
        }
        for (int j = 0; j < ns; j++) {
          CompInfo& c = comp[idx[j]];
          const int ci = idx[j];
          const int nbx = ns == 1 ? 1 : c.h, nby = ns == 1 ? 1 : c.v;
          for (int by = 0; by < nby; by++)
            for (int bx = 0; bx < nbx; bx++) {
              const int X = mx * nbx + bx, Y = my * nby + by;
              const int64_t blk = L.coeff_offset[ci] / 64 + (int64_t)Y * L.blocks_x[ci] + X;
              if (sink.dense)
                block<false>(b, c, sink.dense + blk * 64, blk);
              else
                block<true>(b, c, nullptr, blk);
            }
        }
        if (restart) left--;
      }
    const uint8_t* q = b.p;  // the reader stops in front of a marker, so the next one is at or after b.p
    while (q + 1 < end && !(q[0] == 0xFF && q[1] != 0x00 && q[1] != 0xFF && !(q[1] >= 0xD0 && q[1] <= 0xD7))) q++;
    return q;
  }

  template <bool SPARSE>
  void emit(int16_t* out, int pos, int value) {
    if (SPARSE) {
      if (sink.n >= sink.cap) throw InvalidArg{"jpeg: coefficient entry buffer too small", UVO_ERR_CAPACITY};
      sink.entries[sink.n++] = ((uint32_t)pos << 16) | (uint16_t)(int16_t)value;
    } else {
      out[pos] = (int16_t)value;
    }
  }

  template <bool SPARSE>
  void block(BitReader& b, CompInfo& c, int16_t* out, int64_t blk) {
    const size_t n0 = sink.n;
    int s = huff_decode(b, dc[c.td]);
    if (s > 15) throw InvalidArg{"jpeg: corrupt DC coefficient", UVO_ERR_INVALID};
    // valid streams keep the DC predictor within 16 bits; a corrupt one must not run it into signed overflow
    c.pred = (int)(int16_t)(c.pred + (s ? receive_extend(b, s) : 0));
    if (!SPARSE || (int16_t)c.pred != 0) emit<SPARSE>(out, 0, c.pred);
    const HuffTab& t = ac[c.ta];
    for (int k = 1; k < 64;) {
      if (b.cnt < 16) b.refill();
      const int fa = t.fast_ac[b.peek(FAST_BITS)];
      if (fa) {  // run, size and magnitude bits in one step
        k += (fa >> 8) & 255;
        if (k > 63) throw InvalidArg{"jpeg: corrupt AC run", UVO_ERR_INVALID};
        b.skip(fa & 255);
        emit<SPARSE>(out, kNatural[k++], fa >> 16);
        continue;
      }
      const int rs = huff_decode(b, t);
      s = rs & 15;
      if (s == 0) {
        if ((rs >> 4) != 15) break;  // end of block
        k += 16;
        continue;
      }
      k += rs >> 4;
      if (k > 63) throw InvalidArg{"jpeg: corrupt AC run", UVO_ERR_INVALID};
      emit<SPARSE>(out, kNatural[k], receive_extend(b, s));
      k++;
    }
    if (SPARSE) {
      sink.first[blk] = (uint32_t)n0;
      sink.count[blk] = (uint8_t)(sink.n - n0);  // <= 64
    }
  }
};

// ------------------------------------------------------------------------------------------------ device
// The thread bodies live in jpeg_kernels.cuh (host/device functions of the block and thread index, so that the test
// harness can run the same code on the CPU); the kernels bind them to the launch geometry.
using namespace uvo::jpegk;

__global__ void __launch_bounds__(IDCT_THREADS) k_jpeg_idct(const __grid_constant__ IdctArgs a) {
  __shared__ int s_ws[IDCT_BLOCKS * WS_STRIDE];
  __shared__ int16_t s_tile[IDCT_BLOCKS * TILE_STRIDE];
  idct_clear(a, blockIdx.x, threadIdx.x, s_tile);
  __syncwarp();  // the eight threads of a JPEG block sit in one warp
  idct_scatter(a, blockIdx.x, threadIdx.x, s_tile);
  __syncwarp();
  idct_pass1(a, blockIdx.x, threadIdx.x, s_tile, s_ws);
  __syncwarp();
  idct_pass2(a, blockIdx.x, threadIdx.x, s_ws);
}

__global__ void __launch_bounds__(COLOR_TX* COLOR_TY) k_jpeg_color(const __grid_constant__ ColorArgs a) {
  color_thread(a, blockIdx.x, blockIdx.y, threadIdx.x);
}


// ------------------------------------------------------------------------------------------------ Huffman decode on the GPU
// see jpeg_huff.cuh.  Host side: walk the markers, build the table plan, copy the scan with the stuffed zeros removed.
// Returns false for streams this path does not take (several scans, restart intervals, table ids above 1): the caller
// uses the host decoder for those.  Throws on malformed streams like the host decoder.
struct GpuPlanOut {
  uvo_jpeg_layout L;
  size_t scan_bytes;
};

bool build_gpu_plan(const uint8_t* d, size_t len, JhPlan* plan, uint8_t* scan_out, size_t scan_cap, GpuPlanOut* out) {
  struct Walker : Parser {
    const uint8_t* sos = nullptr;
    size_t sos_n = 0;
    const uint8_t* data = nullptr;
    const uint8_t* end = nullptr;
    int scans = 0;
  } P;
  // the marker walk of Parser::walk, stopping at the first SOS with the position of its data
  if (!d || len < 4 || d[0] != 0xFF || d[1] != 0xD8) throw InvalidArg{"jpeg: not a JPEG stream (no SOI)", UVO_ERR_INVALID};
  size_t i = 2;
  while (i + 4 <= len) {
    if (d[i] != 0xFF || d[i + 1] == 0xFF) {
      i++;
      continue;
    }
    const int m = d[i + 1];
    i += 2;
    if (m == 0xD9) break;
    if (m == 0x01 || m == 0x00 || (m >= 0xD0 && m <= 0xD7)) continue;
    const size_t seg = ((size_t)d[i] << 8) | d[i + 1];
    if (seg < 2 || i + seg > len) throw InvalidArg{"jpeg: truncated marker segment", UVO_ERR_INVALID};
    const uint8_t* s = d + i + 2;
    const size_t n = seg - 2;
    if (m == 0xDB) P.dqt(s, n);
    else if (m == 0xC4) P.dht(s, n);
    else if (m == 0xC0 || m == 0xC1) P.sof(s, n);
    else if (m == 0xDD) {
      if (n < 2) throw InvalidArg{"jpeg: bad DRI", UVO_ERR_INVALID};
      P.restart = (s[0] << 8) | s[1];
    } else if (m == 0xEE) {
      if (n >= 12 && memcmp(s, "Adobe", 5) == 0) P.adobe_transform = s[11];
    } else if (m == 0xDA) {
      if (!P.have_sof) throw InvalidArg{"jpeg: SOS before SOF", UVO_ERR_INVALID};
      P.sos = s;
      P.sos_n = n;
      P.data = d + i + seg;
      P.end = d + len;
      break;
    } else if (m >= 0xC2 && m <= 0xCF && m != 0xC8) {
      throw InvalidArg{"jpeg: only baseline / extended-sequential Huffman streams are supported", UVO_ERR_UNSUPPORTED};
    }
    i += seg;
  }
  if (!P.have_sof || !P.sos) throw InvalidArg{"jpeg: no frame header / no scan", UVO_ERR_INVALID};
  if (P.L.components == 3 && P.adobe_transform == 0)
    throw InvalidArg{"jpeg: Adobe RGB streams are not supported", UVO_ERR_UNSUPPORTED};
  const uvo_jpeg_layout& L = P.L;
  const int ns = P.sos_n ? P.sos[0] : 0;
  if (ns < 1 || ns > L.components || P.sos_n < (size_t)(1 + 2 * ns + 3)) throw InvalidArg{"jpeg: bad SOS", UVO_ERR_INVALID};
  if (ns != L.components || P.restart != 0) return false;  // several scans / restart intervals: host decoder
  memset(plan, 0, sizeof(*plan));
  int order[3];
  for (int j = 0; j < ns; j++) {
    int c = -1;
    for (int q = 0; q < L.components; q++)
      if (P.comp[q].id == P.sos[1 + 2 * j]) c = q;
    if (c < 0) throw InvalidArg{"jpeg: SOS names an unknown component", UVO_ERR_INVALID};
    // every component exactly once: a component named twice would leave another one's block table unwritten
    for (int k = 0; k < j; k++)
      if (order[k] == c) throw InvalidArg{"jpeg: SOS names a component twice", UVO_ERR_INVALID};
    const int td = P.sos[2 + 2 * j] >> 4, ta = P.sos[2 + 2 * j] & 15;
    if (td > 3 || ta > 3 || !P.dc[td].present || !P.ac[ta].present || !P.qt_present[P.comp[c].tq])
      throw InvalidArg{"jpeg: scan refers to a table that was not defined", UVO_ERR_INVALID};
    if (td > 1 || ta > 1) return false;  // baseline allows table ids 0 and 1
    plan->dc_tab[c] = (uint8_t)td;
    plan->ac_tab[c] = (uint8_t)(2 + ta);
    order[j] = c;
  }
  for (int t = 0; t < 4; t++) {
    const HuffTab& h = t < 2 ? P.dc[t] : P.ac[t - 2];
    if (!h.present) continue;
    for (int k = 0; k < (1 << JH_FAST_BITS); k++) {
      const unsigned e = h.fast[k << (FAST_BITS - JH_FAST_BITS)];
      plan->fast[t][k] = (e && (int)(e >> 8) <= JH_FAST_BITS) ? (uint16_t)e : 0;
    }
    memcpy(plan->maxcode[t], h.maxcode, sizeof(h.maxcode));
    memcpy(plan->valoff[t], h.valoff, sizeof(h.valoff));
    memcpy(plan->huffval[t], h.huffval, sizeof(h.huffval));
    // second level: every 16-bit pattern whose 10-bit prefix starts a longer code, resolved with the host decoder's rule
    int n_sub = 0;
    for (int pre = 0; pre < (1 << JH_FAST_BITS); pre++) {
      if (plan->fast[t][pre]) continue;
      bool any = false;
      uint16_t tab[64];
      for (int low = 0; low < 64; low++) {
        const int x16 = (pre << 6) | low;
        tab[low] = 0;
        for (int l = JH_FAST_BITS + 1; l <= 16; l++) {
          const int code = x16 >> (16 - l);
          if (code <= h.maxcode[l]) {
            tab[low] = (uint16_t)((l << 8) | h.huffval[(code + h.valoff[l]) & 255]);
            any = true;
            break;
          }
        }
      }
      if (!any || n_sub >= JH_SUB_TABLES) continue;  // no code there, or out of sub-tables: the slow walk handles it
      memcpy(plan->sub[t] + 64 * n_sub, tab, sizeof(tab));
      plan->fast[t][pre] = (uint16_t)(0x8000 | n_sub);
      n_sub++;
    }
  }
  int bpm = 0;
  for (int j = 0; j < ns; j++) {
    const int c = order[j];
    const int nbx = ns == 1 ? 1 : P.comp[c].h, nby = ns == 1 ? 1 : P.comp[c].v;
    for (int by = 0; by < nby; by++)
      for (int bx = 0; bx < nbx; bx++) {
        if (bpm >= JH_MAX_BPM) return false;
        plan->blk_comp[bpm] = (uint8_t)c;
        plan->blk_v[bpm] = (uint8_t)by;
        plan->blk_h[bpm] = (uint8_t)bx;
        bpm++;
      }
  }
  plan->bpm = bpm;
  plan->components = L.components;
  if (ns == 1) {
    // a single-component frame: one block per MCU, row-major over the blocks that hold real samples -- which are all
    // of its blocks only when the padded block grid equals the sample grid
    plan->mcus_x = div_up(L.samples_x[order[0]], 8);
    plan->mcus_y = div_up(L.samples_y[order[0]], 8);
    if (plan->mcus_x != L.blocks_x[order[0]] || plan->mcus_y != L.blocks_y[order[0]]) return false;
  } else {
    plan->mcus_x = L.blocks_x[order[0]] / P.comp[order[0]].h;
    plan->mcus_y = L.blocks_y[order[0]] / P.comp[order[0]].v;
  }
  plan->total_blocks = plan->mcus_x * plan->mcus_y * bpm;
  if ((int64_t)plan->total_blocks != L.coeff_total / 64) return false;
  for (int c = 0; c < L.components; c++) {
    plan->H[c] = ns == 1 ? 1 : P.comp[c].h;
    plan->V[c] = ns == 1 ? 1 : P.comp[c].v;
    plan->blocks_x[c] = L.blocks_x[c];
    plan->block_off[c] = (int32_t)(L.coeff_offset[c] / 64);
  }
  for (int j = 0; j < ns; j++) memcpy(P.L.quant[order[j]], P.qt[P.comp[order[j]].tq], sizeof(P.L.quant[0]));
  // the scan data: up to the next marker that is neither a stuffed zero nor (there are none) a restart; stuffed
  // zeros removed on the way into the (pinned) output
  const uint8_t* q = P.data;
  size_t o = 0;
  while (q < P.end) {
    const uint8_t* f = (const uint8_t*)memchr(q, 0xFF, (size_t)(P.end - q));
    const size_t run = (size_t)((f ? f : P.end) - q);
    if (o + run + 1 > scan_cap) throw InvalidArg{"jpeg: scan larger than its staging buffer", UVO_ERR_CAPACITY};
    memcpy(scan_out + o, q, run);
    o += run;
    if (!f) {
      q = P.end;
      break;
    }
    if (f + 1 < P.end && f[1] == 0x00) {  // stuffed zero: keep the 0xFF, drop the zero
      scan_out[o++] = 0xFF;
      q = f + 2;
      continue;
    }
    if (f + 1 < P.end && f[1] == 0xFF) {  // fill byte before a marker
      q = f + 1;
      continue;
    }
    q = f;  // a marker: the scan ends here
    break;
  }
  // a second scan (the stream would be multi-scan after all) is not handled here
  if (q + 1 < P.end && q[0] == 0xFF && q[1] != 0xD9) return false;
  if (o + 16 > scan_cap) throw InvalidArg{"jpeg: scan larger than its staging buffer", UVO_ERR_CAPACITY};
  memset(scan_out + o, 0, 16);
  if (o * 8 >= (size_t)1 << 31) return false;
  plan->total_bits = (uint32_t)(o * 8);
  plan->entries_cap = (uint32_t)L.coeff_total;
  out->L = P.L;
  out->scan_bytes = o;
  return true;
}

struct JhShared {
  uint16_t fast[4][1 << JH_FAST_BITS];
  uint16_t sub[4][JH_SUB_TABLES * 64];
  int32_t maxcode[4][18];
  int32_t valoff[4][17];
  uint8_t huffval[4][256];
  uint8_t blk_comp[JH_MAX_BPM + 2], blk_v[JH_MAX_BPM + 2], blk_h[JH_MAX_BPM + 2], dc_tab[4], ac_tab[4];
};

__device__ const uint8_t d_kNatural[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                           12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                           35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                           58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// decoder state: position, block of the MCU, next coefficient index (0 = the DC symbol comes next)
struct JhState {
  unsigned pos;
  int p, z;
};

// the big-endian bit stream from bit `pos` on, through a 64-bit register window refilled one aligned word at a time:
// a symbol (<= 27 bits) costs shifts, not loads
struct JhBits {
  const uint32_t* __restrict__ w;
  unsigned long long buf;  // the next `avail` bits of the stream, left-aligned
  int avail;
  unsigned next;           // index of the word after `ahead`
  uint32_t ahead;          // the next word, loaded one refill early so that its latency hides behind the symbols between
  __device__ __forceinline__ void open(const uint32_t* __restrict__ words, unsigned pos) {
    w = words;
    const unsigned k = pos >> 5, sh = pos & 31;
    const unsigned long long a = __byte_perm(__ldg(w + k), 0, 0x0123), b = __byte_perm(__ldg(w + k + 1), 0, 0x0123);
    buf = ((a << 32) | b) << sh;
    avail = 64 - (int)sh;
    ahead = __ldg(w + k + 2);
    next = k + 3;
  }
  __device__ __forceinline__ uint32_t peek32() {
    if (avail <= 32) {
      buf |= (unsigned long long)__byte_perm(ahead, 0, 0x0123) << (32 - avail);
      avail += 32;
      ahead = __ldg(w + next);
      next++;
    }
    return (uint32_t)(buf >> 32);
  }
  __device__ __forceinline__ void skip(int n) {
    buf <<= n;
    avail -= n;
  }
};

// One symbol (Huffman code + magnitude bits).  kind: 0 = DC difference (a block starts), 1 = AC coefficient at zig-zag
// index zz, 2 = nothing to store (end of block / zero run of 16).  Returns false when the symbol would run past the
// end of the data (the decode stops there: what is left are the padding bits).  Same rules as the host decoder for
// codes that are not in the table (16 bits skipped, symbol 0).
__device__ __forceinline__ bool jh_symbol(const JhShared& T, JhBits& in, unsigned total_bits, int bpm, JhState& s,
                                          int& kind, int& zz, int& val) {
  const uint32_t x = in.peek32();
  const int comp = T.blk_comp[s.p];
  const int t = s.z == 0 ? T.dc_tab[comp] : T.ac_tab[comp];
  int len, sym;
  unsigned e = T.fast[t][x >> (32 - JH_FAST_BITS)];
  if (e & 0x8000u) e = T.sub[t][((e & 0x7fffu) << 6) | ((x >> (32 - JH_FAST_BITS - 6)) & 63u)];
  if (e) {
    len = e >> 8;
    sym = e & 255;
  } else {
    len = 16;
    sym = 0;
    for (int l = JH_FAST_BITS + 1; l <= 16; l++) {
      const int code = (int)(x >> (32 - l));
      if (code <= T.maxcode[t][l]) {
        len = l;
        sym = T.huffval[t][(code + T.valoff[t][l]) & 255];
        break;
      }
    }
  }
  const bool dc = s.z == 0;
  const int mag = dc ? min(sym, 15) : (sym & 15);  // the host decoder rejects DC categories above 15
  const int used = len + mag;
  if (s.pos + used > total_bits) return false;
  int v = 0;
  if (mag) {
    v = (int)((x << len) >> (32 - mag));
    if (v < (1 << (mag - 1))) v += -(1 << mag) + 1;
  }
  val = v;
  if (dc) {
    kind = 0;
    zz = 0;
    s.z = 1;
  } else {
    const int run = sym >> 4;
    if (mag == 0) {
      kind = 2;
      s.z = (run == 15) ? s.z + 16 : 64;  // ZRL / EOB
    } else {
      const int k = s.z + run;
      kind = k <= 63 ? 1 : 2;  // a run past the block is corrupt data: the block ends
      zz = k & 63;
      s.z = k + 1;
    }
  }
  s.pos += used;
  in.skip(used);
  if (s.z >= 64) {
    s.z = 0;
    s.p = s.p + 1 == bpm ? 0 : s.p + 1;
  }
  return true;
}

// exclusive prefix sums of up to three ints per thread over one block; totals returned to everyone
template <int K>
__device__ __forceinline__ void jh_block_scan(int (&v)[K], int (*s_w)[K], int (&tot)[K]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    inc[k] = v[k];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc[k], o);
      if (lane >= o) inc[k] += t;
    }
    if (lane == 31) s_w[wid][k] = inc[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < K; k++) {
      int run = 0;
      for (int w = 0; w < nw; w++) {
        const int t = s_w[w][k];
        s_w[w][k] = run;
        run += t;
      }
      s_w[nw][k] = run;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) {
    tot[k] = s_w[nw][k];
    v[k] = s_w[wid][k] + inc[k] - v[k];
  }
  __syncthreads();
}

// Kernel with grid-wide barriers (jh_soft_sync, or cooperative_groups under UVO_JPEG_COOPERATIVE=1): grid = (blocks
// per image, images), JH_BLOCK threads each, one thread per
// sub-sequence of JH_SUB_BITS bits.  Phases: synchronisation rounds; count; per-image scan of the counts; write;
// DC partial sums; per-image scan of those; DC prediction + block table in plane order.
__device__ __forceinline__ unsigned long long jh_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Grid-wide barrier of a launch whose blocks are all resident (the host bounds how many of these launches are in
// flight, see jpeg_gpu_max_concurrent): a monotonic arrival counter, one poller per block.  A block that waits longer
// than JH_BARRIER_TIMEOUT_NS raises the launch's abort word and every block leaves -- the image is reported corrupt
// instead of the GPU hanging if the residency assumption were ever violated.
constexpr unsigned long long JH_BARRIER_TIMEOUT_NS = 500ull * 1000 * 1000;
__device__ __forceinline__ bool jh_soft_sync(unsigned* bar, unsigned& target, unsigned nblk) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblk;
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v, ab = 0, polls = 0;
    unsigned long long t0 = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if ((++polls & 1023u) == 0) {
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(ab) : "l"(bar + 1) : "memory");
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (!t0) t0 = t;
        if (t - t0 > JH_BARRIER_TIMEOUT_NS) {
          atomicExch(bar + 1, 1u);
          ab = 1;
        }
        if (ab) break;
      }
    }
    __threadfence();
    s_ok = !ab;
  }
  __syncthreads();
  return s_ok != 0;
}

__global__ void __launch_bounds__(JH_BLOCK) k_jpeg_huff(const __grid_constant__ JhArgs args) {
  unsigned bar_target = 0;
  const unsigned n_blk = gridDim.x * gridDim.y;
  bool aborted = false;
  auto grid_sync = [&]() {
    if (args.bar) {
      if (!aborted && !jh_soft_sync(args.bar, bar_target, n_blk)) aborted = true;
    } else {
      cooperative_groups::this_grid().sync();
    }
  };
  __shared__ JhShared T;
  __shared__ int s_w[JH_BLOCK / 32 + 1][3];
  const JhImage& im = args.im[blockIdx.y];
  const JhPlan* __restrict__ P = im.plan;
  const int tid = threadIdx.x;
  for (int i = tid; i < (int)(sizeof(T.fast) / 2); i += JH_BLOCK) (&T.fast[0][0])[i] = (&P->fast[0][0])[i];
  for (int i = tid; i < (int)(sizeof(T.sub) / 2); i += JH_BLOCK) (&T.sub[0][0])[i] = (&P->sub[0][0])[i];
  for (int i = tid; i < 4 * 18; i += JH_BLOCK) (&T.maxcode[0][0])[i] = (&P->maxcode[0][0])[i];
  for (int i = tid; i < 4 * 17; i += JH_BLOCK) (&T.valoff[0][0])[i] = (&P->valoff[0][0])[i];
  for (int i = tid; i < 4 * 256; i += JH_BLOCK) (&T.huffval[0][0])[i] = (&P->huffval[0][0])[i];
  if (tid < JH_MAX_BPM + 2) {
    T.blk_comp[tid] = P->blk_comp[tid];
    T.blk_v[tid] = P->blk_v[tid];
    T.blk_h[tid] = P->blk_h[tid];
  }
  if (tid < 4) {
    T.dc_tab[tid] = P->dc_tab[tid];
    T.ac_tab[tid] = P->ac_tab[tid];
  }
  const unsigned total_bits = P->total_bits;
  const int bpm = P->bpm, total_blocks = P->total_blocks;
  const unsigned cap = P->entries_cap;
  const uint32_t* __restrict__ w = im.scan;
  const int n_sub = (int)((total_bits + JH_SUB_BITS - 1) / JH_SUB_BITS);  // sub-sequences of this image
  const int g = blockIdx.x * JH_BLOCK + tid;                             // this thread's sub-sequence
  const bool live = g < n_sub;
  int* flag = args.flag;
  // this thread's share of the scan-order blocks in the DC passes (and of the block table it must leave decodable)
  const int n_thr = gridDim.x * JH_BLOCK;
  const int bper = (total_blocks + n_thr - 1) / n_thr;
  const int gi = blockIdx.x * JH_BLOCK + tid;
  const int b0 = min(gi * bper, total_blocks), b1 = min(b0 + bper, total_blocks);
  auto leave_empty = [&]() {  // a corrupt stream (or an aborted launch) leaves a decodable, empty image behind
    for (int b = b0; b < b1; b++) {
      im.first[b] = 0;
      im.count[b] = 0;
    }
    if (aborted && tid == 0) {  // every block that gives up says so: the launch is reported as a corrupt image
      im.info[1] = 1;
      if (gi == 0) {
        im.info[0] = 0;
        im.info[2] = JH_MAX_ROUNDS + 1;
      }
    }
  };
#define JH_GRID_SYNC()  \
  do {                  \
    grid_sync();        \
    if (aborted) {      \
      leave_empty();    \
      return;           \
    }                   \
  } while (0)
  __syncthreads();
  // diagnostics: phase time stamps (ns) of block (0, 0) in info[16 .. 63] (tools/jh_time.py)
  unsigned long long* stamps = (unsigned long long*)(args.im[0].info + 16);
  int n_stamp = 0;
  auto stamp = [&]() {
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && n_stamp < 24) stamps[n_stamp++] = jh_now();
  };
  stamp();
  // ---- phase 1: self-synchronisation.  Round r reads the exit states of round r - 1 and writes those of round r
  // (two buffers); a thread whose start state did not change keeps its exit state.  The loop runs until a round
  // changes nothing in ANY image of the launch (every block must take part in every grid barrier).
  const unsigned begin = (unsigned)g * JH_SUB_BITS, end = min(begin + JH_SUB_BITS, total_bits);
  JhState start{begin, 0, 0}, prev{0xffffffffu, -1, -1}, ex{begin, 0, 0};
  unsigned cp_pos0 = 0xffffffffu, cp_pos1 = 0xffffffffu, cp_pos2 = 0xffffffffu, cp_pz0 = 0, cp_pz1 = 0, cp_pz2 = 0;
  int kind, zz, val;
  int rounds = 0;
  for (;; rounds++) {
    uint2* ex_cur = im.exits + (size_t)(rounds & 1) * n_sub;
    const uint2* ex_prev = im.exits + (size_t)((rounds & 1) ^ 1) * n_sub;
    // three flag slots: the one cleared here was last read in round - 2, and every thread has since passed the barrier
    // of round - 1
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) flag[(rounds + 1) % 3] = 0;
    if (live) {
      if (g > 0 && rounds > 0) {
        const uint2 e = ex_prev[g - 1];
        start.pos = e.x;
        start.p = (int)(e.y >> 8);
        start.z = (int)(e.y & 255);
      }
      if (start.pos != prev.pos || start.p != prev.p || start.z != prev.z) {
        prev = start;
        JhState s = start;
        JhBits in;
        in.open(w, s.pos);
        // check-points: the decoder's state at the first symbol boundary at or after each quarter of the sub-sequence.
        // A re-decode that reaches a check-point in the state the previous decode had there is in step with it for
        // good: it stops, and the exit state stands.
        int ck = 0;
        bool in_step = false;
        while (s.pos < end) {
          if (ck < 3 && s.pos >= begin + (unsigned)(JH_SUB_BITS / 4) * (ck + 1)) {
            const unsigned key = (unsigned)((s.p << 8) | s.z);
            unsigned& cpos = ck == 0 ? cp_pos0 : (ck == 1 ? cp_pos1 : cp_pos2);
            unsigned& cpz = ck == 0 ? cp_pz0 : (ck == 1 ? cp_pz1 : cp_pz2);
            if (rounds > 0 && cpos == s.pos && cpz == key) {
              in_step = true;
              break;
            }
            cpos = s.pos;
            cpz = key;
            ck++;
          }
          if (!jh_symbol(T, in, total_bits, bpm, s, kind, zz, val)) break;
        }
        if (!in_step) {
          if (s.pos < end) s.pos = total_bits;  // stopped by the end of the data
          if (rounds == 0 || s.pos != ex.pos || s.p != ex.p || s.z != ex.z) flag[rounds % 3] = 1;
          ex = s;
        }
      }
      ex_cur[g] = make_uint2(ex.pos, (unsigned)((ex.p << 8) | ex.z));
    }
    JH_GRID_SYNC();
    if (rounds < 2) stamp();
    if (!flag[rounds % 3] || rounds >= JH_MAX_ROUNDS) break;
  }
  stamp();
  // (the start state of the last round is the true one: nothing changed in it)
  // ---- phase 2: count what each thread owns (the symbols that start in [its true start, its exit))
  int cnt[2] = {0, 0};  // blocks started, entries
  if (live) {
    JhState s = start;
    JhBits in;
    in.open(w, s.pos);
    while (s.pos < end && jh_symbol(T, in, total_bits, bpm, s, kind, zz, val)) {
      cnt[0] += kind == 0;
      cnt[1] += kind != 2;
    }
    im.counts[g] = make_int2(cnt[0], cnt[1]);
  }
  JH_GRID_SYNC();
  stamp();
  // ---- per-image exclusive scan of the counts by the image's first block
  if (blockIdx.x == 0) {
    const int per = (n_sub + JH_BLOCK - 1) / JH_BLOCK;
    const int a0 = min(tid * per, n_sub), a1 = min(a0 + per, n_sub);
    int v[2] = {0, 0}, tot[2];
    for (int i = a0; i < a1; i++) {
      const int2 c = im.counts[i];
      v[0] += c.x;
      v[1] += c.y;
    }
    jh_block_scan<2>(v, (int(*)[2])s_w, tot);
    for (int i = a0; i < a1; i++) {
      const int2 c = im.counts[i];
      im.counts[i] = make_int2(v[0], v[1]);
      v[0] += c.x;
      v[1] += c.y;
    }
    if (tid == 0) {
      const int err = (tot[0] != total_blocks || (unsigned)tot[1] > cap || rounds >= JH_MAX_ROUNDS) ? 1 : 0;
      im.first_scan[total_blocks] = (uint32_t)min((unsigned)tot[1], cap);
      im.info[0] = (int)min((unsigned)tot[1], cap);
      im.info[1] = err;
      im.info[2] = rounds + 1;
    }
  }
  JH_GRID_SYNC();
  stamp();
  // ---- phase 3: write the entries (DC entries carry the DIFFERENCE for now) and the scan-order block table
  if (live) {
    const int2 base = im.counts[g];
    JhState s = start;
    JhBits in;
    in.open(w, s.pos);
    int sb = base.x, e = base.y;
    while (s.pos < end && jh_symbol(T, in, total_bits, bpm, s, kind, zz, val)) {
      if (kind == 0) {
        if (sb < total_blocks) im.first_scan[sb] = (uint32_t)e;
        sb++;
      }
      if (kind != 2) {
        if ((unsigned)e < cap) im.entries[e] = ((uint32_t)(kind == 0 ? 0 : d_kNatural[zz]) << 16) | (uint16_t)(int16_t)val;
        e++;
      }
    }
  }
  JH_GRID_SYNC();
  stamp();
  const bool err = im.info[1] != 0;
  // ---- phase 4: DC prediction = running sum of the differences per component, in scan order: per-thread partial sums
  // over a contiguous range of scan-order blocks, scanned by the image's first block
  if (!err) {
    int acc[3] = {0, 0, 0};
    for (int sb = b0; sb < b1; sb++) acc[T.blk_comp[sb % bpm]] += (int)(int16_t)(im.entries[im.first_scan[sb]] & 0xffffu);
    im.dc_part[gi] = make_int4(acc[0], acc[1], acc[2], 0);
  }
  JH_GRID_SYNC();
  if (blockIdx.x == 0 && !err) {
    const int per = (n_thr + JH_BLOCK - 1) / JH_BLOCK;
    const int a0 = min(tid * per, n_thr), a1 = min(a0 + per, n_thr);
    int v[3] = {0, 0, 0}, tot[3];
    for (int i = a0; i < a1; i++) {
      const int4 c = im.dc_part[i];
      v[0] += c.x;
      v[1] += c.y;
      v[2] += c.z;
    }
    jh_block_scan<3>(v, s_w, tot);
    for (int i = a0; i < a1; i++) {
      const int4 c = im.dc_part[i];
      im.dc_part[i] = make_int4(v[0], v[1], v[2], 0);
      v[0] += c.x;
      v[1] += c.y;
      v[2] += c.z;
    }
  }
  JH_GRID_SYNC();
  // ---- phase 5: absolute DC values, and the block table in plane order (component-major, row-major), which is how
  // k_jpeg_idct numbers its blocks.  A corrupt stream leaves a decodable (empty) image behind.
  if (err) {
    leave_empty();
    return;
  }
  const int4 pb = im.dc_part[gi];
  int pred[3] = {pb.x, pb.y, pb.z};
  const int mcus_x = P->mcus_x;
  for (int sb = b0; sb < b1; sb++) {
    const int k = sb % bpm, m = sb / bpm, c = T.blk_comp[k];
    const uint32_t f = im.first_scan[sb], f1 = im.first_scan[sb + 1];
    pred[c] += (int)(int16_t)(im.entries[f] & 0xffffu);
    im.entries[f] = (uint32_t)(uint16_t)(int16_t)pred[c];  // natural index 0
    const int mx = m % mcus_x, my = m / mcus_x;
    const int X = mx * P->H[c] + T.blk_h[k], Y = my * P->V[c] + T.blk_v[k];
    const int blk = P->block_off[c] + Y * P->blocks_x[c] + X;
    im.first[blk] = f;
    im.count[blk] = (uint8_t)(f1 - f);
  }
  stamp();
  if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) args.im[0].info[15] = n_stamp;
}
#undef JH_GRID_SYNC
}  // namespace

namespace uvo {

size_t jpeg_sparse_device_bytes(const uvo_jpeg_layout& L, size_t n_entries) {
  const size_t nb = (size_t)L.coeff_total / 64;
  return sizeof(uint32_t) * (nb + std::max<size_t>(n_entries, 1)) + nb;
}

size_t jpeg_plane_bytes(const uvo_jpeg_layout& L) { return jpegk::plane_bytes(L); }

void jpeg_upload_sparse(cudaStream_t copy_stream, const uvo_jpeg_sparse& sp, uint8_t* d_sparse) {
  const size_t nb = (size_t)sp.layout.coeff_total / 64, ne = sp.n_entries;
  uint32_t* d_first = (uint32_t*)d_sparse;
  uint32_t* d_entries = d_first + nb;
  uint8_t* d_count = (uint8_t*)(d_entries + std::max<size_t>(ne, 1));
  UVO_CUDA(cudaMemcpyAsync(d_first, sp.block_first, sizeof(uint32_t) * nb, cudaMemcpyHostToDevice, copy_stream));
  if (ne) UVO_CUDA(cudaMemcpyAsync(d_entries, sp.entries, sizeof(uint32_t) * ne, cudaMemcpyHostToDevice, copy_stream));
  UVO_CUDA(cudaMemcpyAsync(d_count, sp.block_count, nb, cudaMemcpyHostToDevice, copy_stream));
}

static void transform_ptrs(Ctx& c, const uvo_jpeg_layout& L, const uint32_t* d_first, const uint32_t* d_entries,
                           const uint8_t* d_count, uint8_t* d_planes, int bayer_bggr, uint8_t* d_bgr, size_t bgr_pitch);

void jpeg_launch_transform(Ctx& c, const uvo_jpeg_layout& L, size_t n_entries, uint8_t* d_sparse, uint8_t* d_planes,
                           int bayer_bggr, uint8_t* d_bgr, size_t bgr_pitch) {
  const size_t nb = (size_t)L.coeff_total / 64;
  uint32_t* d_first = (uint32_t*)d_sparse;
  uint32_t* d_entries = d_first + nb;
  uint8_t* d_count = (uint8_t*)(d_entries + std::max<size_t>(n_entries, 1));
  transform_ptrs(c, L, d_first, d_entries, d_count, d_planes, bayer_bggr, d_bgr, bgr_pitch);
}

static void transform_ptrs(Ctx& c, const uvo_jpeg_layout& L, const uint32_t* d_first, const uint32_t* d_entries,
                           const uint8_t* d_count, uint8_t* d_planes, int bayer_bggr, uint8_t* d_bgr, size_t bgr_pitch) {
  const int nc = L.components;
  const bool demosaic = bayer_bggr != 0 && nc == 1;
  if (!(nc == 3 || demosaic))
    throw InvalidArg{"jpeg input: a 3-component stream, or a 1-component bayer stream, is required", UVO_ERR_UNSUPPORTED};
  IdctArgs ia;
  ColorArgs ca;
  fill_args(L, d_entries, d_first, d_count, d_planes, nc == 3 ? d_bgr : nullptr, bgr_pitch, ia, ca);
  UVO_KERNEL(c, "k_jpeg_idct");
  k_jpeg_idct<<<div_up(ia.total_blocks, IDCT_BLOCKS), IDCT_THREADS, 0, c.stream>>>(ia);
  UVO_LAUNCH_CHECK(c);
  if (demosaic) {
    launch_demosaic_bggr(c, ia.c[0].plane, (size_t)L.blocks_x[0] * 8, L.width, L.height, d_bgr, bgr_pitch);
  } else {
    UVO_KERNEL(c, "k_jpeg_color");
    k_jpeg_color<<<dim3(div_up(L.width, COLOR_TX), div_up(L.height, COLOR_TY)), COLOR_TX * COLOR_TY, 0, c.stream>>>(ca);
    UVO_LAUNCH_CHECK(c);
  }
}

void jpeg_host_decode_sparse(const uint8_t* jpeg, size_t len, uint32_t* entries, size_t capacity, uint32_t* first,
                             uint8_t* count, size_t* n_entries, uvo_jpeg_layout* layout) {
  Parser H;
  H.run(jpeg, len, nullptr);
  const size_t nb = (size_t)H.L.coeff_total / 64;
  memset(first, 0, sizeof(uint32_t) * nb);  // blocks no scan visits stay empty
  memset(count, 0, nb);
  Parser P;
  P.run_sparse(jpeg, len, entries, capacity, first, count);
  *n_entries = P.sink.n;
  *layout = P.L;
}

static inline size_t up256(size_t n) { return (n + 255) & ~(size_t)255; }

size_t jpeg_gpu_host_bytes(size_t jpeg_len) { return up256(sizeof(JhPlan)) + up256(jpeg_len + 64); }

bool jpeg_gpu_prepare(const uint8_t* jpeg, size_t len, uint8_t* pinned, size_t pinned_cap, JpegGpuJob* job) {
  if (pinned_cap < jpeg_gpu_host_bytes(len)) throw InvalidArg{"jpeg: staging buffer too small", UVO_ERR_CAPACITY};
  JhPlan* plan = (JhPlan*)pinned;
  uint8_t* scan = pinned + up256(sizeof(JhPlan));
  GpuPlanOut o;
  if (!build_gpu_plan(jpeg, len, plan, scan, pinned_cap - up256(sizeof(JhPlan)), &o)) return false;
  job->L = o.L;
  job->upload_bytes = up256(sizeof(JhPlan)) + ((o.scan_bytes + 16 + 3) & ~(size_t)3);
  return true;
}

struct GpuLayout {  // offsets inside the device buffer of one image
  size_t first, entries, count, first_scan, info, exits, counts, dc_part, total;
};
static int jh_blocks(size_t scan_bytes) {  // blocks of the decoder launch an image of this scan length needs
  const size_t n_sub = (scan_bytes * 8 + JH_SUB_BITS - 1) / JH_SUB_BITS;
  return (int)std::max<size_t>((n_sub + JH_BLOCK - 1) / JH_BLOCK, 1);
}
static GpuLayout gpu_layout(const uvo_jpeg_layout& L, size_t upload_bytes) {
  const size_t nb = (size_t)L.coeff_total / 64;
  GpuLayout g;
  size_t o = up256(upload_bytes);
  g.first = o;
  o += up256(nb * 4);
  g.entries = o;
  o += up256((size_t)L.coeff_total * 4);
  g.count = o;
  o += up256(nb);
  g.first_scan = o;
  o += up256((nb + 1) * 4);
  g.info = o;
  o += 512;
  // scratch of the parallel decode, sized from the (upper bound of the) scan length
  const size_t n_thr = (size_t)jh_blocks(upload_bytes) * JH_BLOCK + JH_BLOCK;
  g.exits = o;
  o += up256(2 * n_thr * sizeof(uint2));
  g.counts = o;
  o += up256(n_thr * sizeof(int2));
  g.dc_part = o;
  o += up256(4 * n_thr * sizeof(int4));  // the DC pass runs on the threads of the widest image of the launch
  g.total = o;
  return g;
}

size_t jpeg_gpu_device_bytes(const uvo_jpeg_layout& L, size_t upload_bytes) { return gpu_layout(L, upload_bytes).total; }

// The decoder's grid barrier.  Default: the kernel's own arrival counter on an ordinary launch -- cooperative launches
// do not overlap one another, which would cap compressed input at 1 / (decode time of one pair) however many frames
// are enqueued.  UVO_JPEG_COOPERATIVE=1 selects cudaLaunchCooperativeKernel + cooperative_groups' grid sync instead.
static bool jh_use_cooperative() {
  static const bool v = [] {
    const char* e = getenv("UVO_JPEG_COOPERATIVE");
    return e && e[0] == '1';
  }();
  return v;
}
// blocks of k_jpeg_huff the device holds at once (every block of a launch spins at the barriers, so all must be resident)
static int jh_resident_blocks(Ctx& c) {
  static std::mutex mu;
  static int cached[64] = {};
  std::lock_guard<std::mutex> lk(mu);
  int& v = cached[c.device & 63];
  if (!v) {
    int per_sm = 0, sms = 0;
    UVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_jpeg_huff, JH_BLOCK, 0));
    UVO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device));
    v = std::max(per_sm * sms, 1);
  }
  return v;
}
// how many decoder launches of this width may be in flight at once: together they stay inside half of the device, so
// whatever order the hardware dispatches their blocks in, every launch becomes fully resident
int jpeg_gpu_max_concurrent(Ctx& c, int n, const JpegGpuJob* jobs) {
  if (jh_use_cooperative()) return 1 << 20;
  int blocks = 1;
  for (int i = 0; i < n; i++) blocks = std::max(blocks, jh_blocks(jobs[i].upload_bytes));
  return std::max(jh_resident_blocks(c) / 2 / (blocks * n), 1);
}

__global__ void k_jpeg_huff_status(const int* info0, const int* info1, int* status) {
  status[0] = info0[1];
  status[1] = info0[2];
  if (info1) {
    status[2] = info1[1];
    status[3] = info1[2];
  }
  for (int i = 0; i < 64; i++) status[8 + i] = info0[15 + i];  // phase stamps (diagnostics)
}

void jpeg_gpu_launch(Ctx& c, int n, const JpegGpuJob* jobs, uint8_t* const* d_buf, uint8_t* const* d_planes,
                     int bayer_bggr, uint8_t* const* d_bgr, size_t bgr_pitch, int* d_status) {
  UVO_REQUIRE(n == 1 || n == 2, "jpeg_gpu_launch: one or two images");
  JhArgs a{};
  GpuLayout g[2];
  for (int i = 0; i < n; i++) {
    g[i] = gpu_layout(jobs[i].L, jobs[i].upload_bytes);
    JhImage& im = a.im[i];
    im.plan = (const JhPlan*)d_buf[i];
    im.scan = (const uint32_t*)(d_buf[i] + up256(sizeof(JhPlan)));
    im.first = (uint32_t*)(d_buf[i] + g[i].first);
    im.entries = (uint32_t*)(d_buf[i] + g[i].entries);
    im.count = d_buf[i] + g[i].count;
    im.first_scan = (uint32_t*)(d_buf[i] + g[i].first_scan);
    im.info = (int*)(d_buf[i] + g[i].info);
    im.exits = (uint2*)(d_buf[i] + g[i].exits);
    im.counts = (int2*)(d_buf[i] + g[i].counts);
    im.dc_part = (int4*)(d_buf[i] + g[i].dc_part);
  }
  // the launch is as wide as its longest scan needs; the DC scratch above allows a factor 4 between the two images
  int blocks = 1;
  for (int i = 0; i < n; i++) blocks = std::max(blocks, jh_blocks(jobs[i].upload_bytes));
  for (int i = 0; i < n; i++)
    UVO_REQUIRE(blocks <= 4 * jh_blocks(jobs[i].upload_bytes) + 4, "jpeg pair: the two scans differ too much in length");
  a.flag = (int*)(d_buf[0] + g[0].info) + 8;
  const bool cooperative = jh_use_cooperative();
  a.bar = cooperative ? nullptr : (unsigned*)(a.flag + 4);
  UVO_CUDA(cudaMemsetAsync(a.flag, 0, 6 * sizeof(int), c.stream));
  UVO_KERNEL(c, "k_jpeg_huff");
  if (cooperative) {
    void* kargs[1] = {(void*)&a};
    UVO_CUDA(cudaLaunchCooperativeKernel((const void*)k_jpeg_huff, dim3(blocks, n), dim3(JH_BLOCK), kargs, 0, c.stream));
  } else {
    UVO_REQUIRE(blocks * n <= jh_resident_blocks(c), "jpeg: the scan needs more decoder blocks than the device holds");
    k_jpeg_huff<<<dim3(blocks, n), JH_BLOCK, 0, c.stream>>>(a);
  }
  UVO_LAUNCH_CHECK(c);
  if (d_status) {
    k_jpeg_huff_status<<<1, 1, 0, c.stream>>>(a.im[0].info, n > 1 ? a.im[1].info : nullptr, d_status);
    UVO_CUDA(cudaGetLastError());
  }
  for (int i = 0; i < n; i++)
    transform_ptrs(c, jobs[i].L, a.im[i].first, a.im[i].entries, a.im[i].count, d_planes[i], bayer_bggr, d_bgr[i],
                   bgr_pitch);
}

}  // namespace uvo

extern "C" {

int uvo_jpeg_info(const uint8_t* jpeg, size_t len, uvo_jpeg_layout* layout) {
  if (!layout) return UVO_ERR_INVALID;
  return guarded(nullptr, [&] {
    Parser P;
    P.run(jpeg, len, nullptr);
    *layout = P.L;
  });
}

int uvo_jpeg_entropy_decode(const uint8_t* jpeg, size_t len, int16_t* coeffs_host, size_t capacity,
                            uvo_jpeg_layout* layout) {
  if (!layout || !coeffs_host) return UVO_ERR_INVALID;
  return guarded(nullptr, [&] {
    Parser H;
    H.run(jpeg, len, nullptr);
    if ((size_t)H.L.coeff_total > capacity)
      throw InvalidArg{"uvo_jpeg_entropy_decode: coefficient buffer too small", UVO_ERR_CAPACITY};
    memset(coeffs_host, 0, sizeof(int16_t) * (size_t)H.L.coeff_total);
    Parser P;
    P.run(jpeg, len, coeffs_host);
    *layout = P.L;
  });
}

int uvo_jpeg_entropy_decode_sparse(const uint8_t* jpeg, size_t len, uint32_t* entries_host, size_t capacity,
                                   uint32_t* block_first_host, uint8_t* block_count_host, size_t* n_entries,
                                   uvo_jpeg_layout* layout) {
  if (!layout || !entries_host || !block_first_host || !block_count_host || !n_entries) return UVO_ERR_INVALID;
  return guarded(nullptr, [&] {
    Parser H;
    H.run(jpeg, len, nullptr);
    const size_t nb = (size_t)H.L.coeff_total / 64;
    memset(block_first_host, 0, sizeof(uint32_t) * nb);  // blocks no scan visits stay empty
    memset(block_count_host, 0, nb);
    Parser P;
    P.run_sparse(jpeg, len, entries_host, capacity, block_first_host, block_count_host);
    *n_entries = P.sink.n;
    *layout = P.L;
  });
}

// shared body of uvo_jpeg_decode (out on the host) and uvo_jpeg_decode_device (out in device memory, no copy back)
static void jpeg_decode_impl(uvo_ctx* ctx, const uint8_t* jpeg, size_t len, int bayer_bggr, uint8_t* out,
                             size_t out_pitch, size_t out_capacity, bool out_on_device, int* width, int* height,
                             int* channels) {
  UVO_REQUIRE(jpeg && out && width && height && channels, "uvo_jpeg_decode: bad argument");
  Ctx& c = ctx->c;
  UVO_CUDA(cudaSetDevice(c.device));
  ctx->jpeg_last_route = 0;
  ctx->jpeg_last_rounds = 0;
  if (ctx->jpeg_gpu_entropy) {
    // Huffman decoding on the GPU: only the (unstuffed) scan bytes and the table plan cross PCIe
    UVO_CUDA(cudaStreamSynchronize(c.stream));  // the pinned buffer may still feed the previous call's copy
    const size_t hb = jpeg_gpu_host_bytes(len);
    ctx->jpeg_coef.ensure((hb + 3) / 4);
    JpegGpuJob job;
    if (jpeg_gpu_prepare(jpeg, len, (uint8_t*)ctx->jpeg_coef.p, hb, &job)) {
      const uvo_jpeg_layout& L = job.L;
      const int W = L.width, Hh = L.height, nc = L.components;
      const bool demosaic = bayer_bggr != 0 && nc == 1;
      if (nc == 3 || demosaic) {
        if (demosaic && (W < 3 || Hh < 3))
          throw InvalidArg{"uvo_jpeg_decode: a bayer image needs w, h >= 3", UVO_ERR_INVALID};
        if (out_pitch < (size_t)W * 3 || out_capacity < out_pitch * (size_t)(Hh - 1) + (size_t)W * 3)
          throw InvalidArg{"uvo_jpeg_decode: output buffer too small (see uvo_jpeg_info)", UVO_ERR_CAPACITY};
        StageScratch& s = ctx->scratch;
        s.bytes_a.ensure(jpeg_gpu_device_bytes(L, job.upload_bytes));
        s.bytes_b.ensure(jpegk::plane_bytes(L));
        s.bytes_d.ensure(512);
        size_t dpitch = out_pitch;
        uint8_t* d_out = out;
        if (!out_on_device) {
          dpitch = ((size_t)3 * W + 15) & ~(size_t)15;
          s.bytes_c.ensure(dpitch * Hh);
          d_out = s.bytes_c.get();
        }
        UVO_CUDA(cudaMemcpyAsync(s.bytes_a.get(), ctx->jpeg_coef.p, job.upload_bytes, cudaMemcpyHostToDevice, c.stream));
        uint8_t* d_buf = s.bytes_a.get();
        uint8_t* d_planes = s.bytes_b.get();
        jpeg_gpu_launch(c, 1, &job, &d_buf, &d_planes, bayer_bggr, &d_out, dpitch, (int*)s.bytes_d.get());
        if (!out_on_device)
          UVO_CUDA(cudaMemcpy2DAsync(out, out_pitch, d_out, dpitch, (size_t)3 * W, Hh, cudaMemcpyDeviceToHost, c.stream));
        int status[72] = {0};
        UVO_CUDA(cudaMemcpyAsync(status, s.bytes_d.get(), sizeof(status), cudaMemcpyDeviceToHost, c.stream));
        UVO_CUDA(cudaStreamSynchronize(c.stream));
        ctx->jpeg_last_route = 1;
        ctx->jpeg_last_rounds = status[1];
        memcpy(ctx->jpeg_stamps, status + 8, sizeof(ctx->jpeg_stamps));
        if (status[0]) throw InvalidArg{"jpeg: corrupt or truncated entropy-coded data", UVO_ERR_INVALID};
        *width = W;
        *height = Hh;
        *channels = 3;
        return;
      }
    }
  }
  Parser H;
  H.run(jpeg, len, nullptr);
  const size_t total = (size_t)H.L.coeff_total;
  const int W = H.L.width, Hh = H.L.height, nc = H.L.components;
  // a bayer-format message is a 1-component stream holding the BGGR mosaic: from_ros_to_cv_image runs
  // cvtColor(COLOR_BayerBGGR2BGR) on the decoded image (math_utility.cpp:161-164)
  const bool demosaic = bayer_bggr != 0 && nc == 1;
  if (demosaic && (W < 3 || Hh < 3)) throw InvalidArg{"uvo_jpeg_decode: a bayer image needs w, h >= 3", UVO_ERR_INVALID};
  const int och = (nc == 3 || demosaic) ? 3 : 1;  // channels of the output image
  if (out_pitch < (size_t)W * och || out_capacity < out_pitch * (size_t)(Hh - 1) + (size_t)W * och)
    throw InvalidArg{"uvo_jpeg_decode: output buffer too small (see uvo_jpeg_info)", UVO_ERR_CAPACITY};
  // host: entropy decoding into pinned memory, as the sparse form (one 32-bit entry per non-zero coefficient + a
  // (first, count) pair per block); one pinned buffer: [first: nb x u32][entries: <= total x u32][count: nb x u8]
  UVO_CUDA(cudaStreamSynchronize(c.stream));  // the pinned buffer may still feed the previous call's copy
  const size_t nb = total / 64;
  ctx->jpeg_coef.ensure(nb + total + (nb + 3) / 4);
  uint32_t* h_first = ctx->jpeg_coef.p;
  uint32_t* h_entries = h_first + nb;
  uint8_t* h_count = (uint8_t*)(h_entries + total);
  memset(h_first, 0, sizeof(uint32_t) * nb);  // blocks no scan visits (padding of a non-interleaved scan) stay empty
  memset(h_count, 0, nb);
  Parser P;
  P.run_sparse(jpeg, len, h_entries, total, h_first, h_count);
  const uvo_jpeg_layout& L = P.L;
  const size_t ne = P.sink.n;
  // device: the same three arrays (only the used entries travel), component planes, output
  StageScratch& s = ctx->scratch;
  s.bytes_a.ensure(sizeof(uint32_t) * (nb + std::max<size_t>(ne, 1)) + nb);
  uint32_t* d_first = (uint32_t*)s.bytes_a.get();
  uint32_t* d_entries = d_first + nb;
  uint8_t* d_count = (uint8_t*)(d_entries + std::max<size_t>(ne, 1));
  s.bytes_b.ensure(jpegk::plane_bytes(L));
  // the colour kernel writes straight into the caller's device buffer, or into a staging image that is copied back
  size_t dpitch = out_pitch;
  uint8_t* d_out = out;
  if (!out_on_device && och == 3) {
    dpitch = ((size_t)3 * W + 15) & ~(size_t)15;
    s.bytes_c.ensure(dpitch * Hh);
    d_out = s.bytes_c.get();
  }
  UVO_CUDA(cudaMemcpyAsync(d_first, h_first, sizeof(uint32_t) * (nb + ne), cudaMemcpyHostToDevice, c.stream));
  UVO_CUDA(cudaMemcpyAsync(d_count, h_count, nb, cudaMemcpyHostToDevice, c.stream));
  IdctArgs ia;
  ColorArgs ca;
  fill_args(L, d_entries, d_first, d_count, s.bytes_b.get(), nc == 3 ? d_out : nullptr, dpitch, ia, ca);
  UVO_KERNEL(c, "k_jpeg_idct");
  k_jpeg_idct<<<div_up(ia.total_blocks, IDCT_BLOCKS), IDCT_THREADS, 0, c.stream>>>(ia);
  UVO_LAUNCH_CHECK(c);
  if (demosaic) {  // the luminance plane is the mosaic
    launch_demosaic_bggr(c, ia.c[0].plane, (size_t)L.blocks_x[0] * 8, W, Hh, d_out, dpitch);
    if (!out_on_device)
      UVO_CUDA(cudaMemcpy2DAsync(out, out_pitch, d_out, dpitch, (size_t)3 * W, Hh, cudaMemcpyDeviceToHost, c.stream));
  } else if (nc == 1) {  // the luminance plane is the image
    UVO_CUDA(cudaMemcpy2DAsync(out, out_pitch, ia.c[0].plane, (size_t)L.blocks_x[0] * 8, W, Hh,
                               out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c.stream));
  } else {
    UVO_KERNEL(c, "k_jpeg_color");
    k_jpeg_color<<<dim3(div_up(W, COLOR_TX), div_up(Hh, COLOR_TY)), COLOR_TX * COLOR_TY, 0, c.stream>>>(ca);
    UVO_LAUNCH_CHECK(c);
    if (!out_on_device)
      UVO_CUDA(cudaMemcpy2DAsync(out, out_pitch, d_out, dpitch, (size_t)3 * W, Hh, cudaMemcpyDeviceToHost, c.stream));
  }
  if (!out_on_device) UVO_CUDA(cudaStreamSynchronize(c.stream));  // device output: ordered on the context stream
  *width = W;
  *height = Hh;
  *channels = och;
}

int uvo_jpeg_decode(uvo_ctx* ctx, const uint8_t* jpeg, size_t len, int bayer_bggr, uint8_t* out_host,
                    size_t out_pitch, size_t out_capacity, int* width, int* height, int* channels) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    jpeg_decode_impl(ctx, jpeg, len, bayer_bggr, out_host, out_pitch, out_capacity, false, width, height, channels);
  });
}

int uvo_jpeg_gpu_entropy(uvo_ctx* ctx, int enable, int* last_route, int* last_rounds) {
  if (!ctx) return UVO_ERR_INVALID;
  if (enable >= 0) ctx->jpeg_gpu_entropy = enable != 0;
  if (last_route) *last_route = ctx->jpeg_last_route;
  if (last_rounds) *last_rounds = ctx->jpeg_last_rounds;
  return UVO_OK;
}

int uvo_jpeg_gpu_entropy_stamps(uvo_ctx* ctx, int64_t stamps[24], int* count) {
  if (!ctx || !stamps || !count) return UVO_ERR_INVALID;
  *count = std::min(ctx->jpeg_stamps[0], 24);
  memcpy(stamps, ctx->jpeg_stamps + 1, sizeof(int64_t) * 24);
  return UVO_OK;
}

size_t uvo_jpeg_gpu_staging_bytes(size_t len) { return jpeg_gpu_host_bytes(len); }

int uvo_jpeg_gpu_plan(const uint8_t* jpeg, size_t len, uint8_t* staging, size_t staging_bytes, int* qualifies,
                      size_t* upload_bytes, uint32_t* scan_bits, uvo_jpeg_layout* layout) {
  if (!staging || !qualifies) return UVO_ERR_INVALID;
  *qualifies = 0;
  return guarded(nullptr, [&] {
    JpegGpuJob job{};
    if (!jpeg_gpu_prepare(jpeg, len, staging, staging_bytes, &job)) return;
    *qualifies = 1;
    if (upload_bytes) *upload_bytes = job.upload_bytes;
    if (scan_bits) *scan_bits = ((const JhPlan*)staging)->total_bits;
    if (layout) *layout = job.L;
  });
}

int uvo_jpeg_decode_device(uvo_ctx* ctx, const uint8_t* jpeg, size_t len, int bayer_bggr, uint8_t* out_dev,
                           size_t out_pitch, size_t out_capacity, int* width, int* height, int* channels) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    jpeg_decode_impl(ctx, jpeg, len, bayer_bggr, out_dev, out_pitch, out_capacity, true, width, height, channels);
  });
}

}  // extern "C"
