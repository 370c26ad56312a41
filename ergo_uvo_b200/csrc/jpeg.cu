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
#include <algorithm>
#include <cstring>

#include "capi_internal.cuh"
#include "jpeg.cuh"
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

void jpeg_launch_transform(Ctx& c, const uvo_jpeg_layout& L, size_t n_entries, uint8_t* d_sparse, uint8_t* d_planes,
                           int bayer_bggr, uint8_t* d_bgr, size_t bgr_pitch) {
  const size_t nb = (size_t)L.coeff_total / 64;
  const int nc = L.components;
  const bool demosaic = bayer_bggr != 0 && nc == 1;
  if (!(nc == 3 || demosaic))
    throw InvalidArg{"jpeg input: a 3-component stream, or a 1-component bayer stream, is required", UVO_ERR_UNSUPPORTED};
  uint32_t* d_first = (uint32_t*)d_sparse;
  uint32_t* d_entries = d_first + nb;
  uint8_t* d_count = (uint8_t*)(d_entries + std::max<size_t>(n_entries, 1));
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

int uvo_jpeg_decode_device(uvo_ctx* ctx, const uint8_t* jpeg, size_t len, int bayer_bggr, uint8_t* out_dev,
                           size_t out_pitch, size_t out_capacity, int* width, int* height, int* channels) {
  if (!ctx) return UVO_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    jpeg_decode_impl(ctx, jpeg, len, bayer_bggr, out_dev, out_pitch, out_capacity, true, width, height, channels);
  });
}

}  // extern "C"
