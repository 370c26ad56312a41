// jpeg_kernels.cuh -- the per-thread bodies of k_jpeg_idct and k_jpeg_color (csrc/jpeg.cu) and the argument set-up of
// uvo_jpeg_decode, written as __host__ __device__ functions of (block index, thread index) so that the test harness
// tests/emu/jpeg_emu.cu can execute exactly this code on the CPU over the whole launch grid (tests/test_jpeg_emu.py).
// The library itself only ever runs them inside the two kernels.
// Reference: the cv::imdecode inside from_ros_to_cv_image, math_utility.cpp:154-173 (libjpeg-turbo: jidctint.c,
// jdsample.c, jdcolor.c).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "uvo_c.h"

#ifdef __CUDACC__
#define UVO_HD __host__ __device__ __forceinline__
#else
#define UVO_HD inline
#endif
#ifdef __CUDA_ARCH__
#define UVO_UNROLL _Pragma("unroll")
#else
#define UVO_UNROLL
#endif

namespace uvo {
namespace jpegk {

struct IdctComp {
  uint8_t* plane;  // (blocks_y * 8) rows, pitch = blocks_x * 8
  int blocks_x, n_blocks, first;  // first: index of the component's first block in the launch
  uint16_t quant[64];
};
struct IdctArgs {
  IdctComp c[3];
  int n_comp, total_blocks;
  // sparse coefficients: per block (launch order: component-major, row-major) its first entry and entry count; an
  // entry is (natural index << 16) | (value & 0xffff), non-zero coefficients only
  const uint32_t* entries;
  const uint32_t* first;
  const uint8_t* count;
};
struct ColorComp {
  const uint8_t* plane;
  int pitch, dw, dh;  // real sample counts of the component
  int hexp, vexp;     // expansion to full resolution
};
struct ColorArgs {
  ColorComp c[3];
  int w, h;
  uint8_t* out;
  size_t out_pitch;
};

constexpr int IDCT_THREADS = 256, IDCT_BLOCKS = IDCT_THREADS / 8;  // JPEG blocks per thread block
constexpr int WS_STRIDE = 72;  // ints of workspace per JPEG block: 8 rows 9 apart (rows on different banks in pass 2)
constexpr int TILE_STRIDE = 64;  // int16 coefficients per JPEG block in the dense tile rebuilt in shared memory
constexpr int COLOR_TX = 64, COLOR_TY = 4;

UVO_HD int imin(int a, int b) { return a < b ? a : b; }
UVO_HD int imax(int a, int b) { return a > b ? a : b; }
UVO_HD int jdescale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
// libjpeg's post-IDCT range-limit table, indexed with the value masked to 10 bits
UVO_HD unsigned jrange(int v) {
  const int i = v & 1023;
  return (unsigned)(i < 128 ? i + 128 : i < 512 ? 255 : i < 896 ? 0 : i - 896);
}

// one 1-D pass of jpeg_idct_islow on eight inputs (CONST_BITS 13); results not yet descaled
UVO_HD void islow_1d(const int in[8], int out[8]) {
  constexpr int F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633,
                F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * F_0_541;
  int tmp2 = z1 + z3 * (-F_1_847), tmp3 = z1 + z2 * F_0_765;
  int tmp0 = (in[0] + in[4]) * 8192, tmp1 = (in[0] - in[4]) * 8192;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7];
  tmp1 = in[5];
  tmp2 = in[3];
  tmp3 = in[1];
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
  out[0] = tmp10 + tmp3;
  out[7] = tmp10 - tmp3;
  out[1] = tmp11 + tmp2;
  out[6] = tmp11 - tmp2;
  out[2] = tmp12 + tmp1;
  out[5] = tmp12 - tmp1;
  out[3] = tmp13 + tmp0;
  out[4] = tmp13 - tmp0;
}

// which JPEG block a thread works on, and within which component (eight threads per block, t = column / row)
struct IdctWho {
  bool live;
  int ci, local, t, g;
};
UVO_HD IdctWho idct_who(const IdctArgs& a, int block_idx, int thread_idx) {
  IdctWho w;
  w.t = thread_idx & 7;
  w.g = thread_idx >> 3;
  const int blk = block_idx * IDCT_BLOCKS + w.g;
  w.live = blk < a.total_blocks;
  w.ci = 0;
  if (w.live)
    while (w.ci + 1 < a.n_comp && blk >= a.c[w.ci + 1].first) w.ci++;
  w.local = blk - a.c[w.ci].first;
  return w;
}

// stage 0: the eight threads of a block clear its dense 8x8 tile (thread t: row t)
UVO_HD void idct_clear(const IdctArgs& a, int block_idx, int thread_idx, int16_t* tile /* IDCT_BLOCKS x TILE_STRIDE */) {
  const IdctWho w = idct_who(a, block_idx, thread_idx);
  if (!w.live) return;
  UVO_UNROLL
  for (int k = 0; k < 8; k++) tile[w.g * TILE_STRIDE + 8 * w.t + k] = 0;
}

// stage 1: scatter the block's non-zero coefficients into the tile (thread t: entries t, t + 8, ...); positions are
// distinct within a block, so the writes do not collide
UVO_HD void idct_scatter(const IdctArgs& a, int block_idx, int thread_idx, int16_t* tile) {
  const IdctWho w = idct_who(a, block_idx, thread_idx);
  if (!w.live) return;
  const int blk = block_idx * IDCT_BLOCKS + w.g;
  const uint32_t* e = a.entries + a.first[blk];
  const int n = a.count[blk];
  for (int i = w.t; i < n; i += 8) {
    const uint32_t v = e[i];
    tile[w.g * TILE_STRIDE + (int)((v >> 16) & 63)] = (int16_t)(v & 0xffffu);
  }
}

// pass 1: thread t transforms column t of its block (dequantising on the way in) into the block's workspace.  The
// zero-AC shortcut of the CPU code is not needed: the full column pass yields dc * 4 exactly in that case.
UVO_HD void idct_pass1(const IdctArgs& a, int block_idx, int thread_idx, const int16_t* tile,
                       int* ws /* IDCT_BLOCKS x WS_STRIDE */) {
  const IdctWho w = idct_who(a, block_idx, thread_idx);
  if (!w.live) return;
  const IdctComp& c = a.c[w.ci];
  const int16_t* in = tile + w.g * TILE_STRIDE;
  int v[8], o[8];
UVO_UNROLL
  for (int r = 0; r < 8; r++) v[r] = (int)in[8 * r + w.t] * (int)c.quant[8 * r + w.t];
  islow_1d(v, o);
UVO_UNROLL
  for (int r = 0; r < 8; r++) ws[w.g * WS_STRIDE + 9 * r + w.t] = jdescale(o[r], 11);  // CONST_BITS - PASS1_BITS
}

// pass 2: thread t transforms row t of the workspace and stores the eight samples of that row
UVO_HD void idct_pass2(const IdctArgs& a, int block_idx, int thread_idx, const int* ws) {
  const IdctWho w = idct_who(a, block_idx, thread_idx);
  if (!w.live) return;
  const IdctComp& c = a.c[w.ci];
  int v[8], o[8];
UVO_UNROLL
  for (int k = 0; k < 8; k++) v[k] = ws[w.g * WS_STRIDE + 9 * w.t + k];
  islow_1d(v, o);
  unsigned lo = 0, hi = 0;
UVO_UNROLL
  for (int k = 0; k < 4; k++) {
    lo |= jrange(jdescale(o[k], 18)) << (8 * k);  // CONST_BITS + PASS1_BITS + 3
    hi |= jrange(jdescale(o[4 + k], 18)) << (8 * k);
  }
  const int by = w.local / c.blocks_x, bx = w.local - by * c.blocks_x;
  uint8_t* dst = c.plane + ((size_t)(by * 8 + w.t) * c.blocks_x + bx) * 8;  // 8-byte aligned: pitch = blocks_x * 8
  uint32_t two[2] = {lo, hi};
#ifdef __CUDA_ARCH__
  *reinterpret_cast<uint2*>(dst) = make_uint2(two[0], two[1]);
#else
  memcpy(dst, two, 8);  // little-endian host, as the device
#endif
}

// the component's value at full-resolution pixel (x, y): jdsample.c's fancy filters, replication otherwise
UVO_HD int jsample(const ColorComp& c, int x, int y) {
  if (c.hexp == 1 && c.vexp == 1) return c.plane[(size_t)y * c.pitch + x];
  const bool fancy = c.dw > 2;
  if (c.hexp == 2 && c.vexp == 1 && fancy) {  // h2v1_fancy_upsample
    const uint8_t* p = c.plane + (size_t)y * c.pitch;
    const int i = x >> 1;
    if (x & 1) return i == c.dw - 1 ? p[i] : (p[i] * 3 + p[i + 1] + 2) >> 2;
    return i == 0 ? p[0] : (p[i] * 3 + p[i - 1] + 1) >> 2;
  }
  if (c.hexp == 2 && c.vexp == 2 && fancy) {  // h2v2_fancy_upsample
    const int r = y >> 1, r1 = (y & 1) ? imin(r + 1, c.dh - 1) : imax(r - 1, 0);
    const uint8_t* p0 = c.plane + (size_t)r * c.pitch;
    const uint8_t* p1 = c.plane + (size_t)r1 * c.pitch;
    const int i = x >> 1;
    const int cs = p0[i] * 3 + p1[i];
    if (x & 1) return i == c.dw - 1 ? (cs * 4 + 7) >> 4 : (cs * 3 + p0[i + 1] * 3 + p1[i + 1] + 7) >> 4;
    return i == 0 ? (cs * 4 + 8) >> 4 : (cs * 3 + p0[i - 1] * 3 + p1[i - 1] + 8) >> 4;
  }
  if (c.hexp == 1 && c.vexp == 2) {  // h1v2_fancy_upsample
    const int r = y >> 1, r1 = (y & 1) ? imin(r + 1, c.dh - 1) : imax(r - 1, 0);
    return (c.plane[(size_t)r * c.pitch + x] * 3 + c.plane[(size_t)r1 * c.pitch + x] + ((y & 1) ? 2 : 1)) >> 2;
  }
  return c.plane[(size_t)(y / c.vexp) * c.pitch + x / c.hexp];
}

UVO_HD unsigned jclamp(int v) { return (unsigned)imin(imax(v, 0), 255); }

// one output pixel: jdcolor.c ycc_rgb_convert (SCALEBITS 16), written B, G, R.  Grid: COLOR_TX x COLOR_TY pixels per
// thread block of COLOR_TX * COLOR_TY threads.
UVO_HD void color_thread(const ColorArgs& a, int block_x, int block_y, int thread_idx) {
  const int x = block_x * COLOR_TX + (thread_idx % COLOR_TX), y = block_y * COLOR_TY + thread_idx / COLOR_TX;
  if (x >= a.w || y >= a.h) return;
  const int Y = jsample(a.c[0], x, y), cb = jsample(a.c[1], x, y) - 128, cr = jsample(a.c[2], x, y) - 128;
  constexpr int FIX_1_402 = 91881, FIX_1_772 = 116130, FIX_0_714 = 46802, FIX_0_344 = 22554;
  const int r = Y + ((FIX_1_402 * cr + 32768) >> 16);
  const int g = Y + ((-FIX_0_344 * cb + 32768 - FIX_0_714 * cr) >> 16);
  const int b = Y + ((FIX_1_772 * cb + 32768) >> 16);
  uint8_t* o = a.out + (size_t)y * a.out_pitch + 3 * (size_t)x;
  o[0] = (uint8_t)jclamp(b);
  o[1] = (uint8_t)jclamp(g);
  o[2] = (uint8_t)jclamp(r);
}

// argument set-up shared by uvo_jpeg_decode and the harness: the sparse coefficient arrays, sample planes packed one
// after the other from `planes`
inline size_t plane_bytes(const uvo_jpeg_layout& L) {
  size_t n = 0;
  for (int k = 0; k < L.components; k++) n += (size_t)L.blocks_x[k] * L.blocks_y[k] * 64;
  return n;
}
inline void fill_args(const uvo_jpeg_layout& L, const uint32_t* entries, const uint32_t* first, const uint8_t* count,
                      uint8_t* planes, uint8_t* out, size_t out_pitch, IdctArgs& ia, ColorArgs& ca) {
  memset(&ia, 0, sizeof(ia));
  memset(&ca, 0, sizeof(ca));
  const int nc = L.components;
  ia.n_comp = nc;
  int first_block = 0, hmax = 1, vmax = 1;
  size_t off = 0;
  for (int k = 0; k < nc; k++) {
    hmax = L.h_samp[k] > hmax ? L.h_samp[k] : hmax;
    vmax = L.v_samp[k] > vmax ? L.v_samp[k] : vmax;
  }
  for (int k = 0; k < nc; k++) {
    ia.c[k].plane = planes + off;
    ia.c[k].blocks_x = L.blocks_x[k];
    ia.c[k].n_blocks = L.blocks_x[k] * L.blocks_y[k];
    ia.c[k].first = first_block;
    memcpy(ia.c[k].quant, L.quant[k], sizeof(ia.c[k].quant));
    first_block += ia.c[k].n_blocks;
    off += (size_t)ia.c[k].n_blocks * 64;
    if (k < 3) {
      ca.c[k].plane = ia.c[k].plane;
      ca.c[k].pitch = L.blocks_x[k] * 8;
      ca.c[k].dw = L.samples_x[k];
      ca.c[k].dh = L.samples_y[k];
      ca.c[k].hexp = hmax / L.h_samp[k];
      ca.c[k].vexp = vmax / L.v_samp[k];
    }
  }
  ia.total_blocks = first_block;
  ia.entries = entries;
  ia.first = first;
  ia.count = count;
  ca.w = L.width;
  ca.h = L.height;
  ca.out = out;
  ca.out_pitch = out_pitch;
}

}  // namespace jpegk
}  // namespace uvo
