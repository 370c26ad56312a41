// jpeg_huff.cuh -- Huffman decoding of a baseline JPEG scan ON THE GPU (the serial half of the cv::imdecode inside
// from_ros_to_cv_image, math_utility.cpp:154-173).  The entropy-coded segment is one bit stream with no markers to
// restart at, so the decode is made parallel by SELF-SYNCHRONISATION: the scan is cut into one sub-sequence per
// thread; every thread first decodes its sub-sequence from a guessed state (block position in the MCU 0, coefficient
// index 0), then, round after round, restarts from the state its left neighbour's decode implies at the boundary,
// until no exit state changes.  A Huffman decoder started at a wrong bit falls into step with the true one after a
// few hundred bits (measured: tools/jpeg_sync_probe.py -- median 450 bits at q75, never later than 6.4 kbit), so a few
// rounds reach the fixed point, which is the sequential decode.  Two more passes count and then write the
// coefficients in the sparse form k_jpeg_idct takes; the DC predictions are prefix sums per component.
//
// One launch per image pair (grid-wide barriers between the phases): one thread per sub-sequence of 1024 bits, blocks of 64 threads spread over
// the SMs, grid-wide barriers between the rounds and the passes (a first version ran one 1024-thread block per image:
// correct, but 1.5 ms per 1280x1024 frame -- 32 warps of divergent table walks issue-bound on one SM).
// Supported here: one interleaved scan (every component in it), no restart interval -- what cv::imencode and the
// cameras' encoders produce.  Anything else takes the host decoder.
#pragma once
#include <stdint.h>

namespace uvo {

constexpr int JH_FAST_BITS = 10;  // look-up bits (the host tables are built for 10)
#include <cuda_runtime.h>
constexpr int JH_BLOCK = 64;         // threads per block of the decoder launch
constexpr int JH_SUB_BITS = 1024;    // bits per sub-sequence (= per thread): a decoder started in a wrong state needs a
                                     // few hundred bits to fall into step (median 450 at q75, never more than 6.4 kbit)
constexpr int JH_MAX_ROUNDS = 64;    // more rounds than that without a fixed point: reported as a corrupt stream
constexpr int JH_MAX_BPM = 10;  // blocks per MCU (T.81: at most 10 in an interleaved scan)

constexpr int JH_SUB_TABLES = 16;  // second-level tables (64 entries: the 6 bits after the first 10) per Huffman table

struct JhPlan {  // host-built, read by the kernel (POD; lives in pinned memory next to the scan bytes)
  // tables 0, 1: DC; 2, 3: AC.  (code length << 8) | symbol; 0x8000 | k: a longer code, continue in sub-table k with
  // the next 6 bits; 0: a longer code whose prefix got no sub-table (walk maxcode / valoff, as the host decoder does)
  uint16_t fast[4][1 << JH_FAST_BITS];
  uint16_t sub[4][JH_SUB_TABLES * 64];  // (code length << 8) | symbol, 0 = not a code
  int32_t maxcode[4][18];
  int32_t valoff[4][17];
  uint8_t huffval[4][256];
  int32_t bpm, mcus_x, mcus_y, total_blocks, components;
  int32_t H[3], V[3], blocks_x[3], block_off[3];
  uint8_t blk_comp[JH_MAX_BPM + 2], blk_v[JH_MAX_BPM + 2], blk_h[JH_MAX_BPM + 2];
  uint8_t dc_tab[4], ac_tab[4];  // per component: index into the four tables above
  uint32_t total_bits;           // clean (unstuffed) scan bits
  uint32_t entries_cap;
};

struct JhImage {
  const uint32_t* scan;   // unstuffed scan bytes, zero-padded by >= 16 bytes, 4-byte aligned
  const JhPlan* plan;
  // the sparse form of csrc/jpeg.cu:
  uint32_t* first;        // total_blocks (plane order)
  uint32_t* entries;      // entries_cap
  uint8_t* count;         // total_blocks
  uint32_t* first_scan;   // total_blocks + 1 scratch (scan order)
  int* info;              // [0] entries written, [1] error (0 = ok), [2] rounds of the synchronisation
  uint2* exits;           // 2 x sub-sequences: (exit position, block of the MCU << 8 | coefficient index)
  int2* counts;           // sub-sequences: (blocks started, entries), then their exclusive prefix sums
  int4* dc_part;          // threads of the launch: per-component DC partial sums, then their prefix sums
};

struct JhArgs {
  JhImage im[2];
  int* flag;              // three ints: "some exit state changed in this round", by round number mod 3
  unsigned* bar;          // software grid barrier: [0] arrivals (monotonic), [1] abort; nullptr: cooperative launch
};

}  // namespace uvo
