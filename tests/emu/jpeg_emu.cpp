// tests/emu/jpeg_emu.cpp -- TEST HARNESS (not part of the product, never loaded by it): executes the thread bodies of
// k_jpeg_idct and k_jpeg_color (ergo_uvo_b200/csrc/jpeg_kernels.cuh, compiled here by g++ as plain host functions) on
// the CPU over the same launch grid uvo_jpeg_decode uses, so that the index arithmetic and the integer pipeline of the
// two kernels are checked against the oracle without a GPU (tests/test_jpeg_emu.py).  Warps are run one after the
// other -- pass 1 of a warp, then its pass 2 -- on a poisoned workspace, so a dependency that the kernel's
// __syncwarp() would not cover shows up as a wrong result.  What this cannot show: anything about the real launch
// (occupancy, memory system, timing).
#include <cstring>
#include <vector>

#include "jpeg_kernels.cuh"

using namespace uvo::jpegk;

extern "C" int emu_jpeg_decode(const uint32_t* entries, const uint32_t* first, const uint8_t* count,
                               const uvo_jpeg_layout* L, uint8_t* out, size_t out_pitch) {
  if (!entries || !first || !count || !L || !out) return -1;
  std::vector<uint8_t> planes(plane_bytes(*L), 0xAB);
  IdctArgs ia;
  ColorArgs ca;
  fill_args(*L, entries, first, count, planes.data(), L->components == 3 ? out : nullptr, out_pitch, ia, ca);
  // k_jpeg_idct<<<div_up(total_blocks, IDCT_BLOCKS), IDCT_THREADS>>>
  const int grid = (ia.total_blocks + IDCT_BLOCKS - 1) / IDCT_BLOCKS;
  std::vector<int> ws(IDCT_BLOCKS * WS_STRIDE);
  std::vector<int16_t> tile(IDCT_BLOCKS * TILE_STRIDE);
  for (int b = 0; b < grid; b++) {
    for (auto& v : ws) v = 0x5A5A5A5A;  // shared memory is not initialised
    for (auto& v : tile) v = 0x5A5A;
    for (int warp = 0; warp < IDCT_THREADS / 32; warp++) {  // the stages are what the kernel separates by __syncwarp()
      for (int lane = 0; lane < 32; lane++) idct_clear(ia, b, warp * 32 + lane, tile.data());
      for (int lane = 31; lane >= 0; lane--) idct_scatter(ia, b, warp * 32 + lane, tile.data());
      for (int lane = 0; lane < 32; lane++) idct_pass1(ia, b, warp * 32 + lane, tile.data(), ws.data());
      for (int lane = 31; lane >= 0; lane--) idct_pass2(ia, b, warp * 32 + lane, ws.data());
    }
  }
  if (L->components == 1) {  // cudaMemcpy2D of the luminance plane
    for (int y = 0; y < L->height; y++)
      memcpy(out + (size_t)y * out_pitch, ia.c[0].plane + (size_t)y * L->blocks_x[0] * 8, L->width);
    return 0;
  }
  // k_jpeg_color<<<dim3(div_up(W, COLOR_TX), div_up(H, COLOR_TY)), COLOR_TX * COLOR_TY>>>
  const int gx = (L->width + COLOR_TX - 1) / COLOR_TX, gy = (L->height + COLOR_TY - 1) / COLOR_TY;
  for (int by = 0; by < gy; by++)
    for (int bx = 0; bx < gx; bx++)
      for (int t = 0; t < COLOR_TX * COLOR_TY; t++) color_thread(ca, bx, by, t);
  return 0;
}
