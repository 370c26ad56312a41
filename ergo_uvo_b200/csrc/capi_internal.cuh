// capi_internal.cuh -- definition of the opaque uvo_ctx and of the grow-only scratch the stage-level (host buffer)
// entry points use.  Whole-frame pipelines own their buffers separately (frame.cu).
#pragma once
#include "common.cuh"
#include "frontend.cuh"
#include "imgprep.cuh"

namespace uvo {

struct StageScratch {
  DevBuf<uint8_t> src3, gray, lut;
  DevBuf<int32_t> integral;
  DevBuf<uint8_t> bytes_a, bytes_b, bytes_c, bytes_d, bytes_e;  // generic staging for the other stage calls
};

}  // namespace uvo

struct uvo_ctx {
  uvo::Ctx c;
  uvo::StageScratch scratch;
  uvo::FrontEnd fe;  // front end used by the stage-level uvo_detect_features
  uvo::PinnedBuf<int> pinned_counts;
  int last_match_fallbacks = 0;  // queries of the last matcher call that took the exact full-scan path
  uvo::PinnedBuf<uint32_t> jpeg_coef;  // uvo_jpeg_decode: sparse quantised coefficients, host side of the H2D copy
  int match_exact_only = 0;      // diagnostics (uvo_match_exact_only): stage-level matcher calls skip the tcgen05 pass
  int jpeg_gpu_entropy = 1;      // uvo_jpeg_gpu_entropy: Huffman decoding on the GPU for streams that qualify
  int jpeg_last_route = 0;       // 1: the last uvo_jpeg_decode ran the GPU entropy decoder, 0: the host decoder
  int jpeg_last_rounds = 0;      // synchronisation rounds of that decode
  int jpeg_stamps[64] = {};      // diagnostics: [0] count, then 64-bit globaltimer stamps of the decode's phases
  int pnp_profile = 0;           // diagnostics (uvo_pnp_profile): uvo_solve_pnp_ransac records clock64 phase stamps
  long long pnp_stamps[32] = {};
};
