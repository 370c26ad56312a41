// jpeg.cuh -- internal interface of the JPEG ingest (csrc/jpeg.cu) for the whole-frame handles: the device half of
// the decode (sparse coefficients -> IDCT -> upsampling + colour conversion, or bayer demosaic) launched on a caller's
// stream into a caller's device image, and the host half (Huffman decoding) callable from any thread.
// Reference: the cv::imdecode inside from_ros_to_cv_image, math_utility.cpp:154-173.
#pragma once
#include "common.cuh"

namespace uvo {

// bytes of the device copy of one image's sparse form: block_first (nb x u32), entries (ne x u32), block_count (nb x u8)
size_t jpeg_sparse_device_bytes(const uvo_jpeg_layout& L, size_t n_entries);
// bytes of the component planes the IDCT writes
size_t jpeg_plane_bytes(const uvo_jpeg_layout& L);
// H2D of the three arrays on `copy_stream` into d_sparse (laid out as above)
void jpeg_upload_sparse(cudaStream_t copy_stream, const uvo_jpeg_sparse& sp, uint8_t* d_sparse);
// k_jpeg_idct + (k_jpeg_color | k_demosaic_bggr) on c.stream: d_sparse -> d_planes -> d_bgr (3-channel interleaved).
// 3-component streams, or 1-component streams with bayer_bggr set (the compressed-bayer message of the reference)
void jpeg_launch_transform(Ctx& c, const uvo_jpeg_layout& L, size_t n_entries, uint8_t* d_sparse, uint8_t* d_planes,
                           int bayer_bggr, uint8_t* d_bgr, size_t bgr_pitch);
// host half: throws InvalidArg on malformed / unsupported streams.  `first` (nb) and `count` (nb) are zeroed here.
void jpeg_host_decode_sparse(const uint8_t* jpeg, size_t len, uint32_t* entries, size_t capacity, uint32_t* first,
                             uint8_t* count, size_t* n_entries, uvo_jpeg_layout* layout);


// ---- Huffman decoding on the GPU (jpeg_huff.cuh): qualifying streams never touch the host decoder
struct JpegGpuJob {
  uvo_jpeg_layout L;
  size_t upload_bytes;  // front of the pinned staging to copy to the front of the device buffer: [JhPlan][scan bytes + pad]
};
// pinned staging needed for a stream of `jpeg_len` bytes
size_t jpeg_gpu_host_bytes(size_t jpeg_len);
// host part: marker walk, table plan, unstuffed copy of the scan into `pinned`.  false: the stream is one this path
// does not take (several scans, restart intervals): use the host decoder.  Throws on malformed streams.
bool jpeg_gpu_prepare(const uint8_t* jpeg, size_t len, uint8_t* pinned, size_t pinned_cap, JpegGpuJob* job);
// device buffer for one image: upload region + sparse arrays + scratch
size_t jpeg_gpu_device_bytes(const uvo_jpeg_layout& L, size_t upload_bytes);
// k_jpeg_huff for n (1 or 2) uploaded images in ONE launch, then IDCT + colour / demosaic per image, all on c.stream.
// d_status (nullable): two ints per image on the device, [error, synchronisation rounds]
// how many jpeg_gpu_launch calls of this size may be in flight at once (on different streams); see jpeg.cu
int jpeg_gpu_max_concurrent(Ctx& c, int n, const JpegGpuJob* jobs);
void jpeg_gpu_launch(Ctx& c, int n, const JpegGpuJob* jobs, uint8_t* const* d_buf, uint8_t* const* d_planes,
                     int bayer_bggr, uint8_t* const* d_bgr, size_t bgr_pitch, int* d_status);

}  // namespace uvo
