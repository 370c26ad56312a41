// tma_tile_probe.cu -- stand-alone check of the TMA box k_surf_detect uses: 2-D int32 tensor (w+1) x (h+1) with a
// 16-byte padded pitch, box 64 x 45, loaded at negative and past-the-edge coordinates; compares with a plain copy.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../ergo_uvo_b200/csrc/tma.cuh"
using namespace uvo;
constexpr int BR = 45, BC = 64;
__device__ int g_bc, g_br;
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int r0, int* out) {
  __shared__ __align__(128) int tile[BR * BC];
  __shared__ __align__(8) unsigned long long bar_s;
  const uint32_t bar = smem_u32(&bar_s);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, g_br * g_bc * 4);
    tma_load_2d(smem_u32(tile), &map, c0, r0, bar);
  }
  __syncthreads();
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < BR * BC; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
  const int dtype = argc > 1 ? atoi(argv[1]) : 0, bc = argc > 2 ? atoi(argv[2]) : BC, br = argc > 3 ? atoi(argv[3]) : BR;
  const int a0 = argc > 4 ? atoi(argv[4]) : -14, a1 = argc > 5 ? atoi(argv[5]) : -14;
  cudaMemcpyToSymbol(g_bc, &bc, 4); cudaMemcpyToSymbol(g_br, &br, 4);
  const int w = 1280, h = 1024, pitch = (w + 1 + 3) & ~3;
  std::vector<int> hsum((size_t)pitch * (h + 1));
  for (size_t i = 0; i < hsum.size(); i++) hsum[i] = (int)(i * 2654435761u >> 8);
  int *d, *dout;
  cudaMalloc(&d, hsum.size() * 4);
  cudaMalloc(&dout, BR * BC * 4);
  cudaMemcpy(d, hsum.data(), hsum.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)(w + 1), (cuuint64_t)(h + 1)};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  const cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)br};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_tiled_fn()(&m, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : (dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT32), 2, d, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  const int coords[1][2] = {{a0, a1}};
  for (auto& cr : coords) {
    k<<<1, 256>>>(m, cr[0], cr[1], dout);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<int> out(BR * BC);
    cudaMemcpy(out.data(), dout, BR * BC * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < br; y++)
      for (int x = 0; x < bc; x++) {
        const int gy = cr[1] + y, gx = cr[0] + x;
        const int want = (gy >= 0 && gy <= h && gx >= 0 && gx <= w) ? hsum[(size_t)gy * pitch + gx] : 0;
        bad += out[y * bc + x] != want;
      }
    printf("coord (%d, %d): %s, %d mismatches\n", cr[0], cr[1], cudaGetErrorString(e), bad);
  }
  return 0;
}
