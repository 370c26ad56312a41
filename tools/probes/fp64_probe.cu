// fp64_probe.cu -- measures on the GPU it runs on (a) the dependent-issue latency of DFMA / DADD / DMUL / sqrt / 1/x /
// x/y in SM cycles (one warp, one chain), (b) the sustained DFMA rate of the whole chip (the FP64-issue roofline the
// pose kernels k_pnp_* / k_tv_* are bounded by; written next to MEASURED_PEAKS.json's HBM and bf16 numbers by
// bench.py).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu ; prints JSON.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k_latency(double* out, long long* cyc, double a, double b, int n) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      if (OP == 0) x = fma(x, b, a);
      if (OP == 1) x = __dadd_rn(x, b);
      if (OP == 2) x = __dmul_rn(x, b);
      if (OP == 3) x = sqrt(x) + a;
      if (OP == 4) x = 1.0 / x + a;
      if (OP == 5) x = b / x + a;
      if (OP == 6) x = __fmaf_rn((float)x, 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) {
    *cyc = t1 - t0;
    *out = x;
  }
}

// throughput: every thread runs 8 independent DFMA chains
__global__ void __launch_bounds__(256) k_throughput(double* out, double a, double b, int n) {
  double x[8];
  for (int k = 0; k < 8; k++) x[k] = a + k + threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) x[k] = fma(x[k], b, a);
  }
  double s = 0;
  for (int k = 0; k < 8; k++) s += x[k];
  if (s == 12345.678) out[0] = s;
}

int main() {
  double* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, 64);
  cudaMalloc(&d_cyc, 64);
  const char* names[7] = {"dfma", "dadd", "dmul", "sqrt_plus_add", "rcp_plus_add", "div_plus_add", "ffma_with_cvt"};
  printf("{\"latency_cycles\": {");
  const int n = 2000;
  for (int op = 0; op < 7; op++) {
    long long cyc = 0;
    for (int rep = 0; rep < 2; rep++) {
      switch (op) {
        case 0: k_latency<0><<<1, 32>>>(d_out, d_cyc, 1.000001, 0.999999, n); break;
        case 1: k_latency<1><<<1, 32>>>(d_out, d_cyc, 1.000001, 0.999999, n); break;
        case 2: k_latency<2><<<1, 32>>>(d_out, d_cyc, 1.000001, 0.999999, n); break;
        case 3: k_latency<3><<<1, 32>>>(d_out, d_cyc, 1.000001, 0.999999, n); break;
        case 4: k_latency<4><<<1, 32>>>(d_out, d_cyc, 1.000001, 0.999999, n); break;
        case 5: k_latency<5><<<1, 32>>>(d_out, d_cyc, 1.000001, 0.999999, n); break;
        case 6: k_latency<6><<<1, 32>>>(d_out, d_cyc, 1.000001, 0.999999, n); break;
      }
      cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    }
    printf("%s\"%s\": %.2f", op ? ", " : "", names[op], (double)cyc / (16.0 * n));
  }
  printf("}, ");
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, iters = 20000;
  k_throughput<<<blocks, 256>>>(d_out, 1.000001, 0.999999, 100);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_throughput<<<blocks, 256>>>(d_out, 1.000001, 0.999999, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double fmas = (double)blocks * 256 * 32.0 * iters;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("\"dfma_tflops\": %.2f, \"dfma_per_clk_per_sm\": %.1f, \"sm_count\": %d, \"sm_clock_khz_nominal\": %d, \"gpu\": \"%s\"}\n",
         2.0 * fmas / (best * 1e-3) / 1e12, fmas / (best * 1e-3) / prop.multiProcessorCount / (clk * 1e3),
         prop.multiProcessorCount, clk, prop.name);
  return 0;
}
