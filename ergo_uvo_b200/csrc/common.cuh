// common.cuh -- context, error handling and small device helpers shared by every kernel file of libuvo_b200.so.
// All arithmetic that must be bit-identical to the CPU behaviour of OpenCV is written with explicit
// round-to-nearest intrinsics (no FMA contraction); the library is additionally compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <vector>

#include "uvo_c.h"

namespace uvo {

// optional per-kernel CUDA-event timing (bench.py's roofline leg): every launch is bracketed by two events on the
// launching stream; durations are summed per kernel name when the results are read.
struct KernelTimer {
  bool enabled = false;
  struct Rec {
    const char* name;
    cudaEvent_t a, b;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  int64_t launches = 0;
  std::string err;
  KernelTimer kt;
  const char* pending_name = nullptr;
  cudaEvent_t pending_ev = nullptr;
  // call right before a kernel launch
  void kbegin(const char* name) {
    if (!kt.enabled) return;
    pending_name = name;
    pending_ev = kt.get();
    cudaEventRecord(pending_ev, stream);
  }
  void kend() {
    launches++;
    if (!kt.enabled || !pending_name) return;
    cudaEvent_t b = kt.get();
    cudaEventRecord(b, stream);
    kt.recs.push_back({pending_name, pending_ev, b});
    pending_name = nullptr;
  }
};

struct CudaError {
  cudaError_t e;
  const char* what;
  const char* file;
  int line;
};

#define UVO_CUDA(expr)                                              \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) throw uvo::CudaError{_e, #expr, __FILE__, __LINE__}; \
  } while (0)

#define UVO_LAUNCH_CHECK(ctx)  \
  do {                         \
    (ctx).kend();              \
    UVO_CUDA(cudaGetLastError()); \
  } while (0)
#define UVO_KERNEL(ctx, name) (ctx).kbegin(name)

struct InvalidArg {
  std::string msg;
  int code;
};
#define UVO_REQUIRE(cond, msg) \
  do {                         \
    if (!(cond)) throw uvo::InvalidArg{std::string(msg), UVO_ERR_INVALID}; \
  } while (0)

// runs fn, translating exceptions to status codes + ctx->err
template <class F>
static inline int guarded(Ctx* ctx, F&& fn) {
  try {
    fn();
    return UVO_OK;
  } catch (const CudaError& e) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in %s", (int)e.e, cudaGetErrorString(e.e), e.file, e.line,
             e.what);
    if (ctx) ctx->err = buf;
    cudaGetLastError();
    return UVO_ERR_CUDA;
  } catch (const InvalidArg& e) {
    if (ctx) ctx->err = e.msg;
    return e.code;
  } catch (const std::exception& e) {
    if (ctx) ctx->err = e.what();
    return UVO_ERR_INVALID;
  }
}

// RAII device buffer (grow-only), allocated with cudaMallocAsync-free plain cudaMalloc (allocation is off the
// per-frame path: every pipeline object sizes its buffers once at creation).
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void ensure(size_t count) {
    if (count <= n) return;
    release();
    UVO_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
    n = count;
  }
  T* get() const { return p; }
};

template <class T>
struct PinnedBuf {
  T* p = nullptr;
  size_t n = 0;
  ~PinnedBuf() {
    if (p) cudaFreeHost(p);
  }
  void ensure(size_t count) {
    if (count <= n) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    UVO_CUDA(cudaMallocHost((void**)&p, count * sizeof(T)));
    n = count;
  }
};

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

}  // namespace uvo
