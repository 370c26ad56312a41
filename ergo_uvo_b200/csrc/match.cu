// match.cu -- K8 exact matcher.  Distances are accumulated in exactly the f32 order OpenCV's normL2Sqr_ uses in its
// baseline build (4 accumulators x 4 lanes, separate multiply and add, ((d0+d1)+d2)+d3 then (s0+s2)+(s1+s3)), so the
// distances -- and therefore indices, tie order and ratio-test outcome -- are bit-identical to BFMatcher
// (pinned by tests/golden/matcher_300x400.npz).
//
// k_knn2_partial: grid (query blocks, MATCH_SPLITS train slices); one thread owns one query (64 floats in
// registers), the block streams its slice of the train set through shared memory (broadcast LDS.128).
// k_knn2_merge: merges the slices in ascending train order, applies the ratio test and compacts the survivors in
// query order (single block, ballot scan) -- order-preserving, no atomics.
#include <cfloat>

#include "match.cuh"

namespace uvo {

constexpr int QB = 128;  // queries per block (one per thread)
constexpr int TT = 64;   // train rows staged per tile

__device__ __forceinline__ void knn2_insert(Knn2& b, float d, int j) {
  // BatchDistInvoker insertion: candidates arrive in ascending j; strict comparisons keep the lower index on ties
  if (d < b.d1) {
    if (b.d0 > d) {
      b.d1 = b.d0;
      b.i1 = b.i0;
      b.d0 = d;
      b.i0 = j;
    } else {
      b.d1 = d;
      b.i1 = j;
    }
  }
}

__global__ void __launch_bounds__(QB) k_knn2_partial(const __grid_constant__ MatchArgs a) {
  __shared__ float4 s_t[TT][16];
  const int nq = a.nq_dev ? min(*a.nq_dev, a.nq) : a.nq;
  const int nt = a.nt_dev ? min(*a.nt_dev, a.nt) : a.nt;
  if ((int)(blockIdx.x * QB) >= nq) return;
  const int qi = blockIdx.x * QB + threadIdx.x;
  const bool active = qi < nq;
  // slice of the train set handled by this block (multiple of TT so tiles never straddle slices)
  const int per = ((nt + MATCH_SPLITS - 1) / MATCH_SPLITS + TT - 1) / TT * TT;
  const int t0 = blockIdx.y * per, t1 = min(t0 + per, nt);
  float4 q[16];
  if (active) {
    const float4* qp = reinterpret_cast<const float4*>(a.q + (size_t)qi * 64);
#pragma unroll
    for (int k = 0; k < 16; k++) q[k] = qp[k];
  }
  Knn2 best{FLT_MAX, FLT_MAX, -1, -1};
  for (int base = t0; base < t1; base += TT) {
    const int m = min(TT, t1 - base);
    __syncthreads();
    for (int e = threadIdx.x; e < m * 16; e += QB)
      s_t[e >> 4][e & 15] = reinterpret_cast<const float4*>(a.t + (size_t)base * 64)[e];
    __syncthreads();
    if (active) {
      for (int j = 0; j < m; j++) {
        float4 acc[4];
#pragma unroll
        for (int k = 0; k < 4; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int it = 0; it < 4; it++)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float4 tv = s_t[j][it * 4 + k], qv = q[it * 4 + k];
            float d;
            d = __fsub_rn(qv.x, tv.x); acc[k].x = __fadd_rn(acc[k].x, __fmul_rn(d, d));
            d = __fsub_rn(qv.y, tv.y); acc[k].y = __fadd_rn(acc[k].y, __fmul_rn(d, d));
            d = __fsub_rn(qv.z, tv.z); acc[k].z = __fadd_rn(acc[k].z, __fmul_rn(d, d));
            d = __fsub_rn(qv.w, tv.w); acc[k].w = __fadd_rn(acc[k].w, __fmul_rn(d, d));
          }
        const float s0 = __fadd_rn(__fadd_rn(__fadd_rn(acc[0].x, acc[1].x), acc[2].x), acc[3].x);
        const float s1 = __fadd_rn(__fadd_rn(__fadd_rn(acc[0].y, acc[1].y), acc[2].y), acc[3].y);
        const float s2 = __fadd_rn(__fadd_rn(__fadd_rn(acc[0].z, acc[1].z), acc[2].z), acc[3].z);
        const float s3 = __fadd_rn(__fadd_rn(__fadd_rn(acc[0].w, acc[1].w), acc[2].w), acc[3].w);
        const float d2 = __fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3));
        knn2_insert(best, __fsqrt_rn(d2), base + j);
      }
    }
  }
  if (active) a.partial[(size_t)blockIdx.y * a.nq + qi] = best;
}

__global__ void __launch_bounds__(1024) k_knn2_merge(const __grid_constant__ MatchArgs a) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int nq = a.nq_dev ? min(*a.nq_dev, a.nq) : a.nq;
  const int nt = a.nt_dev ? min(*a.nt_dev, a.nt) : a.nt;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < nq; base += 1024) {
    const int qi = base + tid;
    bool keep = false;
    Knn2 best{FLT_MAX, FLT_MAX, -1, -1};
    if (qi < nq) {
      for (int s = 0; s < MATCH_SPLITS; s++) {  // ascending train order
        const Knn2 p = a.partial[(size_t)s * a.nq + qi];
        if (p.i0 >= 0) knn2_insert(best, p.d0, p.i0);
        if (p.i1 >= 0) knn2_insert(best, p.d1, p.i1);
      }
      a.knn[qi] = best;
      // reference: knn[i][0].distance < ratio * knn[i][1].distance; fewer than 2 train rows => no match (the
      // reference would read out of bounds there)
      keep = nt >= 2 && best.i1 >= 0 && best.d0 < __fmul_rn(a.ratio, best.d1);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int k = 0; k < wid; k++) off += s_warp[k];
    if (keep) {
      uvo_dmatch m;
      m.queryIdx = qi;
      m.trainIdx = best.i0;
      m.imgIdx = 0;
      m.distance = best.d0;
      a.matches[off + __popc(bal & ((1u << lane) - 1))] = m;
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int k = 0; k < 32; k++) tot += s_warp[k];
      s_base += tot;
    }
    __syncthreads();
  }
  if (tid == 0) *a.n_matches = s_base;
}

void launch_match(Ctx& c, const MatchArgs& a) {
  if (a.nq <= 0) {
    UVO_CUDA(cudaMemsetAsync(a.n_matches, 0, sizeof(int), c.stream));
    return;
  }
  UVO_KERNEL(c, "k_knn2_partial");
  k_knn2_partial<<<dim3(div_up(a.nq, QB), MATCH_SPLITS), QB, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_knn2_merge");
  k_knn2_merge<<<1, 1024, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
}

}  // namespace uvo
