// match.cu -- K8 matcher: brute-force L2 kNN (k = 2) + Lowe ratio test, bit-identical to BFMatcher(NORM_L2).knnMatch.
//
// Five kernels per call:
//
//  k_knn_prep    -|t|^2/2 per train row and max |t|^2.
//  k_knn_tc      the Nq x Nt descriptor similarity matrix on the 5th-gen tensor cores.  One CTA owns a 128-query
//                tile and a contiguous chunk of train tiles.  A TMA producer warp streams 128 x D f32 train tiles
//                (D / 32 SWIZZLE_128B atoms of 32 floats; D = 64, or 128 for extended SURF) through a 4-stage (D = 128:
//                2-stage) shared-memory ring; one elected thread issues tcgen05.mma.kind::tf32 (M = 128, N = 128,
//                D / 8 k-steps of 8) into one of four 128-column TMEM
//                accumulators; sixteen epilogue warps read the accumulators back with tcgen05.ld (each thread owns
//                one query row and a quarter of the tile's columns), add -|t|^2/2 and keep the four largest
//                similarities per (row, column quarter) with a branch-free bitonic network on index-carrying keys.
//                The matrix itself is never written anywhere.
//  k_knn_rerank  one warp per query: the candidates (4 per list, <= 128) are re-evaluated with exactly the f32
//                arithmetic OpenCV's normL2Sqr_ uses (4 accumulators x 4 lanes, separate multiply and add,
//                ((d0+d1)+d2)+d3 then (s0+s2)+(s1+s3), pinned by tests/golden/matcher_300x400.npz), and the best two
//                are selected with BatchDistInvoker's tie rule (lower train index first).  The TF32 pass only has to
//                be good enough to *contain* the true top two: every non-candidate j of a list has a key <= that
//                list's 4th key, so its exact squared distance is >= |q|^2 - 2*T - eps, where eps bounds the TF32
//                truncation and key quantisation error.  If the exact second-best does not
//                beat that bound the query is flagged.
//  k_knn_exact   flagged queries (measured: ~0.1 % on the stereo workload) are re-done by an exact scan of the whole
//                train set, one block each -- so the result is exact by construction, not statistically.
//  k_knn_compact ratio test + order-preserving compaction of the survivors in query order (single block, scan;
//                no atomics).
//
// Reference: match_features, VO_utility.cpp:515-573.
#include <cuda.h>

#include <cfloat>
#include <mutex>

#include "match.cuh"
#include "tma.cuh"

namespace uvo {

// ------------------------------------------------------------------------------------------------ geometry
constexpr int TILE = 128;          // queries per CTA == train rows per MMA tile
constexpr int ACC = 4;             // TMEM accumulator buffers of 128 columns (4 x 128 = all 512 columns)
constexpr int ATOM_BYTES = TILE * 128;     // one SWIZZLE_128B atom: 128 rows x 32 floats
constexpr int MAX_CHUNK_TILES = 32;        // train tiles per CTA at most (bounds the -|t|^2/2 table in smem)
constexpr int EPI_WARPS = 16;
constexpr int TC_THREADS = 32 * (2 + EPI_WARPS);  // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..17 epilogue
constexpr int TOPK = MATCH_TOPK;
constexpr int MATCH_MAX_LISTS = MATCH_MAX_CHUNKS * 4;  // (chunk, column quarter) candidate lists per query
// shared-memory plan of k_knn_tc for D floats per descriptor row (64: SURF, 128: extended SURF).  A tile is D / 32
// SWIZZLE_128B atoms; the 128-d ring is two stages deep so that query tile + ring + table stay under 227 KB.
template <int D>
struct TcGeom {
  static_assert(D == 64 || D == 128, "descriptor rows are 64 or 128 floats");
  static constexpr int ATOMS = D / 32;
  static constexpr int TILE_BYTES = ATOMS * ATOM_BYTES;  // 128 rows x D floats
  static constexpr int STAGES = D == 64 ? 4 : 2;         // shared-memory ring depth (train tiles)
  static constexpr int SMEM_A = 0;
  static constexpr int SMEM_B = SMEM_A + TILE_BYTES;
  static constexpr int SMEM_HB = SMEM_B + STAGES * TILE_BYTES;
  static constexpr int SMEM_BAR = SMEM_HB + MAX_CHUNK_TILES * TILE * 4;
  static constexpr int SMEM_TOTAL = SMEM_BAR + 256;
  static constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;  // slack to align the base to 1024 B (SWIZZLE_128B requirement)
  static_assert(SMEM_ALLOC <= 227 * 1024, "k_knn_tc shared memory");
};

// number of train chunks (grid.y CTAs that do work) for ntq query tiles and ntt train tiles; the same formula runs
// on the device in k_knn_tc and k_knn_rerank.
__host__ __device__ inline int chunk_count(int ntq, int ntt, int sms) {
  int n = sms / (ntq > 0 ? ntq : 1);
  if (n < 1) n = 1;
  const int need = (ntt + MAX_CHUNK_TILES - 1) / MAX_CHUNK_TILES;
  if (n < need) n = need;
  if (n > MATCH_MAX_CHUNKS) n = MATCH_MAX_CHUNKS;
  if (n > ntt) n = ntt > 0 ? ntt : 1;
  return n;
}

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs, FP32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups 1024 B apart (SBO), LBO = 1 (unused for
// swizzled K-major), descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
constexpr uint32_t IDESC_TF32_128x128 = (1u << 4) | (2u << 7) | (2u << 10) | ((TILE >> 3) << 17) | ((TILE >> 4) << 24);

// ------------------------------------------------------------------------------------------------ exact arithmetic
// squared L2 distance of two D-float rows in the accumulation order of OpenCV's normL2Sqr_ (baseline SIMD build:
// the same four 4-lane accumulators run over all D / 16 groups of 16 floats)
template <int D>
__device__ __forceinline__ float l2sqr_cv(const float4* __restrict__ q, const float4* __restrict__ t) {
  float4 acc[4];
#pragma unroll
  for (int k = 0; k < 4; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int it = 0; it < D / 16; it++)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float4 tv = t[it * 4 + k], qv = q[it * 4 + k];
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
  return __fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3));
}

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

// (distance, index) packed so that unsigned comparison == (distance asc, index asc); distances are >= 0 or NaN
__device__ __forceinline__ unsigned long long knn_key(float d, int j) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)j;
}
constexpr unsigned long long KEY_NONE = ~0ull;

__device__ __forceinline__ unsigned long long warp_min_key(unsigned long long k) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long x = __shfl_xor_sync(0xffffffffu, k, o);
    k = x < k ? x : k;
  }
  return k;
}

// ------------------------------------------------------------------------------------------------ k_knn_prep
// hb[j] = -|t_j|^2 / 2 for every train row, max |t|^2 (for the error bound); 16 lanes per row, coalesced
template <int D>
__global__ void __launch_bounds__(256) k_knn_prep(const __grid_constant__ MatchArgs a) {
  const int nt = a.nt_dev ? min(*a.nt_dev, a.nt) : a.nt;
  const int sub = threadIdx.x & 15;
  float n2max = 0.f;
  for (int j = (blockIdx.x * 256 + threadIdx.x) >> 4; j < nt; j += (gridDim.x * 256) >> 4) {
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < D / 64; u++) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.t + (size_t)j * D) + sub + 16 * u);
      s += fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (sub == 0) a.hb[j] = -0.5f * s;
    n2max = fmaxf(n2max, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n2max = fmaxf(n2max, __shfl_xor_sync(0xffffffffu, n2max, o));
  if ((threadIdx.x & 31) == 0 && n2max > 0.f) atomicMax(a.tn2max, __float_as_uint(n2max));
}

// ------------------------------------------------------------------------------------------------ k_knn_tc
// Candidate keys: the similarity as an f32 whose low KEY_BITS mantissa bits are replaced by the column's position in
// the list (tile-in-chunk << 5 | column-in-quarter).  Keys order like the similarities up to a relative perturbation
// of 2^-13 (accounted for in the error bound), are unique inside a list, and make top-4 maintenance a branch-free
// FMNMX network.
constexpr int KEY_BITS = 10;
constexpr uint32_t KEY_MASK = (1u << KEY_BITS) - 1;
constexpr float HB_PAD = -1e30f;  // -|t|^2/2 of a column past the end of the train set: finite, below everything

#define UVO_CE(hi, lo)                \
  do {                                \
    const float _h = fmaxf(hi, lo);   \
    lo = fminf(hi, lo);               \
    hi = _h;                          \
  } while (0)

template <int D>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_knn_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t,
         const __grid_constant__ MatchArgs a, const int sms) {
  extern __shared__ uint8_t smem_raw[];
  using G = TcGeom<D>;
  constexpr int STAGES = G::STAGES, TILE_BYTES = G::TILE_BYTES, SMEM_BAR = G::SMEM_BAR;
  const int nq = a.nq_dev ? min(*a.nq_dev, a.nq) : a.nq;
  const int nt = a.nt_dev ? min(*a.nt_dev, a.nt) : a.nt;
  const int ntq = (nq + TILE - 1) / TILE, ntt = (nt + TILE - 1) / TILE;
  if ((int)blockIdx.x >= ntq) return;  // uniform per CTA, before any barrier or TMEM allocation
  const int n_chunks = chunk_count(ntq, ntt, sms);
  if ((int)blockIdx.y >= n_chunks) return;
  const int per = (ntt + n_chunks - 1) / n_chunks;
  const int tile0 = blockIdx.y * per;
  const int ntile = max(0, min(ntt, tile0 + per) - tile0);
  const int row0 = blockIdx.x * TILE;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base + G::SMEM_A, sB = base + G::SMEM_B;
  float* s_hb = reinterpret_cast<float*>(smem + G::SMEM_HB);
  const uint32_t bar = base + SMEM_BAR;
  // barriers (8 B each): full[STAGES], empty[STAGES], A landed, accumulator full[ACC], accumulator empty[ACC]
  const uint32_t bar_full = bar, bar_empty = bar + 8 * STAGES, bar_a = bar + 8 * (2 * STAGES);
  const uint32_t bar_tfull = bar_a + 8, bar_tempty = bar_tfull + 8 * ACC;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + SMEM_BAR + 8 * (2 * STAGES + 1 + 2 * ACC));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_t) : "memory");
    for (int i = 0; i < STAGES; i++) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    mbar_init(bar_a, 1);
    for (int i = 0; i < ACC; i++) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM by launch bounds + shared-memory footprint)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(bar_a, TILE_BYTES);
#pragma unroll
      for (int at = 0; at < G::ATOMS; at++) tma_load_2d(sA + at * ATOM_BYTES, &map_q, 32 * at, row0, bar_a);
      for (int i = 0; i < ntile; i++) {
        const int st = i % STAGES;
        mbar_wait(bar_empty + 8 * st, ((i / STAGES) & 1) ^ 1);
        mbar_expect_tx(bar_full + 8 * st, TILE_BYTES);
        const uint32_t dst = sB + st * TILE_BYTES;
#pragma unroll
        for (int at = 0; at < G::ATOMS; at++)
          tma_load_2d(dst + at * ATOM_BYTES, &map_t, 32 * at, (tile0 + i) * TILE, bar_full + 8 * st);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      mbar_wait(bar_a, 0);
      for (int i = 0; i < ntile; i++) {
        const int st = i % STAGES, ab = i % ACC;
        mbar_wait(bar_tempty + 8 * ab, ((i / ACC) & 1) ^ 1);
        mbar_wait(bar_full + 8 * st, (i / STAGES) & 1);
        tc_fence_after();
        const uint32_t bs = sB + st * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < D / 8; k++) {  // k-steps of 8 floats (32 B); 4 per swizzle atom
          const uint32_t off = (k >> 2) * ATOM_BYTES + (k & 3) * 32;
          tc_mma_tf32(tmem + ab * TILE, smem_desc_sw128(sA + off), smem_desc_sw128(bs + off), IDESC_TF32_128x128,
                      k > 0);
        }
        tc_commit(bar_empty + 8 * st);   // smem stage reusable once these MMAs have read it
        tc_commit(bar_tfull + 8 * ab);   // accumulator complete
      }
    }
  } else {
    // ===== epilogue: 16 warps; warp w reads TMEM lanes 32*(w%4).. (its query rows) and the column quarter
    // (w-2)/4 of every tile: 32 accumulator columns per tile per thread =====
    const int e = threadIdx.x - 64;        // 0..511
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int cq = (warp - 2) >> 2;        // column quarter 0..3
    const int row = quarter * 32 + lane;   // query row inside the tile == TMEM lane
    for (int col = e; col < ntile * TILE; col += EPI_WARPS * 32) {
      const int j = tile0 * TILE + col;
      s_hb[col] = j < nt ? a.hb[j] : HB_PAD;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    float l0 = -INFINITY, l1 = -INFINITY, l2 = -INFINITY, l3 = -INFINITY;  // the list, descending
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16) + cq * 32;
    for (int i = 0; i < ntile; i++) {
      const int ab = i % ACC;
      mbar_wait(bar_tfull + 8 * ab, (i / ACC) & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(lane_addr + ab * TILE, r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * ab);  // values are in registers: the accumulator can be overwritten
      const float4* hb4 = reinterpret_cast<const float4*>(s_hb + i * TILE + cq * 32);
      const uint32_t tbits = (uint32_t)i << 5;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float4 h = hb4[k];
        float x0 = __uint_as_float((__float_as_uint(__uint_as_float(r[4 * k + 0]) + h.x) & ~KEY_MASK) | (tbits + 4 * k + 0));
        float x1 = __uint_as_float((__float_as_uint(__uint_as_float(r[4 * k + 1]) + h.y) & ~KEY_MASK) | (tbits + 4 * k + 1));
        float x2 = __uint_as_float((__float_as_uint(__uint_as_float(r[4 * k + 2]) + h.z) & ~KEY_MASK) | (tbits + 4 * k + 2));
        float x3 = __uint_as_float((__float_as_uint(__uint_as_float(r[4 * k + 3]) + h.w) & ~KEY_MASK) | (tbits + 4 * k + 3));
        // sort the four new keys (descending) -- independent of the list, so off the critical path
        UVO_CE(x0, x1);
        UVO_CE(x2, x3);
        UVO_CE(x0, x2);
        UVO_CE(x1, x3);
        UVO_CE(x1, x2);
        // bitonic merge: the top four of the union, then re-sort the bitonic sequence
        l0 = fmaxf(l0, x3);
        l1 = fmaxf(l1, x2);
        l2 = fmaxf(l2, x1);
        l3 = fmaxf(l3, x0);
        UVO_CE(l0, l2);
        UVO_CE(l1, l3);
        UVO_CE(l0, l1);
        UVO_CE(l2, l3);
      }
    }
    if (row0 + row < nq) {
      const size_t o = ((size_t)(row0 + row) * MATCH_MAX_LISTS + blockIdx.y * 4 + cq) * TOPK;
      *reinterpret_cast<float4*>(a.cand + o) = make_float4(l0, l1, l2, l3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ k_knn_rerank
constexpr int RR_WARPS = 8;

__device__ __forceinline__ Knn2 knn2_from_keys(unsigned long long b0, unsigned long long b1) {
  Knn2 r{FLT_MAX, FLT_MAX, -1, -1};
  if (b0 != KEY_NONE) {
    r.d0 = __uint_as_float((unsigned)(b0 >> 32));
    r.i0 = (int)(unsigned)b0;
  }
  if (b1 != KEY_NONE) {
    r.d1 = __uint_as_float((unsigned)(b1 >> 32));
    r.i1 = (int)(unsigned)b1;
  }
  return r;
}

// per-query bound on |(|q|^2 - 2 key) - D| for every (query, train) pair, D = the f32 value normL2Sqr_ computes:
// tf32 truncation of both operands (2^-8 |q||t|), f32 accumulation on either side, key quantisation (2^-13 |s|,
// |s| <= |q||t| + |t|^2/2); 1 % slack on top
// acc: relative bound of the f32 accumulation over one row (4e-5 for 64 terms, twice that for 128)
__device__ __forceinline__ float match_eps(float nq2, float tn2, float acc) {
  const float qt = sqrtf(nq2 * tn2);
  return 1.01f * (((float)MATCH_TF32_EPS + acc) * qt + acc * (nq2 + tn2) + 2.6e-4f * (qt + 0.5f * tn2));
}

// one warp per query: prune the candidates by their approximate keys, exact distances of the survivors, exact top
// two, completeness check
template <int D>
__global__ void __launch_bounds__(RR_WARPS * 32) k_knn_rerank(const __grid_constant__ MatchArgs a, const int sms) {
  const int nq = a.nq_dev ? min(*a.nq_dev, a.nq) : a.nq;
  const int nt = a.nt_dev ? min(*a.nt_dev, a.nt) : a.nt;
  const int ntq = (nq + TILE - 1) / TILE, ntt = (nt + TILE - 1) / TILE;
  const int n_chunks = chunk_count(ntq, ntt, sms);
  const int per = (ntt + n_chunks - 1) / n_chunks;
  const int n_cand = nt > 0 ? n_chunks * 4 * TOPK : 0;
  const float tn2 = __uint_as_float(*a.tn2max);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int qi = blockIdx.x * RR_WARPS + warp; qi < nq; qi += gridDim.x * RR_WARPS) {
    const float4* qp = reinterpret_cast<const float4*>(a.q + (size_t)qi * D);
    float nq2 = 0.f;
    if (lane < D / 4) {
      const float4 v = qp[lane];
      nq2 = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nq2 += __shfl_xor_sync(0xffffffffu, nq2, o);
    const float eps = match_eps(nq2, tn2, 4e-5f * (D / 64));
    // this lane's candidate keys: c = lane + 32 u  (n_cand <= 128)
    float key[4];
    float m1 = -INFINITY, m2 = -INFINITY;  // lane-local largest two present keys
    float tmax = -INFINITY;                // max over the lists of their 4th key
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int c = lane + 32 * u;
      float k = -INFINITY;
      if (c < n_cand) k = a.cand[(size_t)qi * MATCH_MAX_LISTS * TOPK + c];
      if (!(k > 0.5f * HB_PAD)) k = -INFINITY;  // -inf: empty slot; ~HB_PAD: column past the end; NaN: never
      key[u] = k;
      if ((c & (TOPK - 1)) == TOPK - 1) tmax = fmaxf(tmax, k);
      if (k > m1) {
        m2 = m1;
        m1 = k;
      } else if (k > m2) {
        m2 = k;
      }
    }
    float w1 = m1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      w1 = fmaxf(w1, __shfl_xor_sync(0xffffffffu, w1, o));
    }
    // second largest key of the warp (if the largest occurs twice this under-estimates, which only prunes less)
    float w2 = m1 == w1 ? m2 : m1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w2 = fmaxf(w2, __shfl_xor_sync(0xffffffffu, w2, o));
    // A candidate whose key is more than eps below the second largest key is farther (in exact arithmetic) than
    // both holders of the two largest keys: it cannot be among the best two.
    const float cut = w2 - eps;  // -inf when fewer than two candidates: nothing is pruned
    unsigned long long k0 = KEY_NONE, k1 = KEY_NONE;  // this lane's best two (distance, index) keys
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (key[u] >= cut && key[u] > -INFINITY) {
        const int c = lane + 32 * u;
        const int list = c / TOPK, chunk = list >> 2, cq = list & 3;
        const uint32_t kb = __float_as_uint(key[u]) & KEY_MASK;
        const int j = (chunk * per + (int)(kb >> 5)) * TILE + cq * 32 + (int)(kb & 31);
        const float d = __fsqrt_rn(l2sqr_cv<D>(qp, reinterpret_cast<const float4*>(a.t + (size_t)j * D)));
        const unsigned long long key2 = knn_key(d, j);
        if (key2 < k0) {
          k1 = k0;
          k0 = key2;
        } else if (key2 < k1) {
          k1 = key2;
        }
      }
    }
    const unsigned long long b0 = warp_min_key(k0);
    const unsigned long long b1 = warp_min_key(k0 == b0 ? k1 : k0);  // keys are unique (index in the low word)
    // Completeness: a non-candidate of a list has a key <= that list's 4th key <= tmax, so its exact squared distance
    // is >= |q|^2 - 2 tmax - eps.  tmax == -inf: no list was full, every train row is a candidate.
    bool ok = tmax == -INFINITY;
    if (!ok && b1 != KEY_NONE) {
      const float d1 = __uint_as_float((unsigned)(b1 >> 32));
      ok = d1 * d1 * (1.f + 1e-6f) + eps < nq2 - 2.f * tmax;  // false for NaN anywhere
    }
    if (lane == 0) {
      if (ok)
        a.knn[qi] = knn2_from_keys(b0, b1);
      else
        a.fb_list[atomicAdd(a.n_flagged, 1)] = qi;
    }
  }
}

// ------------------------------------------------------------------------------------------------ k_knn_exact
// exact scan of the whole train set for the (rare) queries whose candidate set could not be proven complete.  Grid
// (EX_SLICES, y): block (s, y) scans slice s of the train rows for flagged queries y, y + gridDim.y, ...; the last
// block to finish a query merges the EX_SLICES partial results (threadfence + counter).
constexpr int EX_THREADS = 256;

template <int D>
__global__ void __launch_bounds__(EX_THREADS) k_knn_exact(const __grid_constant__ MatchArgs a) {
  __shared__ Knn2 s_part[EX_THREADS / 32];
  __shared__ int s_last;
  const int nt = a.nt_dev ? min(*a.nt_dev, a.nt) : a.nt;
  const int nf = *a.n_flagged;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = (nt + MATCH_EX_SLICES - 1) / MATCH_EX_SLICES;
  const int j0 = blockIdx.x * rows, j1 = min(nt, j0 + rows);
  for (int f = blockIdx.y; f < nf; f += gridDim.y) {
    const int fq = a.fb_list[f];
    const float4* qp = reinterpret_cast<const float4*>(a.q + (size_t)fq * D);
    Knn2 best{FLT_MAX, FLT_MAX, -1, -1};
    for (int j = j0 + threadIdx.x; j < j1; j += EX_THREADS)
      knn2_insert(best, __fsqrt_rn(l2sqr_cv<D>(qp, reinterpret_cast<const float4*>(a.t + (size_t)j * D))), j);
    // merge by (distance, index): the global second is the winner's second or somebody else's first
    unsigned long long k0 = best.i0 >= 0 ? knn_key(best.d0, best.i0) : KEY_NONE;
    unsigned long long k1 = best.i1 >= 0 ? knn_key(best.d1, best.i1) : KEY_NONE;
    unsigned long long b0 = warp_min_key(k0);
    unsigned long long b1 = warp_min_key(k0 == b0 && b0 != KEY_NONE ? k1 : k0);
    if (lane == 0) s_part[warp] = knn2_from_keys(b0, b1);
    __syncthreads();
    if (warp == 0) {
      Knn2 p{FLT_MAX, FLT_MAX, -1, -1};
      if (lane < EX_THREADS / 32) p = s_part[lane];
      k0 = p.i0 >= 0 ? knn_key(p.d0, p.i0) : KEY_NONE;
      k1 = p.i1 >= 0 ? knn_key(p.d1, p.i1) : KEY_NONE;
      b0 = warp_min_key(k0);
      b1 = warp_min_key(k0 == b0 && b0 != KEY_NONE ? k1 : k0);
      if (lane == 0) {
        a.ex_part[(size_t)f * MATCH_EX_SLICES + blockIdx.x] = knn2_from_keys(b0, b1);
        __threadfence();
        s_last = atomicAdd(&a.ex_done[f], 1) == MATCH_EX_SLICES - 1;
      }
      __syncwarp();
      if (s_last) {  // warp-uniform (shared): this block saw every other slice's partial
        __threadfence();
        Knn2 p2{FLT_MAX, FLT_MAX, -1, -1};
        if (lane < MATCH_EX_SLICES) p2 = a.ex_part[(size_t)f * MATCH_EX_SLICES + lane];
        k0 = p2.i0 >= 0 ? knn_key(p2.d0, p2.i0) : KEY_NONE;
        k1 = p2.i1 >= 0 ? knn_key(p2.d1, p2.i1) : KEY_NONE;
        b0 = warp_min_key(k0);
        b1 = warp_min_key(k0 == b0 && b0 != KEY_NONE ? k1 : k0);
        if (lane == 0) {
          a.knn[fq] = knn2_from_keys(b0, b1);
          a.ex_done[f] = 0;  // self-cleaning for the next call
        }
      }
    }
    __syncthreads();
  }
}

// diagnostics route (MatchArgs::exact_only): every query is flagged, k_knn_exact does all the work
__global__ void __launch_bounds__(256) k_knn_flag_all(const __grid_constant__ MatchArgs a) {
  const int nq = a.nq_dev ? min(*a.nq_dev, a.nq) : a.nq;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < nq; i += gridDim.x * 256) a.fb_list[i] = i;
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.n_flagged = nq;
}

// ------------------------------------------------------------------------------------------------ k_knn_compact
// ratio test + order-preserving compaction (single block; thread t owns a contiguous run of queries)
__global__ void __launch_bounds__(1024) k_knn_compact(const __grid_constant__ MatchArgs a) {
  __shared__ int s_warp[32];
  const int nq = a.nq_dev ? min(*a.nq_dev, a.nq) : a.nq;
  const int nt = a.nt_dev ? min(*a.nt_dev, a.nt) : a.nt;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = (nq + 1023) / 1024;  // <= 32 (capacity <= 32768)
  const int q0 = tid * per, q1 = min(nq, q0 + per);
  unsigned keep = 0;
  for (int qi = q0; qi < q1; qi++) {
    const Knn2 b = a.knn[qi];
    // reference: knn[i][0].distance < ratio * knn[i][1].distance; fewer than 2 train rows => no match (the
    // reference would read out of bounds there)
    bool ok = nt >= 2 && b.i1 >= 0 && b.d0 < __fmul_rn(a.ratio, b.d1);
    if (ok && a.gate_kq) {
      const uvo_keypoint kq = a.gate_kq[qi], kt = a.gate_kt[b.i0];
      const float dy = fabsf(__fsub_rn(kq.y, kt.y)), disp = __fsub_rn(kq.x, kt.x);
      ok = dy <= a.gate_dy && disp >= a.gate_dmin && disp <= a.gate_dmax;
    }
    if (ok) keep |= 1u << (qi - q0);
  }
  const int cnt = __popc(keep);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    s_warp[lane] = w;  // inclusive totals
  }
  __syncthreads();
  int off = (wid > 0 ? s_warp[wid - 1] : 0) + incl - cnt;
  for (int qi = q0; qi < q1; qi++)
    if (keep >> (qi - q0) & 1u) {
      const Knn2 b = a.knn[qi];
      a.matches[off++] = uvo_dmatch{qi, b.i0, 0, b.d0};
    }
  if (tid == 0) {
    *a.n_matches = s_warp[31];
    *a.n_fallback += *a.n_flagged;  // statistics; then reset the per-call state for the next call on this scratch
    *a.n_flagged = 0;
    *a.tn2max = 0u;
  }
}

// ------------------------------------------------------------------------------------------------ host side
// rows x dim f32 row-major, boxes of 128 rows x 32 floats (one SWIZZLE_128B atom); rows past the end read as zero
static CUtensorMap make_desc_map(const float* base, int rows, int dim) {
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)dim, (cuuint64_t)(rows > 0 ? rows : 1)};
  const cuuint64_t strides[1] = {(cuuint64_t)dim * sizeof(float)};
  const cuuint32_t box[2] = {32, TILE};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw InvalidArg{"cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")", UVO_ERR_CUDA};
  return m;
}

static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

size_t match_scratch_bytes(int cap_q, int cap_t) {
  return align256((size_t)cap_q * MATCH_MAX_LISTS * TOPK * sizeof(float)) + align256((size_t)cap_t * sizeof(float)) +
         align256((size_t)cap_q * sizeof(Knn2)) + 2 * align256((size_t)cap_q * sizeof(int)) +
         align256((size_t)cap_q * MATCH_EX_SLICES * sizeof(Knn2)) + 256;
}

void match_bind_scratch(MatchArgs& a, void* scratch, int cap_q, int cap_t) {
  uint8_t* p = (uint8_t*)scratch;
  a.cand = (float*)p;
  p += align256((size_t)cap_q * MATCH_MAX_LISTS * TOPK * sizeof(float));
  a.hb = (float*)p;
  p += align256((size_t)cap_t * sizeof(float));
  a.knn = (Knn2*)p;
  p += align256((size_t)cap_q * sizeof(Knn2));
  a.fb_list = (int*)p;
  p += align256((size_t)cap_q * sizeof(int));
  a.ex_done = (int*)p;
  p += align256((size_t)cap_q * sizeof(int));
  a.ex_part = (Knn2*)p;
  p += align256((size_t)cap_q * MATCH_EX_SLICES * sizeof(Knn2));
  a.n_flagged = (int*)p;
  a.n_fallback = (int*)(p + 16);
  a.tn2max = (unsigned*)(p + 32);
}

template <int D>
static void launch_match_dim(Ctx& c, const MatchArgs& a) {
  {
    // function attributes are per device (one process may hold contexts on several GPUs) and contexts may be driven
    // from several host threads
    static std::mutex attr_mutex;
    static bool attr_set[64] = {};
    std::lock_guard<std::mutex> lock(attr_mutex);
    if (c.device >= 64 || !attr_set[c.device]) {
      UVO_CUDA(cudaFuncSetAttribute(k_knn_tc<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcGeom<D>::SMEM_ALLOC));
      if (c.device < 64) attr_set[c.device] = true;
    }
  }
  if (a.exact_only) {
    UVO_KERNEL(c, "k_knn_flag_all");
    k_knn_flag_all<<<std::min(div_up(a.nq, 256), c.sm_count), 256, 0, c.stream>>>(a);
    UVO_LAUNCH_CHECK(c);
  } else if (a.nt > 0) {
    UVO_KERNEL(c, "k_knn_prep");
    k_knn_prep<D><<<std::min(div_up(a.nt, 16), 2 * c.sm_count), 256, 0, c.stream>>>(a);
    UVO_LAUNCH_CHECK(c);
    const CUtensorMap mq = make_desc_map(a.q, a.nq, D), mt = make_desc_map(a.t, a.nt, D);
    UVO_KERNEL(c, "k_knn_tc");
    k_knn_tc<D><<<dim3(div_up(a.nq, TILE), MATCH_MAX_CHUNKS), TC_THREADS, TcGeom<D>::SMEM_ALLOC, c.stream>>>(mq, mt, a,
                                                                                                            c.sm_count);
    UVO_LAUNCH_CHECK(c);
  }
  if (!a.exact_only) {
    UVO_KERNEL(c, "k_knn_rerank");
    k_knn_rerank<D><<<std::min(div_up(a.nq, RR_WARPS), 8 * c.sm_count), RR_WARPS * 32, 0, c.stream>>>(a, c.sm_count);
    UVO_LAUNCH_CHECK(c);
  }
  UVO_KERNEL(c, "k_knn_exact");
  k_knn_exact<D><<<dim3(MATCH_EX_SLICES, 16), EX_THREADS, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
  UVO_KERNEL(c, "k_knn_compact");
  k_knn_compact<<<1, 1024, 0, c.stream>>>(a);
  UVO_LAUNCH_CHECK(c);
}

void launch_match(Ctx& c, const MatchArgs& a) {
  if (a.nq <= 0) {
    UVO_CUDA(cudaMemsetAsync(a.n_matches, 0, sizeof(int), c.stream));
    return;
  }
  UVO_REQUIRE(a.nt <= MATCH_MAX_CHUNKS * MAX_CHUNK_TILES * TILE && a.nq <= 32768,
              "matcher: more than 32768 descriptors in one set");
  UVO_REQUIRE(((uintptr_t)a.q & 15) == 0 && ((uintptr_t)a.t & 15) == 0, "matcher: descriptors must be 16-byte aligned");
  if (a.dim == 64)
    launch_match_dim<64>(c, a);
  else if (a.dim == 128)
    launch_match_dim<128>(c, a);
  else
    throw InvalidArg{"matcher: descriptor rows must be 64 (SURF) or 128 (extended SURF) floats", UVO_ERR_UNSUPPORTED};
}

}  // namespace uvo
