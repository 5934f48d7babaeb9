// Top-k building blocks of the search path: transposing warp reduction, per-warp replace-worst
// candidate lists, chunked bitonic selection.  All work on the packed (score, id) keys of
// common.cuh, so one unsigned 64-bit compare carries the whole order (score desc, id asc).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr unsigned FULL = 0xffffffffu;

// ---- transposing reduction ------------------------------------------------------------------
// Each lane holds V partial sums a[0..V).  On return a[0] of lane L is the warp-wide total of
// value index (L >> (5 - log2 V)): 31 shuffles for V = 32 instead of 160.
template <int CNT, int OFF>
struct TransposeStep {
  template <int V>
  static __device__ __forceinline__ void run(float (&a)[V], int lane) {
    if constexpr (OFF >= 1) {
      if constexpr (CNT > 1) {
        constexpr int H = CNT / 2;
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < H; i++) {
          const float send = up ? a[i] : a[i + H];
          const float keep = up ? a[i + H] : a[i];
          a[i] = keep + __shfl_xor_sync(FULL, send, OFF);
        }
        TransposeStep<H, OFF / 2>::run(a, lane);
      } else {
        a[0] += __shfl_xor_sync(FULL, a[0], OFF);
        TransposeStep<1, OFF / 2>::run(a, lane);
      }
    }
  }
};
template <int V>
__device__ __forceinline__ void warp_transpose_reduce(float (&a)[V], int lane) {
  static_assert(V == 1 || V == 2 || V == 4 || V == 8 || V == 16 || V == 32, "V must be a power of two <= 32");
  TransposeStep<V, 16>::run(a, lane);
}
template <int V>
struct Log2;
template <> struct Log2<1> { static constexpr int v = 0; };
template <> struct Log2<2> { static constexpr int v = 1; };
template <> struct Log2<4> { static constexpr int v = 2; };
template <> struct Log2<8> { static constexpr int v = 3; };
template <> struct Log2<16> { static constexpr int v = 4; };
template <> struct Log2<32> { static constexpr int v = 5; };

// ---- per-warp candidate list (replace-worst) ---------------------------------------------------
// `list` holds k keys (shared memory, private to the warp), initialised to 0.  `worst`/`worst_pos`
// are warp-uniform registers caching the minimum key and where it sits.  The whole warp calls
// this with a warp-uniform candidate.
__device__ __forceinline__ void warp_list_insert(unsigned long long* list, int k, unsigned long long cand,
                                                 unsigned long long& worst, int& worst_pos, int lane) {
  if (cand <= worst) return;
  if (lane == 0) list[worst_pos] = cand;
  __syncwarp();
  unsigned long long m = ~0ull;
  int mp = 0x7fffffff;
  for (int j = lane; j < k; j += 32) {
    const unsigned long long t = list[j];
    if (t < m) { m = t; mp = j; }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const unsigned long long om = __shfl_xor_sync(FULL, m, o);
    const int op = __shfl_xor_sync(FULL, mp, o);
    if (om < m || (om == m && op < mp)) { m = om; mp = op; }
  }
  worst = m;
  worst_pos = mp;
  __syncwarp();
}

// ---- block-wide bitonic sort, descending, C a power of two --------------------------------------
__device__ __forceinline__ void block_bitonic_sort_desc(unsigned long long* buf, int C) {
  for (int size = 2; size <= C; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (C >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = buf[lo], b = buf[hi];
        if ((a < b) == up) { buf[lo] = b; buf[hi] = a; }
      }
      __syncthreads();
    }
  }
}

}  // namespace b200
