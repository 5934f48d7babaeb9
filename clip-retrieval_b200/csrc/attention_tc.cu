// K4 on the 5th-gen tensor cores: softmax(Q K^T * hd^-0.5 [+ causal]) V for the
// small-T regime of CLIP (T = 50 / 77 / 257, head dim 64).  One work item = one (sample, head):
//
//   TMA  : K rows of the head (128-row boxes out of the fused qkv buffer) and V^T (64x64 boxes of the
//          [head_dim, keys] buffer the QKV GEMM epilogue writes) once per item; Q in 128-row tiles,
//          double buffered.  Everything lands K-major with the 128-byte swizzle.
//   MMA  : S = Q K^T as tcgen05.mma 128 x (256 + tail) x 16 into TMEM columns [0, 320);
//          O = P V as 128 x 64 x 16 steps over the keys into TMEM columns [320, 384).
//   soft : 4 warps, thread = query row: two passes over the fp32 scores in TMEM (row max, then
//          exp2 / row sum), P written as bf16 straight into the swizzled K-major shared-memory tile
//          the second MMA reads; the output is normalised by 1/l when O is read back.
// All keys of a head fit one S tile (T <= 320), so there is no online rescaling.
// Roofline: MUFU ex2 (T^2 exponentials per head) and the tensor pipe; 4*T^2*hd flops per head.
#include "embed_kernels.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int AT_HD = 64;
constexpr int AT_MAXKEYS = 320;                 // 5 key blocks of 64
constexpr int AT_Q_BYTES = 128 * AT_HD * 2;     // 16 KB
constexpr int AT_K_BYTES = 3 * 128 * AT_HD * 2; // up to 384 key rows
constexpr int AT_VT_BYTES = 3 * 128 * AT_HD * 2; // V as 3 boxes of [128 keys x 64 d] (or V^T: 5 boxes of [64 d x 64 keys])
constexpr int AT_P_BYTES = 5 * 128 * 64 * 2;    // 5 key blocks of [128 rows x 64 keys]
constexpr int AT_SMEM = 2 * AT_Q_BYTES + AT_K_BYTES + AT_VT_BYTES + AT_P_BYTES + 2048 + 256 + 1024;
constexpr int AT_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 softmax (two per SM sub-partition)
constexpr int AT_O_COL = 320;

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVt,
                    __nv_bfloat16* __restrict__ out, int B, int T, int heads, int w, float scale_log2e, int causal,
                    int v_direct) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* base = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = base;                       // [2][16 KB]
  uint8_t* sK = sQ + 2 * AT_Q_BYTES;        // [<=384 rows][128 B]
  uint8_t* sVt = sK + AT_K_BYTES;           // [5][64 rows][128 B]
  uint8_t* sP = sVt + AT_VT_BYTES;          // [5][128 rows][128 B]
  float* s_xm = reinterpret_cast<float*>(sP + AT_P_BYTES);   // [2][128] partial row maxima of the two column halves
  float* s_xl = s_xm + 256;                                  // [2][128] partial row sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_xl + 256);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_free = bars + 1;
  uint64_t* q_full = bars + 2;    // [2]
  uint64_t* q_empty = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* o_empty = bars + 9;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int keys_pad = (T + 15) / 16 * 16;         // keys covered by the MMAs
  const int n1 = keys_pad < 256 ? keys_pad : 256;  // first S MMA
  const int n2 = keys_pad - n1;                    // tail S MMA (0, 16, .., 64)
  const int k_boxes = (keys_pad + 127) / 128;
  const int v_boxes = (keys_pad + 63) / 64;
  const int q_tiles = (T + 127) / 128;
  const int items = B * heads;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmQK);
    ptx::prefetch_tensormap(&tmVt);
  }
  if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_init(kv_full, 1);
      ptx::mbar_init(kv_free, 1);
      for (int i = 0; i < 2; i++) {
        ptx::mbar_init(&q_full[i], 1);
        ptx::mbar_init(&q_empty[i], 1);
      }
      ptx::mbar_init(s_full, 1);
      ptx::mbar_init(p_full, 8);
      ptx::mbar_init(o_full, 1);
      ptx::mbar_init(o_empty, 8);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(s_tmem, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t it = 0, tc = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, it++) {
        const int b = item / heads, h = item - b * heads;
        ptx::mbar_wait(kv_free, (it & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(kv_full, (uint32_t)(k_boxes * 128 * 128 + (v_direct ? k_boxes * 128 * 128 : v_boxes * 64 * 128)));
        for (int i = 0; i < k_boxes; i++)
          ptx::tma_load_2d(sK + i * 128 * 128, &tmQK, kv_full, w + h * AT_HD, b * T + i * 128);
        if (v_direct) {
          // V rows straight from the fused qkv buffer: [keys x 64 d], d contiguous = MN-major B operand
          for (int i = 0; i < k_boxes; i++)
            ptx::tma_load_2d(sVt + i * 128 * 128, &tmQK, kv_full, 2 * w + h * AT_HD, b * T + i * 128);
        } else {
          for (int i = 0; i < v_boxes; i++)
            ptx::tma_load_2d(sVt + i * 64 * 128, &tmVt, kv_full, i * 64, (b * heads + h) * AT_HD);
        }
        for (int mt = 0; mt < q_tiles; mt++, tc++) {
          const int buf = tc & 1;
          ptx::mbar_wait(&q_empty[buf], ((tc >> 1) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&q_full[buf], AT_Q_BYTES);
          ptx::tma_load_2d(sQ + buf * AT_Q_BYTES, &tmQK, &q_full[buf], h * AT_HD, b * T + mt * 128);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc_s1 = ptx::umma_idesc_f16(128, n1, true);
      const uint32_t idesc_s2 = ptx::umma_idesc_f16(128, n2 > 0 ? n2 : 16, true);
      // P.V: B = V^T tile, K-major (keys contiguous), or V itself as an MN-major operand (bit 16 of the
      // instruction descriptor): rows of 64 d = 128 bytes, 8-key swizzle atoms 1024 bytes apart.
      const uint32_t idesc_o = ptx::umma_idesc_f16(128, AT_HD, true) | (v_direct ? (1u << 16) : 0u);
      uint32_t it = 0, tc = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, it++) {
        ptx::mbar_wait(kv_full, it & 1);
        for (int mt = 0; mt < q_tiles; mt++, tc++) {
          const int buf = tc & 1;
          ptx::mbar_wait(&q_full[buf], (tc >> 1) & 1);
          ptx::tc_fence_after();
          // S = Q K^T  (the previous tile's scores were consumed before its P.V was issued)
          const uint64_t dq = ptx::umma_desc_k_sw128(ptx::smem_u32(sQ + buf * AT_Q_BYTES));
          const uint64_t dk1 = ptx::umma_desc_k_sw128(ptx::smem_u32(sK));
          const uint64_t dk2 = ptx::umma_desc_k_sw128(ptx::smem_u32(sK + 256 * 128));
#pragma unroll
          for (int k = 0; k < AT_HD / 16; k++) {
            ptx::umma_f16(tmem_base, dq + (uint64_t)(k * 2), dk1 + (uint64_t)(k * 2), idesc_s1, k != 0 ? 1u : 0u);
            if (n2 > 0)
              ptx::umma_f16(tmem_base + 256, dq + (uint64_t)(k * 2), dk2 + (uint64_t)(k * 2), idesc_s2, k != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&q_empty[buf]);
          ptx::umma_commit(s_full);
          // O = P V
          ptx::mbar_wait(p_full, tc & 1);
          ptx::mbar_wait(o_empty, (tc & 1) ^ 1);
          ptx::tc_fence_after();
          for (int ks = 0; ks < keys_pad / 16; ks++) {
            const uint64_t dp = ptx::umma_desc_k_sw128(ptx::smem_u32(sP + (ks >> 2) * (128 * 128))) + (uint64_t)((ks & 3) * 2);
            const uint64_t dv = v_direct
                ? ptx::umma_desc_k_sw128(ptx::smem_u32(sVt + ks * 16 * 128))     // 16 keys = 2 atoms of 8 rows
                : ptx::umma_desc_k_sw128(ptx::smem_u32(sVt + (ks >> 2) * (64 * 128))) + (uint64_t)((ks & 3) * 2);
            ptx::umma_f16(tmem_base + AT_O_COL, dp, dv, idesc_o, ks != 0 ? 1u : 0u);
          }
          ptx::umma_commit(o_full);
          if (mt == q_tiles - 1) ptx::umma_commit(kv_free);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- softmax + output (warps 2..9) ----------------
    // thread = query row of the tile; the two warps that share a TMEM lane quarter split the key
    // columns in halves and exchange their partial row max / row sum through shared memory.
    const int q4 = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q4 * 32 + lane;                  // row inside the tile
    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int chunks = (keys_pad + 31) / 32;
    const int c_mid = (chunks + 1) / 2;
    const int c_beg = half ? c_mid : 0, c_end = half ? chunks : c_mid;
    uint32_t tc = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int b = item / heads, h = item - b * heads;
      for (int mt = 0; mt < q_tiles; mt++, tc++) {
        const int qrow = mt * 128 + r;             // query index inside the sample
        const int kmax = causal ? (qrow < T ? qrow : T - 1) : T - 1;  // last visible key
        ptx::mbar_wait(s_full, tc & 1);
        ptx::tc_fence_after();
        // pass 1: row maximum over the visible keys (key <= kmax) of this warp's columns
        float m = -INFINITY;
#pragma unroll 1
        for (int c = c_beg; c < c_end; c++) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(lane_base + c * 32, v);
          ptx::tmem_ld_wait();
          const int lim = kmax - c * 32;   // columns j <= lim are visible
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            m0 = fmaxf(m0, j + 0 <= lim ? __uint_as_float(v[j + 0]) : -INFINITY);
            m1 = fmaxf(m1, j + 1 <= lim ? __uint_as_float(v[j + 1]) : -INFINITY);
            m2 = fmaxf(m2, j + 2 <= lim ? __uint_as_float(v[j + 2]) : -INFINITY);
            m3 = fmaxf(m3, j + 3 <= lim ? __uint_as_float(v[j + 3]) : -INFINITY);
          }
          m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
        }
        s_xm[half * 128 + r] = m;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        m = fmaxf(s_xm[r], s_xm[128 + r]);
        // pass 2: p = 2^((s - m) * scale*log2e) (ex2.approx, arguments <= 0), row sum, bf16 P into the
        // swizzled K-major tile
        float l0 = 0.f, l1 = 0.f;
        const float mb = m * scale_log2e;
#pragma unroll 1
        for (int c = c_beg; c < c_end; c++) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(lane_base + c * 32, v);
          ptx::tmem_ld_wait();
          const int lim = kmax - c * 32;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float p0, p1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v[j]), scale_log2e, -mb)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v[j + 1]), scale_log2e, -mb)));
            p0 = j <= lim ? p0 : 0.f;
            p1 = j + 1 <= lim ? p1 : 0.f;
            l0 += p0;
            l1 += p1;
            pk[j >> 1] = pack_bf16x2(p0, p1);
          }
          // 32 keys = 4 chunks of 16 bytes inside key block (c >> 1), chunk index ((c & 1) * 4 + i) ^ (r & 7)
          uint8_t* blk = sP + (c >> 1) * (128 * 128) + r * 128;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int ch = (((c & 1) * 4 + i) ^ (r & 7)) * 16;
            *reinterpret_cast<uint4*>(blk + ch) = make_uint4(pk[i * 4], pk[i * 4 + 1], pk[i * 4 + 2], pk[i * 4 + 3]);
          }
        }
        s_xl[half * 128 + r] = l0 + l1;
        ptx::fence_proxy_async();   // generic-proxy writes of P -> visible to the tensor core
        ptx::tc_fence_before();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (lane == 0) ptx::mbar_arrive(p_full);
        const float l = s_xl[r] + s_xl[128 + r];
        // output: each half normalises and stores 32 of the 64 head columns
        ptx::mbar_wait(o_full, tc & 1);
        ptx::tc_fence_after();
        uint32_t o0[32];
        ptx::tmem_ld_32x32b_x32(lane_base + AT_O_COL + half * 32, o0);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(o_empty);
        if (qrow < T) {
          const float inv = 1.0f / l;
          __nv_bfloat16* op = out + ((size_t)b * T + qrow) * w + (size_t)h * AT_HD + half * 32;
#pragma unroll
          for (int g = 0; g < 4; g++) {
            uint4 a;
            a.x = pack_bf16x2(__uint_as_float(o0[g * 8 + 0]) * inv, __uint_as_float(o0[g * 8 + 1]) * inv);
            a.y = pack_bf16x2(__uint_as_float(o0[g * 8 + 2]) * inv, __uint_as_float(o0[g * 8 + 3]) * inv);
            a.z = pack_bf16x2(__uint_as_float(o0[g * 8 + 4]) * inv, __uint_as_float(o0[g * 8 + 5]) * inv);
            a.w = pack_bf16x2(__uint_as_float(o0[g * 8 + 6]) * inv, __uint_as_float(o0[g * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(op + g * 8) = a;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

bool attention_tc_supported(int T, int heads, int w) {
  return heads > 0 && w % heads == 0 && w / heads == AT_HD && T >= 1 && T <= AT_MAXKEYS;
}

// qkv: [B*T, 3w] (Q and K thirds are read); vt: [B*heads*64, Tp] (V^T, keys contiguous, zero padded).
int attention_tc(const CUtensorMap& tmQK, const CUtensorMap& tmVt, __nv_bfloat16* out, int B, int T, int heads, int w,
                 int causal, int v_direct, int sms, cudaStream_t st) {
  B200_CHECK(attention_tc_supported(T, heads, w), B200_ERR_UNSUPPORTED, "attention_tc: unsupported shape T=%d hd=%d", T,
             heads ? w / heads : 0);
  if (B == 0) return B200_OK;
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!(configured.load() >> (dev & 63) & 1ull)) {
    B200_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    configured.fetch_or(1ull << (dev & 63));
  }
  const float scale_log2e = (1.0f / sqrtf((float)AT_HD)) * 1.4426950408889634f;
  const int items = B * heads;
  const int grid = items < sms ? items : sms;
  attention_tc_kernel<<<grid, AT_THREADS, AT_SMEM, st>>>(tmQK, tmVt, out, B, T, heads, w, scale_log2e, causal, v_direct);
  B200_LAUNCH_OK();
  return B200_OK;
}

}  // namespace b200
