// K4 on tcgen05, third generation (production attention for head dim 64, T <= 264: every CLIP tower at 224 px
// except H/14's vision tower).  Same skeleton as attention_tc2.cu — two query tiles in flight, S and O in TMEM, P
// through swizzled shared memory, V as an MN-major operand — with the three things its profile asked for
// (profiles/r01f: tensor 14 %, XU 30 %, two TMEM read passes ~ 64 B/clk/SM):
//
//  1. ONE pass over the scores.  Softmax is shift invariant, and fp32 / bf16 carry the same relative precision at
//     every magnitude, so the subtracted value need not be the row maximum: any m with  max_j s_ij <= m  cannot
//     overflow, and as long as m - max_j s_ij stays below ~64 (log2 units) every term that matters is far from the
//     flush-to-zero limit.  m_i = |q_i| * max_j |k_j| * scale (Cauchy-Schwarz) is such a bound, costs one 64-term
//     norm per row, and is known BEFORE the scores exist.  The pass also tracks the true row maximum; if a row's
//     slack exceeds 64 the warp repeats the pass with the exact maximum (the scores are still in TMEM): results are
//     those of the two-pass softmax in every case, the second read happens only for pathological rows.
//  2. T = 257 = 2 * 128 + 1 is two tensor-core tiles plus ONE row: the leftover rows (T mod 128 <= 4) are computed on
//     the FMA pipe by the 128 threads of the softmax group that owns the head's last tile, straight from the K / V
//     tiles in shared memory (scores before the tile's own softmax, P.V after its output), instead of a third
//     128-row tile that is 99 % padding.  (A first version gave them to one dedicated warp that held K and V until
//     it was done: 2x slower than attention_tc2 — profiles/r02a.)
//  3. tcgen05.ld of score chunk c+1 is in flight while chunk c is exponentiated (two register buffers).
// Per-sample 3-D tensor maps (rows >= T of a box are zero-filled) keep a sample's boxes from reading its neighbour.
//
//   warp 8      TMA: K, V rows of the head (128-row boxes) once per (sample, head), Q tile per 128 query rows.
//   warp 9      tcgen05.mma issuer: S(t) = Q K^T, O(t) = P V; order S(t), PV(t-1), S(t+1), PV(t), ...
//   (before)    attn_kmax_kernel: max_j |k_j| per (sample, head) for the softmax bound, one warp per head.
//   warps 0..3  softmax group 0 (even tiles), warps 4..7 group 1 (odd tiles): thread = query row.
#include "embed_kernels.cuh"
#include "gemm.cuh"
#include "ptx.cuh"
#include "topk.cuh"
#include <cstdlib>

namespace b200 {

constexpr int A3_HD = 64;
constexpr int A3_MAXT = 264;                      // 256 keys on the tensor core + up to 8 extra keys
constexpr int A3_Q_BYTES = 128 * 128;             // 16 KB
constexpr int A3_KV_MAIN = 2 * 128 * 128;         // 256 rows x 128 B
constexpr int A3_P_BYTES = 4 * 128 * 128;         // 4 key blocks of [128 rows x 64 keys] per group
constexpr int A3_TAIL_MAX = 4;                    // leftover query rows (T mod 128) handled on the FMA pipe
constexpr int A3_TAIL_BYTES = (A3_TAIL_MAX * 264 + 4 * 64 + 16) * 4;   // per group: p[rows][264], partial O [4][64], reductions
constexpr int A3_SMEM = A3_Q_BYTES + 2 * A3_KV_MAIN + 2 * A3_P_BYTES + 2 * A3_TAIL_BYTES + 512 + 1024;
constexpr int A3_THREADS = 320;
constexpr float A3_MAX_SLACK = 64.0f;             // log2 units; above it the pass is repeated with the exact maximum

// barrier over the 128 threads of one softmax group (named barriers 1 and 2)
__device__ __forceinline__ void a3_group_sync(int grp) {
  if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}

// 8 consecutive bf16 of a swizzled K/V row in shared memory -> fp32
__device__ __forceinline__ void a3_unpack8(const uint4& u, float* f) {
  const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
  f[0] = a0.x; f[1] = a0.y; f[2] = a1.x; f[3] = a1.y; f[4] = a2.x; f[5] = a2.y; f[6] = a3.x; f[7] = a3.y;
}

__global__ void __launch_bounds__(A3_THREADS, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tm3, const __nv_bfloat16* __restrict__ qkv,
                     __nv_bfloat16* __restrict__ out, const float* __restrict__ kmax_head, int B, int T, int heads, int w,
                     float scale_log2e, int causal, int onepass, int tail_external) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* base = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = base;
  uint8_t* sK = sQ + A3_Q_BYTES;               // rows 0..255
  uint8_t* sV = sK + A3_KV_MAIN;               // rows 0..255
  uint8_t* sP = sV + A3_KV_MAIN;               // [2 groups][4 key blocks][128 rows][128 B]
  float* sTail = reinterpret_cast<float*>(sP + 2 * A3_P_BYTES);   // [2 groups][A3_TAIL_BYTES / 4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sTail) + 2 * A3_TAIL_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* k_free = bars + 3;    // MMA commit after the last S of the head
  uint64_t* v_full = bars + 4;
  uint64_t* v_free = bars + 5;    // MMA commit after the last P.V of the head
  uint64_t* s_full = bars + 6;    // [2]
  uint64_t* p_full = bars + 8;    // [2]
  uint64_t* o_full = bars + 10;   // [2]
  uint64_t* buf_free = bars + 12; // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int keys_main = T < 256 ? (T + 15) / 16 * 16 : 256;  // keys on the tensor core (multiple of 16)
  const int extra = T > 256 ? T - 256 : 0;                    // keys handled on the FMA pipe
  const int kv_boxes = (keys_main + 127) / 128;
  // query rows: full/partial 128-row tiles on the tensor core, a leftover of <= 8 rows (after at least one full
  // tile) on the FMA pipe
  const int n_full = T / 128, rem = T - n_full * 128;
  const bool tail_rows = rem > 0 && rem <= A3_TAIL_MAX && n_full >= 1;
  const int q_tiles = tail_rows ? n_full : (T + 127) / 128;
  const int fma_rows = (tail_rows && !tail_external) ? rem : 0;   // leftover rows done by attention_tail_kernel otherwise
  const int items = B * heads;

  if (warp == 8 && lane == 0) ptx::prefetch_tensormap(&tm3);
  if (warp == 9) {
    if (lane == 0) {
      ptx::mbar_init(q_full, 1); ptx::mbar_init(q_empty, 1);
      ptx::mbar_init(k_full, 1); ptx::mbar_init(k_free, 1);
      ptx::mbar_init(v_full, 1); ptx::mbar_init(v_free, 1);
      for (int i = 0; i < 2; i++) {
        ptx::mbar_init(&s_full[i], 1);
        ptx::mbar_init(&p_full[i], 4);
        ptx::mbar_init(&o_full[i], 1);
        ptx::mbar_init(&buf_free[i], 4);
      }
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

  if (warp == 8) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t it = 0, tc = 0;
      const uint32_t kv_bytes = (uint32_t)(kv_boxes * 128 * 128);
      for (int item = blockIdx.x; item < items; item += gridDim.x, it++) {
        const int b = item / heads, h = item - b * heads;
        ptx::mbar_wait(k_free, (it & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(k_full, kv_bytes);
        for (int i = 0; i < kv_boxes; i++) ptx::tma_load_3d(sK + i * 128 * 128, &tm3, k_full, w + h * A3_HD, i * 128, b);
        for (int mt = 0; mt < q_tiles; mt++, tc++) {
          ptx::mbar_wait(q_empty, (tc & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(q_full, A3_Q_BYTES);
          ptx::tma_load_3d(sQ, &tm3, q_full, h * A3_HD, mt * 128, b);
          if (mt == 0) {
            ptx::mbar_wait(v_free, (it & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(v_full, kv_bytes);
            for (int i = 0; i < kv_boxes; i++)
              ptx::tma_load_3d(sV + i * 128 * 128, &tm3, v_full, 2 * w + h * A3_HD, i * 128, b);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc_s = ptx::umma_idesc_f16(128, keys_main, true);
      const uint32_t idesc_o = ptx::umma_idesc_f16(128, A3_HD, true) | (1u << 16);  // B (= V) is MN-major
      // Two cursors over this CTA's tile sequence: the next S = Q K^T and the next O = P V.  Whichever has its inputs
      // ready is issued (non-blocking mbarrier tests): a fixed order S(t+1) before PV(t) made every P.V wait for the
      // NEXT head's K load and serialised the two softmax groups (17.8 % of all stall samples on the o_full spin,
      // profiles/r02d_attn3).  P.V first: it unblocks a softmax group and frees a TMEM buffer.
      const int my_items = blockIdx.x < items ? (items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      const uint32_t total = (uint32_t)my_items * (uint32_t)q_tiles;
      uint32_t s_tc = 0, s_it = 0, pv_tc = 0, pv_it = 0;
      int s_mt = 0, pv_mt = 0;
      while (pv_tc < total) {
        bool did = false;
        if (pv_tc < s_tc) {
          const int pb = pv_tc & 1;
          if ((pv_mt != 0 || ptx::mbar_test_wait(v_full, pv_it & 1)) && ptx::mbar_test_wait(&p_full[pb], (pv_tc >> 1) & 1)) {
            ptx::tc_fence_after();
            const uint8_t* sPg = sP + pb * A3_P_BYTES;
            for (int ks = 0; ks < keys_main / 16; ks++) {
              const uint64_t dp = ptx::umma_desc_k_sw128(ptx::smem_u32(sPg + (ks >> 2) * (128 * 128))) + (uint64_t)((ks & 3) * 2);
              const uint64_t dv = ptx::umma_desc_k_sw128(ptx::smem_u32(sV + ks * 16 * 128));  // 16 keys = 2 swizzle atoms
              ptx::umma_f16(tmem_base + pb * 256, dp, dv, idesc_o, ks != 0 ? 1u : 0u);
            }
            ptx::umma_commit(&o_full[pb]);
            if (pv_mt == q_tiles - 1) ptx::umma_commit(v_free);
            pv_tc++;
            if (++pv_mt == q_tiles) { pv_mt = 0; pv_it++; }
            did = true;
          }
        }
        if (!did && s_tc < total && s_tc < pv_tc + 2) {
          const int bsel = s_tc & 1;
          if ((s_mt != 0 || ptx::mbar_test_wait(k_full, s_it & 1)) && ptx::mbar_test_wait(q_full, s_tc & 1) &&
              ptx::mbar_test_wait(&buf_free[bsel], ((s_tc >> 1) & 1) ^ 1)) {   // O(s_tc - 2) has been read out of this buffer
            ptx::tc_fence_after();
            const uint64_t dq = ptx::umma_desc_k_sw128(ptx::smem_u32(sQ));
            const uint64_t dk = ptx::umma_desc_k_sw128(ptx::smem_u32(sK));
#pragma unroll
            for (int k = 0; k < A3_HD / 16; k++)
              ptx::umma_f16(tmem_base + bsel * 256, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
            ptx::umma_commit(q_empty);
            ptx::umma_commit(&s_full[bsel]);
            if (s_mt == q_tiles - 1) ptx::umma_commit(k_free);
            s_tc++;
            if (++s_mt == q_tiles) { s_mt = 0; s_it++; }
            did = true;
          }
        }
        // nothing ready: yield the issue slots of this sub-partition to the two softmax warps that share it (a
        // spinning issuer executed 7 M instructions per SM and slowed exactly the warps every tile waits for)
        if (!did) __nanosleep(64);
      }
    }
    __syncwarp();
  } else {
    // ---------------- softmax groups ----------------
    const int grp = warp >> 2;                     // 0: even tiles, 1: odd tiles
    const int q4 = warp & 3;                       // TMEM lane quarter
    const int r = q4 * 32 + lane;                  // row inside the tile
    const uint32_t tbase = tmem_base + grp * 256 + ((uint32_t)(q4 * 32) << 16);
    uint8_t* sPg = sP + grp * A3_P_BYTES;
    float* sTp = sTail + grp * (A3_TAIL_BYTES / 4);      // p of the leftover rows [A3_TAIL_MAX][264]
    float* sTo = sTp + A3_TAIL_MAX * 264;                 // partial O of one leftover row [4 warps][64]
    float* sTr = sTo + 4 * 64;                            // cross-warp reductions [8]
    const int chunks = (keys_main + 31) / 32;
    uint32_t tc = 0, it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, it++) {
      const int b = item / heads, h = item - b * heads;
      for (int mt = 0; mt < q_tiles; mt++, tc++) {
        if ((int)(tc & 1) != grp) continue;
        const uint32_t n = tc >> 1;
        const int qrow = mt * 128 + r;
        const int kmax = causal ? (qrow < T ? qrow : T - 1) : T - 1;   // last visible key
        // ---- before the scores exist: |q|, the extra keys' scores (FMA pipe), the bound m ----
        float se[8];
#pragma unroll
        for (int e = 0; e < 8; e++) se[e] = -INFINITY;
        float qn2 = 0.f;
        if (onepass || extra > 0) {
          float qf[A3_HD];
          const int qr = qrow < T ? qrow : T - 1;
          const uint4* qp = reinterpret_cast<const uint4*>(qkv + ((size_t)b * T + qr) * 3 * w + (size_t)h * A3_HD);
#pragma unroll
          for (int c = 0; c < 8; c++) a3_unpack8(qp[c], qf + c * 8);
#pragma unroll
          for (int c = 0; c < A3_HD; c++) qn2 = fmaf(qf[c], qf[c], qn2);
          for (int e = 0; e < extra; e++) {
            const uint4* kp = reinterpret_cast<const uint4*>(qkv + ((size_t)b * T + 256 + e) * 3 * w + w + (size_t)h * A3_HD);
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 8; c++) {
              float f[8];
              a3_unpack8(__ldg(kp + c), f);
#pragma unroll
              for (int e2 = 0; e2 < 8; e2++) acc = fmaf(qf[c * 8 + e2], f[e2], acc);
            }
#pragma unroll
            for (int ee = 0; ee < 8; ee++)
              if (ee == e) se[ee] = (256 + e <= kmax) ? acc : -INFINITY;
          }
        }
        float mb = 0.f;
        if (onepass) mb = sqrtf(qn2) * 1.0001f * __ldg(kmax_head + item) * scale_log2e;   // >= every score of this row (log2 units)
        // ---- leftover query rows of this head (T mod 128 <= 4): scores and probabilities now, while K is resident ----
        const bool last_tile = mt == q_tiles - 1;
        float tail_l[A3_TAIL_MAX];
        if (last_tile) {
          if (fma_rows > 0) {
            // K and V rows of the head come from the qkv buffer (L2: the TMA boxes of this head were just read from
            // there), not from the shared-memory tiles: nothing here holds the tiles back from the next head
            for (int fr = 0; fr < fma_rows; fr++) {
              const int trow = n_full * 128 + fr;
              const int tkmax = causal ? trow : T - 1;
              float qf[A3_HD];
              const uint4* qp = reinterpret_cast<const uint4*>(qkv + ((size_t)b * T + trow) * 3 * w + (size_t)h * A3_HD);
#pragma unroll
              for (int c = 0; c < 8; c++) a3_unpack8(__ldg(qp + c), qf + c * 8);
              // thread r: keys r, r + 128, r + 256
              float sc[3];
#pragma unroll
              for (int u = 0; u < 3; u++) {
                const int j = r + u * 128;
                float acc = -INFINITY;
                if (j <= tkmax) {
                  acc = 0.f;
                  const uint4* kp = reinterpret_cast<const uint4*>(qkv + ((size_t)b * T + j) * 3 * w + w + (size_t)h * A3_HD);
#pragma unroll
                  for (int c = 0; c < 8; c++) {
                    float f[8];
                    a3_unpack8(__ldg(kp + c), f);
#pragma unroll
                    for (int e = 0; e < 8; e++) acc = fmaf(qf[c * 8 + e], f[e], acc);
                  }
                }
              sc[u] = acc;
              }
              // exact two-pass softmax over the 128 threads of the group
              float mx = fmaxf(fmaxf(sc[0], sc[1]), sc[2]);
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
              if (lane == 0) sTr[q4] = mx;
              a3_group_sync(grp);
              mx = fmaxf(fmaxf(sTr[0], sTr[1]), fmaxf(sTr[2], sTr[3]));
              const float mbt = mx * scale_log2e;
              float ls = 0.f;
#pragma unroll
              for (int u = 0; u < 3; u++) {
                float pv;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv) : "f"(fmaf(sc[u], scale_log2e, -mbt)));
                pv = sc[u] == -INFINITY ? 0.f : pv;
                ls += pv;
                const int j = r + u * 128;
                // P is rounded to bf16 before it multiplies V, as on the tensor-core rows; the row sum keeps fp32
                if (j < 264) sTp[fr * 264 + j] = __bfloat162float(__float2bfloat16_rn(pv));
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(0xffffffffu, ls, o);
              if (lane == 0) sTr[4 + q4] = ls;
              a3_group_sync(grp);
              tail_l[fr] = (sTr[4] + sTr[5]) + (sTr[6] + sTr[7]);
              a3_group_sync(grp);   // sTr is reused by the next row
            }
          }
        }
        ptx::mbar_wait(&s_full[grp], n & 1);
        ptx::tc_fence_after();

        float l = 0.f, tmax = -INFINITY;
        float pe[8];
        if (!onepass) {
          // two-pass mode: the exact row maximum first (one more read of the scores out of TMEM)
          float m = -INFINITY;
#pragma unroll
          for (int e = 0; e < 8; e++) m = fmaxf(m, se[e]);
#pragma unroll 1
          for (int c = 0; c < chunks; c++) {
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(tbase + c * 32, v);
            ptx::tmem_ld_wait();
            const int lim = kmax - c * 32;
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
          mb = m * scale_log2e;
        }
        // exponentials against mb (the bound, or the exact maximum), row sum, true maximum, P -> swizzled K-major tile.
        // One-pass mode: if the bound sits too far above the true maximum of some row of this warp, the pass is
        // repeated with exact maxima (S is still in TMEM, P is simply rewritten).
        for (int attempt = 0; attempt < 2; attempt++) {
          float l0 = 0.f, l1 = 0.f, t0 = -INFINITY, t1 = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < chunks; c++) {
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(tbase + c * 32, v);
            ptx::tmem_ld_wait();
            const int lim = kmax - c * 32;
            uint32_t pk[16];
            if (lim >= 31) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float s0 = __uint_as_float(v[j]), s1 = __uint_as_float(v[j + 1]);
                float p0, p1;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(s0, scale_log2e, -mb)));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(s1, scale_log2e, -mb)));
                t0 = fmaxf(t0, s0);
                t1 = fmaxf(t1, s1);
                l0 += p0;
                l1 += p1;
                pk[j >> 1] = pack_bf16x2(p0, p1);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float s0 = j <= lim ? __uint_as_float(v[j]) : -INFINITY;
                const float s1 = j + 1 <= lim ? __uint_as_float(v[j + 1]) : -INFINITY;
                float p0, p1;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(s0, scale_log2e, -mb)));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(s1, scale_log2e, -mb)));
                p0 = j <= lim ? p0 : 0.f;
                p1 = j + 1 <= lim ? p1 : 0.f;
                t0 = fmaxf(t0, s0);
                t1 = fmaxf(t1, s1);
                l0 += p0;
                l1 += p1;
                pk[j >> 1] = pack_bf16x2(p0, p1);
              }
            }
            uint8_t* blk = sPg + (c >> 1) * (128 * 128) + r * 128;
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int ch = (((c & 1) * 4 + i) ^ (r & 7)) * 16;
              *reinterpret_cast<uint4*>(blk + ch) = make_uint4(pk[i * 4], pk[i * 4 + 1], pk[i * 4 + 2], pk[i * 4 + 3]);
            }
          }
#pragma unroll
          for (int e = 0; e < 8; e++) {
            float p;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(se[e], scale_log2e, -mb)));
            pe[e] = se[e] == -INFINITY ? 0.f : p;
            l0 += pe[e];
            t0 = fmaxf(t0, se[e]);
          }
          l = l0 + l1;
          tmax = fmaxf(t0, t1);
          if (!onepass) break;
          const float exact = tmax * scale_log2e;
          if (!__any_sync(0xffffffffu, (mb - exact > A3_MAX_SLACK) && tmax != -INFINITY)) break;
          mb = tmax == -INFINITY ? mb : exact;
        }
        ptx::fence_proxy_async();   // P (generic-proxy stores) -> visible to the tensor core
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[grp]);
        // output: O from TMEM (+ the extra keys' p_e * v_e), * 1/l, bf16
        ptx::mbar_wait(&o_full[grp], n & 1);
        ptx::tc_fence_after();
        uint32_t o0[32], o1[32];
        ptx::tmem_ld_32x32b_x32(tbase, o0);
        ptx::tmem_ld_32x32b_x32(tbase + 32, o1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&buf_free[grp]);
        if (qrow < T) {
          float of[A3_HD];
#pragma unroll
          for (int j = 0; j < 32; j++) { of[j] = __uint_as_float(o0[j]); of[32 + j] = __uint_as_float(o1[j]); }
          for (int e = 0; e < extra; e++) {
            float p = 0.f;
#pragma unroll
            for (int ee = 0; ee < 8; ee++) if (ee == e) p = pe[ee];
            p = __bfloat162float(__float2bfloat16_rn(p));   // same rounding as the P that went through the tensor core
            const uint4* vp = reinterpret_cast<const uint4*>(qkv + ((size_t)b * T + 256 + e) * 3 * w + 2 * w + (size_t)h * A3_HD);
#pragma unroll
            for (int c = 0; c < 8; c++) {
              float f[8];
              a3_unpack8(__ldg(vp + c), f);
#pragma unroll
              for (int e2 = 0; e2 < 8; e2++) of[c * 8 + e2] = fmaf(p, f[e2], of[c * 8 + e2]);
            }
          }
          const float inv = 1.0f / l;
          __nv_bfloat16* op = out + ((size_t)b * T + qrow) * w + (size_t)h * A3_HD;
#pragma unroll
          for (int g = 0; g < 8; g++) {
            uint4 a;
            a.x = pack_bf16x2(of[g * 8 + 0] * inv, of[g * 8 + 1] * inv);
            a.y = pack_bf16x2(of[g * 8 + 2] * inv, of[g * 8 + 3] * inv);
            a.z = pack_bf16x2(of[g * 8 + 4] * inv, of[g * 8 + 5] * inv);
            a.w = pack_bf16x2(of[g * 8 + 6] * inv, of[g * 8 + 7] * inv);
            *reinterpret_cast<uint4*>(op + g * 8) = a;
          }
        }
        // ---- leftover query rows: O = P V on the FMA pipe, V still resident; then V is released ----
        if (last_tile) {
          if (fma_rows > 0) {
            const int dp = lane;                              // dims 2 dp, 2 dp + 1
            for (int fr = 0; fr < fma_rows; fr++) {
              const int trow = n_full * 128 + fr;
              const int tkmax = causal ? trow : T - 1;
              const float* pr = sTp + fr * 264;
              float a0 = 0.f, a1 = 0.f;
              // warp q4 of the group: keys 64 q4 .. 64 q4 + 63; warp 3 also the keys >= 256 (V rows from L2, 128 B per warp)
              const uint32_t* vbase = reinterpret_cast<const uint32_t*>(qkv + (size_t)b * T * 3 * w + 2 * w + (size_t)h * A3_HD) + dp;
              const size_t vstride = (size_t)3 * w / 2;       // row pitch in 32-bit words
              const int j0 = q4 * 64;
#pragma unroll 1
              for (int half64 = 0; half64 < 2; half64++) {
                const int jb = j0 + half64 * 32;
                uint32_t vr[32];
#pragma unroll
                for (int jj = 0; jj < 32; jj++)   // every load issued before the first is used: one L2 latency, not 32
                  vr[jj] = (jb + jj <= tkmax) ? __ldg(vbase + (size_t)(jb + jj) * vstride) : 0u;
#pragma unroll
                for (int jj = 0; jj < 32; jj++) {
                  const float2 vv = unpack_bf16x2(vr[jj]);
                  const float pj = (jb + jj <= tkmax) ? pr[jb + jj] : 0.f;
                  a0 = fmaf(pj, vv.x, a0);
                  a1 = fmaf(pj, vv.y, a1);
                }
              }
              if (q4 == 3) {
                for (int j = 256; j <= tkmax; j++) {
                  const float2 vv = unpack_bf16x2(__ldg(vbase + (size_t)j * vstride));
                  const float pj = pr[j];
                  a0 = fmaf(pj, vv.x, a0);
                  a1 = fmaf(pj, vv.y, a1);
                }
              }
              sTo[q4 * 64 + 2 * dp] = a0;
              sTo[q4 * 64 + 2 * dp + 1] = a1;
              a3_group_sync(grp);
              if (q4 == 0) {
                const float inv = 1.0f / tail_l[fr];
                const float o0 = (sTo[2 * dp] + sTo[64 + 2 * dp]) + (sTo[128 + 2 * dp] + sTo[192 + 2 * dp]);
                const float o1 = (sTo[2 * dp + 1] + sTo[64 + 2 * dp + 1]) + (sTo[128 + 2 * dp + 1] + sTo[192 + 2 * dp + 1]);
                *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + trow) * w + (size_t)h * A3_HD + 2 * dp) = pack_bf16x2(o0 * inv, o1 * inv);
              }
              a3_group_sync(grp);   // sTo is reused by the next row
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) ptx::tmem_dealloc(tmem_base, 512);
}

// max_j |k_j| (Euclidean norm of a key row of the head, times a hair for the rounding of the norms) per (sample, head):
// the factor of the softmax bound that does not depend on the query.  One warp per head, the K third of the qkv buffer
// read once (180 MB per ViT-L/14 layer at batch 1024, ~30 us).
__global__ void __launch_bounds__(256)
attn_kmax_kernel(const __nv_bfloat16* __restrict__ qkv, int B, int T, int heads, int w, float* __restrict__ kmax_head) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= B * heads) return;
  const int b = item / heads, h = item - b * heads;
  float kn2 = 0.f;
  for (int j = lane; j < T; j += 32) {
    const uint4* kp = reinterpret_cast<const uint4*>(qkv + ((size_t)b * T + j) * 3 * w + w + (size_t)h * A3_HD);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; c++) {
      float f[8];
      a3_unpack8(__ldg(kp + c), f);
#pragma unroll
      for (int e = 0; e < 8; e++) acc = fmaf(f[e], f[e], acc);
    }
    kn2 = fmaxf(kn2, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kn2 = fmaxf(kn2, __shfl_xor_sync(0xffffffffu, kn2, o));
  if (lane == 0) kmax_head[item] = sqrtf(kn2) * 1.0001f + 1e-30f;
}

// The leftover query rows (T mod 128 <= 4 rows per head; ONE for T = 257) as their own small kernel: one warp per
// (sample, head, row), K and V of the head straight from the qkv buffer.  Inside the tensor-core kernel these rows sat
// on one softmax group's critical path (+0.76 ms per ViT-L/14 layer, more than the padded third tile they replace,
// profiles/r02g_attention_T_sweep.txt); here they run on a second stream next to it — the big kernel leaves the SMs'
// FMA pipes, registers and most of the DRAM bandwidth idle (ncu: issue 42 %, DRAM 21 %).
// Scores: lane = dimension pair, 32 keys at a time, one transposing reduction per 32 keys (lane i ends up with key
// base + i).  P.V: lane = dimension pair again, p_j broadcast by shuffle.  All loads are 128-byte rows.
__global__ void __launch_bounds__(128)
attention_tail_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int B, int T, int heads, int w,
                      float scale_log2e, int causal, int row0, int nrows) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= B * heads * nrows) return;
  const int fr = wid % nrows, item = wid / nrows;
  const int b = item / heads, h = item - b * heads;
  const int trow = row0 + fr;
  const int tkmax = causal ? trow : T - 1;
  const size_t pitch = (size_t)3 * w / 2;                                  // row pitch in 32-bit words
  const uint32_t* base = reinterpret_cast<const uint32_t*>(qkv + (size_t)b * T * 3 * w + (size_t)h * A3_HD) + lane;
  const float2 ql = unpack_bf16x2(__ldg(base + (size_t)trow * pitch));    // this lane's two dimensions of q
  const uint32_t* kb = base + w / 2;
  const uint32_t* vb = base + w;
  constexpr int NB = (A3_MAXT + 31) / 32;                                  // key blocks of 32: 9
  float s[NB];
  float mx = -INFINITY;
#pragma unroll
  for (int blk = 0; blk < NB; blk++) {
    s[blk] = -INFINITY;
    if (blk * 32 <= tkmax) {
      float part[32];
      uint32_t kr[32];
#pragma unroll
      for (int i = 0; i < 32; i++) kr[i] = (blk * 32 + i <= tkmax) ? __ldg(kb + (size_t)(blk * 32 + i) * pitch) : 0u;
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const float2 kk = unpack_bf16x2(kr[i]);
        part[i] = fmaf(ql.x, kk.x, ql.y * kk.y);
      }
      warp_transpose_reduce<32>(part, lane);                               // lane i: score of key blk * 32 + i
      s[blk] = (blk * 32 + lane <= tkmax) ? part[0] : -INFINITY;
      mx = fmaxf(mx, s[blk]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float mb = mx * scale_log2e;
  float l = 0.f;
#pragma unroll
  for (int blk = 0; blk < NB; blk++) {
    float p;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(s[blk], scale_log2e, -mb)));
    p = s[blk] == -INFINITY ? 0.f : p;
    l += p;
    s[blk] = __bfloat162float(__float2bfloat16_rn(p));   // P is rounded to bf16 before it multiplies V, as on the tensor-core rows
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int blk = 0; blk < NB; blk++) {
    if (blk * 32 <= tkmax) {
      uint32_t vr[32];
#pragma unroll
      for (int i = 0; i < 32; i++) vr[i] = (blk * 32 + i <= tkmax) ? __ldg(vb + (size_t)(blk * 32 + i) * pitch) : 0u;
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const float pj = __shfl_sync(0xffffffffu, s[blk], i);
        const float2 vv = unpack_bf16x2(vr[i]);
        o0 = fmaf(pj, vv.x, o0);
        o1 = fmaf(pj, vv.y, o1);
      }
    }
  }
  const float inv = 1.0f / l;
  *reinterpret_cast<uint32_t*>(out + ((size_t)b * T + trow) * w + (size_t)h * A3_HD + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
}

int attention_tail_rows(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int heads, int w, int causal, int row0,
                        int nrows, cudaStream_t st) {
  B200_CHECK(heads > 0 && w / heads == A3_HD && T <= A3_MAXT && nrows >= 1 && row0 + nrows <= T, B200_ERR_UNSUPPORTED,
             "attention_tail_rows: unsupported shape T=%d hd=%d rows %d..%d", T, heads ? w / heads : 0, row0, row0 + nrows);
  if (B == 0) return B200_OK;
  const float scale_log2e = (1.0f / sqrtf((float)A3_HD)) * 1.4426950408889634f;
  const unsigned nw = (unsigned)((size_t)B * heads * nrows);
  attention_tail_kernel<<<(nw + 3) / 4, 128, 0, st>>>(qkv, out, B, T, heads, w, scale_log2e, causal, row0, nrows);
  B200_LAUNCH_OK();
  return B200_OK;
}

bool attention_tc3_supported(int T, int heads, int w) {
  return heads > 0 && w % heads == 0 && w / heads == A3_HD && T >= 1 && T <= A3_MAXT;
}

static int attn3_onepass_default() {
  const char* e = getenv("B200_ATTN_ONEPASS");
  return e ? atoi(e) : 0;
}

int attention_tc3(const CUtensorMap& tm3, const __nv_bfloat16* qkv, __nv_bfloat16* out, float* kmax_scratch, int B, int T,
                  int heads, int w, int causal, int sms, cudaStream_t st, cudaStream_t side, cudaEvent_t ev_fork,
                  cudaEvent_t ev_join) {
  static const int onepass = attn3_onepass_default();
  static const int tail_inside = getenv("B200_ATTN_TAIL_INSIDE") != nullptr;   // A/B: leftover rows inside the big kernel
  B200_CHECK(kmax_scratch != nullptr, B200_ERR_INVALID, "attention_tc3: needs a [B * heads] fp32 scratch");
  B200_CHECK(attention_tc3_supported(T, heads, w), B200_ERR_UNSUPPORTED, "attention_tc3: unsupported shape T=%d hd=%d", T,
             heads ? w / heads : 0);
  if (B == 0) return B200_OK;
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!(configured.load() >> (dev & 63) & 1ull)) {
    B200_CUDA(cudaFuncSetAttribute(attention_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM));
    configured.fetch_or(1ull << (dev & 63));
  }
  const float scale_log2e = (1.0f / sqrtf((float)A3_HD)) * 1.4426950408889634f;
  const int items = B * heads;
  const int grid = items < sms ? items : sms;
  if (onepass) {
    attn_kmax_kernel<<<(items + 7) / 8, 256, 0, st>>>(qkv, B, T, heads, w, kmax_scratch);
    B200_LAUNCH_OK();
  }
  const int n_full = T / 128, rem = T - n_full * 128;
  const bool tail_rows = rem > 0 && rem <= A3_TAIL_MAX && n_full >= 1;
  const bool external = tail_rows && !tail_inside;
  if (external) {
    // leftover rows next to the tensor-core kernel: fork to the side stream (when the caller has one), join after
    const unsigned nw = (unsigned)((size_t)items * rem);
    cudaStream_t ts = (side != nullptr && ev_fork != nullptr && ev_join != nullptr) ? side : st;
    if (ts != st) {
      B200_CUDA(cudaEventRecord(ev_fork, st));
      B200_CUDA(cudaStreamWaitEvent(ts, ev_fork, 0));
    }
    attention_tail_kernel<<<(nw + 3) / 4, 128, 0, ts>>>(qkv, out, B, T, heads, w, scale_log2e, causal, n_full * 128, rem);
    B200_LAUNCH_OK();
    if (ts != st) B200_CUDA(cudaEventRecord(ev_join, ts));
  }
  attention_tc3_kernel<<<grid, A3_THREADS, A3_SMEM, st>>>(tm3, qkv, out, kmax_scratch, B, T, heads, w, scale_log2e, causal, onepass,
                                                          external ? 1 : 0);
  B200_LAUNCH_OK();
  if (external && side != nullptr && ev_fork != nullptr && ev_join != nullptr) B200_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
  return B200_OK;
}

}  // namespace b200
