// FROZEN MEASUREMENT REFERENCE — not a production path.  This is attention_tc2.cu as it stood in round-2 session i
// (commit b1984e3), compiled into the same library under other symbol names so that a later build of the kernel can be
// timed against it in ONE process on ONE box (boxes of the pool differ by more than the changes being measured):
// tools/attn_variants.py ("old" column, profiles/r02o_*) through b200_attention_tc_bf16_device(Tp = -4), and
// B200_ATTN_GEN=4 in the model.  Nothing selects it by default.
//
// K4 on tcgen05, two query tiles in flight (the production attention for head dim 64, T <= 264:
// every CLIP tower at 224 px except H/14's vision tower).
//
// attention_tc.cu runs the chain  S = QK^T -> softmax -> O = PV -> store  strictly in sequence for
// one 128-row query tile at a time; its profile (profiles/r01b) shows the softmax warps waiting on
// the two MMAs ~35% of the time and the tensor pipe 11% busy.  Here the 512 TMEM columns are two
// 256-column buffers: tile t uses buffer t & 1 for its scores S (keys 0..255) and, once its softmax
// has consumed them, for its output O (columns 0..63 of the same buffer).  Two softmax groups of 4
// warps alternate tiles, so S(t+1) and P.V(t-1) run on the tensor pipe while group (t & 1) is in its
// exp pass.  Keys past 256 (the cls token makes T = 257) never touch the tensor core: their scores
// q.k_e and their p_e*v_e contributions are a 64-term dot product per row on the FMA pipe; the q rows
// come out of the (double-buffered) Q tile in shared memory — read from global memory they put ~1 us
// of uncoalesced-load latency in front of every tile's softmax (profiles/r02h: the 257th KEY cost as
// much as the padded third query tile).
//
//   warp 8      TMA: K, V rows of the head (128-row boxes) once per (sample, head),
//               Q tile per 128 query rows; all straight out of the fused qkv buffer, 128B swizzle.
//   warp 9      tcgen05.mma issuer: S(t) = Q K^T (128 x keys x 16, K-major), O(t) = P V with V as an
//               MN-major operand; order S(t), PV(t-1), S(t+1), PV(t), ...
//   warps 0..3  softmax group 0 (even tiles), warps 4..7 group 1 (odd tiles): thread = query row:
//               row max, ex2.approx, row sum, P as bf16 into the swizzled K-major smem tile of the
//               group, then O * (1/l) -> bf16 -> global.
#include "embed_kernels.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int A2_HD = 64;
constexpr int A2_MAXT = 264;                      // 256 keys on the tensor core + up to 8 extra keys
constexpr int A2_Q_BYTES = 128 * 128;             // 16 KB
constexpr int A2_KV_MAIN = 2 * 128 * 128;         // 256 rows x 128 B
constexpr int A2_P_BYTES = 4 * 128 * 128;         // 4 key blocks of [128 rows x 64 keys] per group
constexpr int A2_SMEM = 2 * A2_Q_BYTES + 2 * A2_KV_MAIN + 2 * A2_P_BYTES + 256 + 1024;   // Q tile double-buffered
constexpr int A2_THREADS = 320;

__global__ void __launch_bounds__(A2_THREADS, 1)
attention_tc2_r02i_kernel(const __grid_constant__ CUtensorMap tmBig,
                     const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int B, int T, int heads,
                     int w, float scale_log2e, int causal) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* base = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sQ = base;                          // [2] query tiles: tile tc lives in buffer tc & 1 (= its softmax group)
  uint8_t* sK = sQ + 2 * A2_Q_BYTES;           // rows 0..255
  uint8_t* sV = sK + A2_KV_MAIN;               // rows 0..255
  uint8_t* sP = sV + A2_KV_MAIN;               // [2 groups][4 key blocks][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * A2_P_BYTES);
  uint64_t* q_full = bars + 0;    // [2]
  uint64_t* q_empty = bars + 2;   // [2]  MMA commit (S of the tile done) [+ the 4 softmax warps that read their q rows]
  uint64_t* k_full = bars + 4;
  uint64_t* k_free = bars + 5;
  uint64_t* v_full = bars + 6;
  uint64_t* v_free = bars + 7;
  uint64_t* s_full = bars + 8;    // [2]
  uint64_t* p_full = bars + 10;   // [2]
  uint64_t* o_full = bars + 12;   // [2]
  uint64_t* buf_free = bars + 14; // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int keys_main = T < 256 ? (T + 15) / 16 * 16 : 256;  // keys on the tensor core (multiple of 16)
  const int extra = T > 256 ? T - 256 : 0;                    // keys handled on the FMA pipe
  const int kv_boxes = (keys_main + 127) / 128;
  const int q_tiles = (T + 127) / 128;
  const int items = B * heads;

  if (warp == 8 && lane == 0) {
    ptx::prefetch_tensormap(&tmBig);
  }
  if (warp == 9) {
    if (lane == 0) {
      for (int i = 4; i < 8; i++) ptx::mbar_init(&bars[i], 1);
      for (int i = 0; i < 2; i++) {
        ptx::mbar_init(&q_full[i], 1);
        // keys past 256 are scored on the FMA pipe from the q rows in shared memory: the tile's softmax warps hold it too
        ptx::mbar_init(&q_empty[i], extra > 0 ? 5 : 1);
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

  if (warp == 8) {  // the single-thread roles use the highest warp ids (arbiter priority)
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t it = 0, tc = 0;
      const uint32_t kv_bytes = (uint32_t)(kv_boxes * 128 * 128);
      for (int item = blockIdx.x; item < items; item += gridDim.x, it++) {
        const int b = item / heads, h = item - b * heads;
        // K of this head (free once the previous head's last S MMA has completed)
        ptx::mbar_wait(k_free, (it & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(k_full, kv_bytes);
        for (int i = 0; i < kv_boxes; i++)
          ptx::tma_load_2d(sK + i * 128 * 128, &tmBig, k_full, w + h * A2_HD, b * T + i * 128);
        for (int mt = 0; mt < q_tiles; mt++, tc++) {
          ptx::mbar_wait(&q_empty[tc & 1], ((tc >> 1) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&q_full[tc & 1], A2_Q_BYTES);
          ptx::tma_load_2d(sQ + (tc & 1) * A2_Q_BYTES, &tmBig, &q_full[tc & 1], h * A2_HD, b * T + mt * 128);
          if (mt == 0) {
            // V of this head (free once the previous head's last P.V has completed)
            ptx::mbar_wait(v_free, (it & 1) ^ 1);
            ptx::mbar_arrive_expect_tx(v_full, kv_bytes);
            for (int i = 0; i < kv_boxes; i++)
              ptx::tma_load_2d(sV + i * 128 * 128, &tmBig, v_full, 2 * w + h * A2_HD, b * T + i * 128);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc_s = ptx::umma_idesc_f16(128, keys_main, true);
      const uint32_t idesc_o = ptx::umma_idesc_f16(128, A2_HD, true) | (1u << 16);  // B (= V) is MN-major
      uint32_t it = 0, tc = 0;
      // one-tile software pipeline: PV of the previous tile is issued after S of the current one
      bool have_prev = false;
      uint32_t prev_tc = 0;
      bool prev_last = false, prev_first = false;
      uint32_t prev_it = 0;
      auto issue_pv = [&](uint32_t ptc, bool first_of_item, bool last_of_item, uint32_t pit) {
        const int pb = ptc & 1;
        if (first_of_item) ptx::mbar_wait(v_full, pit & 1);
        ptx::mbar_wait(&p_full[pb], (ptc >> 1) & 1);
        ptx::tc_fence_after();
        const uint8_t* sPg = sP + pb * A2_P_BYTES;
        for (int ks = 0; ks < keys_main / 16; ks++) {
          const uint64_t dp = ptx::umma_desc_k_sw128(ptx::smem_u32(sPg + (ks >> 2) * (128 * 128))) + (uint64_t)((ks & 3) * 2);
          const uint64_t dv = ptx::umma_desc_k_sw128(ptx::smem_u32(sV + ks * 16 * 128));  // 16 keys = 2 swizzle atoms
          ptx::umma_f16(tmem_base + pb * 256, dp, dv, idesc_o, ks != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&o_full[pb]);
        if (last_of_item) ptx::umma_commit(v_free);
      };
      for (int item = blockIdx.x; item < items; item += gridDim.x, it++) {
        for (int mt = 0; mt < q_tiles; mt++, tc++) {
          const int bsel = tc & 1;
          if (mt == 0) ptx::mbar_wait(k_full, it & 1);
          ptx::mbar_wait(&q_full[bsel], (tc >> 1) & 1);
          ptx::mbar_wait(&buf_free[bsel], ((tc >> 1) & 1) ^ 1);   // O(tc-2) has been read out of this buffer
          ptx::tc_fence_after();
          const uint64_t dq = ptx::umma_desc_k_sw128(ptx::smem_u32(sQ + bsel * A2_Q_BYTES));
          const uint64_t dk = ptx::umma_desc_k_sw128(ptx::smem_u32(sK));
#pragma unroll
          for (int k = 0; k < A2_HD / 16; k++)
            ptx::umma_f16(tmem_base + bsel * 256, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k != 0 ? 1u : 0u);
          ptx::umma_commit(&q_empty[bsel]);
          ptx::umma_commit(&s_full[bsel]);
          if (mt == q_tiles - 1) ptx::umma_commit(k_free);
          if (have_prev) issue_pv(prev_tc, prev_first, prev_last, prev_it);
          have_prev = true;
          prev_tc = tc;
          prev_first = mt == 0;
          prev_last = mt == q_tiles - 1;
          prev_it = it;
        }
      }
      if (have_prev) issue_pv(prev_tc, prev_first, prev_last, prev_it);
    }
    __syncwarp();
  } else {
    // ---------------- softmax groups ----------------
    const int grp = warp >> 2;                     // 0: even tiles, 1: odd tiles
    const int q4 = warp & 3;                       // TMEM lane quarter
    const int r = q4 * 32 + lane;                  // row inside the tile
    const uint32_t tbase = tmem_base + grp * 256 + ((uint32_t)(q4 * 32) << 16);
    uint8_t* sPg = sP + grp * A2_P_BYTES;
    const int chunks = (keys_main + 31) / 32;
    uint32_t tc = 0, it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, it++) {
      const int b = item / heads, h = item - b * heads;
      for (int mt = 0; mt < q_tiles; mt++, tc++) {
        if ((int)(tc & 1) != grp) continue;
        const uint32_t n = tc >> 1;
        const int qrow = mt * 128 + r;
        if (mt * 128 + q4 * 32 >= T) {
          // every row of this warp lies past the sequence (T = 257: three of the four warps of the third tile):
          // no scores to read, no exponentials, nothing to store — the P rows it would have written only feed
          // output rows nobody keeps.  It still takes part in the hand-shakes of the tile.
          if (extra > 0 && lane == 0) ptx::mbar_arrive(&q_empty[grp]);   // never reads the q rows: its share is free
          ptx::mbar_wait(&s_full[grp], n & 1);
          ptx::tc_fence_after();
          ptx::fence_proxy_async();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&p_full[grp]);
          ptx::mbar_wait(&o_full[grp], n & 1);
          ptx::tc_fence_after();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&buf_free[grp]);
          continue;
        }
        const int kmax = causal ? (qrow < T ? qrow : T - 1) : T - 1;   // last visible key
        // scores of the extra keys (>= 256) on the FMA pipe: q row from global (L2), k rows from smem
        float se[8];
#pragma unroll
        for (int e = 0; e < 8; e++) se[e] = -INFINITY;
        if (extra > 0) {
          // this thread's query row out of the Q tile the TMA staged for the MMA (buffer = group): ~30 cycles of
          // shared-memory latency instead of a ~1 us uncoalesced global read in front of every tile's softmax
          float qf[A2_HD];
          ptx::mbar_wait(&q_full[grp], n & 1);
          const uint8_t* qrow_s = sQ + grp * A2_Q_BYTES + r * 128;
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const uint4 u = *reinterpret_cast<const uint4*>(qrow_s + ((c ^ (r & 7)) * 16));
            const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
            qf[c * 8 + 0] = a0.x; qf[c * 8 + 1] = a0.y; qf[c * 8 + 2] = a1.x; qf[c * 8 + 3] = a1.y;
            qf[c * 8 + 4] = a2.x; qf[c * 8 + 5] = a2.y; qf[c * 8 + 6] = a3.x; qf[c * 8 + 7] = a3.y;
          }
          for (int e = 0; e < extra; e++) {
            // key row 256+e of this head, straight from the qkv buffer (same address for the whole warp)
            const uint4* kp = reinterpret_cast<const uint4*>(qkv + ((size_t)b * T + 256 + e) * 3 * w + w + (size_t)h * A2_HD);
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 8; c++) {
              const uint4 u = __ldg(kp + c);
              const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
              acc = fmaf(qf[c * 8 + 0], a0.x, acc); acc = fmaf(qf[c * 8 + 1], a0.y, acc);
              acc = fmaf(qf[c * 8 + 2], a1.x, acc); acc = fmaf(qf[c * 8 + 3], a1.y, acc);
              acc = fmaf(qf[c * 8 + 4], a2.x, acc); acc = fmaf(qf[c * 8 + 5], a2.y, acc);
              acc = fmaf(qf[c * 8 + 6], a3.x, acc); acc = fmaf(qf[c * 8 + 7], a3.y, acc);
            }
#pragma unroll
            for (int ee = 0; ee < 8; ee++)
              if (ee == e) se[ee] = (256 + e <= kmax) ? acc : -INFINITY;
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&q_empty[grp]);   // this warp's q rows are in registers
        }
        ptx::mbar_wait(&s_full[grp], n & 1);
        ptx::tc_fence_after();
        // pass 1: row maximum
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
        // pass 2: exponentials, row sum, P -> swizzled K-major tile of this group
        const float mb = m * scale_log2e;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
        for (int c = 0; c < chunks; c++) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(tbase + c * 32, v);
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
          uint8_t* blk = sPg + (c >> 1) * (128 * 128) + r * 128;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int ch = (((c & 1) * 4 + i) ^ (r & 7)) * 16;
            *reinterpret_cast<uint4*>(blk + ch) = make_uint4(pk[i * 4], pk[i * 4 + 1], pk[i * 4 + 2], pk[i * 4 + 3]);
          }
        }
        float pe[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
          float p;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(se[e], scale_log2e, -mb)));
          pe[e] = se[e] == -INFINITY ? 0.f : p;
          l0 += pe[e];
        }
        const float l = l0 + l1;
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
          float of[A2_HD];
#pragma unroll
          for (int j = 0; j < 32; j++) { of[j] = __uint_as_float(o0[j]); of[32 + j] = __uint_as_float(o1[j]); }
          for (int e = 0; e < extra; e++) {
            float p = 0.f;
#pragma unroll
            for (int ee = 0; ee < 8; ee++) if (ee == e) p = pe[ee];
            const uint4* vp = reinterpret_cast<const uint4*>(qkv + ((size_t)b * T + 256 + e) * 3 * w + 2 * w + (size_t)h * A2_HD);
#pragma unroll
            for (int c = 0; c < 8; c++) {
              const uint4 u = __ldg(vp + c);
              const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
              of[c * 8 + 0] = fmaf(p, a0.x, of[c * 8 + 0]); of[c * 8 + 1] = fmaf(p, a0.y, of[c * 8 + 1]);
              of[c * 8 + 2] = fmaf(p, a1.x, of[c * 8 + 2]); of[c * 8 + 3] = fmaf(p, a1.y, of[c * 8 + 3]);
              of[c * 8 + 4] = fmaf(p, a2.x, of[c * 8 + 4]); of[c * 8 + 5] = fmaf(p, a2.y, of[c * 8 + 5]);
              of[c * 8 + 6] = fmaf(p, a3.x, of[c * 8 + 6]); of[c * 8 + 7] = fmaf(p, a3.y, of[c * 8 + 7]);
            }
          }
          const float inv = 1.0f / l;
          __nv_bfloat16* op = out + ((size_t)b * T + qrow) * w + (size_t)h * A2_HD;
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
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) ptx::tmem_dealloc(tmem_base, 512);
}

bool attention_tc2_r02i_supported(int T, int heads, int w) {
  return heads > 0 && w % heads == 0 && w / heads == A2_HD && T >= 1 && T <= A2_MAXT;
}

int attention_tc2_r02i(const CUtensorMap& tmBig, const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int heads, int w,
                  int causal, int sms, cudaStream_t st) {
  B200_CHECK(attention_tc2_r02i_supported(T, heads, w), B200_ERR_UNSUPPORTED, "attention_tc2: unsupported shape T=%d hd=%d", T,
             heads ? w / heads : 0);
  if (B == 0) return B200_OK;
  static std::atomic<unsigned long long> configured{0};
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (!(configured.load() >> (dev & 63) & 1ull)) {
    B200_CUDA(cudaFuncSetAttribute(attention_tc2_r02i_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM));
    configured.fetch_or(1ull << (dev & 63));
  }
  const float scale_log2e = (1.0f / sqrtf((float)A2_HD)) * 1.4426950408889634f;
  const int items = B * heads;
  const int grid = items < sms ? items : sms;
  attention_tc2_r02i_kernel<<<grid, A2_THREADS, A2_SMEM, st>>>(tmBig, qkv, out, B, T, heads, w, scale_log2e, causal);
  B200_LAUNCH_OK();
  return B200_OK;
}

}  // namespace b200
