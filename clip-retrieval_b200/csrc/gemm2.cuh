// tcgen05 GEMM on CTA PAIRS (cta_group::2): the large-shape variant of gemm.cuh.
//
// Why: with one CTA per 128x256 tile every SM pulls 48 KB of operands from L2 per 64-wide k-block
// (96 B/clk/SM at the MMA rate) — more than the L2 can feed 148 SMs.  A pair of CTAs on one TPC
// computes a 256x256 tile with ONE tcgen05.mma.cta_group::2 per k-step: each CTA stages its own
// 128 rows of A and only HALF of the B tile (128 of the 256 weight rows); the tensor core reads
// both halves across the pair.  Operand traffic drops to 32 KB per k-block per SM and the ring
// deepens to 6 stages in the same shared memory.
//
// Roles per CTA (gemm.cuh): warp 8 TMA producer (its A half + its B half; completion bytes are credited
// to the leader's full barrier), warp 9 = MMA issuer in the leader CTA only, warps 0..7 epilogue over
// this CTA's 128 accumulator rows (two warps per TMEM lane quarter, each draining half of the columns).  tcgen05.commit multicasts "slot free" / "accumulator ready" to
// the barriers of both CTAs; epilogue warps of both CTAs release the accumulator stage on the
// leader's barrier (remote mbarrier arrive for the second CTA).
#pragma once
#include "gemm.cuh"

namespace b200 {

constexpr int G2_BM = 256;          // rows per pair tile (128 per CTA)
constexpr int G2_BN = 256;
// Two variants.  TMAST = false: 6-stage ring, results stored from registers (16 B per thread and row).  TMAST = true
// (production): 5-stage ring + 64 KB of staging — every epilogue warp owns two 32-row x 64-column boxes (128-byte
// swizzle); the residual tile is PREFETCHED into them by TMA before the accumulator is ready, results overwrite it in
// place and leave through cp.async.bulk.tensor stores.  Why: with register stores each warp instruction touches 32
// different rows (32 half-used sectors); the LSU work of one tile was ~12 % of the K=1024 GEMMs and the per-thread
// residual loads another 20 % of out-proj (tools/gemm_exp.sh, profiles/r02b_gemm_exp.txt).
template <bool TMAST> struct G2Cfg {
  static constexpr int STAGES = TMAST ? 5 : 6;   // 7 stages (224 KB) measured no better than 6
  static constexpr int STAGING = TMAST ? 8 * 2 * 4096 : 0;
  static constexpr int ALIGN_SLACK = TMAST ? 768 : 1024;   // dynamic shared memory starts 1024-aligned in practice; checked
  static constexpr int SMEM_BYTES = STAGES * (2 * 128 * GEMM_BK * 2) + STAGING + 2 * 256 * 4 + 256 + ALIGN_SLACK;
};
constexpr int G2_A_BYTES = 128 * GEMM_BK * 2;   // 16 KB
constexpr int G2_B_BYTES = 128 * GEMM_BK * 2;   // 16 KB (half of the 256-wide B tile)
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;

// SPEC: see epi_pack8 (gemm.cuh): -1 = every epilogue feature decided at run time, 0..7 = activation | residual << 2 fixed.
template <bool TMAST, int SPEC = -1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, int M,
                              int N, int K, GemmEpilogue ep) {
  constexpr int G2_STAGES = G2Cfg<TMAST>::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  if (TMAST && pad > (uint32_t)G2Cfg<TMAST>::ALIGN_SLACK) __trap();   // never on sm_100 (the dynamic window starts at 1024)
  uint8_t* base = smem_raw + pad;
  uint8_t* sA = base;
  uint8_t* sB = base + G2_STAGES * G2_A_BYTES;
  uint8_t* sStage = base + G2_STAGES * G2_STAGE_BYTES;                  // [8 warps][2 boxes][32 rows x 128 B]
  float* s_bias = reinterpret_cast<float*>(sStage + G2Cfg<TMAST>::STAGING);
  float* s_c = s_bias + G2_BN;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_c + G2_BN);
  uint64_t* empty = full + G2_STAGES;
  uint64_t* tfull = empty + G2_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* rbar = tempty + 2;                                           // [8] residual boxes landed (one per epilogue warp)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(rbar + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int mb = (M + G2_BM - 1) / G2_BM, nb = (N + G2_BN - 1) / G2_BN, kb = (K + GEMM_BK - 1) / GEMM_BK;
  const int tiles = mb * nb;

  if (warp == GEMM_WARP_TMA && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    if (TMAST) {
      ptx::prefetch_tensormap(&tmC);
      ptx::prefetch_tensormap(&tmR);
    }
  }
  if (warp == GEMM_WARP_MMA) {
    if (lane == 0) {
      for (int s = 0; s < G2_STAGES; s++) {
        ptx::mbar_init(&full[s], 1);   // the leader's producer arrives; both CTAs' TMA bytes are credited here
        ptx::mbar_init(&empty[s], 1);  // multicast commit from the leader's MMA
      }
      for (int a = 0; a < 2; a++) {
        ptx::mbar_init(&tfull[a], 1);
        ptx::mbar_init(&tempty[a], 16);  // 8 epilogue warps x 2 CTAs (used in the leader only)
      }
      for (int i = 0; i < 8; i++) ptx::mbar_init(&rbar[i], 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc_pair(s_tmem, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == GEMM_WARP_TMA) {
    // ---------------- TMA producer (both CTAs) ----------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < tiles; tile += npairs) {
        int m_blk, n_blk;
        gemm_tile_coords(tile, mb, nb, m_blk, n_blk);
        for (int kbi = 0; kbi < kb; kbi++) {
          const long long t0 = (ep.dbg && leader) ? clock64() : 0;
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          if (ep.dbg && leader) atomicAdd(ep.dbg + 3, (unsigned long long)(clock64() - t0));
          ptx::tma_load_2d_pair(sA + stage * G2_A_BYTES, &tmA, &full[stage], kbi * GEMM_BK,
                                m_blk * G2_BM + (int)rank * 128);
          ptx::tma_load_2d_pair(sB + stage * G2_B_BYTES, &tmB, &full[stage], kbi * GEMM_BK,
                                n_blk * G2_BN + (int)rank * 128);
          // The peer's loads complete on the leader's barrier too (peer bit cleared in the TMA); the peer
          // cannot run a ring cycle ahead because its empty[] is released by the leader's MMA commit.
#ifdef B200_TIMING_EXPERIMENTS
          if (leader)
            ptx::mbar_arrive_expect_tx(&full[stage], 2 * (G2_A_BYTES + (ep.exp_b_bytes ? ep.exp_b_bytes : G2_B_BYTES)));
#else
          if (leader) ptx::mbar_arrive_expect_tx(&full[stage], 2 * G2_STAGE_BYTES);
#endif
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == GEMM_WARP_MMA) {
    // ---------------- MMA issuer (leader CTA only) ----------------
    if (leader && lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(G2_BM, G2_BN, true);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long w_full = 0, w_acc = 0;
      const long long t_begin = ep.dbg ? clock64() : 0;
      for (int tile = pair; tile < tiles; tile += npairs) {
        long long t0 = ep.dbg ? clock64() : 0;
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        if (ep.dbg) w_acc += clock64() - t0;
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * G2_BN;
        for (int kbi = 0; kbi < kb; kbi++) {
          t0 = ep.dbg ? clock64() : 0;
          ptx::mbar_wait(&full[stage], phase);
          if (ep.dbg) w_full += clock64() - t0;
          ptx::tc_fence_after();
          const uint64_t da = ptx::umma_desc_k_sw128(ptx::smem_u32(sA + stage * G2_A_BYTES));
          const uint64_t db = ptx::umma_desc_k_sw128(ptx::smem_u32(sB + stage * G2_B_BYTES));
#pragma unroll
          for (int k = 0; k < GEMM_BK / GEMM_UMMA_K; k++)
            ptx::umma_f16_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kbi | k) != 0 ? 1u : 0u);
          ptx::umma_commit_pair(&empty[stage], 3);
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_pair(&tfull[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (ep.dbg) {
        atomicAdd(ep.dbg + 0, (unsigned long long)w_full);
        atomicAdd(ep.dbg + 1, (unsigned long long)w_acc);
        atomicAdd(ep.dbg + 2, (unsigned long long)(clock64() - t_begin));
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue (warps 0..7 of both CTAs) ----------------
    const int q = warp & 3;
    const int et = warp * 32 + lane;  // 0..255
    const int half = warp >> 2;        // which half of the tile columns this warp drains
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t rphase = 0;
    const uint32_t tempty0_remote = ptx::mapa_u32(ptx::smem_u32(&tempty[0]), 0);
    for (int tile = pair; tile < tiles; tile += npairs) {
      int m_blk, n_blk;
      gemm_tile_coords(tile, mb, nb, m_blk, n_blk);
      const int n0 = n_blk * G2_BN;
      const bool prof = ep.dbg != nullptr && leader && warp == 0 && lane == 0;   // B200_GEMM_DEBUG: where this warp's tile time goes
      const long long tp0 = prof ? clock64() : 0;
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
      for (int j = et; j < G2_BN; j += 256) {
        s_bias[j] = (ep.bias != nullptr && n0 + j < N) ? ep.bias[n0 + j] : 0.0f;
        if (SPEC < 0) s_c[j] = (ep.ln_c != nullptr && n0 + j < N) ? ep.ln_c[n0 + j] : 0.0f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");

      const int row = m_blk * G2_BM + (int)rank * 128 + q * 32 + lane;
      const bool row_ok = row < M;
      EpiRow er = epi_row(ep, row, n0);
      if (SPEC < 0 && ep.ln_stats != nullptr && row_ok) ln_row_coeffs(ep.ln_stats + (int64_t)row * (ep.ln_w >> 6), ep.ln_w >> 6, ep.ln_w, er.ln_a, er.ln_b);
      float st_k = 0.f, st_s = 0.f, st_q = 0.f;
      constexpr int CPW = G2_BN / 64;   // 32-column chunks per epilogue warp
      if constexpr (TMAST) {
        // ---- results through shared memory and TMA stores ----
        uint8_t* stage = sStage + warp * 8192;                 // two boxes of 32 rows x 64 columns (128 B rows, swizzled)
        const int row0 = m_blk * G2_BM + (int)rank * 128 + q * 32;      // first row of this warp's slab
        const int col0 = n0 + half * 128;                                // first column of this warp's half tile
        const bool has_res = SPEC >= 0 ? ((SPEC >> 2) & 1) != 0 : ep.residual != nullptr;
        const bool use_stats = SPEC < 0 && ep.stats_out != nullptr;
        const bool use_ln = SPEC < 0 && ep.ln_stats != nullptr;
        const int act = SPEC >= 0 ? (SPEC & 3) : ep.act;
        int out_row0 = row0, res_row0 = row0;
        if (ep.out_group > 0) out_row0 = (row0 / ep.out_group) * (ep.out_group + 1) + 1 + row0 % ep.out_group;
        if (ep.res_row_mod > 0) res_row0 = ep.res_row_off + row0 % ep.res_row_mod;
        // the boxes are free once the stores of the previous tile have read them
        if (lane == 0) ptx::bulk_wait_group_read<0>();
        __syncwarp();
        if (has_res && lane == 0 && row0 < M) {
          // residual prefetch: both boxes in flight while the main loop of this tile runs
          ptx::mbar_arrive_expect_tx(&rbar[warp], 8192);
          ptx::tma_load_2d(stage, &tmR, &rbar[warp], col0, res_row0);
          ptx::tma_load_2d(stage + 4096, &tmR, &rbar[warp], col0 + 64, res_row0);
        }
        const long long tp1 = prof ? clock64() : 0;
        ptx::mbar_wait(&tfull[acc], acc_phase);
        ptx::tc_fence_after();
        const long long tp2 = prof ? clock64() : 0;
        long long tp3 = 0;
        if (has_res && row0 < M) ptx::mbar_wait(&rbar[warp], rphase);
        // the accumulator chunk c+1 is in flight (tcgen05.ld) while chunk c is finished and staged
        uint32_t rbuf[2][32];
        const uint32_t tacc = tmem_base + acc * G2_BN + half * CPW * 32 + ((uint32_t)(q * 32) << 16);
        ptx::tmem_ld_32x32b_x32(tacc, rbuf[0]);
#pragma unroll
        for (int ci = 0; ci < CPW; ci++) {
          const int c = half * CPW + ci;
          uint32_t(&r)[32] = rbuf[ci & 1];
          ptx::tmem_ld_wait();
          if (ci + 1 < CPW) ptx::tmem_ld_32x32b_x32(tacc + (ci + 1) * 32, rbuf[(ci + 1) & 1]);
          if (ci == CPW - 1) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (leader) ptx::mbar_arrive(&tempty[acc]);
              else ptx::mbar_arrive_cluster(tempty0_remote + (uint32_t)acc * 8u);
            }
            if (prof) tp3 = clock64();
          }
          if (use_stats && (c & 1) == 0) {
            st_s = 0.f;
            st_q = 0.f;
            float v0;
            const int colb = c * 32;
            if (use_ln) v0 = act_apply(fmaf(__uint_as_float(r[0]), er.ln_a, fmaf(er.ln_b, s_c[colb], s_bias[colb])), act);
            else v0 = act_apply(__uint_as_float(r[0]) + s_bias[colb], act);
            if (has_res) {
              const uint4 r0 = *reinterpret_cast<const uint4*>(stage + (ci >> 1) * 4096 + lane * 128 + ((((ci & 1) * 4) ^ (lane & 7)) * 16));
              v0 = __fadd_rn(v0, unpack_bf16x2(r0.x).x);
            }
            st_k = __bfloat162float(__float2bfloat16_rn(v0));
          }
          uint8_t* box = stage + (ci >> 1) * 4096 + lane * 128;          // this thread's row of the box
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const int col = c * 32 + g * 8;                               // column inside the tile
            uint8_t* slot = box + (((((ci & 1) * 4) + g) ^ (lane & 7)) * 16);
            uint4 rr = make_uint4(0u, 0u, 0u, 0u);
            if (has_res) rr = *reinterpret_cast<const uint4*>(slot);
            uint4 o = make_uint4(0u, 0u, 0u, 0u);                        // rows past M store zeros
            if (row_ok && n0 + col < N) o = epi_pack8<SPEC>(ep, er, r + g * 8, s_bias, s_c, col, has_res, rr, st_k, st_s, st_q);
            *reinterpret_cast<uint4*>(slot) = o;
          }
          if (use_stats && (c & 1) == 1 && row_ok && n0 + c * 32 < N) {
            const float mean = st_k + st_s * (1.0f / 64.0f);
            const float m2 = fmaxf(st_q - st_s * st_s * (1.0f / 64.0f), 0.f);
            ep.stats_out[(int64_t)row * (N >> 6) + ((n0 + c * 32) >> 6)] = make_float2(mean, m2);
          }
          if (ci & 1) {
            // a 64-column box is complete: generic-proxy writes -> visible to the TMA engine, then one store
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0 && row0 < M && col0 + (ci >> 1) * 64 < N) {
              ptx::tma_store_2d(&tmC, stage + (ci >> 1) * 4096, col0 + (ci >> 1) * 64, out_row0);
              ptx::bulk_commit_group();
            }
          }
        }
        if (prof) {
          const long long tp4 = clock64();
          atomicAdd(ep.dbg + 4, (unsigned long long)(tp1 - tp0));
          atomicAdd(ep.dbg + 5, (unsigned long long)(tp2 - tp1));
          atomicAdd(ep.dbg + 6, (unsigned long long)(tp3 - tp2));
          atomicAdd(ep.dbg + 7, (unsigned long long)(tp4 - tp3));
        }
        if (has_res && row0 < M) rphase ^= 1;
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      // residual prefetch: issued before the accumulator wait so its latency overlaps the main loop
      uint4 res[CPW][4];
#pragma unroll
      for (int ci = 0; ci < CPW; ci++) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
          const int col = (half * CPW + ci) * 32 + g * 8;
          res[ci][g] = (row_ok && er.res_ptr != nullptr && n0 + col < N)
                           ? *reinterpret_cast<const uint4*>(er.res_ptr + col) : make_uint4(0u, 0u, 0u, 0u);
        }
      }

      ptx::mbar_wait(&tfull[acc], acc_phase);
      ptx::tc_fence_after();

#pragma unroll
      for (int ci = 0; ci < CPW; ci++) {
        const int c = half * CPW + ci;
        uint32_t r[32];
#ifdef B200_TIMING_EXPERIMENTS
        if (ep.exp_skip_tmem) {
#pragma unroll
          for (int j = 0; j < 32; j++) r[j] = 0x3f800000u;
        } else
#endif
        {
          ptx::tmem_ld_32x32b_x32(tmem_base + acc * G2_BN + c * 32 + ((uint32_t)(q * 32) << 16), r);
          ptx::tmem_ld_wait();
        }
        if (ci == CPW - 1) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) ptx::mbar_arrive(&tempty[acc]);
            else ptx::mbar_arrive_cluster(tempty0_remote + (uint32_t)acc * 8u);
          }
        }
        if (row_ok) epi_chunk<SPEC>(ep, er, r, s_bias, s_c, c, n0, N, row, res[ci], st_k, st_s, st_q);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (TMAST && lane == 0) ptx::bulk_wait_group<0>();   // every store of this thread has been performed before the CTA exits
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == GEMM_WARP_MMA) ptx::tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace b200
