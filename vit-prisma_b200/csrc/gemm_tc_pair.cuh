// gemm_tc_pair.cuh -- CTA-pair (tcgen05 cta_group::2) variant of k_gemm_tc2; included by gemm_tc.cu inside its anonymous namespace.
//
// Why: the persistent 128 x 256 kernel is limited by the per-SM operand ingest (bf16 plain GEMM 0.106 ms vs cuBLAS 0.096,
// profiles/r01_gemm_notes.md).  A pair of CTAs on one TPC computes a 256 x BN tile with ONE tcgen05.mma.cta_group::2 stream issued by
// the leader: each CTA loads its own 128 rows of A and HALF of the B tile, so the B bytes per SM halve.
//
// Protocol (differences from k_gemm_tc2 are marked [2CTA]):
//   * __cluster_dims__(2,1,1); rank = %cluster_ctarank; rank 0 = leader.  A cluster walks tiles cluster_id, cluster_id + n_clusters ...
//   * smem ring per CTA: A [128 x BK] (+lo) | B-half [BN/2 x BK] (+lo); identical offsets in both CTAs (the MMA reads the peer's
//     shared memory at the same offsets).
//   * full[s] lives in the LEADER: the leader arms expect_tx with the bytes of BOTH CTAs; both producers issue
//     cp.async.bulk.tensor ... .cta_group::2 with the barrier address masked to the leader (bit 24 cleared)            [2CTA]
//   * empty[s] and tfull[ab] live in EACH CTA and are signalled by tcgen05.commit.cta_group::2 ... multicast::cluster, mask 0b11 [2CTA]
//   * tempty[ab] lives in the leader, count 2*NEPI: every epilogue warp of either CTA arrives on it through mapa          [2CTA]
//   * TMEM: tcgen05.alloc.cta_group::2 executed by warp 1 of both CTAs; each CTA's 128 lanes x (2*BN) columns hold ITS 128 rows.
//   * teardown: cluster barrier before tcgen05.dealloc.cta_group::2                                                     [2CTA]
// PTX forms follow the CUTLASS headers vendored in this image (cute/arch/copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_2D,
// cute/arch/mma_sm100_umma.hpp SM100_MMA_*_2x1SM_SS, cutlass/arch/barrier.h umma_arrive_multicast_2x1SM, ClusterBarrier::arrive).
#pragma once


// (the cta_group::2 PTX wrappers -- tma_load_2d_pair, tc_mma_pair, tc_commit_pair, mbar_arrive_remote ... -- live in tc_common.cuh)

template <typename T, int NPASS, int BN, int STAGES, int NEPI>
struct PairCfg {
  static constexpr int ES = sizeof(T);
  static constexpr int BK = 128 / ES;
  static constexpr int NOP = NPASS == 3 ? 2 : 1;
  static constexpr int A_BYTES = TC_BM * 128;              // this CTA's 128 rows of the 256-row tile
  static constexpr int BH_BYTES = (BN / 2) * 128;          // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = NOP * (A_BYTES + BH_BYTES);
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int EPI_WARP_FLOATS = sizeof(T) == 2 ? 32 * 33 : 32 * 36;
  static constexpr int EPI_STAGE_BYTES = NEPI * EPI_WARP_FLOATS * 4;
  static constexpr int SMEM_BYTES = RING_BYTES + EPI_STAGE_BYTES + 1024 + 256;
  static constexpr int THREADS = 64 + NEPI * 32;
  static constexpr int TMEM_COLS = 2 * BN;                 // two accumulators of BN fp32 columns (power of two: BN in {128, 256})
  static constexpr uint32_t FMT = sizeof(T) == 2 ? 1u : 2u;
  // M = 256 across the pair, N = BN
  static constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
  static_assert(TMEM_COLS <= 512, "TMEM budget exceeded");
};

template <typename T, int NPASS, int BN, int STAGES, int NEPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + NEPI * 32, 1)
k_gemm_tc2_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmAlo,
                const __grid_constant__ CUtensorMap tmBhlo, int K, EpiParams ep, int num_m_tiles /* of 256 rows */, int num_n_tiles, int m_fast) {
  using C = PairCfg<T, NPASS, BN, STAGES, NEPI>;
  auto tile_m = [&](int tile) { return m_fast ? tile % num_m_tiles : tile / num_n_tiles; };
  auto tile_n = [&](int tile) { return m_fast ? tile / num_m_tiles : tile % num_n_tiles; };
  constexpr int KIND = sizeof(T) == 2 ? 0 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t ring = (smem0 + 1023u) & ~1023u;
  const uint32_t epi_stage = ring + C::RING_BYTES;
  const uint32_t bar_base = epi_stage + C::EPI_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // used in the leader only
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };           // one per CTA
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };       // one per CTA
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };  // used in the leader only
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_ptr_generic = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int num_kb = (K + C::BK - 1) / C::BK;
  const int num_tiles = num_m_tiles * num_n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmBh);
    if (NPASS == 3) { prefetch_tmap(&tmAlo); prefetch_tmap(&tmBhlo); }
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * NEPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // peer barriers initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int m0 = tile_m(tile) * 256 + (int)rank * TC_BM;          // this CTA's rows of A
        const int n0 = tile_n(tile) * BN + (int)rank * (BN / 2);        // this CTA's half of B
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);                   // own slot free (signalled by the leader's multicast commit)
          if (leader) mbar_expect_tx(full_bar(s), 2u * C::STAGE_BYTES);   // bytes of BOTH CTAs land on the leader's barrier
          const uint32_t sa = ring + s * C::STAGE_BYTES;
          const int kc = kb * C::BK;
          tma_load_2d_pair(sa, &tmA, full_bar(s), kc, m0);
          if (NPASS == 3) tma_load_2d_pair(sa + C::A_BYTES, &tmAlo, full_bar(s), kc, m0);
          const uint32_t sb = sa + C::NOP * C::A_BYTES;
          tma_load_2d_pair(sb, &tmBh, full_bar(s), kc, n0);
          if (NPASS == 3) tma_load_2d_pair(sb + C::BH_BYTES, &tmBhlo, full_bar(s), kc, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (leader && lane == 0) {
      uint32_t it = 0;
      int li = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++li) {
        const int ab = li & 1;
        const uint32_t aph = (li >> 1) & 1;
        mbar_wait(tempty_bar(ab), aph ^ 1);                  // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = ring + s * C::STAGE_BYTES;
          const uint32_t sb = sa + C::NOP * C::A_BYTES;
#pragma unroll
          for (int k = 0; k < 128 / 32; ++k) {
            const uint32_t koff = k * 32;
            const uint64_t a_hi = make_smem_desc(sa + koff);
            const uint64_t b_hi = make_smem_desc(sb + koff);
            const uint32_t first = (kb | k) != 0 ? 1u : 0u;
            if (NPASS == 3) {
              const uint64_t a_lo = make_smem_desc(sa + C::A_BYTES + koff);
              const uint64_t b_lo = make_smem_desc(sb + C::BH_BYTES + koff);
              tc_mma_pair<KIND>(d_tmem, a_lo, b_hi, C::IDESC, first);
              tc_mma_pair<KIND>(d_tmem, a_hi, b_lo, C::IDESC, 1u);
              tc_mma_pair<KIND>(d_tmem, a_hi, b_hi, C::IDESC, 1u);
            } else {
              tc_mma_pair<KIND>(d_tmem, a_hi, b_hi, C::IDESC, first);
            }
          }
          tc_commit_pair(empty_bar(s));                      // slot free in both CTAs
        }
        tc_commit_pair(tfull_bar(ab));                       // accumulator complete: wake both epilogues
      }
    }
  } else {
    // ===================== epilogue (both CTAs, their own 128 rows) =====================
    const int e = warp - 2;
    const int quarter = warp & 3;
    constexpr int CPW = BN / (NEPI / 4);
    const int cbase = (e / 4) * CPW;
    float* stage = reinterpret_cast<float*>(smem_raw + (epi_stage - smem0)) + e * C::EPI_WARP_FLOATS;
    int li = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++li) {
      const int m0 = tile_m(tile) * 256 + (int)rank * TC_BM, n0 = tile_n(tile) * BN;
      const int ab = li & 1;
      const uint32_t aph = (li >> 1) & 1;
      mbar_wait(tfull_bar(ab), aph);
      tc_fence_after();
      const int row0 = m0 + quarter * 32;
      const int nrows = min(32, ep.M - row0);
#pragma unroll 1
      for (int c = 0; c < CPW / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * BN + cbase + c * 32), r);
        tmem_ld_wait();
        if (c == CPW / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(tempty_bar(ab), 0);   // the leader's MMA warp counts 2*NEPI arrivals
        }
        const int col0 = n0 + cbase + c * 32;
        if (ep.vec16_ok && col0 + 32 <= ep.N) {
          if (nrows > 0) epi_chunk_vec<T>(ep, r, reinterpret_cast<uint8_t*>(stage), lane, row0, nrows, col0);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = __uint_as_float(r[j]);
          __syncwarp();
          if (nrows > 0) {
            if (ep.vec_ok) epi_rows_pair<T>(ep, stage, lane, row0, nrows, col0);
            else if (col0 + lane < ep.N) epi_rows_scalar<T>(ep, stage, lane, row0, nrows, col0 + lane);
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // the peer may still be reading its half of TMEM / our smem
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

template <typename T, int NPASS, int BN, int STAGES, int NEPI>
int launch_tc2_pair(const PbGemm* g, cudaStream_t st) {
  using C = PairCfg<T, NPASS, BN, STAGES, NEPI>;
  CUtensorMap tmA, tmBh, tmAlo, tmBhlo;
  PB_TRY(make_map(&tmA, g->A, g->dtype, g->M, g->K, g->lda, TC_BM));
  PB_TRY(make_map(&tmBh, g->B, g->dtype, g->N, g->K, g->ldb, BN / 2));          // box = half of the B tile
  if (NPASS == 3) {
    PB_TRY(make_map(&tmAlo, g->A_lo, g->dtype, g->M, g->K, g->lda, TC_BM));
    PB_TRY(make_map(&tmBhlo, g->B_lo, g->dtype, g->N, g->K, g->ldb, BN / 2));
  } else {
    tmAlo = tmA;
    tmBhlo = tmBh;
  }
  auto kern = k_gemm_tc2_pair<T, NPASS, BN, STAGES, NEPI>;
  static bool attr_done = false;
  if (!attr_done) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  EpiParams ep = pb_make_epi(g);
  const int num_m = (g->M + 255) / 256, num_n = (g->N + BN - 1) / BN;
  int clusters = pb_sm_count() / 2;
  if (clusters > num_m * num_n) clusters = num_m * num_n;
  const size_t planes = NPASS > 1 ? 2 : 1;
  const size_t a_bytes = (size_t)g->M * g->K * sizeof(T) * planes, b_bytes = (size_t)g->N * g->K * sizeof(T) * planes;
  const int m_fast = (b_bytes > ((size_t)48 << 20) && a_bytes < b_bytes) ? 1 : 0;
  kern<<<2 * clusters, C::THREADS, C::SMEM_BYTES, st>>>(tmA, tmBh, tmAlo, tmBhlo, g->K, ep, num_m, num_n, m_fast);
  PB_LAUNCH_CHECK();
  return PB_OK;
}


