// CTA-pair tcgen05 GEMM for sm_100a: out[M,N] = epi(A[M,K] . B[N,K]^T), bf16 -> fp32 in TMEM.
//
// The production-shape path of llc_gemm_bf16_tn (token counts in the thousands, N % 256 == 0).
// A cluster of two CTAs (one SM each) owns a 256 x 256 output tile: tcgen05.mma.cta_group::2 with
// M = 256 reads A rows [128 r, 128 r + 128) and B rows [128 r, 128 r + 128) from CTA r's shared
// memory and leaves 128 accumulator rows x 256 columns in each CTA's TMEM. Per k-block each CTA
// fetches 32 KB through TMA for 2 x 128 x 256 x 64 MACs of its own - half the L2->SM bytes per
// FLOP of the single-CTA 128 x 256 kernel, which measured L2-feed-bound (profiles/).
//
//   warp 0      TMA producer (one lane per CTA): 5-stage ring; both CTAs complete bytes on the
//               LEADER's full barrier
//   warp 1      MMA issuer (leader CTA, one lane); commits multicast to both CTAs' barriers
//   warps 2..9  epilogue (gemm_epilogue.cuh): two warps per TMEM lane quarter, 128 columns each
// TMEM holds two 256-column accumulators, so a tile's epilogue overlaps the next tile's mainloop.
#include <stdlib.h>

#include "gemm_epilogue.cuh"

namespace {

constexpr int BM = 256;       // per pair; 128 rows per CTA
constexpr int BN = 256;
constexpr int BK = 64;        // 128 B swizzle row
constexpr int kStages = 5;
constexpr int kPrefetchKb = 8;  // A blocks requested into L2 this many k-blocks ahead of the ring
constexpr int kABytes = 128 * BK * 2;
constexpr int kBBytes = 128 * BK * 2;
constexpr int kStageBytes = kABytes + kBBytes;       // per CTA
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kEpiBytes = kEpiWarps * kEpiWarpBytes;
constexpr int kBarBytes = 256;
constexpr int kSmem = 1024 + kStages * kStageBytes + kEpiBytes + kBarBytes;
constexpr int kTmemCols = 512;

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
             int M, int N, int K, EpiParams ep, int dbg) {
  // dbg (LLC_GEMM_DBG, development only): 1 = epilogue does no work, 2 = producer issues no TMA,
  // 4 = no MMA is issued, 8 = no L2 prefetch, 32 = producer does not wait for free slots,
  // 64 = MMA issuer does not wait for data, 128 = no per-stage commit (wrong results: timing only)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* smem_ab = smem;
  uint8_t* smem_epi = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kEpiBytes);
  uint64_t* full_bar = bars;                   // [kStages]  used in the leader CTA only
  uint64_t* empty_bar = bars + kStages;        // [kStages]  one per CTA (multicast commit)
  uint64_t* tfull_bar = bars + 2 * kStages;    // [2]        one per CTA (multicast commit)
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]     leader only: both CTAs' epilogues
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* epi_bar = bars + 2 * kStages + 5;     // [2 per epilogue warp] residual tiles landed

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int tiles_m = (M + BM - 1) / BM;
  const int tiles_n = N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      // one arrival (the leader's expect_tx, armed with BOTH CTAs' bytes); the peer's TMA only
      // completes bytes on it. Early peer bytes just drive the tx-count negative until the
      // leader arms the phase, and a peer TMA can never reach the next phase early because its
      // own empty barrier fires only after the MMAs that waited on this phase retired.
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tfull_bar[b]), 1);
      mbar_init(smem_u32(&tempty_bar[b]), 2 * kEpiWarps);
    }
    for (int b = 0; b < 2 * kEpiWarps; ++b) mbar_init(smem_u32(&epi_bar[b]), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_cg2<kTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  cluster_sync_all();   // barrier inits and the TMEM allocation are visible to both CTAs
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();   // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    // the whole warp runs the loop; one elected lane issues (operands stay uniform)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m0 = (tile / tiles_n) * BM + (int)rank * 128;
      const int n0 = (tile % tiles_n) * BN + (int)rank * 128;
      const int next_tile = tile + num_pairs;
      const int m0_next = (next_tile / tiles_n) * BM + (int)rank * 128;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (!(dbg & 32)) mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb_local = smem_u32(&full_bar[stage]);
        const uint32_t fb_leader = mapa_shared(fb_local, 0);
        const uint32_t sa = smem_u32(smem_ab + stage * kStageBytes);
        if (elect_one()) {
          // A (activations) streams from HBM once per GEMM: pull it into L2 well ahead of the
          // smem ring so the ring's own loads see L2 latency. B (weights) stays L2-resident.
          if (!(dbg & 8)) {
            const int pk = kb + kPrefetchKb;
            if (pk < num_kb) tma_prefetch_2d(&tmA, pk * BK, m0);
            else if (next_tile < num_tiles && pk - num_kb < num_kb)
              tma_prefetch_2d(&tmA, (pk - num_kb) * BK, m0_next);
          }
          if (dbg & 2) {
            if (rank == 0) mbar_arrive(fb_local);
          } else {
            if (rank == 0) mbar_expect_tx(fb_local, 2 * kStageBytes);
            if (dbg & 2048) {
              tma_load_2d_cg2(sa, &tmA, fb_leader, kb * BK, m0);
              tma_load_2d_cg2(sa + kABytes, &tmB, fb_leader, kb * BK, n0);
            } else {
              // operands are re-read by the other n-/m-tiles within microseconds: keep them in L2
              // ahead of the single-use output / residual streams
              tma_load_2d_cg2_hint(sa, &tmA, fb_leader, kb * BK, m0, kEvictLast);
              tma_load_2d_cg2_hint(sa + kABytes, &tmB, fb_leader, kb * BK, n0, kEvictLast);
            }
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        mbar_wait(smem_u32(&tempty_bar[buf]), bphase ^ 1);  // both epilogues drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!(dbg & 64)) mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_ab + stage * kStageBytes);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + kABytes);
          const int ksteps = min(BK / 16, (K - kb * BK) / 16);
          if (elect_one()) {
            if (!(dbg & 4))
              for (int k = 0; k < ksteps; ++k)
                umma_bf16_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            if (!(dbg & 128))
              umma_commit_mc(smem_u32(&empty_bar[stage]), 0x3);  // frees the slot in both CTAs
            if (kb == num_kb - 1)
              umma_commit_mc(smem_u32(&tfull_bar[buf]), 0x3);  // accumulator complete (both CTAs)
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int ew = warp - 2;
    const int q = warp & 3;    // TMEM lane quarter this warp may access
    const int hh = ew >> 2;    // column half
    uint8_t* tile_s = smem_epi + ew * kEpiWarpBytes;
    const uint32_t te_leader = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    uint64_t* my_bar = epi_bar + 2 * ew;
    EpiF32State f32st;
    f32st.ph[0] = f32st.ph[1] = 0;
    const bool f32_resid = MODE == EPI_F32 && ep.resid != nullptr;
    if (f32_resid && pair < num_tiles && lane == 0)   // residual tiles of the first output tile
      epi_f32_prime(&tmO2, tile_s, my_bar, (pair / tiles_n) * BM + (int)rank * 128 + q * 32,
                    (pair % tiles_n) * BN + hh * (BN / 2));
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const int row0 = (tile / tiles_n) * BM + (int)rank * 128 + q * 32;
      const int col0 = (tile % tiles_n) * BN + hh * (BN / 2);
      {  // while this tile's mainloop runs: pull the NEXT tile's aux / residual rows into L2
        const int nt = tile + num_pairs;
        if (nt < num_tiles)
          epi_l2_prefetch<MODE, BN / 2 / 32>(ep, (nt / tiles_n) * BM + (int)rank * 128 + q * 32,
                                             (nt % tiles_n) * BN + hh * (BN / 2), M, lane);
      }
      mbar_wait_relaxed(smem_u32(&tfull_bar[buf]), bphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + buf * BN + hh * (BN / 2) + ((uint32_t)(q * 32) << 16);
      if (!(dbg & 1)) {
        if (MODE == EPI_F32) {
          epi_warp_tile_f32_tma<BN / 2 / 32>(ep, &tmO, &tmO2, t_addr, tile_s, my_bar, f32st, row0,
                                             col0, N, lane);
          const int nt = tile + num_pairs;
          if (f32_resid && nt < num_tiles && lane == 0)
            epi_f32_prime(&tmO2, tile_s, my_bar, (nt / tiles_n) * BM + (int)rank * 128 + q * 32,
                          (nt % tiles_n) * BN + hh * (BN / 2));
        } else
          epi_warp_tile_tma<MODE, BN / 2 / 32>(ep, &tmO, &tmO2, t_addr, tile_s, row0, col0, M, N,
                                               lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(te_leader + buf * 8);
    }
  }

  if (warp >= 2 && lane == 0) tma_store_wait<0>();  // bulk stores landed
  // no CTA may exit (or free TMEM) while its peer can still signal its barriers / read its smem
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2<kTmemCols>(tmem_base);
  }
}

template <int MODE>
int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                 const CUtensorMap& tmO2, int M, int N, int K, const EpiParams& ep,
                 cudaStream_t stream) {
  LLC_CONFIGURE_SMEM(gemm2_kernel<MODE>, kSmem);
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  const int pairs = llc_num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  LLC_PROF_BEGIN(LLC_K_GEMM, M, N, K, 2.0 * M * N * K,
                 2.0 * ((double)M * K + (double)N * K) + (double)M * N * (ep.out_fp32 ? 4 : 2),
                 stream);
  static const int dbg = llc_dev_env("LLC_GEMM_DBG") ? atoi(llc_dev_env("LLC_GEMM_DBG")) : 0;
  EpiParams ep2 = ep;
  ep2.dbg = dbg;
  // a bf16 output that fits L2 (dh, d_o: 77 MB) is read again by the next kernel(s): let it stay
  static const bool nokeep = llc_dev_env("LLC_GEMM_NOKEEP") != nullptr;
  ep2.keep_out = (!nokeep && !ep.out_fp32 && (double)M * N * 2.0 <= 100e6) ? 1 : 0;
  LLC_CUDA(llc_launch_pdl(gemm2_kernel<MODE>, dim3(grid), dim3(kThreads), kSmem, stream, tmA, tmB, tmO,
                          tmO2, M, N, K, ep2, dbg));
  LLC_PROF_END(stream);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("gemm2_kernel");
  return 0;
}

}  // namespace

// Chosen by llc_gemm_bf16_tn for shapes that fill the machine with 256 x 256 pair tiles.
bool llc_gemm2_eligible(int M, int N, int K) {
  if (N % BN != 0 || K < BK) return false;
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  return tiles >= llc_num_sms() / 2;
}

int llc_gemm2_launch(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                     const EpiParams& ep, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  int rc = llc_encode_tmap_2d(&tmA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K,
                              (uint64_t)M, (uint64_t)lda * 2, BK, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = llc_encode_tmap_2d(&tmB, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)N,
                          (uint64_t)ldb * 2, BK, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  // bf16 outputs leave through TMA stores: [32 rows x 64 columns] boxes, 128 B swizzle
  CUtensorMap tmO = tmA, tmO2 = tmA;
  const int mode = epi_mode_of(ep);
  if (mode == EPI_F32) {
    // fp32 out (tmO) and residual (tmO2): [32 rows x 32 columns] boxes = 128 B lines
    rc = llc_encode_tmap_2d(&tmO, ep.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)N,
                            (uint64_t)M, (uint64_t)ep.ld_out * 4, 32, 32,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (ep.resid != nullptr) {
      rc = llc_encode_tmap_2d(&tmO2, ep.resid, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)N,
                              (uint64_t)M, (uint64_t)ep.ld_resid * 4, 32, 32,
                              CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  } else {
    if (ep.out != nullptr) {
      rc = llc_encode_tmap_2d(&tmO, ep.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)N,
                              (uint64_t)M, (uint64_t)ep.ld_out * 2, 64, 32,
                              CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    if (mode == EPI_GELU) {
      rc = llc_encode_tmap_2d(&tmO2, ep.out2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)N,
                              (uint64_t)M, (uint64_t)ep.ld_out2 * 2, 64, 32,
                              CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  }
  switch (mode) {
    case EPI_GELU: return launch_gemm2<EPI_GELU>(tmA, tmB, tmO, tmO2, M, N, K, ep, stream);
    case EPI_DGELU: return launch_gemm2<EPI_DGELU>(tmA, tmB, tmO, tmO2, M, N, K, ep, stream);
    case EPI_F32: return launch_gemm2<EPI_F32>(tmA, tmB, tmO, tmO2, M, N, K, ep, stream);
    default: return launch_gemm2<EPI_BF16>(tmA, tmB, tmO, tmO2, M, N, K, ep, stream);
  }
}
