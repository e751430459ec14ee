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
//
// Stream-K schedule (SK = true): with T = 197 x images tokens the tile count rarely divides the 74
// CTA pairs (32 images, N = 768: 75 tiles -> a second wave holding ONE tile, 51 % efficiency). The
// (tile, k-block) units are then cut into 74 equal contiguous ranges. A range that starts inside
// a tile (k0 > 0) leaves its fp32 partial in a caller-provided workspace and raises a flag; the
// pair whose range holds the tile's k-block 0 owns the tile: it adds the partials of the pairs
// after it into its TMEM accumulator and runs the normal epilogue. A contributor's partial is the
// FIRST thing it computes and an owner's tile the LAST, so nobody waits in practice, and the
// result is deterministic (fixed ranges, fixed summation order).
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
constexpr int kSkFlagBytes = 8192;                    // flags[pair][rank][epilogue warp] (ints)
constexpr int kSkSlotFloats = 32 * (BN / 2);          // one epilogue warp's 32 rows x 128 columns

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait_all() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Work list of one CTA pair. !SK: whole tiles pair, pair + P, ...; SK: the pair's contiguous
// range of (tile, k-block) units, cut at tile boundaries into segments [kb0, kb1).
template <bool SK>
struct WorkIter {
  int num_pairs, num_tiles, num_kb, tile_next;
  long long u, u1;
  __device__ WorkIter(int pair, int num_pairs_, int num_tiles_, int num_kb_)
      : num_pairs(num_pairs_), num_tiles(num_tiles_), num_kb(num_kb_), tile_next(pair) {
    const long long total = (long long)num_tiles_ * num_kb_;
    u = SK ? total * pair / num_pairs_ : 0;
    u1 = SK ? total * (pair + 1) / num_pairs_ : 0;
  }
  __device__ bool next(int& tile, int& kb0, int& kb1) {
    if (SK) {
      if (u >= u1) return false;
      tile = (int)(u / num_kb);
      kb0 = (int)(u - (long long)tile * num_kb);
      const long long len = min((long long)(num_kb - kb0), u1 - u);
      kb1 = kb0 + (int)len;
      u += len;
      return true;
    }
    if (tile_next >= num_tiles) return false;
    tile = tile_next;
    tile_next += num_pairs;
    kb0 = 0;
    kb1 = num_kb;
    return true;
  }
  // next segment / next segment this pair OWNS (kb0 == 0), without advancing; tile -1 if none
  __device__ void peek(int& tile, int& kb0) const {
    WorkIter c = *this;
    int kb1;
    if (!c.next(tile, kb0, kb1)) tile = -1;
  }
  __device__ int peek_owned() const {
    WorkIter c = *this;
    int t, a, b;
    while (c.next(t, a, b))
      if (a == 0) return t;
    return -1;
  }
};

template <int MODE, bool SK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmO2,
             int M, int N, int K, EpiParams ep, int dbg, uint8_t* sk_ws) {
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
    WorkIter<SK> wi(pair, num_pairs, num_tiles, num_kb);
    int tile, kb0, kb1;
    while (wi.next(tile, kb0, kb1)) {
      const int m0 = (tile / tiles_n) * BM + (int)rank * 128;
      const int n0 = (tile % tiles_n) * BN + (int)rank * 128;
      int next_tile, next_kb0;
      wi.peek(next_tile, next_kb0);
      const int m0_next = (next_tile / tiles_n) * BM + (int)rank * 128;
      if ((dbg & 4096) && next_tile < 0 && elect_one()) pdl_trigger();   // last work item
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!(dbg & 32)) mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const uint32_t fb_local = smem_u32(&full_bar[stage]);
        const uint32_t fb_leader = mapa_shared(fb_local, 0);
        const uint32_t sa = smem_u32(smem_ab + stage * kStageBytes);
        if (elect_one()) {
          // A (activations) streams from HBM once per GEMM: pull it into L2 well ahead of the
          // smem ring so the ring's own loads see L2 latency. B (weights) stays L2-resident.
          if (!(dbg & 8)) {
            const int pk = kb + kPrefetchKb;
            if (pk < kb1) tma_prefetch_2d(&tmA, pk * BK, m0);
            else if (next_tile >= 0 && next_kb0 + pk - kb1 < num_kb)
              tma_prefetch_2d(&tmA, (next_kb0 + pk - kb1) * BK, m0_next);
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
      WorkIter<SK> wi(pair, num_pairs, num_tiles, num_kb);
      int tile, kb0, kb1;
      for (; wi.next(tile, kb0, kb1); ++it) {
        const int buf = it & 1;
        const uint32_t bphase = (it >> 1) & 1;
        mbar_wait(smem_u32(&tempty_bar[buf]), bphase ^ 1);  // both epilogues drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!(dbg & 64)) mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_ab + stage * kStageBytes);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + kABytes);
          const int ksteps = min(BK / 16, (K - kb * BK) / 16);
          if (elect_one()) {
            if (!(dbg & 4))
              for (int k = 0; k < ksteps; ++k)
                umma_bf16_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0);
            if (!(dbg & 128))
              umma_commit_mc(smem_u32(&empty_bar[stage]), 0x3);  // frees the slot in both CTAs
            if (kb == kb1 - 1)
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
    WorkIter<SK> wi(pair, num_pairs, num_tiles, num_kb);
    {
      const int ft = wi.peek_owned();   // residual tiles of the first output tile
      if (f32_resid && ft >= 0 && lane == 0)
        epi_f32_prime(&tmO2, tile_s, my_bar, (ft / tiles_n) * BM + (int)rank * 128 + q * 32,
                      (ft % tiles_n) * BN + hh * (BN / 2));
    }
    int* sk_flags = reinterpret_cast<int*>(sk_ws);
    float* sk_part = reinterpret_cast<float*>(sk_ws + kSkFlagBytes);
    const long long sk_total = (long long)num_tiles * num_kb;
    int it = 0;
    int tile, kb0, kb1;
    for (; wi.next(tile, kb0, kb1); ++it) {
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const int row0 = (tile / tiles_n) * BM + (int)rank * 128 + q * 32;
      const int col0 = (tile % tiles_n) * BN + hh * (BN / 2);
      const int nt = wi.peek_owned();
      if (nt >= 0)   // while this tile's mainloop runs: pull the NEXT tile's aux / residual rows into L2
        epi_l2_prefetch<MODE, BN / 2 / 32>(ep, (nt / tiles_n) * BM + (int)rank * 128 + q * 32,
                                           (nt % tiles_n) * BN + hh * (BN / 2), M, lane);
      mbar_wait_relaxed(smem_u32(&tfull_bar[buf]), bphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + buf * BN + hh * (BN / 2) + ((uint32_t)(q * 32) << 16);
      if (SK && kb0 > 0) {
        // contributor: this pair's range starts inside the tile. Leave the fp32 partial and raise
        // the flag
        const size_t slot = ((size_t)(pair * 2 + (int)rank) * kEpiWarps + ew);
        // slot layout: float4 (chunk c, quad j, lane) at (c * 8 + j) * 32 + lane - every store /
        // load instruction of the warp moves 512 contiguous bytes
        float4* dst = reinterpret_cast<float4*>(sk_part + slot * kSkSlotFloats);
#pragma unroll 1
        for (int c = 0; c < BN / 2 / 32; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32(t_addr + c * 32, acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(dst + (c * 8 + j) * 32 + lane,
                   make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                               __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3])));
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_gpu(sk_flags + slot, 1);
      } else {
        if (SK && kb1 < num_kb) {
          // owner of a split tile: add the partials of the pairs whose ranges start inside it
          const long long tile_end = (long long)(tile + 1) * num_kb;
          int q1 = pair + 1;
          while (q1 < num_pairs && sk_total * q1 / num_pairs < tile_end) ++q1;   // contributors: (pair, q1)
          for (int cq = pair + 1; cq < q1; ++cq) {
            const int* f = sk_flags + ((size_t)(cq * 2 + (int)rank) * kEpiWarps + ew);
            if (lane == 0) {
              unsigned spins = 0;
              while (ld_acquire_gpu(f) == 0) {
                __nanosleep(64);
                if (++spins > (1u << 25)) __trap();   // seconds: a lost partial, not a slow one
              }
            }
            __syncwarp();
          }
#pragma unroll 1
          for (int c = 0; c < BN / 2 / 32; ++c) {
            uint32_t acc[32];
            tmem_ld_32x32(t_addr + c * 32, acc);
            tmem_ld_wait();
            for (int cq = pair + 1; cq < q1; ++cq) {
              const float4* src = reinterpret_cast<const float4*>(
                  sk_part + ((size_t)(cq * 2 + (int)rank) * kEpiWarps + ew) * kSkSlotFloats);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 v = __ldcg(src + (c * 8 + j) * 32 + lane);
                acc[4 * j] = __float_as_uint(__uint_as_float(acc[4 * j]) + v.x);
                acc[4 * j + 1] = __float_as_uint(__uint_as_float(acc[4 * j + 1]) + v.y);
                acc[4 * j + 2] = __float_as_uint(__uint_as_float(acc[4 * j + 2]) + v.z);
                acc[4 * j + 3] = __float_as_uint(__uint_as_float(acc[4 * j + 3]) + v.w);
              }
            }
            tmem_st_32x32(t_addr + c * 32, acc);
          }
          tmem_st_wait_all();
          __syncwarp();
          if (lane == 0)
            for (int cq = pair + 1; cq < q1; ++cq)
              sk_flags[(size_t)(cq * 2 + (int)rank) * kEpiWarps + ew] = 0;   // clean for the next launch
        }
        if (!(dbg & 1)) {
          if (MODE == EPI_F32) {
            epi_warp_tile_f32_tma<BN / 2 / 32>(ep, &tmO, &tmO2, t_addr, tile_s, my_bar, f32st, row0,
                                               col0, N, lane);
            if (f32_resid && nt >= 0 && lane == 0)
              epi_f32_prime(&tmO2, tile_s, my_bar, (nt / tiles_n) * BM + (int)rank * 128 + q * 32,
                            (nt % tiles_n) * BN + hh * (BN / 2));
          } else
            epi_warp_tile_tma<MODE, BN / 2 / 32>(ep, &tmO, &tmO2, t_addr, tile_s, row0, col0, M, N,
                                                 lane);
        }
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

int g_stream_k = 1;   // llc_gemm_set_stream_k

// stream-K pays when the plain schedule leaves a poorly filled last wave and every pair still gets
// a few k-blocks
bool gemm2_wants_sk(int M, int N, int K) {
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  const int pairs = llc_num_sms() / 2;
  const int num_kb = (K + BK - 1) / BK;
  const int waves = (tiles + pairs - 1) / pairs;
  const double eff = (double)tiles / ((double)waves * pairs);
  // Measured per shape (tools/sk_bench.py -> profiles/r02_streamk_ab.txt): the schedule pays when
  // the whole problem is about ONE wave (39 / 75 tiles at 16 / 32 images: 1.2-1.5x at K = 3072,
  // 1.1-1.3x at K = 2320) and K is long enough to amortise the partial's trip through L2 and the
  // un-overlapped fix-up epilogue (~6 us; the K = 768 / 784 GEMMs, whose whole tile takes 2 us,
  // got slower). From two waves on it ties (64 images) or loses (128 images: 0.86x): contiguous
  // unit ranges make the pairs walk far-apart tiles, so the three n-tiles that share an A row
  // block are no longer in flight together and A is re-read from HBM.
  return eff < 0.92 && num_kb >= 32 && tiles <= pairs + pairs / 4 &&
         (long long)tiles * num_kb >= 4LL * pairs;
}

template <int MODE>
int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                 const CUtensorMap& tmO2, int M, int N, int K, const EpiParams& ep,
                 cudaStream_t stream) {
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  const int pairs = llc_num_sms() / 2;
  const bool sk = g_stream_k && ep.ws != nullptr && ep.ws_bytes >= llc_gemm_ws_bytes() &&
                  gemm2_wants_sk(M, N, K);
  const int grid = 2 * ((sk || tiles >= pairs) ? pairs : tiles);
  LLC_PROF_BEGIN(LLC_K_GEMM, M, N, K, 2.0 * M * N * K,
                 2.0 * ((double)M * K + (double)N * K) + (double)M * N * (ep.out_fp32 ? 4 : 2),
                 stream);
  static const int dbg = llc_dev_env("LLC_GEMM_DBG") ? atoi(llc_dev_env("LLC_GEMM_DBG")) : 0;
  EpiParams ep2 = ep;
  ep2.dbg = dbg;
  const int kdbg = dbg | (g_llc_pdl_trigger ? 4096 : 0);
  // a bf16 output that fits L2 (dh, d_o: 77 MB) is read again by the next kernel(s): let it stay
  static const bool nokeep = llc_dev_env("LLC_GEMM_NOKEEP") != nullptr;
  ep2.keep_out = (!nokeep && !ep.out_fp32 && (double)M * N * 2.0 <= 100e6) ? 1 : 0;
  if (sk) {
    LLC_CONFIGURE_SMEM((gemm2_kernel<MODE, true>), kSmem);
    LLC_CUDA(llc_launch_pdl(gemm2_kernel<MODE, true>, dim3(grid), dim3(kThreads), kSmem, stream,
                            tmA, tmB, tmO, tmO2, M, N, K, ep2, kdbg,
                            reinterpret_cast<uint8_t*>(ep.ws)));
  } else {
    LLC_CONFIGURE_SMEM((gemm2_kernel<MODE, false>), kSmem);
    LLC_CUDA(llc_launch_pdl(gemm2_kernel<MODE, false>, dim3(grid), dim3(kThreads), kSmem, stream,
                            tmA, tmB, tmO, tmO2, M, N, K, ep2, kdbg, (uint8_t*)nullptr));
  }
  LLC_PROF_END(stream);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("gemm2_kernel");
  return 0;
}

}  // namespace

extern "C" int llc_gemm_set_stream_k(int on) {
  const int old = g_stream_k;
  g_stream_k = on != 0;
  return old;
}

extern "C" size_t llc_gemm_ws_bytes(void) {
  return (size_t)kSkFlagBytes +
         (size_t)(llc_num_sms() / 2) * 2 * kEpiWarps * kSkSlotFloats * sizeof(float);
}

// Chosen by llc_gemm_bf16_tn for shapes that fill the machine with 256 x 256 pair tiles, or -
// with a stream-K workspace - whose (tile, k-block) units do.
bool llc_gemm2_eligible(int M, int N, int K, bool have_ws) {
  if (N % BN != 0 || K < BK) return false;
  const int tiles = ((M + BM - 1) / BM) * (N / BN);
  if (tiles >= llc_num_sms() / 2) return true;
  // one or two k-blocks (the adapter's up-projection, K = 64): the launch is its epilogue, and
  // this kernel's TMA-store epilogue beats the single-CTA kernel's even on a partly filled grid
  // (M = 7700, N = 512, K = 64: 51 us there)
  if (K <= 2 * BK && tiles >= 16) return true;
  return g_stream_k && have_ws && tiles >= 16 && gemm2_wants_sk(M, N, K);
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
