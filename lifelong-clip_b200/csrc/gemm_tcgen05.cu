// TMA-fed tcgen05/TMEM GEMM for sm_100a: out[M,N] = epi(A[M,K] . B[N,K]^T), bf16 -> fp32.
//
// Replaces every dense contraction of the reference's ViT block (F.linear calls at
// models/clip/lora.py:837-839,1072-1074 and models/clip/model.py:219-222) and their
// activation-gradient transposes. The rank-r LoRA update rides in 16 extra K columns of both
// operands, so the same accumulator holds W.x + s.B.(A.x).
//
// Persistent, warp-specialised:
//   warp 0      TMA producer   (one lane): A/B 128B-swizzled K-major tiles -> smem ring
//   warp 1      MMA issuer     (one lane): tcgen05.mma 128 x BN x 16, accumulators in TMEM
//   warps 2..5  epilogue: tcgen05.ld -> smem transpose -> coalesced global I/O
// TMEM holds two BN-column accumulators so tile i's epilogue overlaps tile i+1's mainloop.
#include <stdlib.h>

#include "gemm_epilogue.cuh"

bool llc_gemm2_eligible(int M, int N, int K, bool have_ws);
int llc_gemm2_launch(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                     const EpiParams& ep, cudaStream_t stream);

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kThreads = 192;
constexpr int kEpiWarps = 4;
constexpr int kEpiPad = 33;  // 32x32 fp32 transpose tile, padded against bank conflicts

template <int BN>
struct Cfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 10);  // BN = 32: skinny, HBM-bound
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBytes = kEpiWarps * 32 * kEpiPad * 4;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmem = 1024 /*align slack*/ + kStages * kStageBytes + kEpiBytes + kBarBytes;
  static constexpr int kTmemCols = 2 * BN;  // 256 or 512: power of two
};

// Coalesced-form epilogue on V consecutive columns of one row.
template <int V>
__device__ __forceinline__ void epi_store(const EpiParams& p, int row, int col, float (&v)[V]) {
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (p.act == 1) {
    if (p.out != nullptr) {
      uint32_t* zp = reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                                 (size_t)row * p.ld_out + col);
      if (V == 8) {
        uint4 q = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                             pack_bf16(v[6], v[7]));
        *reinterpret_cast<uint4*>(zp) = q;
      } else {
        *reinterpret_cast<uint2*>(zp) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
      }
    }
    // the activation is applied to the bf16-rounded pre-activation so that backward (which
    // only sees the saved bf16 z) differentiates exactly the function forward evaluated
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = quick_gelu(__bfloat162float(__float2bfloat16_rn(v[i])));
    uint32_t* gp = reinterpret_cast<uint32_t*>(p.out2 + (size_t)row * p.ld_out2 + col);
    if (V == 8) {
      *reinterpret_cast<uint4*>(gp) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                                 pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    } else {
      *reinterpret_cast<uint2*>(gp) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    }
    return;
  }
  if (p.act == 2) {
    const uint32_t* zp =
        reinterpret_cast<const uint32_t*>(p.aux + (size_t)row * p.ld_aux + col);
#pragma unroll
    for (int i = 0; i < V; i += 2) {
      float2 z = unpack_bf16(__ldg(zp + i / 2));
      v[i] *= quick_gelu_grad(z.x);
      v[i + 1] *= quick_gelu_grad(z.y);
    }
  }
  if (p.resid != nullptr) {
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      float4 r = *reinterpret_cast<const float4*>(p.resid + (size_t)row * p.ld_resid + col + i);
      v[i] += r.x; v[i + 1] += r.y; v[i + 2] += r.z; v[i + 3] += r.w;
    }
  }
  if (p.out_fp32) {
    float* op = reinterpret_cast<float*>(p.out) + (size_t)row * p.ld_out + col;
#pragma unroll
    for (int i = 0; i < V; i += 4)
      *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
    uint32_t* op = reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                               (size_t)row * p.ld_out + col);
    if (V == 8) {
      *reinterpret_cast<uint4*>(op) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                                 pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    } else {
      *reinterpret_cast<uint2*>(op) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               int M, int N, int K, EpiParams ep) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* smem_ab = smem;
  float* smem_epi = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes + C::kEpiBytes);
  uint64_t* full_bar = bars;                      // [kStages]
  uint64_t* empty_bar = bars + C::kStages;        // [kStages]
  uint64_t* tfull_bar = bars + 2 * C::kStages;    // [2]
  uint64_t* tempty_bar = bars + 2 * C::kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  const int tiles_m = (M + BM - 1) / BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tfull_bar[b]), 1);
      mbar_init(smem_u32(&tempty_bar[b]), kEpiWarps * 32);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // (warp-uniform control flow, one elected lane issues: a divergent `if (lane == 0)` makes
    // ptxas wrap every TMA / MMA operand in a register->uniform-register waterfall loop)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM;
      const int n0 = (tile % tiles_n) * BN;
      if ((ep.dbg & 4096) && tile + (int)gridDim.x >= num_tiles && elect_one()) pdl_trigger();
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, C::kStageBytes);
          const uint32_t sa = smem_u32(smem_ab + stage * C::kStageBytes);
          tma_load_2d(sa, &tmA, fb, kb * BK, m0);
          tma_load_2d(sa + C::kABytes, &tmB, fb, kb * BK, n0);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      mbar_wait(smem_u32(&tempty_bar[buf]), bphase ^ 1);  // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem_ab + stage * C::kStageBytes);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + C::kABytes);
          const int ksteps = min(BK / 16, (K - kb * BK) / 16);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row (start address is in 16 B units)
            if (k < ksteps) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(smem_u32(&empty_bar[stage]));  // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) umma_commit(smem_u32(&tfull_bar[buf]));  // accumulator complete
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    float* tile_s = smem_epi + (warp - 2) * 32 * kEpiPad;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t bphase = (it >> 1) & 1;
      const int m0 = (tile / tiles_n) * BM + q * 32;
      const int n0 = (tile % tiles_n) * BN;
      mbar_wait(smem_u32(&tfull_bar[buf]), bphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_addr + c0, r);
        tmem_ld_wait();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) tile_s[lane * kEpiPad + j] = __uint_as_float(r[j]);
        __syncwarp();
        if (n0 + c0 < N) {
          if (ep.out_fp32) {
            // 4 columns per lane: 8 lanes cover one 128 B row segment, 4 rows per pass
#pragma unroll
            for (int pass = 0; pass < 8; ++pass) {
              const int rr = pass * 4 + (lane >> 3);
              const int cc = (lane & 7) * 4;
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = tile_s[rr * kEpiPad + cc + e];
              const int row = m0 + rr, col = n0 + c0 + cc;
              if (row < M && col < N) epi_store<4>(ep, row, col, v);
            }
          } else {
            // 8 columns per lane (16 B of bf16): 4 lanes per row, 8 rows per pass
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
              const int rr = pass * 8 + (lane >> 2);
              const int cc = (lane & 3) * 8;
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = tile_s[rr * kEpiPad + cc + e];
              const int row = m0 + rr, col = n0 + c0 + cc;
              if (row < M && col < N) epi_store<8>(ep, row, col, v);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty_bar[buf]));
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

template <int BN>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K,
                const EpiParams& ep, cudaStream_t stream) {
  using C = Cfg<BN>;
  LLC_CONFIGURE_SMEM(gemm_tn_kernel<BN>, C::kSmem);
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < llc_num_sms() ? tiles : llc_num_sms();
  LLC_PROF_BEGIN(LLC_K_GEMM, M, N, K, 2.0 * M * N * K,
                 2.0 * ((double)M * K + (double)N * K) + (double)M * N * (ep.out_fp32 ? 4 : 2),
                 stream);
  LLC_CUDA(llc_launch_pdl(gemm_tn_kernel<BN>, dim3(grid), dim3(kThreads), C::kSmem, stream, tmA, tmB, M,
                          N, K, ep));
  LLC_PROF_END(stream);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("gemm_tn_kernel");
  return 0;
}

}  // namespace

extern "C" int llc_gemm_bf16_tn(const void* A, int lda, const void* B, int ldb, int M, int N,
                                int K, const llc_gemm_epi* e, void* stream) {
  LLC_REQUIRE(A && B && e, "llc_gemm_bf16_tn: null operand");
  LLC_REQUIRE(M > 0 && N > 0 && K > 0, "llc_gemm_bf16_tn: empty problem M=%d N=%d K=%d", M, N, K);
  LLC_REQUIRE(K % 16 == 0, "llc_gemm_bf16_tn: K=%d must be a multiple of 16", K);
  LLC_REQUIRE(N % 8 == 0, "llc_gemm_bf16_tn: N=%d must be a multiple of 8", N);
  LLC_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= K,
              "llc_gemm_bf16_tn: lda=%d ldb=%d must be multiples of 8 and >= K=%d", lda, ldb, K);
  LLC_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0,
              "llc_gemm_bf16_tn: operands must be 16-byte aligned");
  LLC_REQUIRE(e->act >= 0 && e->act <= 2, "llc_gemm_bf16_tn: bad act %d", e->act);
  if (e->act == 1) {
    LLC_REQUIRE(e->out2 && e->ld_out2 % 8 == 0 && !e->out_fp32 && !e->resid,
                "llc_gemm_bf16_tn: act=1 needs bf16 out2 and no residual");
    LLC_REQUIRE(e->out == nullptr || e->ld_out % 8 == 0, "llc_gemm_bf16_tn: ld_out %% 8");
  } else {
    LLC_REQUIRE(e->out && e->ld_out % 8 == 0, "llc_gemm_bf16_tn: out missing / ld_out %% 8");
  }
  if (e->act == 2) LLC_REQUIRE(e->aux && e->ld_aux % 8 == 0, "llc_gemm_bf16_tn: act=2 needs aux");
  if (e->resid) LLC_REQUIRE(e->ld_resid % 4 == 0, "llc_gemm_bf16_tn: ld_resid %% 4");

  EpiParams ep;
  ep.bias = e->bias; ep.resid = e->resid; ep.ld_resid = e->ld_resid; ep.act = e->act;
  ep.aux = reinterpret_cast<const __nv_bfloat16*>(e->aux); ep.ld_aux = e->ld_aux;
  ep.out = e->out; ep.ld_out = e->ld_out; ep.out_fp32 = e->out_fp32;
  ep.out2 = reinterpret_cast<__nv_bfloat16*>(e->out2); ep.ld_out2 = e->ld_out2;
  ep.dbg = g_llc_pdl_trigger ? 4096 : 0;
  ep.keep_out = 0;
  ep.ws = e->ws; ep.ws_bytes = e->ws_bytes;
  LLC_REQUIRE(e->ws == nullptr || ((uintptr_t)e->ws & 15) == 0, "llc_gemm_bf16_tn: ws misaligned");

  // production shapes: 256 x 256 tiles on CTA pairs (gemm2_tcgen05.cu); LLC_GEMM_1CTA=1 forces
  // the single-CTA kernel below (debugging / A-B comparison)
  static const bool force_1cta = llc_dev_env("LLC_GEMM_1CTA") != nullptr;
  if (!force_1cta &&
      llc_gemm2_eligible(M, N, K, e->ws != nullptr && e->ws_bytes >= llc_gemm_ws_bytes()))
    return llc_gemm2_launch(A, lda, B, ldb, M, N, K, ep, reinterpret_cast<cudaStream_t>(stream));

  // BN=256 keeps smem traffic per MMA lowest; fall back to 128-wide tiles when the problem would
  // leave most SMs without a tile.
  const int tiles256 = ((M + BM - 1) / BM) * ((N + 255) / 256);
  const bool use256 = (N % 256 == 0) && tiles256 >= llc_num_sms();
  // N <= 32: rank-r row products (N = 16), stream A at HBM rate. M <= 512 (the class-token-only
  // last block: M = images): 32-wide tiles spread the few rows over 4x more CTAs
  static const bool small_m32 = llc_dev_env("LLC_GEMM_SMALLM_BN128") == nullptr;
  const bool use32 = N <= 32 || (small_m32 && M <= 512 && N % 32 == 0 && !use256);
  const int BN = use256 ? 256 : (use32 ? 32 : 128);

  CUtensorMap tmA, tmB;
  int rc = llc_encode_tmap_2d(&tmA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K,
                              (uint64_t)M, (uint64_t)lda * 2, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = llc_encode_tmap_2d(&tmB, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)N,
                          (uint64_t)ldb * 2, BK, BN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (use32) return launch_gemm<32>(tmA, tmB, M, N, K, ep, st);
  return use256 ? launch_gemm<256>(tmA, tmB, M, N, K, ep, st)
                : launch_gemm<128>(tmA, tmB, M, N, K, ep, st);
}
