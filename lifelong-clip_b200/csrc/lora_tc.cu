// LoRA weight-gradient reductions on the tensor cores.
//   P[c, j] = sum_t X[t, c] * w[t, j]      X bf16 [T, C] (row pitch ld_x), w bf16 [T, 16]
// (dB = s g^T u, dA = du^T h of models/clip/lora.py:838-839,1072-1074 under autograd). The op is a
// [C x T] . [T x 16] contraction that streams X once: HBM-bound. Each CTA owns 128 columns of X
// and a contiguous slice of the tokens; TMA streams [64 tokens x 64 columns] boxes through an
// 8-stage ring, and tcgen05.mma consumes them AS THEY LIE: X tiles are the MN-major A operand
// (M = columns), the 16-wide w rows are the MN-major B operand. The per-slice results go out as
// partials that llc_lora_colsum_finish adds in a fixed order (bit-deterministic).
#include "common.cuh"

namespace {

constexpr int kKB = 64;                      // tokens per k-block
constexpr int kATile = 2 * kKB * 128;        // two 64-column atoms
constexpr int kBTile = kKB * 128;            // w rows, 32 B used per 128 B row
constexpr int kStage = kATile + kBTile;
constexpr int kStages = 8;
constexpr int kSmem = 1024 + kStages * kStage + 256;
constexpr int kThreads = 6 * 32;             // TMA, MMA, 4 x (w loader + epilogue)

__global__ void __launch_bounds__(kThreads, 1)
lora_colsum_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __nv_bfloat16* __restrict__ w,
                      int ld_w, int T, int C, int R, int tok_per_split,
                      float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStage);
  uint64_t* full_bar = bars;               // [kStages] TMA bytes + the 4 loader warps
  uint64_t* empty_bar = bars + kStages;    // [kStages]
  uint64_t* done_bar = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 128;
  const int t_begin = blockIdx.y * tok_per_split;
  const int t_end = min(T, t_begin + tok_per_split);
  const int num_kb = t_end > t_begin ? (t_end - t_begin + kKB - 1) / kKB : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1 + 128);   // TMA issuer + every loader thread
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(done_bar), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<32>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&full_bar[stage]);
        const uint32_t sa = smem_u32(smem + stage * kStage);
        mbar_expect_tx(fb, kATile);
        tma_load_2d(sa, &tmX, fb, c0, t_begin + kb * kKB);
        tma_load_2d(sa + kKB * 128, &tmX, fb, c0 + 64, t_begin + kb * kKB);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, 16, 1, 1);  // A and B both MN-major
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      fence_proxy_async_smem();   // the w rows were written through the generic proxy (cp.async)
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * kStage);
#pragma unroll
        for (int ks = 0; ks < kKB / 16; ++ks)
          umma_bf16(tmem_base, umma_desc_mn_sw128(sa + ks * 2048, kKB * 128, 1024),
                    umma_desc_mn_sw128(sa + kATile + ks * 2048, 8192, 1024), idesc, (kb | ks) != 0);
        umma_commit(smem_u32(&empty_bar[stage]));
        if (kb == num_kb - 1) umma_commit(smem_u32(done_bar));
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // w loader: warp i stages token rows [16 i, 16 i + 16) of every k-block (two 16 B pieces per
    // row, TMA's 128 B swizzle pattern so the tile reads as an MN-major operand)
    const int wi = warp - 2;
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
      const int row = wi * 16 + (lane >> 1), piece = lane & 1;
      const int t = t_begin + kb * kKB + row;
      // asynchronous 16 B copy (zero-filled past the slice end); its completion arrives on the
      // stage's full barrier, so the loader never waits for a load
      const uint32_t dst = smem_u32(smem + stage * kStage + kATile + row * 128 +
                                    ((piece ^ (row & 7)) << 4));
      const __nv_bfloat16* src = w + (size_t)(t < t_end ? t : t_begin) * ld_w + piece * 8;
      const uint32_t nbytes = t < t_end ? 16u : 0u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes)
                   : "memory");
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                       smem_u32(&full_bar[stage]))
                   : "memory");
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    // epilogue: accumulator row = column c0 + 32 q + lane of X
    float* out = partial + ((size_t)blockIdx.y * C + c0) * R;
    const int q = warp & 3;
    const int col = q * 32 + lane;
    if (num_kb > 0) {
      mbar_wait(smem_u32(done_bar), 0);
      tc_fence_after();
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
            "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(tmem_base + ((uint32_t)(q * 32) << 16))
          : "memory");
      tmem_ld_wait();
      for (int j = 0; j < R; ++j) out[(size_t)col * R + j] = __uint_as_float(v[j]);
    } else {
      for (int j = 0; j < R; ++j) out[(size_t)col * R + j] = 0.f;
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<32>(tmem_base);
  }
}

}  // namespace

bool llc_colsum_tc_eligible(const void* X, int ld_x, int T, int C, const void* w, int ld_w) {
  return C % 128 == 0 && T >= 1024 && ld_x % 8 == 0 && ld_w % 8 == 0 &&
         ((uintptr_t)X & 15) == 0 && ((uintptr_t)w & 15) == 0;
}

int llc_colsum_tc(const void* X, int ld_x, int T, int C, int R, const void* w, int ld_w,
                  float* partial, int* n_partials, cudaStream_t st) {
  CUtensorMap tm;
  if (int rc = llc_encode_tmap_2d(&tm, X, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)C,
                                  (uint64_t)T, (uint64_t)ld_x * 2, 64, kKB,
                                  CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  const int mtiles = C / 128;
  int splits = llc_num_sms() / mtiles;
  if (splits < 1) splits = 1;
  int per = (T + splits - 1) / splits;
  per = (per + kKB - 1) / kKB * kKB;          // k-blocks never straddle two slices
  splits = (T + per - 1) / per;
  static bool configured = false;
  if (!configured) {
    LLC_CUDA(cudaFuncSetAttribute(lora_colsum_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kSmem));
    configured = true;
  }
  LLC_PROF_BEGIN(LLC_K_LORA_SIDE, T, C, 2, 2.0 * T * C * 16, 2.0 * T * C, st);
  LLC_CUDA(llc_launch_pdl(lora_colsum_tc_kernel, dim3(mtiles, splits), dim3(kThreads), kSmem, st, tm,
                          reinterpret_cast<const __nv_bfloat16*>(w), ld_w, T, C, R, per, partial));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("lora_colsum_tc_kernel");
  *n_partials = splits;
  return 0;
}
