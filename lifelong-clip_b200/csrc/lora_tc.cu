// LoRA weight-gradient reductions on the tensor cores.
//   P[c, j] = sum_t X[t, c] * w[t, j]      X bf16 [T, C] (row pitch ld_x), w bf16 [T, 16]
// (dB = s g^T u, dA = du^T h of models/clip/lora.py:838-839,1072-1074 under autograd). The op is a
// [C x T] . [T x 16] contraction that streams X once: HBM-bound. Each CTA owns 128 columns of X
// and a contiguous slice of the tokens; TMA streams [64 tokens x 64 columns] boxes through an
// 8-stage ring, and tcgen05.mma consumes them AS THEY LIE: X tiles are the MN-major A operand
// (M = columns), the 16-wide w rows are the MN-major B operand. The per-slice results go out as
// partials that llc_lora_colsum_finish adds in a fixed order (bit-deterministic).
// Optional rider (dA_o of the attention block, X = the attention output O): the same O tiles are
// also what delta = rowsum(dO o O) of the attention backward needs, so the kernel can TMA-load the
// matching dO tiles next to them and two warps form delta[token, head] from shared memory - the
// separate delta pass (its own read of O from HBM) disappears.
#include "common.cuh"

namespace {

constexpr int kKB = 64;                      // tokens per k-block
constexpr int kATile = 2 * kKB * 128;        // two 64-column atoms
constexpr int kBTile = kKB * 128;            // w rows, 32 B used per 128 B row
constexpr int kDTile = 2 * kKB * 128;        // dO tile of the delta rider (same shape as A)
constexpr int kStage = kATile + kBTile + kDTile;
constexpr int kStages = 5;
constexpr int kSmem = 1024 + kStages * kStage + 256;
constexpr int kThreads = 6 * 32;             // TMA, MMA, 2 x w loader, 2 x delta (all 4: epilogue)

__global__ void __launch_bounds__(kThreads, 1)
lora_colsum_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD,
                      const __nv_bfloat16* __restrict__ w, int ld_w, int T, int C, int R,
                      int tok_per_split, float* __restrict__ partial, float* __restrict__ delta,
                      int delta_ld) {
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
      mbar_init(smem_u32(&full_bar[s]), 1 + 64);    // TMA issuer + every loader thread
      mbar_init(smem_u32(&empty_bar[s]), delta ? 3 : 1);   // MMA commit (+ the two delta warps)
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
        mbar_expect_tx(fb, delta ? kATile + kDTile : kATile);
        tma_load_2d(sa, &tmX, fb, c0, t_begin + kb * kKB);
        tma_load_2d(sa + kKB * 128, &tmX, fb, c0 + 64, t_begin + kb * kKB);
        if (delta) {
          tma_load_2d(sa + kATile + kBTile, &tmD, fb, c0, t_begin + kb * kKB);
          tma_load_2d(sa + kATile + kBTile + kKB * 128, &tmD, fb, c0 + 64, t_begin + kb * kKB);
        }
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
    if (warp < 4) {
      // w loader: warps 2, 3 stage the 64 token rows of every k-block (thread = row, two 16 B
      // pieces, TMA's 128 B swizzle pattern so the tile reads as an MN-major operand)
      const int row = (warp - 2) * 32 + lane;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        const int t = t_begin + kb * kKB + row;
        // asynchronous 16 B copies (zero-filled past the slice end); their completion arrives on
        // the stage's full barrier, so the loader never waits for a load
        const __nv_bfloat16* src = w + (size_t)(t < t_end ? t : t_begin) * ld_w;
        const uint32_t nbytes = t < t_end ? 16u : 0u;
        const uint32_t dst = smem_u32(smem + stage * kStage + kATile + row * 128);
#pragma unroll
        for (int piece = 0; piece < 2; ++piece)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                           dst + ((piece ^ (row & 7)) << 4)),
                       "l"(src + piece * 8), "r"(nbytes)
                       : "memory");
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                         smem_u32(&full_bar[stage]))
                     : "memory");
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    } else if (delta) {
      // delta rider: warps 4, 5, thread = token row of the k-block, both heads of the column tile:
      // delta[t, head] = sum over the head's 64 columns of O[t, c] * dO[t, c], from the landed tiles
      const int row = (warp - 4) * 32 + lane;
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        const uint8_t* sO = smem + stage * kStage;
        const uint8_t* sD = sO + kATile + kBTile;
        const int t = t_begin + kb * kKB + row;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          float d = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t off = (uint32_t)(a * kKB * 128 + row * 128 + ((c ^ (row & 7)) << 4));
            const uint4 ov = *reinterpret_cast<const uint4*>(sO + off);
            const uint4 dv = *reinterpret_cast<const uint4*>(sD + off);
            const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 x = unpack_bf16(ow[e]), y = unpack_bf16(dw[e]);
              d = fmaf(x.x, y.x, d);
              d = fmaf(x.y, y.y, d);
            }
          }
          if (t < t_end) delta[(size_t)t * delta_ld + (c0 >> 6) + a] = d;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));   // this warp is done reading
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
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


// ------------------------------------------------------------------------------------------------
// Fused pass over X (bf16 [T, C]) for one LoRA projection's backward: both reductions that read X,
//   column sums   P[c, j] = sum_t X[t, c] * w[t, j]          (dB = s g^T u)
//   row products  U[t, j] = sum_c X[t, c] * F[j, c]          (du = s g B, F = s B^T packed [16, C])
// from ONE stream of X through shared memory (they were a colsum launch plus a skinny GEMM, each
// reading X from HBM). A CTA owns 128-token slabs and ALL columns: per slab it walks the C / 128
// column tiles; the [128 tokens x 128 columns] tile (two 64-column TMA boxes) feeds
//   8 MMAs with the tile as MN-major A (M = columns), w rows as MN-major B  -> TMEM cols 16 ct..
//   8 MMAs with the tile as K-major  A (M = tokens),  F block as K-major B  -> U accumulator
// (the factor rows of a tile's columns are loaded with it: 4 KB from L2 next to 32 KB from HBM).
// The column-sum accumulators (C / 128 x 16 TMEM columns, <= 288) live for the whole kernel and
// go out as one partial per CTA; U (2 x 16 columns, double-buffered) leaves after every slab.
constexpr int kFXTile = 2 * 128 * 128;     // 128 tokens x two 64-column atoms
constexpr int kFFTile = 2 * 16 * 128;      // the 16 factor rows of the same 128 columns
constexpr int kFStage = kFXTile + kFFTile;
constexpr int kFWTile = 128 * 128;         // w rows of a slab, 32 B used per 128 B row
constexpr int kFStages = 5;
constexpr uint32_t kFUCol = 288;           // U accumulators behind the column sums

__global__ void __launch_bounds__(kThreads, 1)
lora_fused_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmF,
                     const __nv_bfloat16* __restrict__ w, int ld_w, __nv_bfloat16* __restrict__ U,
                     int ld_u, int T, int C, int R, float* __restrict__ partial, int rev) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int nct = C / 128;
  uint8_t* sW = smem;                                   // w rows of 2 slabs
  uint8_t* sX = sW + 2 * kFWTile;                       // kFStages x (X tile | factor block)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + kFStages * kFStage);
  uint64_t* full_bar = bars;                  // [kFStages] X tile landed
  uint64_t* empty_bar = bars + kFStages;      // [kFStages] its MMAs retired
  uint64_t* w_full = bars + 2 * kFStages;     // [2] w rows of a slab landed (128 loader threads)
  uint64_t* u_done = w_full + 2;              // [2] the slab's MMAs retired (U complete, w free)
  uint64_t* u_free = u_done + 2;              // [2] U accumulator drained (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_free + 2);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int nslabs = (T + 127) / 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmF);
    for (int s = 0; s < kFStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&w_full[b]), 128);
      mbar_init(smem_u32(&u_done[b]), 1);
      mbar_init(smem_u32(&u_free[b]), 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int slab = blockIdx.x; slab < nslabs; slab += gridDim.x) {
      for (int ct = 0; ct < nct; ++ct) {
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&full_bar[stage]);
          const uint32_t sa = smem_u32(sX + stage * kFStage);
          mbar_expect_tx(fb, kFStage);
          // rev: walk the slabs downwards (llc_set_traversal bit 2)
          const int srow = (rev ? nslabs - 1 - slab : slab) * 128;
          tma_load_2d(sa, &tmX, fb, ct * 128, srow);
          tma_load_2d(sa + 128 * 128, &tmX, fb, ct * 128 + 64, srow);
          // the factor rows of these columns ride along (72 KB in all: L2-resident)
          tma_load_2d(sa + kFXTile, &tmF, fb, ct * 128, 0);
          tma_load_2d(sa + kFXTile + 16 * 128, &tmF, fb, ct * 128 + 64, 0);
        }
        __syncwarp();
        if (++stage == kFStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_c = umma_idesc_bf16(128, 16, 1, 1);   // tile^T . w : A, B MN-major
    const uint32_t idesc_r = umma_idesc_bf16(128, 16, 0, 0);   // tile . F^T : A, B K-major
    int stage = 0, k = 0;
    uint32_t phase = 0;
    for (int slab = blockIdx.x; slab < nslabs; slab += gridDim.x, ++k) {
      const int wb = k & 1;
      const uint32_t wph = (k >> 1) & 1;
      mbar_wait(smem_u32(&w_full[wb]), wph);
      mbar_wait(smem_u32(&u_free[wb]), wph ^ 1);
      fence_proxy_async_smem();   // the w rows were written through the generic proxy (cp.async)
      const uint32_t swb = smem_u32(sW + wb * kFWTile);
      for (int ct = 0; ct < nct; ++ct) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(sX + stage * kFStage);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)     // K = the slab's tokens
            umma_bf16(tmem_base + ct * 16, umma_desc_mn_sw128(sa + ks * 2048, 128 * 128, 1024),
                      umma_desc_mn_sw128(swb + ks * 2048, 8192, 1024), idesc_c, (k | ks) != 0);
#pragma unroll
          for (int a = 0; a < 2; ++a)        // K = the tile's columns
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16(tmem_base + kFUCol + wb * 16, umma_desc_k_sw128(sa + a * 128 * 128) + 2 * k4,
                        umma_desc_k_sw128(sa + kFXTile + a * 2048) + 2 * k4, idesc_r,
                        (ct | a | k4) != 0);
          umma_commit(smem_u32(&empty_bar[stage]));
          if (ct == nct - 1) umma_commit(smem_u32(&u_done[wb]));
        }
        __syncwarp();
        if (++stage == kFStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;            // token within the slab / column within a tile
    const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16);
    // this thread's w row of a slab: two 16 B pieces, TMA's 128 B swizzle (MN-major operand)
    auto load_w = [&](int slab, int buf) {
      const int t = (rev ? nslabs - 1 - slab : slab) * 128 + row;
      const __nv_bfloat16* src = w + (size_t)(t < T ? t : 0) * ld_w;
      const uint32_t nbytes = t < T ? 16u : 0u;
      const uint32_t dst = smem_u32(sW + buf * kFWTile + row * 128);
#pragma unroll
      for (int piece = 0; piece < 2; ++piece)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                         dst + ((piece ^ (row & 7)) << 4)),
                     "l"(src + piece * 8), "r"(nbytes)
                     : "memory");
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                       smem_u32(&w_full[buf]))
                   : "memory");
    };
    load_w(blockIdx.x, 0);
    if (blockIdx.x + (int)gridDim.x < nslabs) load_w(blockIdx.x + gridDim.x, 1);
    int k = 0;
    for (int slab = blockIdx.x; slab < nslabs; slab += gridDim.x, ++k) {
      const int wb = k & 1;
      mbar_wait(smem_u32(&u_done[wb]), (k >> 1) & 1);
      tc_fence_after();
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
            "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(tb + kFUCol + wb * 16)
          : "memory");
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&u_free[wb]));
      const int t = (rev ? nslabs - 1 - slab : slab) * 128 + row;
      if (t < T) {
        uint4* dst = reinterpret_cast<uint4*>(U + (size_t)t * ld_u);
        dst[0] = make_uint4(pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1])),
                            pack_bf16(__uint_as_float(v[2]), __uint_as_float(v[3])),
                            pack_bf16(__uint_as_float(v[4]), __uint_as_float(v[5])),
                            pack_bf16(__uint_as_float(v[6]), __uint_as_float(v[7])));
        dst[1] = make_uint4(pack_bf16(__uint_as_float(v[8]), __uint_as_float(v[9])),
                            pack_bf16(__uint_as_float(v[10]), __uint_as_float(v[11])),
                            pack_bf16(__uint_as_float(v[12]), __uint_as_float(v[13])),
                            pack_bf16(__uint_as_float(v[14]), __uint_as_float(v[15])));
      }
      // every MMA that read this slab's w rows has retired: the buffer takes the slab after next
      const int next = slab + 2 * (int)gridDim.x;
      if (next < nslabs) load_w(next, wb);
    }
    // the last u_done covered every MMA of the CTA: the column sums are final
    float* out = partial + (size_t)blockIdx.x * C * R;
    for (int ct = 0; ct < nct; ++ct) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
            "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(tb + ct * 16)
          : "memory");
      tmem_ld_wait();
      for (int j = 0; j < R; ++j) out[(size_t)(ct * 128 + row) * R + j] = __uint_as_float(v[j]);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

bool llc_lora_fused_eligible(const void* X, int ld_x, int T, int C, int R, const void* w, int ld_w,
                             const void* F, int ld_f, const void* U, int ld_u) {
  return C % 128 == 0 && C / 128 <= 18 && T >= 1024 && R >= 1 && R <= 16 && ld_x % 8 == 0 &&
         ld_w % 8 == 0 && ld_f % 8 == 0 && ld_u % 8 == 0 &&
         (((uintptr_t)X | (uintptr_t)w | (uintptr_t)F | (uintptr_t)U) & 15) == 0;
}

// partial: [n_partials][C][R] floats, one slice per CTA (n_partials <= number of SMs)
int llc_lora_fused_tc(const void* X, int ld_x, int T, int C, int R, const void* w, int ld_w,
                      const void* F, int ld_f, void* U, int ld_u, float* partial, int* n_partials,
                      cudaStream_t st) {
  CUtensorMap tm;
  if (int rc = llc_encode_tmap_2d(&tm, X, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)C,
                                  (uint64_t)T, (uint64_t)ld_x * 2, 64, 128,
                                  CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  CUtensorMap tf;
  if (int rc = llc_encode_tmap_2d(&tf, F, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)C, 16,
                                  (uint64_t)ld_f * 2, 64, 16, CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  const int nslabs = (T + 127) / 128;
  const int grid = nslabs < llc_num_sms() ? nslabs : llc_num_sms();
  const int smem = 1024 + 2 * kFWTile + kFStages * kFStage + 256;
  LLC_CONFIGURE_SMEM(lora_fused_tc_kernel, smem);
  LLC_PROF_BEGIN(LLC_K_LORA_SIDE, T, C, 3, 4.0 * T * C * 16, 2.0 * T * C, st);
  LLC_CUDA(llc_launch_pdl(lora_fused_tc_kernel, dim3(grid), dim3(kThreads), (size_t)smem, st, tm, tf,
                          reinterpret_cast<const __nv_bfloat16*>(w), ld_w,
                          reinterpret_cast<__nv_bfloat16*>(U), ld_u, T, C, R, partial,
                          (g_llc_traversal >> 2) & 1));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("lora_fused_tc_kernel");
  *n_partials = grid;
  return 0;
}

bool llc_colsum_tc_eligible(const void* X, int ld_x, int T, int C, const void* w, int ld_w) {
  return C % 128 == 0 && T >= 1024 && ld_x % 8 == 0 && ld_w % 8 == 0 &&
         ((uintptr_t)X & 15) == 0 && ((uintptr_t)w & 15) == 0;
}

// d_o / delta: optional rider (see the header comment): delta[t * delta_ld + c / 64] =
// sum over head c / 64 of X[t, c] * d_o[t, c]; requires C % 128 == 0 (it always is here)
int llc_colsum_tc_delta(const void* X, int ld_x, int T, int C, int R, const void* w, int ld_w,
                        float* partial, int* n_partials, const void* d_o, int ld_do, float* delta,
                        int delta_ld, cudaStream_t st) {
  CUtensorMap tm, td;
  if (int rc = llc_encode_tmap_2d(&tm, X, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)C,
                                  (uint64_t)T, (uint64_t)ld_x * 2, 64, kKB,
                                  CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  td = tm;
  if (delta != nullptr)
    if (int rc = llc_encode_tmap_2d(&td, d_o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)C,
                                    (uint64_t)T, (uint64_t)ld_do * 2, 64, kKB,
                                    CU_TENSOR_MAP_SWIZZLE_128B))
      return rc;
  const int mtiles = C / 128;
  int splits = llc_num_sms() / mtiles;
  if (splits < 1) splits = 1;
  int per = (T + splits - 1) / splits;
  per = (per + kKB - 1) / kKB * kKB;          // k-blocks never straddle two slices
  splits = (T + per - 1) / per;
  LLC_CONFIGURE_SMEM(lora_colsum_tc_kernel, kSmem);
  LLC_PROF_BEGIN(LLC_K_LORA_SIDE, T, C, delta ? 4 : 2, 2.0 * T * C * 16,
                 (delta ? 4.0 : 2.0) * T * C, st);
  LLC_CUDA(llc_launch_pdl(lora_colsum_tc_kernel, dim3(mtiles, splits), dim3(kThreads), kSmem, st, tm,
                          td, reinterpret_cast<const __nv_bfloat16*>(w), ld_w, T, C, R, per, partial,
                          delta, delta_ld));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("lora_colsum_tc_kernel");
  *n_partials = splits;
  return 0;
}

int llc_colsum_tc(const void* X, int ld_x, int T, int C, int R, const void* w, int ld_w,
                  float* partial, int* n_partials, cudaStream_t st) {
  return llc_colsum_tc_delta(X, ld_x, T, C, R, w, ld_w, partial, n_partials, nullptr, 0, nullptr, 0,
                             st);
}

extern "C" int llc_lora_side_fused(const void* X, int ld_x, int T, int C, int r, const void* w,
                                   int ld_w, const void* F, int ld_f, void* U, int ld_u,
                                   float* partial, int* n_partials, void* stream) {
  LLC_REQUIRE(X && w && F && U && partial && n_partials, "llc_lora_side_fused: null pointer");
  LLC_REQUIRE(llc_lora_fused_eligible(X, ld_x, T, C, r, w, ld_w, F, ld_f, U, ld_u),
              "llc_lora_side_fused: unsupported shape/alignment (T=%d C=%d r=%d)", T, C, r);
  // partial slices use the padded rank llc_lora_colsum_finish expects (4 or 8 columns)
  return llc_lora_fused_tc(X, ld_x, T, C, r <= 4 ? 4 : 8, w, ld_w, F, ld_f, U, ld_u, partial,
                           n_partials, (cudaStream_t)stream);
}
