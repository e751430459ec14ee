// tcgen05/TMEM attention core for sm_100a (sequence length <= 256; the mma.sync kernels in
// attention.cu remain for longer sequences).
// Reference math: models/clip/lora.py:950 (q * hd^-0.5), :1002-1006 (head n*H+h), :1043 bmm(q,k^T),
// :1063 softmax, :1068 bmm(P,v), :1070-1071 merge heads. P [N*H, L, L] is never materialised.
//
// One CTA walks (sample, head) pairs. Per pair, with LK = L rounded up to 32 and M-tiles of 128
// query rows:
//   S_t  = Q_t K^T        tcgen05.mma SS, M=128, N=LK, K=64      -> TMEM region t, fp32
//   P_t  = exp2(..)       one thread per query row straight out of TMEM (no shuffles), written
//                         back to TMEM IN PLACE as packed bf16 pairs (tcgen05.st)
//   O_t  = P_t V          tcgen05.mma with A from TMEM, B = V as loaded ([key][hd], MN-major)
// Warp roles: 0 = TMA producer (Q,K,V tiles of the next pair land while this one computes),
// 1 = MMA issuer, 2..5 / 6..9 = softmax + epilogue for M-tile 0 / 1.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int kThreads = 320;
constexpr int kTileBytes = 128 * 128;          // 128 rows x 64 bf16
constexpr int kMatBytes = 2 * kTileBytes;      // up to 256 rows
constexpr int kStageBytes = 3 * kMatBytes;     // Q | K | V
constexpr int kFwdSmem = 1024 + 2 * kStageBytes + 256;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,"
      "%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct FwdParams {
  __nv_bfloat16* o;
  int ld_o;
  float* lse;
  int N, L, H, LK, NT, sn, sl, causal, dbg;
};

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStageBytes);
  uint64_t* kv_full = bars;        // [2] TMA landed Q,K,V of a pair
  uint64_t* kv_empty = bars + 2;   // [2] all MMAs reading the stage retired
  uint64_t* s_full = bars + 4;     // [2] per M-tile: S complete
  uint64_t* p_ready = bars + 6;    // [2] per M-tile: P written to TMEM (128 arrivals)
  uint64_t* o_full = bars + 8;     // [2] per M-tile: O complete
  uint64_t* s_free = bars + 10;    // [2] per M-tile: O drained, region reusable (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int pairs = p.N * p.H;
  const int D = p.H * HD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&kv_full[i]), 1);
      mbar_init(smem_u32(&kv_empty[i]), 1);
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_ready[i]), 128);
      mbar_init(smem_u32(&o_full[i]), 1);
      mbar_init(smem_u32(&s_free[i]), 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int it = 0;
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int n = pr / p.H, h = pr % p.H;
      mbar_wait(smem_u32(&kv_empty[st]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&kv_full[st]);
        mbar_expect_tx(fb, 3 * p.NT * kTileBytes);
        const uint32_t base = smem_u32(smem + st * kStageBytes);
        for (int m = 0; m < 3; ++m)        // Q, K, V: column blocks h*64 + {0, D, 2D}
          for (int t = 0; t < p.NT; ++t)
            tma_load_3d(base + m * kMatBytes + t * kTileBytes, &tmQKV, fb, m * D + h * HD, t * 128,
                        n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_s = umma_idesc_bf16(128, p.LK, 0, 0);
    const uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);  // B = V is MN-major
    const int ksteps_o = p.LK / 16;
    int it = 0;
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1, tp = it & 1;
      const uint32_t base = smem_u32(smem + st * kStageBytes);
      mbar_wait(smem_u32(&kv_full[st]), ph);
      tc_fence_after();
      for (int t = 0; t < p.NT; ++t) {
        mbar_wait(smem_u32(&s_free[t]), tp ^ 1);   // previous pair's O_t drained
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = umma_desc_k_sw128(base + t * kTileBytes);
          const uint64_t bdesc = umma_desc_k_sw128(base + kMatBytes);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16(tmem_base + t * 256, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
          umma_commit(smem_u32(&s_full[t]));
        }
        __syncwarp();
      }
      for (int t = 0; t < p.NT; ++t) {
        mbar_wait(smem_u32(&p_ready[t]), tp);
        tc_fence_after();
        if (elect_one()) {
          for (int ks = 0; ks < ksteps_o; ++ks) {
            // V rows [16 ks, 16 ks + 16): two 8-row groups of 1024 B
            const uint64_t bdesc = umma_desc_mn_sw128(base + 2 * kMatBytes + ks * 2048, 8192, 1024);
            umma_bf16_ts(tmem_base + t * 256 + 128, tmem_base + t * 256 + ks * 8, bdesc, idesc_o,
                         ks != 0);
          }
          umma_commit(smem_u32(&o_full[t]));
          if (t == p.NT - 1) umma_commit(smem_u32(&kv_empty[st]));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue
    const int t = (warp - 2) >> 2;   // M-tile of this warp group
    const int q = warp & 3;          // TMEM lane quarter
    if (t < p.NT) {
      const int row = t * 128 + q * 32 + lane;   // query index within the pair
      const uint32_t treg = tmem_base + t * 256 + ((uint32_t)(q * 32) << 16);
      const float c2 = 0.125f * kLog2e;          // hd^-0.5 = 1/8 for hd = 64
      const int nch = p.LK / 32;
      int it = 0;
      for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
        const uint32_t tp = it & 1;
        const int n = pr / p.H, h = pr % p.H;
        const int kmax = p.causal ? min(p.L, row + 1) : p.L;   // keys [0, kmax) are visible
        mbar_wait_relaxed(smem_u32(&s_full[t]), tp);
        tc_fence_after();
        // pass 1: row maximum. Chunks that lie entirely below kmax take the predicate-free
        // path; the TMEM load of chunk c+1 is in flight while chunk c is reduced.
        float m = -INFINITY;
        {
          uint32_t va[32], vb[32];
          tmem_ld_32x32(treg, va);
          for (int c = 0; c < nch; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nch) tmem_ld_32x32(treg + (c + 1) * 32, vb);
            if (c * 32 + 32 <= kmax) {
#pragma unroll
              for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(va[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c * 32 + j < kmax) m = fmaxf(m, __uint_as_float(va[j]));
            }
            if (c + 1 < nch) {
              tmem_ld_wait();
              if (c + 2 < nch) tmem_ld_32x32(treg + (c + 2) * 32, va);
              if (c * 32 + 64 <= kmax) {
#pragma unroll
                for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(vb[j]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (c * 32 + 32 + j < kmax) m = fmaxf(m, __uint_as_float(vb[j]));
              }
            }
          }
        }
        const float mc = (m == -INFINITY) ? 0.f : m * c2;
        // pass 2: p = exp2(s c2 - m c2), row sum, packed bf16 pairs back into the S columns
        float l = 0.f;
        {
          uint32_t va[32], vb[32], w[16];
          auto chunk = [&](const uint32_t (&v)[32], int c) {
            if (c * 32 + 32 <= kmax) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float p0 = ex2(fmaf(__uint_as_float(v[j]), c2, -mc));
                const float p1 = ex2(fmaf(__uint_as_float(v[j + 1]), c2, -mc));
                l += p0 + p1;
                w[j >> 1] = pack_bf16(p0, p1);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float p0 =
                    (c * 32 + j < kmax) ? ex2(fmaf(__uint_as_float(v[j]), c2, -mc)) : 0.f;
                const float p1 =
                    (c * 32 + j + 1 < kmax) ? ex2(fmaf(__uint_as_float(v[j + 1]), c2, -mc)) : 0.f;
                l += p0 + p1;
                w[j >> 1] = pack_bf16(p0, p1);
              }
            }
            tmem_st_x16(treg + c * 16, w);
          };
          tmem_ld_32x32(treg, va);
          for (int c = 0; c < nch; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nch) tmem_ld_32x32(treg + (c + 1) * 32, vb);
            chunk(va, c);
            if (c + 1 < nch) {
              tmem_ld_wait();
              if (c + 2 < nch) tmem_ld_32x32(treg + (c + 2) * 32, va);
              chunk(vb, c + 1);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&p_ready[t]));
        if (p.lse != nullptr && row < p.L)
          p.lse[(size_t)pr * p.L + row] = m * 0.125f + logf(l);
        const float inv = 1.0f / l;
        // epilogue: O row (64 fp32) -> bf16 -> one 128 B row segment per thread
        mbar_wait_relaxed(smem_u32(&o_full[t]), tp);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(treg + 128, o0);
        tmem_ld_32x32(treg + 160, o1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(&s_free[t]));
        if (row < p.L && !(p.dbg & 2)) {
          uint4* dst = reinterpret_cast<uint4*>(p.o + (size_t)(n * p.sn + row * p.sl) * p.ld_o +
                                                h * HD);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            dst[j] = make_uint4(
                pack_bf16(__uint_as_float(o0[8 * j]) * inv, __uint_as_float(o0[8 * j + 1]) * inv),
                pack_bf16(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv),
                pack_bf16(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv),
                pack_bf16(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv));
            dst[4 + j] = make_uint4(
                pack_bf16(__uint_as_float(o1[8 * j]) * inv, __uint_as_float(o1[8 * j + 1]) * inv),
                pack_bf16(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv),
                pack_bf16(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv),
                pack_bf16(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv));
          }
        }
      }
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


// ------------------------------------------------------------------------------------------------
// Backward. With P = softmax(S), S = q k^T / 8, delta_q = sum_c dO[q,c] O[q,c]:
//   dP = dO V^T, dS = P o (dP - delta), dQ = dS K / 8, dK = dS^T Q / 8, dV = P^T dO.
// Per (sample, head) the CTA runs 2 NT tile-phases of 128 accumulator rows, all on tcgen05:
//   type A (query rows):  R0 = Q_t K^T, R1 = dO_t V^T  -> dS (bf16, into R1)   -> dQ_t = dS K
//   type B (key rows):    R0 = K_t Q^T, R1 = V_t dO^T  -> P^T (R0), dS^T (R1)  -> dV_t = P^T dO,
//                                                                                 dK_t = dS^T Q
// R0/R1 = TMEM columns [0,256) / [256,512); LK = L rounded up to 16 columns are used. 256 threads
// work on a phase: thread = (row, column half; half 0 = columns [0, c0), half 1 = [c0, LK)). Each
// half writes its packed bf16 operands over columns of ITS OWN half that it has already consumed
// (words at [0, c0/2) and [c0, c0 + (LK-c0)/2)), so no thread overwrites data another thread
// still has to read; the MMA walks the two areas k-step by k-step. The second operands (K, dO, Q
// as "[k][hd]" MN-major B) are the tiles exactly as TMA loaded them. Q, K, V, dO of the next pair
// land in a second smem stage while this pair computes (two stages fit up to LK = 208); outputs
// leave through a 16 KB staging tile and TMA stores (rows >= L are clipped by the tensor map).
struct BwdParams {
  const __nv_bfloat16* o;
  int ld_o;
  const float* lse;
  int N, L, H, LK, NT, c0, sn, sl, causal, nstages, mat_bytes, dbg;
};
constexpr int kStagingBytes = 128 * 128;

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1,
                                             int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(m),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 16 B chunk c16 of row r in a [rows x 128 B] tile laid out in TMA's 128 B swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int c16) {
  return (uint32_t)(row * 128 + ((c16 ^ (row & 7)) << 4));
}

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ0, const __grid_constant__ CUtensorMap tmQ1,
                   const __grid_constant__ CUtensorMap tmD0, const __grid_constant__ CUtensorMap tmD1,
                   const __grid_constant__ CUtensorMap tmOut, BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int LK = p.LK, NT = p.NT, c0 = p.c0, mat = p.mat_bytes;
  const int stage_bytes = 4 * mat;
  uint8_t* staging = smem + p.nstages * stage_bytes;
  float* sLse = reinterpret_cast<float*>(staging + kStagingBytes);   // [256] lse * log2e
  float* sDelta = sLse + 256;                                        // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  uint64_t* ld_full = bars;       // [2] TMA landed Q,K,V,dO of a pair
  uint64_t* ld_empty = bars + 2;  // [2] every MMA of the pair retired
  uint64_t* s_full = bars + 4;    // R0/R1 hold the phase's S-type products
  uint64_t* p_ready = bars + 5;   // operands written back to TMEM (256 arrivals)
  uint64_t* o_full = bars + 6;    // output accumulators complete
  uint64_t* acc_free = bars + 7;  // accumulators drained (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int pairs = p.N * p.H;
  const int D = p.H * HD;
  // dV / dK accumulator columns inside R0 / R1: past both operand areas
  const int out_b = (c0 + (LK - c0) / 2 + 31) & ~31;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();   // swizzled tiles need the 1024 B alignment asked for
    tma_prefetch_desc(&tmQ0); tma_prefetch_desc(&tmQ1);
    tma_prefetch_desc(&tmD0); tma_prefetch_desc(&tmD1);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&ld_full[i]), 1);
      mbar_init(smem_u32(&ld_empty[i]), 1);
    }
    mbar_init(smem_u32(s_full), 1);
    mbar_init(smem_u32(p_ready), 256);
    mbar_init(smem_u32(o_full), 1);
    mbar_init(smem_u32(acc_free), 256);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t smem0 = smem_u32(smem);
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int it = 0;
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const int st = it % p.nstages;
      const uint32_t ph = (it / p.nstages) & 1;
      const int n = pr / p.H, h = pr % p.H;
      mbar_wait(smem_u32(&ld_empty[st]), ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = smem_u32(&ld_full[st]);
        const uint32_t base = smem0 + st * stage_bytes;
        mbar_expect_tx(fb, 4 * mat);
        for (int m = 0; m < 3; ++m) {   // Q, K, V: column blocks h*64 + {0, D, 2D}
          tma_load_3d(base + m * mat, &tmQ0, fb, m * D + h * HD, 0, n);
          if (NT > 1) tma_load_3d(base + m * mat + kTileBytes, &tmQ1, fb, m * D + h * HD, 128, n);
        }
        tma_load_3d(base + 3 * mat, &tmD0, fb, h * HD, 0, n);
        if (NT > 1) tma_load_3d(base + 3 * mat + kTileBytes, &tmD1, fb, h * HD, 128, n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_s = umma_idesc_bf16(128, LK, 0, 0);
    const uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);
    const uint32_t R0 = tmem_base, R1 = tmem_base + 256;
    const int ksteps = LK / 16, kh = c0 / 16;
    int it = 0, g = 0;
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const int st = it % p.nstages;
      const uint32_t sQ = smem0 + st * stage_bytes, sK = sQ + mat, sV = sQ + 2 * mat,
                     sD = sQ + 3 * mat;
      mbar_wait(smem_u32(&ld_full[st]), (it / p.nstages) & 1);
      tc_fence_after();
      for (int ph = 0; ph < 2 * NT; ++ph, ++g) {
        const bool type_a = ph < NT;
        const int t = type_a ? ph : ph - NT;
        mbar_wait(smem_u32(acc_free), (g & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a0 = umma_desc_k_sw128((type_a ? sQ : sK) + t * kTileBytes);
          const uint64_t b0 = umma_desc_k_sw128(type_a ? sK : sQ);
          const uint64_t a1 = umma_desc_k_sw128((type_a ? sD : sV) + t * kTileBytes);
          const uint64_t b1 = umma_desc_k_sw128(type_a ? sV : sD);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(R0, a0 + 2 * k, b0 + 2 * k, idesc_s, k != 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(R1, a1 + 2 * k, b1 + 2 * k, idesc_s, k != 0);
          umma_commit(smem_u32(s_full));
        }
        __syncwarp();
        mbar_wait(smem_u32(p_ready), g & 1);
        tc_fence_after();
        if (elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t aoff = ks < kh ? ks * 8 : c0 + (ks - kh) * 8;
            if (type_a) {   // dQ_t = dS K
              umma_bf16_ts(R0, R1 + aoff, umma_desc_mn_sw128(sK + ks * 2048, 8192, 1024), idesc_o,
                           ks != 0);
            } else {        // dV_t = P^T dO ; dK_t = dS^T Q
              umma_bf16_ts(R0 + out_b, R0 + aoff, umma_desc_mn_sw128(sD + ks * 2048, 8192, 1024),
                           idesc_o, ks != 0);
              umma_bf16_ts(R1 + out_b, R1 + aoff, umma_desc_mn_sw128(sQ + ks * 2048, 8192, 1024),
                           idesc_o, ks != 0);
            }
          }
          umma_commit(smem_u32(o_full));
          if (ph == 2 * NT - 1) umma_commit(smem_u32(&ld_empty[st]));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ element-wise + epilogue
    const int q = warp & 3;             // TMEM lane quarter
    const int hh = (warp - 2) >> 2;     // column half
    const int r = q * 32 + lane;        // accumulator row within a phase
    const int tid2 = (warp - 2) * 32 + lane;   // 0..255
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t R0 = tmem_base + lane_sel, R1 = tmem_base + 256 + lane_sel;
    const float c2 = 0.125f * kLog2e;
    const int col_base = hh ? c0 : 0;
    const int nchunk = (hh ? LK - c0 : c0) / 16;
    int it = 0, g = 0;
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const int st = it % p.nstages;
      const int n = pr / p.H, h = pr % p.H;
      const int tok0 = n * p.sn;
      const uint8_t* sDO = smem + st * stage_bytes + 3 * mat;
      mbar_wait(smem_u32(&ld_full[st]), (it / p.nstages) & 1);
      // delta_q = dO_q . O_q for the 32 queries of this warp: 8 lanes per row, 16 B each, the O
      // rows read straight from global (coalesced), the dO rows from the landed tile
      {
        const int rb = (warp - 2) * 32;
        uint4 ov[8];
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
          const int row = rb + ps * 4 + (lane >> 3);
          ov[ps] = make_uint4(0, 0, 0, 0);
          if (row < p.L)
            ov[ps] = *reinterpret_cast<const uint4*>(
                p.o + (size_t)(tok0 + row * p.sl) * p.ld_o + h * HD + (lane & 7) * 8);
        }
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
          const int row = rb + ps * 4 + (lane >> 3);
          float d = 0.f;
          if (row < p.L) {
            const uint4 dv = *reinterpret_cast<const uint4*>(sDO + sw128_off(row, lane & 7));
            const uint32_t dw[4] = {dv.x, dv.y, dv.z, dv.w};
            const uint32_t ow[4] = {ov[ps].x, ov[ps].y, ov[ps].z, ov[ps].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 a = unpack_bf16(dw[e]), b = unpack_bf16(ow[e]);
              d += a.x * b.x + a.y * b.y;
            }
          }
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          if ((lane & 7) == 0) sDelta[row] = d;
        }
        sLse[tid2] = tid2 < p.L ? p.lse[(size_t)pr * p.L + tid2] * kLog2e : 0.f;
      }
      named_bar_sync(1, 256);
      for (int ph = 0; ph < 2 * NT; ++ph, ++g) {
        const bool type_a = ph < NT;
        const int t = type_a ? ph : ph - NT;
        const int gr = t * 128 + r;     // global row: query (A) or key (B) index
        mbar_wait_relaxed(smem_u32(s_full), g & 1);
        tc_fence_after();
        const float lse_r = sLse[gr & 255], del_r = sDelta[gr & 255];
        // Out-of-range rows/columns need no masking in the non-causal case: TMA zero-filled the
        // rows >= L of Q, K, V and dO, so their products vanish in the output MMAs; only the
        // chunk that straddles L is masked (p could overflow there), and causal chunks are
        // masked element by element.
        auto process = [&](const uint32_t (&sv)[16], const uint32_t (&dv)[16], int c) {
          const int col0 = col_base + c * 16;
          uint32_t wp[8], wd[8];
          const bool fast = !p.causal && col0 + 16 <= p.L;
          if (type_a) {
            if (fast) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float p0 = ex2(fmaf(__uint_as_float(sv[j]), c2, -lse_r));
                const float p1 = ex2(fmaf(__uint_as_float(sv[j + 1]), c2, -lse_r));
                wd[j >> 1] = pack_bf16(p0 * (__uint_as_float(dv[j]) - del_r),
                                       p1 * (__uint_as_float(dv[j + 1]) - del_r));
              }
            } else {
              const int kmax = (gr < p.L) ? (p.causal ? min(p.L, gr + 1) : p.L) : 0;
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float ds[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const float pe = (col0 + j + e < kmax)
                                       ? ex2(fmaf(__uint_as_float(sv[j + e]), c2, -lse_r)) : 0.f;
                  ds[e] = pe * (__uint_as_float(dv[j + e]) - del_r);
                }
                wd[j >> 1] = pack_bf16(ds[0], ds[1]);
              }
            }
            tmem_st_x8(R1 + col_base + c * 8, wd);
          } else {
            float lq[16], dq[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 a = *reinterpret_cast<const float4*>(&sLse[col0 + j]);
              const float4 b = *reinterpret_cast<const float4*>(&sDelta[col0 + j]);
              lq[j] = a.x; lq[j + 1] = a.y; lq[j + 2] = a.z; lq[j + 3] = a.w;
              dq[j] = b.x; dq[j + 1] = b.y; dq[j + 2] = b.z; dq[j + 3] = b.w;
            }
            if (fast) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float p0 = ex2(fmaf(__uint_as_float(sv[j]), c2, -lq[j]));
                const float p1 = ex2(fmaf(__uint_as_float(sv[j + 1]), c2, -lq[j + 1]));
                wp[j >> 1] = pack_bf16(p0, p1);
                wd[j >> 1] = pack_bf16(p0 * (__uint_as_float(dv[j]) - dq[j]),
                                       p1 * (__uint_as_float(dv[j + 1]) - dq[j + 1]));
              }
            } else {
              const int qmin = p.causal ? gr : 0;   // visible iff qmin <= query < L
              const bool krow = gr < p.L;
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                float pe[2], ds[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int qi = col0 + j + e;
                  const bool ok = krow && qi < p.L && qi >= qmin;
                  pe[e] = ok ? ex2(fmaf(__uint_as_float(sv[j + e]), c2, -lq[j + e])) : 0.f;
                  ds[e] = pe[e] * (__uint_as_float(dv[j + e]) - dq[j + e]);
                }
                wp[j >> 1] = pack_bf16(pe[0], pe[1]);
                wd[j >> 1] = pack_bf16(ds[0], ds[1]);
              }
            }
            tmem_st_x8(R0 + col_base + c * 8, wp);
            tmem_st_x8(R1 + col_base + c * 8, wd);
          }
        };
        if (!(p.dbg & 1)) {
          // chunk c+1 is in flight from TMEM while chunk c is processed
          uint32_t sA[16], dA[16], sB[16], dB[16];
          tmem_ld_x16(R0 + col_base, sA);
          tmem_ld_x16(R1 + col_base, dA);
          for (int c = 0; c < nchunk; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nchunk) {
              tmem_ld_x16(R0 + col_base + (c + 1) * 16, sB);
              tmem_ld_x16(R1 + col_base + (c + 1) * 16, dB);
            }
            process(sA, dA, c);
            if (c + 1 < nchunk) {
              tmem_ld_wait();
              if (c + 2 < nchunk) {
                tmem_ld_x16(R0 + col_base + (c + 2) * 16, sA);
                tmem_ld_x16(R1 + col_base + (c + 2) * 16, dA);
              }
              process(sB, dB, c + 1);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(p_ready));
        // ---- epilogue of the phase: accumulators -> staging tile -> TMA store
        mbar_wait_relaxed(smem_u32(o_full), g & 1);
        tc_fence_after();
        if (type_a) {
          uint32_t a[32];
          tmem_ld_32x32(R0 + hh * 32, a);   // dQ columns [32 hh, 32 hh + 32)
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(acc_free));
          if (tid2 == 0) tma_store_wait_read<0>();   // staging no longer read by an older store
          named_bar_sync(1, 256);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(staging + sw128_off(r, hh * 4 + j)) = make_uint4(
                pack_bf16(__uint_as_float(a[8 * j]) * 0.125f, __uint_as_float(a[8 * j + 1]) * 0.125f),
                pack_bf16(__uint_as_float(a[8 * j + 2]) * 0.125f, __uint_as_float(a[8 * j + 3]) * 0.125f),
                pack_bf16(__uint_as_float(a[8 * j + 4]) * 0.125f, __uint_as_float(a[8 * j + 5]) * 0.125f),
                pack_bf16(__uint_as_float(a[8 * j + 6]) * 0.125f, __uint_as_float(a[8 * j + 7]) * 0.125f));
          fence_proxy_async_smem();
          named_bar_sync(1, 256);
          if (tid2 == 0 && !(p.dbg & 8)) {
            tma_store_3d(&tmOut, smem_u32(staging), h * HD, t * 128, n);
            tma_store_commit();
          }
        } else {
          // half 0 holds dV (R0), half 1 holds dK (R1, scaled by hd^-0.5); the staging tile is
          // used twice, dV first
          uint32_t a[32], b[32];
          const uint32_t src = (hh == 0 ? R0 : R1) + out_b;
          tmem_ld_32x32(src, a);
          tmem_ld_32x32(src + 32, b);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(acc_free));
          const float sc = hh == 0 ? 1.0f : 0.125f;
          for (int pass = 0; pass < 2; ++pass) {
            if (tid2 == 0) tma_store_wait_read<0>();
            named_bar_sync(1, 256);
            if (hh == pass) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                *reinterpret_cast<uint4*>(staging + sw128_off(r, j)) = make_uint4(
                    pack_bf16(__uint_as_float(a[8 * j]) * sc, __uint_as_float(a[8 * j + 1]) * sc),
                    pack_bf16(__uint_as_float(a[8 * j + 2]) * sc, __uint_as_float(a[8 * j + 3]) * sc),
                    pack_bf16(__uint_as_float(a[8 * j + 4]) * sc, __uint_as_float(a[8 * j + 5]) * sc),
                    pack_bf16(__uint_as_float(a[8 * j + 6]) * sc, __uint_as_float(a[8 * j + 7]) * sc));
                *reinterpret_cast<uint4*>(staging + sw128_off(r, 4 + j)) = make_uint4(
                    pack_bf16(__uint_as_float(b[8 * j]) * sc, __uint_as_float(b[8 * j + 1]) * sc),
                    pack_bf16(__uint_as_float(b[8 * j + 2]) * sc, __uint_as_float(b[8 * j + 3]) * sc),
                    pack_bf16(__uint_as_float(b[8 * j + 4]) * sc, __uint_as_float(b[8 * j + 5]) * sc),
                    pack_bf16(__uint_as_float(b[8 * j + 6]) * sc, __uint_as_float(b[8 * j + 7]) * sc));
              }
              fence_proxy_async_smem();
            }
            named_bar_sync(1, 256);
            if (tid2 == 0 && !(p.dbg & 8)) {
              tma_store_3d(&tmOut, smem_u32(staging), (pass == 0 ? 2 * D : D) + h * HD, t * 128, n);
              tma_store_commit();
            }
          }
        }
      }
      // sLse / sDelta are rewritten for the next pair only after every thread left the last phase
      named_bar_sync(1, 256);
    }
    if (tid2 == 0) tma_store_wait<0>();
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

// 3-D view of a token-major activation buffer: (column, token l of a sample, sample n)
static int encode_tokens_map(CUtensorMap* tm, const void* base, int cols, int ld, int L, int N,
                             int sn, int sl) {
  return llc_encode_tmap_3d(tm, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)cols,
                            (uint64_t)L, (uint64_t)N, (uint64_t)ld * 2 * sl, (uint64_t)ld * 2 * sn,
                            HD, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B);
}

bool llc_attn_tc_eligible(int L) { return L <= 256; }

int llc_attn_fwd_tc(const void* qkv, int ld_qkv, void* o, int ld_o, float* lse, int N, int L, int H,
                    int sn, int sl, int causal, cudaStream_t st) {
  CUtensorMap tm;
  // a sample-major [N, L, .] buffer with ONE sample still needs a non-zero outer stride
  if (int rc = encode_tokens_map(&tm, qkv, 3 * H * HD, ld_qkv, L, N, sn, sl)) return rc;
  FwdParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(o); p.ld_o = ld_o; p.lse = lse;
  p.N = N; p.L = L; p.H = H; p.LK = (L + 31) / 32 * 32; p.NT = (L + 127) / 128;
  p.sn = sn; p.sl = sl; p.causal = causal;
  static const int dbg = getenv("LLC_ATTN_DBG") ? atoi(getenv("LLC_ATTN_DBG")) : 0;
  p.dbg = dbg;
  static bool configured = false;
  if (!configured) {
    LLC_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kFwdSmem));
    configured = true;
  }
  const int grid = N * H < llc_num_sms() ? N * H : llc_num_sms();
  LLC_PROF_BEGIN(LLC_K_ATTN_FWD, N * H, L, 0, 4.0 * N * H * (double)L * L * HD,
                 8.0 * N * H * (double)L * HD, st);
  LLC_CUDA(llc_launch_pdl(attn_fwd_tc_kernel, dim3(grid), dim3(kThreads), kFwdSmem, st, tm, p));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_fwd_tc_kernel");
  return 0;
}

static int encode_tokens_map_rows(CUtensorMap* tm, const void* base, int cols, int ld, int L, int N,
                                  int sn, int sl, int box_rows) {
  return llc_encode_tmap_3d(tm, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)cols,
                            (uint64_t)L, (uint64_t)N, (uint64_t)ld * 2 * sl, (uint64_t)ld * 2 * sn,
                            HD, box_rows, 1, CU_TENSOR_MAP_SWIZZLE_128B);
}

int llc_attn_bwd_tc(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o, int ld_do,
                    const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H, int sn, int sl,
                    int causal, cudaStream_t st) {
  BwdParams p;
  p.o = reinterpret_cast<const __nv_bfloat16*>(o); p.ld_o = ld_o; p.lse = lse;
  p.N = N; p.L = L; p.H = H; p.LK = (L + 15) / 16 * 16; p.NT = (L + 127) / 128;
  p.c0 = (p.LK / 2 + 15) / 16 * 16;
  p.sn = sn; p.sl = sl; p.causal = causal;
  p.mat_bytes = p.LK * 128;
  if (p.mat_bytes < kTileBytes && p.NT == 1) p.mat_bytes = p.LK * 128;  // one box of LK rows
  const int fixed = kStagingBytes + 2 * 256 * 4 + 256;
  p.nstages = (2 * 4 * p.mat_bytes + fixed <= 227 * 1024) ? 2 : 1;
  const int smem = p.nstages * 4 * p.mat_bytes + fixed;
  static const int dbg = getenv("LLC_ATTN_DBG") ? atoi(getenv("LLC_ATTN_DBG")) : 0;
  p.dbg = dbg;
  const int rows0 = p.NT > 1 ? 128 : p.LK, rows1 = p.NT > 1 ? p.LK - 128 : 16;
  CUtensorMap q0, q1, d0, d1, to;
  if (int rc = encode_tokens_map_rows(&q0, qkv, 3 * H * HD, ld_qkv, L, N, sn, sl, rows0)) return rc;
  if (int rc = encode_tokens_map_rows(&q1, qkv, 3 * H * HD, ld_qkv, L, N, sn, sl, rows1)) return rc;
  if (int rc = encode_tokens_map_rows(&d0, d_o, H * HD, ld_do, L, N, sn, sl, rows0)) return rc;
  if (int rc = encode_tokens_map_rows(&d1, d_o, H * HD, ld_do, L, N, sn, sl, rows1)) return rc;
  if (int rc = encode_tokens_map_rows(&to, dqkv, 3 * H * HD, ld_dqkv, L, N, sn, sl, 128)) return rc;
  static int configured = 0;
  if (configured < smem) {
    LLC_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  smem));
    configured = smem;
  }
  const int grid = N * H < llc_num_sms() ? N * H : llc_num_sms();
  LLC_PROF_BEGIN(LLC_K_ATTN_BWD, N * H, L, 0, 8.0 * N * H * (double)L * L * HD,
                 16.0 * N * H * (double)L * HD, st);
  LLC_CUDA(llc_launch_pdl(attn_bwd_tc_kernel, dim3(grid), dim3(kThreads), (size_t)smem, st, q0, q1, d0, d1,
                          to, p));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_bwd_tc_kernel");
  return 0;
}
