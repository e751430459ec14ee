// Attention forward on tcgen05/TMEM (sequence length <= 256, head dim 64), tile-staggered.
// Reference math: models/clip/lora.py:950 (q * hd^-0.5), :1002-1006 (head n*H+h), :1043 bmm(q,k^T),
// :1063 softmax, :1068 bmm(P,v), :1070-1071 merge heads. P [N*H, L, L] is never materialised.
//
// One persistent CTA per SM walks (sample, head) pairs. Per pair and 128-query tile t:
//   S_t  = Q_t K^T        tcgen05.mma SS, M=128, N=LK (L rounded up to 16), K=64 -> TMEM region t
//   P_t  = exp2(..)       one thread per query row straight out of TMEM, written back IN PLACE as
//                         packed bf16 pairs (tcgen05.st)
//   O_t  = P_t V          tcgen05.mma with A from TMEM, B = V as loaded ([key][hd], MN-major)
//   O_t / l -> bf16 staging tile -> TMA store (rows >= L are clipped by the tensor map)
// The exponentials bound the kernel (16 MUFU results per clock and SM): the previous version
// (attention_tc.cu) issued both tiles' MMAs in lock step, so both softmax groups sat in their MUFU
// phase together and then waited together. Here the MMA warp is event driven: it issues whichever
// of {S_t of the next pair, P_t V} has its inputs ready, so the two groups drift apart until one
// group's exp phase overlaps the other's row-max / PV / epilogue phases.
// Warp roles: 0 = TMA producer (Q,K,V of the next pair land while this one computes),
// 1 = MMA issuer, 2..5 / 6..9 = softmax + epilogue for tile 0 / 1.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int kThreads = 320;
constexpr int kTileBytes = 128 * 128;          // 128 rows x 64 bf16
constexpr int kMatBytes = 2 * kTileBytes;      // up to 256 rows
constexpr int kStageBytes = 3 * kMatBytes;     // Q | K | V
constexpr int kSmem = 1024 + 2 * kStageBytes + 2 * kTileBytes + 256;
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
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,"
      "%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
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
// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
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
__device__ __forceinline__ uint32_t sw128_off(int row, int c16) {
  return (uint32_t)(row * 128 + ((c16 ^ (row & 7)) << 4));
}

struct Fwd2Params {
  float* lse;
  int N, L, H, LK, NT, causal, dbg;
  int lse_ld;   // elements between two pairs' lse rows (L, or the full sequence length when this
                // launch covers the first 256 tokens of a longer sequence: attention_long.cu)
};

__global__ void __launch_bounds__(kThreads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                 Fwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* staging = smem + 2 * kStageBytes;     // one O tile per softmax group
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 2 * kTileBytes);
  uint64_t* kv_full = bars;        // [2] TMA landed Q,K,V of a pair
  uint64_t* kv_empty = bars + 2;   // [2] all MMAs reading the stage retired
  uint64_t* s_full = bars + 4;     // [2] per tile: S complete
  uint64_t* p_ready = bars + 6;    // [2] per tile: P written to TMEM (4 warps)
  uint64_t* o_full = bars + 8;     // [2] per tile: O complete
  uint64_t* s_free = bars + 10;    // [2] per tile: O drained, region reusable (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  uint32_t* trace = reinterpret_cast<uint32_t*>(bars + 14);   // [24] debug timeline (dbg & 128)
#define TRF(cond, i)                                                                        \
  do {                                                                                      \
    if ((p.dbg & 128) && blockIdx.x == 0 && (cond) && lane == 0) trace[i] = (uint32_t)clock64(); \
  } while (0)

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int pairs = p.N * p.H;
  const int D = p.H * HD;
  const int n_it = blockIdx.x < pairs ? (pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&kv_full[i]), 1);
      mbar_init(smem_u32(&kv_empty[i]), 1);
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_ready[i]), 4);
      mbar_init(smem_u32(&o_full[i]), 1);
      mbar_init(smem_u32(&s_free[i]), 4);
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
    for (int it = 0; it < n_it; ++it) {
      const int pr = blockIdx.x + it * gridDim.x;
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int n = pr / p.H, h = pr % p.H;
      if ((p.dbg & 4096) && it == n_it - 1 && elect_one()) pdl_trigger();   // last pair
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
    // ------------------------------------------------------------------ MMA issuer (event driven)
    const uint32_t idesc_s = umma_idesc_bf16(128, p.LK, 0, 0);
    const uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);  // B = V is MN-major
    const int ksteps_o = p.LK / 16;
    int n_s[2] = {0, 0};     // pairs whose S_t has been issued
    int n_o[2] = {0, 0};     // pairs whose P_t V has been issued
    if (p.NT == 1) n_s[1] = n_o[1] = n_it;
    while (n_o[0] < n_it || n_o[1] < n_it) {
      bool progress = false;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (n_s[t] == n_o[t]) {
          // next for this tile: S_t of pair n_s[t] (needs the pair's operands and a drained region)
          const int i = n_s[t];
          // (the votes make the outcome provably warp-uniform: otherwise every MMA operand below
          // goes through a register -> uniform-register waterfall loop)
          if (i < n_it &&
              __all_sync(0xffffffffu, mbar_test(smem_u32(&kv_full[i & 1]), (i >> 1) & 1) &&
                                          mbar_test(smem_u32(&s_free[t]), (i & 1) ^ 1))) {
            tc_fence_after();
            if (elect_one()) {
              const uint32_t base = smem_u32(smem + (i & 1) * kStageBytes);
              const uint64_t adesc = umma_desc_k_sw128(base + t * kTileBytes);
              const uint64_t bdesc = umma_desc_k_sw128(base + kMatBytes);
#pragma unroll
              for (int k = 0; k < HD / 16; ++k)
                umma_bf16(tmem_base + t * 256, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
              umma_commit(smem_u32(&s_full[t]));
            }
            __syncwarp();
            TRF(i == 4 || i == 5, (i - 4) * 12 + t);
            ++n_s[t];
            progress = true;
          }
        } else {
          // next for this tile: O_t = P_t V of pair n_o[t]
          const int i = n_o[t];
          if (__all_sync(0xffffffffu, mbar_test(smem_u32(&p_ready[t]), i & 1))) {
            tc_fence_after();
            if (elect_one()) {
              const uint32_t base = smem_u32(smem + (i & 1) * kStageBytes);
              for (int ks = 0; ks < ksteps_o; ++ks) {
                // V rows [16 ks, 16 ks + 16): two 8-row groups of 1024 B
                const uint64_t bdesc =
                    umma_desc_mn_sw128(base + 2 * kMatBytes + ks * 2048, 8192, 1024);
                umma_bf16_ts(tmem_base + t * 256 + 128, tmem_base + t * 256 + ks * 8, bdesc, idesc_o,
                             ks != 0);
              }
              umma_commit(smem_u32(&o_full[t]));
              // the stage is free once BOTH tiles' MMAs of the pair have retired: the tile that
              // issues its P V last commits (a commit covers every earlier MMA of this thread)
              if (n_o[t ^ 1] > i) umma_commit(smem_u32(&kv_empty[i & 1]));
            }
            __syncwarp();
            TRF(i == 4 || i == 5, (i - 4) * 12 + 2 + t);
            ++n_o[t];
            progress = true;
          }
        }
      }
      if (!progress) __nanosleep(100);   // leave the issue slots to the softmax warps on this SMSP
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue
    const int t = (warp - 2) >> 2;   // tile of this warp group
    const int q = warp & 3;          // TMEM lane quarter
    if (t < p.NT) {
      const int r = q * 32 + lane;               // row within the tile
      const int row = t * 128 + r;               // query index within the pair
      const int gtid = (warp - 2 - 4 * t) * 32 + lane;
      const bool live = t * 128 + q * 32 < p.L;  // warp-uniform: some row of this warp exists
      const uint32_t treg = tmem_base + t * 256 + ((uint32_t)(q * 32) << 16);
      uint8_t* stg = staging + t * kTileBytes;
      const float c2 = 0.125f * kLog2e;          // hd^-0.5 = 1/8 for hd = 64
      const int nch = p.LK / 32;                 // full 32-column chunks
      const bool tail = (p.LK & 16) != 0;        // ... plus one 16-column chunk
      for (int it = 0; it < n_it; ++it) {
        const int pr = blockIdx.x + it * gridDim.x;
        const uint32_t tp = it & 1;
        const int n = pr / p.H, h = pr % p.H;
        // visible keys. Rows >= L of a live warp (zero Q rows, never stored) keep kmax = L so the
        // warp does not diverge between the predicate-free and the masked chunk path
        const int kmax = p.causal ? min(p.L, row + 1) : p.L;
        mbar_wait_relaxed(smem_u32(&s_full[t]), tp);
        tc_fence_after();
        TRF((it == 4 || it == 5) && q == 0, (it - 4) * 12 + 4 + t);
        float m = -INFINITY, l = 0.f;
        if (live) {
          // pass 1: row maximum (four independent running maxima; the TMEM load of chunk c+1 is
          // in flight while chunk c is reduced)
          {
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
            uint32_t va[32], vb[32];
            tmem_ld_32x32(treg, va);
            for (int c = 0; c < nch; c += 2) {
              tmem_ld_wait();
              if (c + 1 < nch) tmem_ld_32x32(treg + (c + 1) * 32, vb);
              if (c * 32 + 32 <= kmax) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  m0 = fmaxf(m0, __uint_as_float(va[j]));
                  m1 = fmaxf(m1, __uint_as_float(va[j + 1]));
                  m2 = fmaxf(m2, __uint_as_float(va[j + 2]));
                  m3 = fmaxf(m3, __uint_as_float(va[j + 3]));
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (c * 32 + j < kmax) m0 = fmaxf(m0, __uint_as_float(va[j]));
              }
              if (c + 1 < nch) {
                tmem_ld_wait();
                if (c + 2 < nch) tmem_ld_32x32(treg + (c + 2) * 32, va);
                if (c * 32 + 64 <= kmax) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4) {
                    m0 = fmaxf(m0, __uint_as_float(vb[j]));
                    m1 = fmaxf(m1, __uint_as_float(vb[j + 1]));
                    m2 = fmaxf(m2, __uint_as_float(vb[j + 2]));
                    m3 = fmaxf(m3, __uint_as_float(vb[j + 3]));
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (c * 32 + 32 + j < kmax) m0 = fmaxf(m0, __uint_as_float(vb[j]));
                }
              }
            }
            if (tail) {
              uint32_t vt[16];
              tmem_ld_x16(treg + nch * 32, vt);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (nch * 32 + j < kmax) m1 = fmaxf(m1, __uint_as_float(vt[j]));
            }
            m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          }
          TRF((it == 4 || it == 5) && q == 0, (it - 4) * 12 + 6 + t);
          const float mc = (m == -INFINITY) ? 0.f : m * c2;
          // pass 2: p = exp2(s c2 - m c2), row sum, packed bf16 pairs back into the S columns
          {
            float l0 = 0.f, l1 = 0.f;
            uint32_t va[32], vb[32], w[16];
            auto chunk = [&](const uint32_t (&v)[32], int c) {
              if (c * 32 + 32 <= kmax) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float p0 = ex2(fmaf(__uint_as_float(v[j]), c2, -mc));
                  const float p1 = ex2(fmaf(__uint_as_float(v[j + 1]), c2, -mc));
                  l0 += p0;
                  l1 += p1;
                  w[j >> 1] = pack_bf16(p0, p1);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float p0 =
                      (c * 32 + j < kmax) ? ex2(fmaf(__uint_as_float(v[j]), c2, -mc)) : 0.f;
                  const float p1 =
                      (c * 32 + j + 1 < kmax) ? ex2(fmaf(__uint_as_float(v[j + 1]), c2, -mc)) : 0.f;
                  l0 += p0;
                  l1 += p1;
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
            if (tail) {
              uint32_t vt[16], wt[8];
              tmem_ld_x16(treg + nch * 32, vt);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float p0 =
                    (nch * 32 + j < kmax) ? ex2(fmaf(__uint_as_float(vt[j]), c2, -mc)) : 0.f;
                const float p1 =
                    (nch * 32 + j + 1 < kmax) ? ex2(fmaf(__uint_as_float(vt[j + 1]), c2, -mc)) : 0.f;
                l0 += p0;
                l1 += p1;
                wt[j >> 1] = pack_bf16(p0, p1);
              }
              tmem_st_x8(treg + nch * 16, wt);
            }
            l = l0 + l1;
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&p_ready[t]));
        TRF((it == 4 || it == 5) && q == 0, (it - 4) * 12 + 8 + t);
        TRF(it == 4, 24 + t * 4 + q);
        if (p.lse != nullptr && row < p.L) p.lse[(size_t)pr * p.lse_ld + row] = m * 0.125f + logf(l);
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        // epilogue: O row (64 fp32) -> bf16 -> staging tile -> TMA store
        mbar_wait_relaxed(smem_u32(&o_full[t]), tp);
        tc_fence_after();
        uint32_t o0[32], o1[32];   // (unconditional: a conditional load sends the arrays to local memory)
        tmem_ld_32x32(treg + 128, o0);
        tmem_ld_32x32(treg + 160, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_free[t]));
        if (gtid == 0) tma_store_wait_read<0>();   // the previous pair's store has read `stg`
        if (t == 0) named_bar_sync(1, 128); else named_bar_sync(2, 128);
        if (live) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<uint4*>(stg + sw128_off(r, j)) = make_uint4(
                pack_bf16(__uint_as_float(o0[8 * j]) * inv, __uint_as_float(o0[8 * j + 1]) * inv),
                pack_bf16(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv),
                pack_bf16(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv),
                pack_bf16(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv));
            *reinterpret_cast<uint4*>(stg + sw128_off(r, 4 + j)) = make_uint4(
                pack_bf16(__uint_as_float(o1[8 * j]) * inv, __uint_as_float(o1[8 * j + 1]) * inv),
                pack_bf16(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv),
                pack_bf16(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv),
                pack_bf16(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv));
          }
          fence_proxy_async_smem();
        }
        if (t == 0) named_bar_sync(1, 128); else named_bar_sync(2, 128);
        if (gtid == 0) {
          tma_store_3d(&tmO, smem_u32(stg), h * HD, t * 128, n);
          tma_store_commit();
        }
        TRF((it == 4 || it == 5) && q == 0, (it - 4) * 12 + 10 + t);
        if ((p.dbg & 128) && blockIdx.x == 0 && it == 6 && gtid == 0 && t == 0) {
          const uint32_t t0 = trace[0];
          printf("pair 4 p_ready arrivals by lane quarter: g0 %d %d %d %d | g1 %d %d %d %d\n", (int)(trace[24] - t0),
                 (int)(trace[25] - t0), (int)(trace[26] - t0), (int)(trace[27] - t0), (int)(trace[28] - t0),
                 (int)(trace[29] - t0), (int)(trace[30] - t0), (int)(trace[31] - t0));
          for (int k = 0; k < 2; ++k)
            printf("pair %d: S0 %d S1 %d PV0 %d PV1 %d | g0: sfull %d max %d pdone %d stored %d | g1: sfull %d max %d pdone %d stored %d\n",
                   4 + k, (int)(trace[k * 12] - t0), (int)(trace[k * 12 + 1] - t0), (int)(trace[k * 12 + 2] - t0),
                   (int)(trace[k * 12 + 3] - t0), (int)(trace[k * 12 + 4] - t0), (int)(trace[k * 12 + 6] - t0),
                   (int)(trace[k * 12 + 8] - t0), (int)(trace[k * 12 + 10] - t0), (int)(trace[k * 12 + 5] - t0),
                   (int)(trace[k * 12 + 7] - t0), (int)(trace[k * 12 + 9] - t0), (int)(trace[k * 12 + 11] - t0));
        }
      }
      if (gtid == 0) tma_store_wait<0>();
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

int llc_attn_fwd_tc2(const void* qkv, int ld_qkv, void* o, int ld_o, float* lse, int N, int L,
                     int H, int sn, int sl, int causal, cudaStream_t st, int lse_ld) {
  CUtensorMap tm, to;
  // tokens are (sample n, position l) at row n sn + l sl: [columns, L, N] with strides (sl, sn)
  if (int rc = llc_encode_tmap_3d(&tm, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                                  (uint64_t)3 * H * HD, (uint64_t)L, (uint64_t)N,
                                  (uint64_t)ld_qkv * 2 * sl, (uint64_t)ld_qkv * 2 * sn, HD, 128, 1,
                                  CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  if (int rc = llc_encode_tmap_3d(&to, o, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)H * HD,
                                  (uint64_t)L, (uint64_t)N, (uint64_t)ld_o * 2 * sl,
                                  (uint64_t)ld_o * 2 * sn, HD, 128, 1, CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  Fwd2Params p;
  p.lse = lse;
  p.lse_ld = lse_ld > 0 ? lse_ld : L;
  p.N = N; p.L = L; p.H = H; p.LK = (L + 15) / 16 * 16; p.NT = (L + 127) / 128;
  p.causal = causal;
  static const int dbg = llc_dev_env("LLC_ATTN_DBG") ? atoi(llc_dev_env("LLC_ATTN_DBG")) : 0;
  p.dbg = dbg | (g_llc_pdl_trigger ? 4096 : 0);
  LLC_CONFIGURE_SMEM(attn_fwd2_kernel, kSmem);
  const int grid = N * H < llc_num_sms() ? N * H : llc_num_sms();
  LLC_PROF_BEGIN(LLC_K_ATTN_FWD, N * H, L, 0, 4.0 * N * H * (double)L * L * HD,
                 8.0 * N * H * (double)L * HD, st);
  LLC_CUDA(llc_launch_pdl(attn_fwd2_kernel, dim3(grid), dim3(kThreads), (size_t)kSmem, st, tm, to, p));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_fwd2_kernel");
  return 0;
}
