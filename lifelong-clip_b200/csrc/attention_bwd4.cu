// Attention backward, unit-pipelined (sequence length <= 256, head dim 64), all on tcgen05.
//   dP = dO V^T, dS = P o (dP - delta), dQ = dS K / 8, dK = dS^T Q / 8, dV = P^T dO
// (autograd of models/clip/lora.py:950,1043,1063,1068). Every score is computed ONCE, in the
// TRANSPOSED orientation: a unit is (key tile j of <= 128 keys) x (64 queries u).
//   S-type MMAs   RS[b] = K_j Q_u^T, RP[b] = V_j dO_u^T      (TMEM fp32, 64 columns each, b = unit & 1)
//   element-wise  256 threads, thread = (key row, half of the unit's queries): P^T and dS^T.
//                 Both go back to TMEM in place as packed bf16 (A operands of dV, dK: an MMA whose
//                 A comes from shared memory pays 64 clk for the 4 KB A slice whatever N is, twice
//                 the N = 64 floor); dS^T also goes to a shared-memory tile [128 keys x 64 queries]
//   output MMAs   dV_j += P^T dO_u          A = P^T from TMEM, B = dO_u as loaded (MN-major)
//                 dK_j += dS^T Q_u          A = dS^T from TMEM, B = Q_u as loaded
//                 dQ_t += dS K_j            once per two units: A = the two dS^T tiles read
//                                           TRANSPOSED (MN-major A), B = K_j as loaded
// One persistent CTA per SM walks over (sample, head) pairs; everything that is not the
// element-wise math is kept off its critical path:
//   warp 0      TMA producer. The operands of the NEXT pair are loaded piecewise into the regions
//               the current pair has finished with (K_0/V_0 after key tile 0, Q_u/dO_u after the
//               unit's last MMAs), so a pair starts without waiting for HBM.
//   warp 1      MMA issuer. RS/RP are double-buffered: the tensor pipe works on unit g+1 and on
//               the output MMAs of unit g-1 while the element-wise warps are busy with unit g.
//   warps 2-9   element-wise.
//   warps 10-13 auxiliary: lse / delta = rowsum(dO o O) of the next pair (straight from global),
//               and the accumulator drains (TMEM -> bf16 staging tile -> TMA store).
// TMEM: RS 2x64 + RP 2x64 + dQ_0, dQ_1, dV_j, dK_j 4x64 = 512 columns. tcgen05 executes one
// thread's MMAs in order, which is what orders buffer reuse against the MMAs that read them.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int kThreads = 14 * 32;
constexpr int kTileBytes = 128 * 128;           // 128 rows x 128 B: operand tile / dS^T unit tile
constexpr int kUnitRows = 64 * 128;             // 64 operand rows (one query unit) in bytes
constexpr int kStagingBytes = 2 * kTileBytes;   // two output tiles in flight
constexpr float kLog2e = 1.4426950408889634f;
// TMEM columns
constexpr uint32_t kRS = 0, kRP = 128, kDQ = 256, kDV = 384, kDK = 448;
// Shared-memory descriptors are built from a 32-bit low word (start address >> 4 | LBO field)
// and one common high word (SBO = 1024 B, descriptor version 1, 128 B swizzle): stepping through
// an operand is then a single 32-bit add on the issuing thread (addresses never carry out of
// the 14-bit field).
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo_k(uint32_t addr) {          // K-major, SW128
  return ((addr & 0x3FFFFu) >> 4) | (1u << 16);
}
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t addr, uint32_t lbo) {   // MN-major, SW128
  return ((addr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo) {
  return ((uint64_t)kDescHi << 32) | lo;
}

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


struct Bwd4Params {
  const float* lse;     // [N*H, L] log-sum-exp of the scaled scores (natural log)
  const float* delta;   // [tokens, H] rowsum(dO o O) per (token, head)
  int N, L, H, LK, NT, NU, sn, sl, causal, mat_bytes, dbg;
};

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd4_kernel(const __grid_constant__ CUtensorMap tmQ64, const __grid_constant__ CUtensorMap tmQ16,
                 const __grid_constant__ CUtensorMap tmD64, const __grid_constant__ CUtensorMap tmD16,
                 const __grid_constant__ CUtensorMap tmOut, Bwd4Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int LK = p.LK, NT = p.NT, NU = p.NU, mat = p.mat_bytes;
  uint8_t* sdS = smem + 4 * mat;                  // 4 unit tiles: (block & 1) * 2 + (unit & 1)
  uint8_t* staging = sdS + 4 * kTileBytes;
  float* sLse = reinterpret_cast<float*>(staging + kStagingBytes);   // [2][256] lse * log2e
  float* sDelta = sLse + 512;                                        // [2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 512);
  uint64_t* kv_full = bars;         // [2] K_j, V_j landed
  uint64_t* kv_empty = bars + 2;    // [2] every MMA reading K_j, V_j of the pair retired
  uint64_t* q_full = bars + 4;      // [4] Q_u, dO_u landed
  uint64_t* q_empty = bars + 8;     // [4] every MMA reading Q_u, dO_u of the pair retired
  uint64_t* s_full = bars + 12;     // [2] RS/RP[b] hold a unit's S-type products
  uint64_t* p_ready = bars + 14;    // [2] P^T in TMEM, dS^T tile in smem (8 warps)
  uint64_t* acc_done = bars + 16;   // dV_j / dK_j complete
  uint64_t* acc_free = bars + 17;   // ...and drained (4 warps)
  uint64_t* dq_done = bars + 18;    // dQ_0 / dQ_1 complete
  uint64_t* dq_free = bars + 19;    // ...and drained (4 warps)
  uint64_t* dl_full = bars + 20;    // [2] lse / delta of a pair ready (4 warps)
  uint64_t* dl_empty = bars + 22;   // [2] ...and consumed (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  uint32_t* trace = reinterpret_cast<uint32_t*>(bars + 26);   // [96] debug timeline (dbg & 128)
#define TR(i)                                                                                     \
  do {                                                                                            \
    if ((p.dbg & 128) && blockIdx.x == 0 && it == 2 && lane == 0) trace[i] = (uint32_t)clock64(); \
  } while (0)

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int pairs = p.N * p.H;
  const int D = p.H * HD;
  const int total = NT * NU;      // units per pair

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    tma_prefetch_desc(&tmQ64); tma_prefetch_desc(&tmQ16);
    tma_prefetch_desc(&tmD64); tma_prefetch_desc(&tmD16);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&kv_full[i]), 1);
      mbar_init(smem_u32(&kv_empty[i]), 1);
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_ready[i]), 8);
      mbar_init(smem_u32(&dl_full[i]), 4);
      mbar_init(smem_u32(&dl_empty[i]), 8);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(smem_u32(&q_full[i]), 1);
      mbar_init(smem_u32(&q_empty[i]), 1);
    }
    mbar_init(smem_u32(acc_done), 1);
    mbar_init(smem_u32(acc_free), 4);
    mbar_init(smem_u32(dq_done), 1);
    mbar_init(smem_u32(dq_free), 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t sQ = smem_u32(smem), sK = sQ + mat, sV = sQ + 2 * mat, sD = sQ + 3 * mat;
  pdl_wait();

  // register budget per warpgroup (512 x 128 in all): issuers 56, element-wise 88, auxiliary 104
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // rows [row0, row0 + nrows) of one operand: 64-row boxes, then 16-row boxes
    auto load_rows = [&](uint32_t dst, const CUtensorMap* m64, const CUtensorMap* m16, uint32_t bar,
                         int col, int row0, int nrows, int n) {
      int r = 0;
      for (; r + 64 <= nrows; r += 64) tma_load_3d(dst + (row0 + r) * 128, m64, bar, col, row0 + r, n);
      for (; r < nrows; r += 16) tma_load_3d(dst + (row0 + r) * 128, m16, bar, col, row0 + r, n);
    };
    int it = 0;
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const int n = pr / p.H, h = pr % p.H;
      const uint32_t par = (it & 1) ^ 1;
      if ((p.dbg & 4096) && pr + (int)gridDim.x >= pairs && elect_one()) pdl_trigger();
      auto load_kv = [&](int j) {
        const int Kj = min(128, LK - 128 * j);
        mbar_wait(smem_u32(&kv_empty[j]), par);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&kv_full[j]);
          mbar_expect_tx(fb, 2 * Kj * 128);
          load_rows(sK, &tmQ64, &tmQ16, fb, D + h * HD, 128 * j, Kj, n);
          load_rows(sV, &tmQ64, &tmQ16, fb, 2 * D + h * HD, 128 * j, Kj, n);
        }
        __syncwarp();
      };
      // in the order the previous pair releases the regions and this pair needs them
      load_kv(0);
      for (int u = 0; u < NU; ++u) {
        const int Wu = min(64, LK - 64 * u);
        mbar_wait(smem_u32(&q_empty[u]), par);
        if (elect_one()) {
          const uint32_t fb = smem_u32(&q_full[u]);
          mbar_expect_tx(fb, 2 * Wu * 128);
          load_rows(sQ, &tmQ64, &tmQ16, fb, h * HD, 64 * u, Wu, n);
          load_rows(sD, &tmD64, &tmD16, fb, h * HD, 64 * u, Wu, n);
        }
        __syncwarp();
      }
      for (int j = 1; j < NT; ++j) load_kv(j);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);      // A K-major / TMEM, B MN-major
    const uint32_t idesc_t = umma_idesc_bf16(128, HD, 1, 1);      // A MN-major (transposed tiles)
    const uint32_t kQ = desc_lo_k(sQ), kK = desc_lo_k(sK), kV = desc_lo_k(sV), kD = desc_lo_k(sD);
    const uint32_t mQ = desc_lo_mn(sQ, 8192), mK = desc_lo_mn(sK, 8192), mD = desc_lo_mn(sD, 8192);
    const uint32_t mS = desc_lo_mn(smem_u32(sdS), kTileBytes);
    int it = 0, g = 0, gj = 0, blk = 0;
    // RS[b] = K_j Q_u^T, RP[b] = V_j dO_u^T
    auto issue_s = [&](int j, int u, int b) {
      const int Wu = min(64, LK - 64 * u);
      const uint32_t idesc_s = umma_idesc_bf16(128, Wu, 0, 0);
      const uint32_t a0 = kK + j * 1024, b0 = kQ + u * 512, a1 = kV + j * 1024, b1 = kD + u * 512;
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        umma_bf16(tmem_base + kRS + b * 64, mk_desc(a0 + 2 * k), mk_desc(b0 + 2 * k), idesc_s, k != 0);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        umma_bf16(tmem_base + kRP + b * 64, mk_desc(a1 + 2 * k), mk_desc(b1 + 2 * k), idesc_s, k != 0);
      umma_commit(smem_u32(&s_full[b]));
    };
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      TR(0);
      // the first two units' S-type products
      mbar_wait(smem_u32(&kv_full[0]), par);
      mbar_wait(smem_u32(&q_full[0]), par);
      tc_fence_after();
      if (elect_one()) issue_s(0, 0, g & 1);
      __syncwarp();
      int j2 = 0, u2 = 1;                 // the next unit whose S-type products are to be issued
      if (u2 >= NU) { u2 = 0; j2 = 1; }
      if (j2 < NT) {
        if (j2 == 0) mbar_wait(smem_u32(&q_full[u2]), par);
        else mbar_wait(smem_u32(&kv_full[j2]), par);
        tc_fence_after();
        if (elect_one()) issue_s(j2, u2, (g + 1) & 1);
        __syncwarp();
        if (++u2 >= NU) { u2 = 0; ++j2; }
      }
      TR(1);
      mbar_wait(smem_u32(dq_free), par ^ 1);   // previous pair's dQ drained
      int n = 0;
      for (int j = 0; j < NT; ++j) {
        const int Kj = min(128, LK - 128 * j);
        mbar_wait(smem_u32(acc_free), (gj & 1) ^ 1);   // previous dV / dK drained
        for (int u = 0; u < NU; ++u, ++g, ++n) {
          const int b = g & 1;
          const int nks = min(64, LK - 64 * u) / 16;
          const bool block_end = (u & 1) || u == NU - 1;
          mbar_wait(smem_u32(&p_ready[b]), (g >> 1) & 1);
          if (j2 < NT) {    // operands of the unit after next (first use in this pair only)
            if (j2 == 0) mbar_wait(smem_u32(&q_full[u2]), par);
            else if (u2 == 0) mbar_wait(smem_u32(&kv_full[j2]), par);
          }
          tc_fence_after();
          TR(2 + 2 * n);
          if (elect_one()) {
            const uint32_t aP = tmem_base + kRS + b * 64, aS = tmem_base + kRP + b * 64;
            // dV_j (+)= P^T dO_u ; dK_j (+)= dS^T Q_u : K dimension = the unit's queries
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < nks) {
                umma_bf16_ts(tmem_base + kDV, aP + ks * 16,
                             mk_desc(mD + u * 512 + ks * 128), idesc_o, (u | ks) != 0);
                umma_bf16_ts(tmem_base + kDK, aS + ks * 16, mk_desc(mQ + u * 512 + ks * 128), idesc_o,
                             (u | ks) != 0);
              }
            if (j == NT - 1) umma_commit(smem_u32(&q_empty[u]));   // Q_u, dO_u: no reader left
            // dQ_t (+)= dS K_j once the block's (<= 2) units are in: K dimension = keys
            if (block_end) {
              const uint32_t a = mS + (uint32_t)((blk & 1) * 2) * 1024, bb = mK + j * 1024;
              const int nkk = Kj / 16;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                if (ks < nkk)
                  umma_bf16(tmem_base + kDQ + (u >> 1) * 64, mk_desc(a + ks * 128),
                            mk_desc(bb + ks * 128), idesc_t, (j | ks) != 0);
            }
            if (u == NU - 1) {
              umma_commit(smem_u32(acc_done));
              if (j2 > j || j2 >= NT) umma_commit(smem_u32(&kv_empty[j]));   // see below
              if (j == NT - 1) umma_commit(smem_u32(dq_done));
            }
            // the unit after next reuses this unit's buffers (same in-order pipe)
            if (j2 < NT) issue_s(j2, u2, b);
          }
          __syncwarp();
          TR(3 + 2 * n);
          if (j2 < NT && ++u2 >= NU) { u2 = 0; ++j2; }
          if (block_end) ++blk;
        }
        ++gj;
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ element-wise
    const int q4 = warp & 3;            // TMEM lane quarter
    const int hh = (warp - 2) >> 2;     // which half of the unit's (<= 4) 16-query chunks
    const int r = q4 * 32 + lane;       // key row within the tile
    const uint32_t tb = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const float c2 = 0.125f * kLog2e;
    int it = 0, g = 0, blk = 0;
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const float* ls_p = sLse + (it & 1) * 256;
      const float* de_p = sDelta + (it & 1) * 256;
      if (warp == 2) TR(32);
      mbar_wait(smem_u32(&dl_full[it & 1]), (it >> 1) & 1);
      if (warp == 2) TR(33);
      int n = 0;
      for (int j = 0; j < NT; ++j) {
        const int Kj = min(128, LK - 128 * j);
        const int key = j * 128 + r;
        const bool warp_on = q4 * 32 < Kj;       // warp-uniform: any row of this quarter in range
        for (int u = 0; u < NU; ++u, ++g, ++n) {
          const int b = g & 1;
          const int nc = min(64, LK - 64 * u) / 16;    // 16-query chunks in this unit
          uint8_t* tile = sdS + ((blk & 1) * 2 + (u & 1)) * kTileBytes;
          mbar_wait(smem_u32(&s_full[b]), (g >> 1) & 1);
          tc_fence_after();
          if (warp == 2) TR(34 + 2 * n);
          if (warp_on && 2 * hh < nc && !(p.dbg & 1)) {
            const bool two = 2 * hh + 1 < nc;        // this half has a second chunk
            uint32_t sv[2][16], dv[2][16];
            tmem_ld_x16(tb + kRS + b * 64 + hh * 32, sv[0]);
            tmem_ld_x16(tb + kRP + b * 64 + hh * 32, dv[0]);
            if (two) {
              tmem_ld_x16(tb + kRS + b * 64 + hh * 32 + 16, sv[1]);
              tmem_ld_x16(tb + kRP + b * 64 + hh * 32 + 16, dv[1]);
            }
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c == 1 && !two) break;
              // chunk ks of the unit: queries q0 .. q0 + 15
              const int ks = 2 * hh + c;
              const int q0 = 64 * u + ks * 16;
              uint32_t wp[8], wd[8];
              if (key >= p.L || (p.dbg & 2)) {
#pragma unroll
                for (int e = 0; e < 8; ++e) wp[e] = wd[e] = 0u;
              } else if (q0 + 16 <= p.L && (!p.causal || key <= q0)) {
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                  const float4 l4 = *reinterpret_cast<const float4*>(&ls_p[q0 + e]);
                  const float4 d4 = *reinterpret_cast<const float4*>(&de_p[q0 + e]);
                  const float p0 = ex2(fmaf(__uint_as_float(sv[c][e]), c2, -l4.x));
                  const float p1 = ex2(fmaf(__uint_as_float(sv[c][e + 1]), c2, -l4.y));
                  const float p2 = ex2(fmaf(__uint_as_float(sv[c][e + 2]), c2, -l4.z));
                  const float p3 = ex2(fmaf(__uint_as_float(sv[c][e + 3]), c2, -l4.w));
                  wp[e >> 1] = pack_bf16(p0, p1);
                  wp[(e >> 1) + 1] = pack_bf16(p2, p3);
                  wd[e >> 1] = pack_bf16(p0 * (__uint_as_float(dv[c][e]) - d4.x),
                                         p1 * (__uint_as_float(dv[c][e + 1]) - d4.y));
                  wd[(e >> 1) + 1] = pack_bf16(p2 * (__uint_as_float(dv[c][e + 2]) - d4.z),
                                               p3 * (__uint_as_float(dv[c][e + 3]) - d4.w));
                }
              } else {
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                  const bool v0 = q0 + e < p.L && (!p.causal || key <= q0 + e);
                  const bool v1 = q0 + e + 1 < p.L && (!p.causal || key <= q0 + e + 1);
                  const float p0 = v0 ? ex2(fmaf(__uint_as_float(sv[c][e]), c2, -ls_p[q0 + e])) : 0.f;
                  const float p1 =
                      v1 ? ex2(fmaf(__uint_as_float(sv[c][e + 1]), c2, -ls_p[q0 + e + 1])) : 0.f;
                  wp[e >> 1] = pack_bf16(p0, p1);
                  wd[e >> 1] =
                      pack_bf16(v0 ? p0 * (__uint_as_float(dv[c][e]) - de_p[q0 + e]) : 0.f,
                                v1 ? p1 * (__uint_as_float(dv[c][e + 1]) - de_p[q0 + e + 1]) : 0.f);
                }
              }
              if (!(p.dbg & 4)) {
                // in place over the chunk's own (already read) S / dP columns
                tmem_st_x8(tb + kRS + b * 64 + ks * 16, wp);          // A operand of dV
                tmem_st_x8(tb + kRP + b * 64 + ks * 16, wd);          // A operand of dK
                *reinterpret_cast<uint4*>(tile + sw128_off(r, 2 * ks)) =
                    make_uint4(wd[0], wd[1], wd[2], wd[3]);
                *reinterpret_cast<uint4*>(tile + sw128_off(r, 2 * ks + 1)) =
                    make_uint4(wd[4], wd[5], wd[6], wd[7]);
              }
            }
            tmem_st_wait();
            fence_proxy_async_smem();     // dS^T tile -> visible to the tensor core
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&p_ready[b]));
          if (warp == 2) TR(35 + 2 * n);
          if ((u & 1) || u == NU - 1) ++blk;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&dl_empty[it & 1]));   // lse / delta buffer consumed
    }
  } else {
    // ------------------------------------------------------------------ auxiliary
    const int aw = warp - 10;           // 0..3
    const int q4 = warp & 3;            // TMEM lane quarter
    const int r = q4 * 32 + lane;       // accumulator row
    const int tid3 = aw * 32 + lane;
    const uint32_t tb = tmem_base + ((uint32_t)(q4 * 32) << 16);
    int sbuf = 0;   // staging tiles alternate: a store may still be reading the other one

    // this thread's row of an accumulator (64 fp32 columns) -> bf16 staging tile -> TMA store
    auto drain_tile = [&](uint32_t src, uint64_t* free_bar, bool release, float sc, int col,
                          int row0, int n) {
      uint32_t a[32], b[32];
      tmem_ld_32x32(src, a);
      tmem_ld_32x32(src + 32, b);
      tmem_ld_wait();
      if (release) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(free_bar));
      }
      uint8_t* stg = staging + sbuf * kTileBytes;
      sbuf ^= 1;
      if (tid3 == 0) tma_store_wait_read<1>();   // the store issued two tiles ago has read `stg`
      named_bar_sync(1, 128);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<uint4*>(stg + sw128_off(r, j)) = make_uint4(
            pack_bf16(__uint_as_float(a[8 * j]) * sc, __uint_as_float(a[8 * j + 1]) * sc),
            pack_bf16(__uint_as_float(a[8 * j + 2]) * sc, __uint_as_float(a[8 * j + 3]) * sc),
            pack_bf16(__uint_as_float(a[8 * j + 4]) * sc, __uint_as_float(a[8 * j + 5]) * sc),
            pack_bf16(__uint_as_float(a[8 * j + 6]) * sc, __uint_as_float(a[8 * j + 7]) * sc));
        *reinterpret_cast<uint4*>(stg + sw128_off(r, 4 + j)) = make_uint4(
            pack_bf16(__uint_as_float(b[8 * j]) * sc, __uint_as_float(b[8 * j + 1]) * sc),
            pack_bf16(__uint_as_float(b[8 * j + 2]) * sc, __uint_as_float(b[8 * j + 3]) * sc),
            pack_bf16(__uint_as_float(b[8 * j + 4]) * sc, __uint_as_float(b[8 * j + 5]) * sc),
            pack_bf16(__uint_as_float(b[8 * j + 6]) * sc, __uint_as_float(b[8 * j + 7]) * sc));
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (tid3 == 0 && !(p.dbg & 8)) {
        tma_store_3d(&tmOut, smem_u32(stg), col, row0, n);
        tma_store_commit();
      }
    };
    // lse * log2e and delta of one pair -> buffer (i & 1). The four global loads are issued
    // early (prepare_load) and only consumed after the next drain (prepare_store), so their
    // latency never sits in front of an accumulator drain.
    float pl0 = 0.f, pl1 = 0.f, pd0 = 0.f, pd1 = 0.f;
    auto prepare_load = [&](int pr) {
      const size_t base = (size_t)pr * p.L;
      const int n = pr / p.H, h = pr % p.H;
      const size_t tok = (size_t)n * p.sn;
      pl0 = tid3 < p.L ? p.lse[base + tid3] : 0.f;
      pd0 = tid3 < p.L ? p.delta[(tok + (size_t)tid3 * p.sl) * p.H + h] : 0.f;
      pl1 = tid3 + 128 < p.L ? p.lse[base + tid3 + 128] : 0.f;
      pd1 = tid3 + 128 < p.L ? p.delta[(tok + (size_t)(tid3 + 128) * p.sl) * p.H + h] : 0.f;
    };
    auto prepare_store = [&](int i) {
      float* ls_p = sLse + (i & 1) * 256;
      float* de_p = sDelta + (i & 1) * 256;
      mbar_wait(smem_u32(&dl_empty[i & 1]), ((i >> 1) & 1) ^ 1);
      ls_p[tid3] = pl0 * kLog2e;
      ls_p[tid3 + 128] = pl1 * kLog2e;
      de_p[tid3] = pd0;
      de_p[tid3 + 128] = pd1;
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&dl_full[i & 1]));
    };

    int it = 0, gj = 0;
    if (blockIdx.x < pairs) {
      prepare_load(blockIdx.x);
      prepare_store(0);
    }
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const int n = pr / p.H, h = pr % p.H;
      const bool has_next = pr + (int)gridDim.x < pairs;
      if (has_next) prepare_load(pr + gridDim.x);
      for (int j = 0; j < NT; ++j, ++gj) {
        // dV_j, then dK_j (scaled by hd^-0.5); the accumulators are released after the second read
        mbar_wait(smem_u32(acc_done), gj & 1);
        tc_fence_after();
        if (aw == 0) TR(52 + 2 * j);
        drain_tile(tb + kDV, acc_free, false, 1.0f, 2 * D + h * HD, j * 128, n);
        drain_tile(tb + kDK, acc_free, true, 0.125f, D + h * HD, j * 128, n);
        if (j == 0 && has_next) prepare_store(it + 1);
        if (aw == 0) TR(53 + 2 * j);
      }
      mbar_wait(smem_u32(dq_done), it & 1);
      tc_fence_after();
      if (aw == 0) TR(56);
      drain_tile(tb + kDQ, dq_free, NT == 1, 0.125f, h * HD, 0, n);
      if (NT > 1) drain_tile(tb + kDQ + 64, dq_free, true, 0.125f, h * HD, 128, n);
      if (aw == 0) TR(57);
      if ((p.dbg & 128) && blockIdx.x == 0 && it == 2 && tid3 == 0) {
        const uint32_t t0 = trace[0];
        printf("MMA: start 0 prologue %u |", trace[1] - t0);
        for (int i = 0; i < total; ++i)
          printf(" u%d: pready %u issued %u |", i, trace[2 + 2 * i] - t0, trace[3 + 2 * i] - t0);
        printf("\nELT: start %d delta %d |", (int)(trace[32] - t0), (int)(trace[33] - t0));
        for (int i = 0; i < total; ++i)
          printf(" u%d: sfull %d arrived %d |", i, (int)(trace[34 + 2 * i] - t0), (int)(trace[35 + 2 * i] - t0));
        printf("\n kv0 %d..%d kv1 %d..%d dq %d end %d\n", (int)(trace[52] - t0), (int)(trace[53] - t0),
               (int)(trace[54] - t0), (int)(trace[55] - t0), (int)(trace[56] - t0), (int)(trace[57] - t0));
      }
    }
    if (tid3 == 0) tma_store_wait<0>();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
#undef TR
}

// delta[tok H + h] = sum_d dO[tok, h, d] * O[tok, h, d], tok = n sn + l sl. Eight lanes per
// (token, head): one 16 B piece of each row, three shuffles. blockIdx.y = sample; HC = H as a
// compile-time constant (0: runtime) keeps the index arithmetic off the XU pipe.
template <int HC>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, int ld_o, const __nv_bfloat16* __restrict__ d_o,
                  int ld_do, float* __restrict__ delta, int L, int Hrt, int sn, int sl) {
  pdl_wait();
  const int H = HC ? HC : Hrt;
  const int piece = threadIdx.x & 7;
  const int item = blockIdx.x * 32 + (threadIdx.x >> 3);     // (l, h) of this sample
  const int n = blockIdx.y;
  float d = 0.f;
  int l = 0, h = 0;
  const bool ok = item < L * H;
  if (ok) {
    l = item / H;
    h = item - l * H;
    const size_t tok = (size_t)n * sn + (size_t)l * sl;
    const uint4 ov = *reinterpret_cast<const uint4*>(o + tok * ld_o + h * HD + piece * 8);
    const uint4 dv = *reinterpret_cast<const uint4*>(d_o + tok * ld_do + h * HD + piece * 8);
    const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 a = unpack_bf16(dw[e]), b = unpack_bf16(ow[e]);
      d += a.x * b.x + a.y * b.y;
    }
  }
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  d += __shfl_xor_sync(0xffffffffu, d, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 4);
  if (ok && piece == 0) delta[((size_t)n * sn + (size_t)l * sl) * H + h] = d;
}

int encode_rows(CUtensorMap* tm, const void* base, int cols, int ld, int L, int N, int sn, int sl,
                int box_rows) {
  return llc_encode_tmap_3d(tm, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)cols,
                            (uint64_t)L, (uint64_t)N, (uint64_t)ld * 2 * sl, (uint64_t)ld * 2 * sn,
                            HD, box_rows, 1, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace

// Shared memory the kernel needs for sequence length L (the caller falls back to the
// block-structured kernel when it does not fit).
int llc_attn_bwd_tc4_smem(int L) {
  const int LK = (L + 15) / 16 * 16;
  return 4 * LK * 128 + 4 * kTileBytes + kStagingBytes + 4 * 256 * 4 + 1024;
}

int llc_attn_bwd_tc4(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o,
                     int ld_do, const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H,
                     int sn, int sl, int causal, float* delta_ws, int delta_ready, cudaStream_t st) {
  // delta_ready: the caller already filled delta_ws[token * H + head] (llc_colsum_tc_delta)
  LLC_REQUIRE(delta_ws, "llc_attn_bwd: delta scratch ([N*H*L] floats) is required");
  Bwd4Params p;
  p.lse = lse;
  p.delta = delta_ws;
  p.N = N; p.L = L; p.H = H; p.LK = (L + 15) / 16 * 16; p.NT = (L + 127) / 128;
  p.NU = (p.LK + 63) / 64;
  p.sn = sn; p.sl = sl; p.causal = causal;
  p.mat_bytes = p.LK * 128;
  static const int dbg = llc_dev_env("LLC_ATTN_DBG") ? atoi(llc_dev_env("LLC_ATTN_DBG")) : 0;
  p.dbg = dbg | (g_llc_pdl_trigger ? 4096 : 0);
  const int smem = llc_attn_bwd_tc4_smem(L);
  CUtensorMap q64, q16, d64, d16, to;
  if (int rc = encode_rows(&q64, qkv, 3 * H * HD, ld_qkv, L, N, sn, sl, 64)) return rc;
  if (int rc = encode_rows(&q16, qkv, 3 * H * HD, ld_qkv, L, N, sn, sl, 16)) return rc;
  if (int rc = encode_rows(&d64, d_o, H * HD, ld_do, L, N, sn, sl, 64)) return rc;
  if (int rc = encode_rows(&d16, d_o, H * HD, ld_do, L, N, sn, sl, 16)) return rc;
  if (int rc = encode_rows(&to, dqkv, 3 * H * HD, ld_dqkv, L, N, sn, sl, 128)) return rc;
  LLC_CONFIGURE_SMEM(attn_bwd4_kernel, smem);
  const int grid = N * H < llc_num_sms() ? N * H : llc_num_sms();
  LLC_PROF_BEGIN(LLC_K_ATTN_BWD, N * H, L, 0, 8.0 * N * H * (double)L * L * HD,
                 16.0 * N * H * (double)L * HD, st);
  if (!(delta_ready && delta_ws)) {
    const dim3 dgrid((unsigned)((L * H + 31) / 32), (unsigned)N);
    const __nv_bfloat16* ob = reinterpret_cast<const __nv_bfloat16*>(o);
    const __nv_bfloat16* db = reinterpret_cast<const __nv_bfloat16*>(d_o);
    if (H == 12)
      LLC_CUDA(llc_launch_pdl(attn_delta_kernel<12>, dgrid, dim3(256), (size_t)0, st, ob, ld_o, db,
                              ld_do, delta_ws, L, H, sn, sl));
    else if (H == 16)
      LLC_CUDA(llc_launch_pdl(attn_delta_kernel<16>, dgrid, dim3(256), (size_t)0, st, ob, ld_o, db,
                              ld_do, delta_ws, L, H, sn, sl));
    else
      LLC_CUDA(llc_launch_pdl(attn_delta_kernel<0>, dgrid, dim3(256), (size_t)0, st, ob, ld_o, db,
                              ld_do, delta_ws, L, H, sn, sl));
    LLC_COUNT_LAUNCH();
    LLC_LAUNCH_CHECK("attn_delta_kernel");
  }
  LLC_CUDA(llc_launch_pdl(attn_bwd4_kernel, dim3(grid), dim3(kThreads), (size_t)smem, st, q64, q16,
                          d64, d16, to, p));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_bwd4_kernel");
  return 0;
}
