// Attention backward, block-structured (sequence length <= 256, head dim 64), all on tcgen05.
//   dP = dO V^T, dS = P o (dP - delta), dQ = dS K / 8, dK = dS^T Q / 8, dV = P^T dO
// (autograd of models/clip/lora.py:950,1043,1063,1068). Every score is computed ONCE: the pair
// (sample, head) is cut into blocks (query tile t, key tile j) of <= 128 x 128.
//   S-type MMAs   R0 = Q_t K_j^T, R1 = dO_t V_j^T                      (TMEM, fp32)
//   element-wise  256 threads, thread = (query row, column half): P, dS. dS goes back to TMEM as
//                 packed bf16 (A operand of dQ); P and dS also go to shared memory as bf16 rows
//   output MMAs   dQ_t += dS K_j           A = dS from TMEM, B = K_j as loaded ([key][hd], MN-major)
//                 dV_j += P^T dO_t         A = the P tile read TRANSPOSED (MN-major A descriptor)
//                 dK_j += dS^T Q_t         A = the dS tile read transposed
// The accumulators dQ_0, dQ_1 (live over the whole pair) and dV_j, dK_j (live over the inner t
// loop) sit next to R0/R1: 128 + 128 + 4 x 64 = 512 TMEM columns exactly. tcgen05 executes the
// issuing thread's MMAs in order, so a block's output MMAs and the next block's S-type MMAs are
// issued back to back and need no barrier between them.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int kThreads = 320;
constexpr int kTileBytes = 128 * 128;
constexpr int kPTileBytes = 2 * kTileBytes;   // [128 q rows] x [2 atoms of 64 keys]
constexpr int kStagingBytes = 2 * kTileBytes;   // two output tiles in flight
constexpr float kLog2e = 1.4426950408889634f;
// TMEM columns
constexpr uint32_t kR0 = 0, kR1 = 128, kDQ = 256, kDV = 384, kDK = 448;

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

struct Bwd3Params {
  const __nv_bfloat16* o;
  int ld_o;
  const float* lse;
  int N, L, H, LK, NT, sn, sl, causal, mat_bytes, dbg;
  int lse_ld;   // elements between two pairs' lse rows (see Fwd2Params)
};

__global__ void __launch_bounds__(kThreads, 1)
attn_bwd3_kernel(const __grid_constant__ CUtensorMap tmQ0, const __grid_constant__ CUtensorMap tmQ1,
                 const __grid_constant__ CUtensorMap tmD0, const __grid_constant__ CUtensorMap tmD1,
                 const __grid_constant__ CUtensorMap tmOut, Bwd3Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int LK = p.LK, NT = p.NT, mat = p.mat_bytes;
  uint8_t* sP = smem + 4 * mat;
  uint8_t* sdS = sP + kPTileBytes;
  uint8_t* staging = sdS + kPTileBytes;
  float* sLse = reinterpret_cast<float*>(staging + kStagingBytes);   // [256] lse * log2e
  float* sDelta = sLse + 256;                                        // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  uint64_t* ld_full = bars;        // TMA landed the first 128 rows of Q,K,V,dO
  uint64_t* ld_full1 = bars + 9;   // ...and the rest (NT > 1)
  uint64_t* ld_empty = bars + 1;   // every MMA of the pair retired
  uint64_t* s_full = bars + 2;     // R0/R1 hold a block's S-type products
  uint64_t* p_ready = bars + 3;    // dS in TMEM, P/dS tiles in smem (256 arrivals)
  uint64_t* kv_done = bars + 4;    // dV_j / dK_j complete
  uint64_t* kv_free = bars + 5;    // ...and drained (256 arrivals)
  uint64_t* dq_done = bars + 6;    // dQ_0 / dQ_1 complete
  uint64_t* dq_free = bars + 7;    // ...and drained (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int pairs = p.N * p.H;
  const int D = p.H * HD;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023) __trap();
    tma_prefetch_desc(&tmQ0); tma_prefetch_desc(&tmQ1);
    tma_prefetch_desc(&tmD0); tma_prefetch_desc(&tmD1);
    tma_prefetch_desc(&tmOut);
    mbar_init(smem_u32(ld_full), 1);
    mbar_init(smem_u32(ld_full1), 1);
    mbar_init(smem_u32(ld_empty), 1);
    mbar_init(smem_u32(s_full), 1);
    mbar_init(smem_u32(p_ready), 256);
    mbar_init(smem_u32(kv_done), 1);
    mbar_init(smem_u32(kv_free), 256);
    mbar_init(smem_u32(dq_done), 1);
    mbar_init(smem_u32(dq_free), 256);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const uint32_t sQ = smem_u32(smem), sK = sQ + mat, sV = sQ + 2 * mat, sD = sQ + 3 * mat;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    int it = 0;
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const int n = pr / p.H, h = pr % p.H;
      mbar_wait(smem_u32(ld_empty), (it & 1) ^ 1);
      if (elect_one()) {
        // first the 128-row tiles every pair starts with, then the remainder on its own barrier
        const uint32_t fb = smem_u32(ld_full), fb1 = smem_u32(ld_full1);
        const int rows0 = NT > 1 ? 128 : LK;
        mbar_expect_tx(fb, 4 * rows0 * 128);
        for (int m = 0; m < 3; ++m)     // Q, K, V: column blocks h*64 + {0, D, 2D}
          tma_load_3d(sQ + m * mat, &tmQ0, fb, m * D + h * HD, 0, n);
        tma_load_3d(sD, &tmD0, fb, h * HD, 0, n);
        if (NT > 1) {
          mbar_expect_tx(fb1, 4 * (LK - 128) * 128);
          for (int m = 0; m < 3; ++m)
            tma_load_3d(sQ + m * mat + kTileBytes, &tmQ1, fb1, m * D + h * HD, 128, n);
          tma_load_3d(sD + kTileBytes, &tmD1, fb1, h * HD, 128, n);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc_o = umma_idesc_bf16(128, HD, 0, 1);      // A from TMEM, B MN-major
    const uint32_t idesc_t = umma_idesc_bf16(128, HD, 1, 1);      // A MN-major (transposed tile)
    const uint32_t sPa = smem_u32(sP), sSa = smem_u32(sdS);
    int it = 0, g = 0, gj = 0;
    auto issue_s = [&](int t, int j) {   // R0 = Q_t K_j^T, R1 = dO_t V_j^T
      const int Nj = min(128, LK - 128 * j);
      const uint32_t idesc_s = umma_idesc_bf16(128, Nj, 0, 0);
      const uint64_t a0 = umma_desc_k_sw128(sQ + t * kTileBytes);
      const uint64_t b0 = umma_desc_k_sw128(sK + j * kTileBytes);
      const uint64_t a1 = umma_desc_k_sw128(sD + t * kTileBytes);
      const uint64_t b1 = umma_desc_k_sw128(sV + j * kTileBytes);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        umma_bf16(tmem_base + kR0, a0 + 2 * k, b0 + 2 * k, idesc_s, k != 0);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        umma_bf16(tmem_base + kR1, a1 + 2 * k, b1 + 2 * k, idesc_s, k != 0);
      umma_commit(smem_u32(s_full));
    };
    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      mbar_wait(smem_u32(ld_full), it & 1);
      tc_fence_after();
      if (elect_one()) issue_s(0, 0);
      __syncwarp();
      mbar_wait(smem_u32(dq_free), (it & 1) ^ 1);   // previous pair's dQ drained
      for (int j = 0; j < NT; ++j) {
        const int Nj = min(128, LK - 128 * j);
        const int c0 = (Nj / 2 + 15) / 16 * 16;
        mbar_wait(smem_u32(kv_free), (gj & 1) ^ 1);  // previous dV / dK drained
        for (int t = 0; t < NT; ++t, ++g) {
          const int Kt = min(128, LK - 128 * t);    // query rows of this tile that can matter
          mbar_wait(smem_u32(p_ready), g & 1);
          if (NT > 1 && j == 0 && t == 0) mbar_wait(smem_u32(ld_full1), it & 1);  // rest landed
          tc_fence_after();
          if (elect_one()) {
            // dQ_t (+)= dS K_j : K dimension = keys of the block
            for (int ks = 0; ks < Nj / 16; ++ks) {
              const uint32_t aoff = ks < c0 / 16 ? ks * 8 : c0 + (ks - c0 / 16) * 8;
              umma_bf16_ts(tmem_base + kDQ + t * 64, tmem_base + kR1 + aoff,
                           umma_desc_mn_sw128(sK + j * kTileBytes + ks * 2048, 8192, 1024), idesc_o,
                           (j | ks) != 0);
            }
            // dV_j (+)= P^T dO_t ; dK_j (+)= dS^T Q_t : K dimension = query rows of the tile
            for (int ks = 0; ks < Kt / 16; ++ks) {
              umma_bf16(tmem_base + kDV, umma_desc_mn_sw128(sPa + ks * 2048, kTileBytes, 1024),
                        umma_desc_mn_sw128(sD + t * kTileBytes + ks * 2048, 8192, 1024), idesc_t,
                        (t | ks) != 0);
              umma_bf16(tmem_base + kDK, umma_desc_mn_sw128(sSa + ks * 2048, kTileBytes, 1024),
                        umma_desc_mn_sw128(sQ + t * kTileBytes + ks * 2048, 8192, 1024), idesc_t,
                        (t | ks) != 0);
            }
            // next block's S-type products queue right behind (same in-order pipe)
            if (t + 1 < NT) issue_s(t + 1, j);
            else if (j + 1 < NT) issue_s(0, j + 1);
            if (t == NT - 1) umma_commit(smem_u32(kv_done));
            if (t == NT - 1 && j == NT - 1) {
              umma_commit(smem_u32(dq_done));
              umma_commit(smem_u32(ld_empty));
            }
          }
          __syncwarp();
        }
        ++gj;
      }
    }
  } else {
    // ------------------------------------------------------------------ element-wise + epilogues
    const int q = warp & 3;             // TMEM lane quarter
    const int hh = (warp - 2) >> 2;     // column half
    const int r = q * 32 + lane;        // row within a tile
    const int tid2 = (warp - 2) * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t tb = tmem_base + lane_sel;
    const float c2 = 0.125f * kLog2e;
    int it = 0, g = 0, gj = 0;
    int sbuf = 0;   // staging tiles alternate: a store may still be reading the other one

    // accumulator tile (this thread's row, 32 or 64 columns) -> staging -> TMA store
    auto store_tile = [&](const uint32_t (&a)[32], const uint32_t (&b)[32], bool wide, bool mine,
                          int c16_base, float sc, int col, int row0, int n) {
      uint8_t* stg = staging + sbuf * kTileBytes;
      sbuf ^= 1;
      if (tid2 == 0) tma_store_wait_read<1>();   // the store issued two tiles ago has read `stg`
      named_bar_sync(1, 256);
      if (mine) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint4*>(stg + sw128_off(r, c16_base + j)) = make_uint4(
              pack_bf16(__uint_as_float(a[8 * j]) * sc, __uint_as_float(a[8 * j + 1]) * sc),
              pack_bf16(__uint_as_float(a[8 * j + 2]) * sc, __uint_as_float(a[8 * j + 3]) * sc),
              pack_bf16(__uint_as_float(a[8 * j + 4]) * sc, __uint_as_float(a[8 * j + 5]) * sc),
              pack_bf16(__uint_as_float(a[8 * j + 6]) * sc, __uint_as_float(a[8 * j + 7]) * sc));
          if (wide)
            *reinterpret_cast<uint4*>(stg + sw128_off(r, 4 + j)) = make_uint4(
                pack_bf16(__uint_as_float(b[8 * j]) * sc, __uint_as_float(b[8 * j + 1]) * sc),
                pack_bf16(__uint_as_float(b[8 * j + 2]) * sc, __uint_as_float(b[8 * j + 3]) * sc),
                pack_bf16(__uint_as_float(b[8 * j + 4]) * sc, __uint_as_float(b[8 * j + 5]) * sc),
                pack_bf16(__uint_as_float(b[8 * j + 6]) * sc, __uint_as_float(b[8 * j + 7]) * sc));
        }
        fence_proxy_async_smem();
      }
      named_bar_sync(1, 256);
      if (tid2 == 0 && !(p.dbg & 8)) {
        tma_store_3d(&tmOut, smem_u32(stg), col, row0, n);
        tma_store_commit();
      }
    };

    for (int pr = blockIdx.x; pr < pairs; pr += gridDim.x, ++it) {
      const int n = pr / p.H, h = pr % p.H;
      const int tok0 = n * p.sn;
      // delta_q = dO_q . O_q for 128 queries at a time: warp w takes 16 rows, 8 lanes per row
      // (16 B each), O rows straight from global (coalesced), dO rows from the landed tile
      auto compute_delta = [&](int row_begin) {
        const uint8_t* sDO = smem + 3 * mat;
        const int rb = row_begin + (warp - 2) * 16;
        uint4 ov[4];
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
          const int row = rb + ps * 4 + (lane >> 3);
          ov[ps] = make_uint4(0, 0, 0, 0);
          if (row < p.L)
            ov[ps] = *reinterpret_cast<const uint4*>(
                p.o + (size_t)(tok0 + row * p.sl) * p.ld_o + h * HD + (lane & 7) * 8);
        }
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
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
      };
      sLse[tid2] = tid2 < p.L ? p.lse[(size_t)pr * p.lse_ld + tid2] * kLog2e : 0.f;
      mbar_wait(smem_u32(ld_full), it & 1);
      compute_delta(0);
      named_bar_sync(1, 256);
      for (int j = 0; j < NT; ++j) {
        const int Nj = min(128, LK - 128 * j);
        const int c0 = (Nj / 2 + 15) / 16 * 16;
        const int col_base = hh ? c0 : 0;
        const int nchunk = (hh ? Nj - c0 : c0) / 16;
        for (int t = 0; t < NT; ++t, ++g) {
          if (j == 0 && t == 1) {   // the second query tile's rows arrived on their own barrier
            mbar_wait(smem_u32(ld_full1), it & 1);
            compute_delta(128);
            named_bar_sync(1, 256);
          }
          const int qi = t * 128 + r;                    // this thread's query
          const float lse_r = sLse[qi], del_r = sDelta[qi];
          const int kmax = qi < p.L ? (p.causal ? min(p.L, qi + 1) : p.L) : 0;   // visible keys
          mbar_wait(smem_u32(s_full), g & 1);
          tc_fence_after();
          auto process = [&](const uint32_t (&sv)[16], const uint32_t (&dv)[16], int c) {
            const int col0 = col_base + c * 16;          // column inside the block
            const int key0 = j * 128 + col0;
            uint32_t wp[8], wd[8];
            if (key0 + 16 <= kmax) {
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                const float p0 = ex2(fmaf(__uint_as_float(sv[e]), c2, -lse_r));
                const float p1 = ex2(fmaf(__uint_as_float(sv[e + 1]), c2, -lse_r));
                wp[e >> 1] = pack_bf16(p0, p1);
                wd[e >> 1] = pack_bf16(p0 * (__uint_as_float(dv[e]) - del_r),
                                       p1 * (__uint_as_float(dv[e + 1]) - del_r));
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                const float p0 =
                    key0 + e < kmax ? ex2(fmaf(__uint_as_float(sv[e]), c2, -lse_r)) : 0.f;
                const float p1 =
                    key0 + e + 1 < kmax ? ex2(fmaf(__uint_as_float(sv[e + 1]), c2, -lse_r)) : 0.f;
                wp[e >> 1] = pack_bf16(p0, p1);
                wd[e >> 1] = pack_bf16(p0 * (__uint_as_float(dv[e]) - del_r),
                                       p1 * (__uint_as_float(dv[e + 1]) - del_r));
              }
            }
            tmem_st_x8(tb + kR1 + col_base + c * 8, wd);          // A operand of dQ
            // row r of the P / dS tiles: 64-key atoms of [128 rows x 128 B], TMA-style swizzle
            const uint32_t off = (uint32_t)(col0 >> 6) * kTileBytes;
            const int k8 = (col0 & 63) >> 3;
            *reinterpret_cast<uint4*>(sP + off + sw128_off(r, k8)) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
            *reinterpret_cast<uint4*>(sP + off + sw128_off(r, k8 + 1)) = make_uint4(wp[4], wp[5], wp[6], wp[7]);
            *reinterpret_cast<uint4*>(sdS + off + sw128_off(r, k8)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
            *reinterpret_cast<uint4*>(sdS + off + sw128_off(r, k8 + 1)) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
          };
          if (!(p.dbg & 1)) {
            uint32_t sA[16], dA[16], sB[16], dB[16];
            tmem_ld_x16(tb + kR0 + col_base, sA);
            tmem_ld_x16(tb + kR1 + col_base, dA);
            for (int c = 0; c < nchunk; c += 2) {
              tmem_ld_wait();
              if (c + 1 < nchunk) {
                tmem_ld_x16(tb + kR0 + col_base + (c + 1) * 16, sB);
                tmem_ld_x16(tb + kR1 + col_base + (c + 1) * 16, dB);
              }
              process(sA, dA, c);
              if (c + 1 < nchunk) {
                tmem_ld_wait();
                if (c + 2 < nchunk) {
                  tmem_ld_x16(tb + kR0 + col_base + (c + 2) * 16, sA);
                  tmem_ld_x16(tb + kR1 + col_base + (c + 2) * 16, dA);
                }
                process(sB, dB, c + 1);
              }
            }
          }
          tmem_st_wait();
          fence_proxy_async_smem();     // P / dS tiles -> visible to the tensor core
          tc_fence_before();
          mbar_arrive(smem_u32(p_ready));
        }
        // dV_j / dK_j are complete: half 0 stores dV, half 1 stores dK (scaled by hd^-0.5)
        {
          mbar_wait(smem_u32(kv_done), gj & 1);
          tc_fence_after();
          uint32_t a[32], b[32];
          const uint32_t src = tb + (hh == 0 ? kDV : kDK);
          tmem_ld_32x32(src, a);
          tmem_ld_32x32(src + 32, b);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(smem_u32(kv_free));
          store_tile(a, b, true, hh == 0, 0, 1.0f, 2 * D + h * HD, j * 128, n);
          store_tile(a, b, true, hh == 1, 0, 0.125f, D + h * HD, j * 128, n);
          ++gj;
        }
      }
      // dQ_0 / dQ_1: columns [32 hh, 32 hh + 32) of each row
      {
        mbar_wait(smem_u32(dq_done), it & 1);
        tc_fence_after();
        uint32_t a0[32], a1[32];
        tmem_ld_32x32(tb + kDQ + hh * 32, a0);
        if (NT > 1) tmem_ld_32x32(tb + kDQ + 64 + hh * 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(smem_u32(dq_free));
        store_tile(a0, a0, false, true, hh * 4, 0.125f, h * HD, 0, n);
        if (NT > 1) store_tile(a1, a1, false, true, hh * 4, 0.125f, h * HD, 128, n);
      }
      // sLse / sDelta are rewritten for the next pair only after every thread is done with them
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

int encode_rows(CUtensorMap* tm, const void* base, int cols, int ld, int L, int N, int sn, int sl,
                int box_rows) {
  return llc_encode_tmap_3d(tm, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)cols,
                            (uint64_t)L, (uint64_t)N, (uint64_t)ld * 2 * sl, (uint64_t)ld * 2 * sn,
                            HD, box_rows, 1, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace

int llc_attn_bwd_tc3(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o,
                     int ld_do, const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H,
                     int sn, int sl, int causal, cudaStream_t st, int lse_ld) {
  Bwd3Params p;
  p.o = reinterpret_cast<const __nv_bfloat16*>(o); p.ld_o = ld_o; p.lse = lse;
  p.lse_ld = lse_ld > 0 ? lse_ld : L;
  p.N = N; p.L = L; p.H = H; p.LK = (L + 15) / 16 * 16; p.NT = (L + 127) / 128;
  p.sn = sn; p.sl = sl; p.causal = causal;
  p.mat_bytes = p.LK * 128;
  static const int dbg = llc_dev_env("LLC_ATTN_DBG") ? atoi(llc_dev_env("LLC_ATTN_DBG")) : 0;
  p.dbg = dbg;
  const int smem = 4 * p.mat_bytes + 2 * kPTileBytes + kStagingBytes + 2 * 256 * 4 + 256;
  const int rows0 = p.NT > 1 ? 128 : p.LK, rows1 = p.NT > 1 ? p.LK - 128 : 16;
  CUtensorMap q0, q1, d0, d1, to;
  if (int rc = encode_rows(&q0, qkv, 3 * H * HD, ld_qkv, L, N, sn, sl, rows0)) return rc;
  if (int rc = encode_rows(&q1, qkv, 3 * H * HD, ld_qkv, L, N, sn, sl, rows1)) return rc;
  if (int rc = encode_rows(&d0, d_o, H * HD, ld_do, L, N, sn, sl, rows0)) return rc;
  if (int rc = encode_rows(&d1, d_o, H * HD, ld_do, L, N, sn, sl, rows1)) return rc;
  if (int rc = encode_rows(&to, dqkv, 3 * H * HD, ld_dqkv, L, N, sn, sl, 128)) return rc;
  LLC_CONFIGURE_SMEM(attn_bwd3_kernel, smem);
  const int grid = N * H < llc_num_sms() ? N * H : llc_num_sms();
  LLC_PROF_BEGIN(LLC_K_ATTN_BWD, N * H, L, 0, 8.0 * N * H * (double)L * L * HD,
                 16.0 * N * H * (double)L * HD, st);
  LLC_CUDA(llc_launch_pdl(attn_bwd3_kernel, dim3(grid), dim3(kThreads), (size_t)smem, st, q0, q1, d0,
                          d1, to, p));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_bwd3_kernel");
  return 0;
}
