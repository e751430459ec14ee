// Fused attention core of the reference's lora.multi_head_attention_forward
// (models/clip/lora.py:950 q*=hd^-0.5, :1002-1006 head split n*H+h, :1043 bmm(q,k^T), :1063 softmax,
//  :1068 bmm(P,v), :1070-1071 merge heads): softmax(q k^T / sqrt(hd)) v, forward and backward.
// P [N*H, L, L] is never written; forward saves only the row log-sum-exp.
//
// Sequence lengths on this path are tiny and fixed (197 / 257 image tokens, 77 text tokens), so one
// CTA owns one (sample, head): Q, K, V (and dO) for the whole sequence sit in shared memory, each
// warp owns 16-row tiles and keeps scores in registers. Backward runs two phases over the same
// staged tiles (query-owned rows -> dQ, key-owned rows -> dK, dV), so nothing is accumulated
// across warps: no atomics, bit-deterministic.
#include <stdlib.h>

#include "common.cuh"

// tcgen05/TMEM path (attention_fwd2.cu, attention_bwd4.cu, attention_bwd3.cu)
static bool llc_attn_tc_eligible(int L) { return L <= 256; }
int llc_attn_bwd_tc4(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o,
                     int ld_do, const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H,
                     int sn, int sl, int causal, float* delta_ws, int delta_ready, cudaStream_t st);
int llc_attn_bwd_tc4_smem(int L);
int llc_attn_bwd_tc3(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o,
                     int ld_do, const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H,
                     int sn, int sl, int causal, cudaStream_t st, int lse_ld);
int llc_attn_fwd_tc2(const void* qkv, int ld_qkv, void* o, int ld_o, float* lse, int N, int L,
                     int H, int sn, int sl, int causal, cudaStream_t st, int lse_ld);
// 256 < L <= 264 (ViT-L/14: 257): TMEM kernels on the first 256 tokens + side kernels
bool llc_attn_long_eligible(int L, int causal);
int llc_attn_fwd_long(const void* qkv, int ld_qkv, void* o, int ld_o, float* lse, int N, int L,
                      int H, int sn, int sl, cudaStream_t st);
int llc_attn_bwd_long(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o,
                      int ld_do, const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H,
                      int sn, int sl, cudaStream_t st);

namespace {

constexpr int HD = 64;
constexpr int kWarps = 4;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// tile [rows][64] bf16, 128 B rows, 16 B chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}
// A fragment (16 rows x 16 k) at (row0, k0) of a [row][k] tile
__device__ __forceinline__ void load_a(uint32_t base, int row0, int k0, int lane, uint32_t (&a)[4]) {
  ldsm_x4(tile_addr(base, row0 + (lane & 7) + ((lane >> 3) & 1) * 8, (k0 >> 3) + (lane >> 4)), a);
}
// B fragments for two n8 tiles (n0..n0+15) x k16 from a [n][k] tile: {b0,b1 | b0,b1}
__device__ __forceinline__ void load_b_nk(uint32_t base, int n0, int k0, int lane,
                                          uint32_t (&b)[4]) {
  ldsm_x4(tile_addr(base, n0 + (lane & 7) + (lane >> 4) * 8, (k0 >> 3) + ((lane >> 3) & 1)), b);
}
// B fragments for two n8 tiles (n0..n0+15) x k16 from a [k][n] tile (transposing load)
__device__ __forceinline__ void load_b_kn(uint32_t base, int k0, int n0, int lane,
                                          uint32_t (&b)[4]) {
  ldsm_x4_t(tile_addr(base, k0 + (lane & 7) + ((lane >> 3) & 1) * 8, (n0 >> 3) + (lane >> 4)), b);
}

// stage one [L x 64] head slice (rows >= L zero) into a swizzled tile
template <int LP>
__device__ __forceinline__ void stage_tile(uint32_t sbase, uint8_t* sptr,
                                           const __nv_bfloat16* g, int ld, int tok0, int sl, int L,
                                           int col0, int tid, int nthreads) {
  for (int idx = tid; idx < LP * 8; idx += nthreads) {
    const int row = idx >> 3, c = idx & 7;
    const int off = row * 128 + ((c ^ (row & 7)) << 4);
    if (row < L) {
      cp_async16(sbase + off, g + (size_t)(tok0 + row * sl) * ld + col0 + c * 8);
    } else {
      *reinterpret_cast<uint4*>(sptr + off) = make_uint4(0, 0, 0, 0);
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <int LP>
__global__ void __launch_bounds__(kWarps * 32)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, __nv_bfloat16* __restrict__ o,
                int ld_o, float* __restrict__ lse, int L, int H, int sn, int sl, int causal) {
  constexpr int NT = LP / 8;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* pQ = smem;
  uint8_t* pK = smem + LP * 128;
  uint8_t* pV = smem + 2 * LP * 128;
  const uint32_t sQ = smem_u32(pQ), sK = smem_u32(pK), sV = smem_u32(pV);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * HD;
  const int tok0 = n * sn;
  stage_tile<LP>(sQ, pQ, qkv, ld_qkv, tok0, sl, L, h * HD, tid, kWarps * 32);
  stage_tile<LP>(sK, pK, qkv, ld_qkv, tok0, sl, L, D + h * HD, tid, kWarps * 32);
  stage_tile<LP>(sV, pV, qkv, ld_qkv, tok0, sl, L, 2 * D + h * HD, tid, kWarps * 32);
  cp_async_wait_all();
  __syncthreads();

  const float scale_log2 = 0.125f * kLog2e;  // hd^-0.5 = 1/8 for hd = 64
  const int g = lane >> 2, t = lane & 3;

  for (int mt = warp; mt < LP / 16; mt += kWarps) {
    const int row0 = mt * 16;
    if (row0 >= L) break;
    uint32_t qf[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) load_a(sQ, row0, kk * 16, lane, qf[kk]);
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int j2 = 0; j2 < NT / 2; ++j2) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t b[4];
        load_b_nk(sK, j2 * 16, kk * 16, lane, b);
        mma16816(s[2 * j2], qf[kk], b[0], b[1]);
        mma16816(s[2 * j2 + 1], qf[kk], b[2], b[3]);
      }
    }
    // mask padded keys (and future keys when causal), row max
    const int q0 = row0 + g, q1 = row0 + g + 8;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = j * 8 + 2 * t + e;
        const bool dead0 = key >= L || (causal && key > q0);
        const bool dead1 = key >= L || (causal && key > q1);
        if (dead0) s[j][e] = -INFINITY;
        if (dead1) s[j][2 + e] = -INFINITY;
        m0 = fmaxf(m0, s[j][e]);
        m1 = fmaxf(m1, s[j][2 + e]);
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float ms0 = m0 * scale_log2, ms1 = m1 * scale_log2;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = exp2f(s[j][0] * scale_log2 - ms0);
      s[j][1] = exp2f(s[j][1] * scale_log2 - ms0);
      s[j][2] = exp2f(s[j][2] * scale_log2 - ms1);
      s[j][3] = exp2f(s[j][3] * scale_log2 - ms1);
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    if (lse != nullptr && t == 0) {
      float* lrow = lse + (size_t)blockIdx.x * L;
      if (q0 < L) lrow[q0] = m0 * 0.125f + logf(l0);
      if (q1 < L) lrow[q1] = m1 * 0.125f + logf(l1);
    }
    // O = P V
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        uint32_t b[4];
        load_b_kn(sV, kk * 16, n2 * 16, lane, b);
        mma16816(acc[2 * n2], a, b[0], b[1]);
        mma16816(acc[2 * n2 + 1], a, b[2], b[3]);
      }
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    // stage the 16x64 output tile in this warp's (now dead) Q rows, then write 128 B rows
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = j;  // 16 B chunk index = n8 tile index
      const uint32_t a0 = tile_addr(sQ, row0 + g, c) + t * 4;
      const uint32_t a1 = tile_addr(sQ, row0 + g + 8, c) + t * 4;
      const uint32_t v0 = pack_bf16(acc[j][0] * inv0, acc[j][1] * inv0);
      const uint32_t v1 = pack_bf16(acc[j][2] * inv1, acc[j][3] * inv1);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(v0) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(v1) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int rr = row0 + pass * 4 + (lane >> 3), c = lane & 7;
      if (rr < L) {
        const uint4 v = *reinterpret_cast<const uint4*>(pQ + rr * 128 + ((c ^ (rr & 7)) << 4));
        *reinterpret_cast<uint4*>(o + (size_t)(tok0 + rr * sl) * ld_o + h * HD + c * 8) = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Backward. With P = softmax(S), S = scale q k^T, delta_q = sum_c dO[q,c] O[q,c]:
//   dP = dO V^T, dS = P o (dP - delta), dQ = scale dS K, dK = scale dS^T Q, dV = P^T dO.
// Phase 1 (query-owned 16-row tiles): S, dP by key chunks -> dQ.
// Phase 2 (key-owned 16-row tiles): S^T = K Q^T, dP^T = V dO^T by query chunks -> dV, dK.
template <int LP>
__global__ void __launch_bounds__(kWarps * 32)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv,
                const __nv_bfloat16* __restrict__ o, int ld_o,
                const __nv_bfloat16* __restrict__ d_o, int ld_do, const float* __restrict__ lse,
                __nv_bfloat16* __restrict__ dqkv, int ld_dqkv, int L, int H, int sn, int sl,
                int causal) {
  constexpr int CH = (LP % 64 == 0) ? 64 : ((LP % 32 == 0) ? 32 : 16);  // chunk of the other dim
  constexpr int CT = CH / 8;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* pQ = smem;
  uint8_t* pK = smem + LP * 128;
  uint8_t* pV = smem + 2 * LP * 128;
  uint8_t* pD = smem + 3 * LP * 128;
  float* sLse = reinterpret_cast<float*>(smem + 4 * LP * 128);  // [LP] lse * log2e
  float* sDelta = sLse + LP;                                    // [LP]
  const uint32_t sQ = smem_u32(pQ), sK = smem_u32(pK), sV = smem_u32(pV), sD = smem_u32(pD);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * HD;
  const int tok0 = n * sn;
  stage_tile<LP>(sQ, pQ, qkv, ld_qkv, tok0, sl, L, h * HD, tid, kWarps * 32);
  stage_tile<LP>(sK, pK, qkv, ld_qkv, tok0, sl, L, D + h * HD, tid, kWarps * 32);
  stage_tile<LP>(sV, pV, qkv, ld_qkv, tok0, sl, L, 2 * D + h * HD, tid, kWarps * 32);
  stage_tile<LP>(sD, pD, d_o, ld_do, tok0, sl, L, h * HD, tid, kWarps * 32);
  // delta and lse: 8 lanes per row, 16 B of O and dO each
  for (int idx = tid; idx < LP * 8; idx += kWarps * 32) {
    const int row = idx >> 3, c = idx & 7;
    float d = 0.f;
    if (row < L) {
      const size_t tok = (size_t)(tok0 + row * sl);
      const uint4 ov = *reinterpret_cast<const uint4*>(o + tok * ld_o + h * HD + c * 8);
      const uint4 dv = *reinterpret_cast<const uint4*>(d_o + tok * ld_do + h * HD + c * 8);
      const uint32_t ow[4] = {ov.x, ov.y, ov.z, ov.w}, dw[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = unpack_bf16(ow[e]), b = unpack_bf16(dw[e]);
        d += a.x * b.x + a.y * b.y;
      }
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    if (c == 0) {
      sDelta[row] = d;
      sLse[row] = (row < L) ? lse[(size_t)blockIdx.x * L + row] * kLog2e : 0.f;
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const float scale_log2 = 0.125f * kLog2e;
  const int g = lane >> 2, t = lane & 3;
  constexpr int MT = LP / 16;

  for (int job = warp; job < 2 * MT; job += kWarps) {
    if (job < MT) {
      // ---------------- phase 1: queries row0..row0+15 -> dQ
      const int row0 = job * 16;
      if (row0 >= L) continue;
      uint32_t qf[4][4], df[4][4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        load_a(sQ, row0, kk * 16, lane, qf[kk]);
        load_a(sD, row0, kk * 16, lane, df[kk]);
      }
      const int q0 = row0 + g, q1 = row0 + g + 8;
      const float lse0 = sLse[q0], lse1 = sLse[q1];
      const float dl0 = sDelta[q0], dl1 = sDelta[q1];
      float dq[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
#pragma unroll 1
      for (int k0 = 0; k0 < LP; k0 += CH) {
        if (k0 >= L) break;
        float s[CT][4], dp[CT][4];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
          dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
        }
#pragma unroll
        for (int j2 = 0; j2 < CT / 2; ++j2) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint32_t b[4];
            load_b_nk(sK, k0 + j2 * 16, kk * 16, lane, b);
            mma16816(s[2 * j2], qf[kk], b[0], b[1]);
            mma16816(s[2 * j2 + 1], qf[kk], b[2], b[3]);
            load_b_nk(sV, k0 + j2 * 16, kk * 16, lane, b);
            mma16816(dp[2 * j2], df[kk], b[0], b[1]);
            mma16816(dp[2 * j2 + 1], df[kk], b[2], b[3]);
          }
        }
        // dS = P o (dP - delta)  (scale applied once at the end)
#pragma unroll
        for (int j = 0; j < CT; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int key = k0 + j * 8 + 2 * t + e;
            const bool dead0 = key >= L || (causal && key > q0);
            const bool dead1 = key >= L || (causal && key > q1);
            const float p0 = dead0 ? 0.f : exp2f(s[j][e] * scale_log2 - lse0);
            const float p1 = dead1 ? 0.f : exp2f(s[j][2 + e] * scale_log2 - lse1);
            s[j][e] = p0 * (dp[j][e] - dl0);
            s[j][2 + e] = p1 * (dp[j][2 + e] - dl1);
          }
        }
#pragma unroll
        for (int kk = 0; kk < CT / 2; ++kk) {
          uint32_t a[4];
          a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
          a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
          a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
          for (int n2 = 0; n2 < 4; ++n2) {
            uint32_t b[4];
            load_b_kn(sK, k0 + kk * 16, n2 * 16, lane, b);
            mma16816(dq[2 * n2], a, b[0], b[1]);
            mma16816(dq[2 * n2 + 1], a, b[2], b[3]);
          }
        }
      }
      // dq rows are private to this thread quad: write 4 B pairs straight to global
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = h * HD + j * 8 + 2 * t;
        if (q0 < L)
          *reinterpret_cast<uint32_t*>(dqkv + (size_t)(tok0 + q0 * sl) * ld_dqkv + col) =
              pack_bf16(dq[j][0] * 0.125f, dq[j][1] * 0.125f);
        if (q1 < L)
          *reinterpret_cast<uint32_t*>(dqkv + (size_t)(tok0 + q1 * sl) * ld_dqkv + col) =
              pack_bf16(dq[j][2] * 0.125f, dq[j][3] * 0.125f);
      }
    } else {
      // ---------------- phase 2: keys row0..row0+15 -> dK, dV
      const int row0 = (job - MT) * 16;
      if (row0 >= L) continue;
      uint32_t kf[4][4], vf[4][4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        load_a(sK, row0, kk * 16, lane, kf[kk]);
        load_a(sV, row0, kk * 16, lane, vf[kk]);
      }
      const int k0r = row0 + g, k1r = row0 + g + 8;
      float dk[8][4], dv[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
        dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < LP; c0 += CH) {
        if (c0 >= L) break;
        float s[CT][4], dp[CT][4];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
          dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
        }
#pragma unroll
        for (int j2 = 0; j2 < CT / 2; ++j2) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint32_t b[4];
            load_b_nk(sQ, c0 + j2 * 16, kk * 16, lane, b);
            mma16816(s[2 * j2], kf[kk], b[0], b[1]);
            mma16816(s[2 * j2 + 1], kf[kk], b[2], b[3]);
            load_b_nk(sD, c0 + j2 * 16, kk * 16, lane, b);
            mma16816(dp[2 * j2], vf[kk], b[0], b[1]);
            mma16816(dp[2 * j2 + 1], vf[kk], b[2], b[3]);
          }
        }
        // p^T and dS^T; s <- P^T, dp <- dS^T
#pragma unroll
        for (int j = 0; j < CT; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int qi = c0 + j * 8 + 2 * t + e;
            const float lq = sLse[qi], dl = sDelta[qi];
            const bool dead0 = qi >= L || k0r >= L || (causal && k0r > qi);
            const bool dead1 = qi >= L || k1r >= L || (causal && k1r > qi);
            const float p0 = dead0 ? 0.f : exp2f(s[j][e] * scale_log2 - lq);
            const float p1 = dead1 ? 0.f : exp2f(s[j][2 + e] * scale_log2 - lq);
            s[j][e] = p0;
            s[j][2 + e] = p1;
            dp[j][e] = p0 * (dp[j][e] - dl);
            dp[j][2 + e] = p1 * (dp[j][2 + e] - dl);
          }
        }
#pragma unroll
        for (int kk = 0; kk < CT / 2; ++kk) {
          uint32_t ap[4], as[4];
          ap[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
          ap[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
          ap[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          ap[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
          as[0] = pack_bf16(dp[2 * kk][0], dp[2 * kk][1]);
          as[1] = pack_bf16(dp[2 * kk][2], dp[2 * kk][3]);
          as[2] = pack_bf16(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
          as[3] = pack_bf16(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
          for (int n2 = 0; n2 < 4; ++n2) {
            uint32_t b[4];
            load_b_kn(sD, c0 + kk * 16, n2 * 16, lane, b);
            mma16816(dv[2 * n2], ap, b[0], b[1]);
            mma16816(dv[2 * n2 + 1], ap, b[2], b[3]);
            load_b_kn(sQ, c0 + kk * 16, n2 * 16, lane, b);
            mma16816(dk[2 * n2], as, b[0], b[1]);
            mma16816(dk[2 * n2 + 1], as, b[2], b[3]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = h * HD + j * 8 + 2 * t;
        if (k0r < L) {
          __nv_bfloat16* r = dqkv + (size_t)(tok0 + k0r * sl) * ld_dqkv;
          *reinterpret_cast<uint32_t*>(r + D + col) =
              pack_bf16(dk[j][0] * 0.125f, dk[j][1] * 0.125f);
          *reinterpret_cast<uint32_t*>(r + 2 * D + col) = pack_bf16(dv[j][0], dv[j][1]);
        }
        if (k1r < L) {
          __nv_bfloat16* r = dqkv + (size_t)(tok0 + k1r * sl) * ld_dqkv;
          *reinterpret_cast<uint32_t*>(r + D + col) =
              pack_bf16(dk[j][2] * 0.125f, dk[j][3] * 0.125f);
          *reinterpret_cast<uint32_t*>(r + 2 * D + col) = pack_bf16(dv[j][2], dv[j][3]);
        }
      }
    }
  }
}

template <int LP>
int launch_fwd(const __nv_bfloat16* qkv, int ld_qkv, __nv_bfloat16* o, int ld_o, float* lse, int N,
               int L, int H, int sn, int sl, int causal, cudaStream_t st) {
  constexpr int smem = 3 * LP * 128;
  LLC_CONFIGURE_SMEM(attn_fwd_kernel<LP>, smem);
  LLC_PROF_BEGIN(LLC_K_ATTN_FWD, N * H, L, 0, 4.0 * N * H * (double)L * L * HD,
                 8.0 * N * H * (double)L * HD, st);
  attn_fwd_kernel<LP><<<N * H, kWarps * 32, smem, st>>>(qkv, ld_qkv, o, ld_o, lse, L, H, sn, sl,
                                                        causal);
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_fwd_kernel");
  return 0;
}

template <int LP>
int launch_bwd(const __nv_bfloat16* qkv, int ld_qkv, const __nv_bfloat16* o, int ld_o,
               const __nv_bfloat16* d_o, int ld_do, const float* lse, __nv_bfloat16* dqkv,
               int ld_dqkv, int N, int L, int H, int sn, int sl, int causal, cudaStream_t st) {
  constexpr int smem = 4 * LP * 128 + 2 * LP * 4;
  LLC_CONFIGURE_SMEM(attn_bwd_kernel<LP>, smem);
  LLC_PROF_BEGIN(LLC_K_ATTN_BWD, N * H, L, 0, 8.0 * N * H * (double)L * L * HD,
                 16.0 * N * H * (double)L * HD, st);
  attn_bwd_kernel<LP><<<N * H, kWarps * 32, smem, st>>>(qkv, ld_qkv, o, ld_o, d_o, ld_do, lse,
                                                        dqkv, ld_dqkv, L, H, sn, sl, causal);
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_bwd_kernel");
  return 0;
}

#define DISPATCH_LP(L, CALL)                                                   \
  do {                                                                         \
    const int lp_ = ((L) + 15) / 16 * 16;                                      \
    if (lp_ <= 16) { constexpr int LP = 16; return CALL; }                     \
    if (lp_ <= 32) { constexpr int LP = 32; return CALL; }                     \
    if (lp_ <= 64) { constexpr int LP = 64; return CALL; }                     \
    if (lp_ <= 80) { constexpr int LP = 80; return CALL; }                     \
    if (lp_ <= 208) { constexpr int LP = 208; return CALL; }                   \
    if (lp_ <= 272) { constexpr int LP = 272; return CALL; }                   \
    llc_set_error("attention: sequence length %d unsupported (max 272)", (L)); \
    return LLC_ERR_ARG;                                                        \
  } while (0)

int check_common(const void* qkv, int ld_qkv, int N, int L, int H, const char* who) {
  LLC_REQUIRE(qkv && N > 0 && L > 0 && H > 0, "%s: empty problem", who);
  LLC_REQUIRE(ld_qkv % 8 == 0 && ld_qkv >= 3 * H * HD, "%s: ld_qkv=%d too small / unaligned", who,
              ld_qkv);
  LLC_REQUIRE(((uintptr_t)qkv & 15) == 0, "%s: qkv must be 16-byte aligned", who);
  return 0;
}

}  // namespace

extern "C" int llc_attn_fwd(const void* qkv, int ld_qkv, void* o, int ld_o, float* lse, int N,
                            int L, int H, int tok_stride_n, int tok_stride_l, int causal,
                            void* stream) {
  if (int rc = check_common(qkv, ld_qkv, N, L, H, "llc_attn_fwd")) return rc;
  LLC_REQUIRE(o && ld_o % 8 == 0 && ld_o >= H * HD && ((uintptr_t)o & 15) == 0,
              "llc_attn_fwd: bad output");
  // sequences up to 256 tokens run on the tensor cores through TMEM; LLC_ATTN_LEGACY=1 keeps the
  // mma.sync kernels (development A/B only)
  static const bool legacy = llc_dev_env("LLC_ATTN_LEGACY") != nullptr;
  if (!legacy && llc_attn_tc_eligible(L))
    return llc_attn_fwd_tc2(qkv, ld_qkv, o, ld_o, lse, N, L, H, tok_stride_n, tok_stride_l, causal,
                            (cudaStream_t)stream, 0);
  if (!legacy && llc_attn_long_eligible(L, causal))
    return llc_attn_fwd_long(qkv, ld_qkv, o, ld_o, lse, N, L, H, tok_stride_n, tok_stride_l,
                             (cudaStream_t)stream);
  DISPATCH_LP(L, (launch_fwd<LP>((const __nv_bfloat16*)qkv, ld_qkv, (__nv_bfloat16*)o, ld_o, lse,
                                 N, L, H, tok_stride_n, tok_stride_l, causal,
                                 (cudaStream_t)stream)));
}

// true when llc_attn_bwd_ws runs the unit-pipelined kernel for this L (the one that reads delta)
bool llc_attn_bwd_uses_delta(int L) {
  static const bool other = llc_dev_env("LLC_ATTN_LEGACY") || llc_dev_env("LLC_ATTN_BWD3");
  return !other && llc_attn_tc_eligible(L) && llc_attn_bwd_tc4_smem(L) <= 227 * 1024;
}

// llc_attn_bwd with the state of the caller-provided delta = rowsum(dO o O) scratch ([N*H*L]
// floats) made explicit: delta_ready = 1 when the caller already filled it (llc_colsum_tc_delta)
int llc_attn_bwd_ws(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o,
                    int ld_do, const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H,
                    int tok_stride_n, int tok_stride_l, int causal, float* delta_ws, int delta_ready,
                    void* stream) {
  if (int rc = check_common(qkv, ld_qkv, N, L, H, "llc_attn_bwd")) return rc;
  LLC_REQUIRE(o && d_o && lse && dqkv, "llc_attn_bwd: null pointer");
  LLC_REQUIRE(ld_o % 8 == 0 && ld_do % 8 == 0 && ld_dqkv % 2 == 0 && ld_dqkv >= 3 * H * HD,
              "llc_attn_bwd: bad leading dimension");
  LLC_REQUIRE((((uintptr_t)o | (uintptr_t)d_o) & 15) == 0 && ((uintptr_t)dqkv & 3) == 0,
              "llc_attn_bwd: misaligned pointer");
  static const bool legacy = llc_dev_env("LLC_ATTN_LEGACY") != nullptr;
  static const bool v3 = llc_dev_env("LLC_ATTN_BWD3") != nullptr;   // block-structured kernel (A/B)
  if (!legacy && !v3 && llc_attn_tc_eligible(L) && ld_dqkv % 8 == 0 &&
      ((uintptr_t)dqkv & 15) == 0 && llc_attn_bwd_tc4_smem(L) <= 227 * 1024)
    return llc_attn_bwd_tc4(qkv, ld_qkv, o, ld_o, d_o, ld_do, lse, dqkv, ld_dqkv, N, L, H,
                            tok_stride_n, tok_stride_l, causal, delta_ws, delta_ready,
                            (cudaStream_t)stream);
  if (!legacy && llc_attn_tc_eligible(L) && ld_dqkv % 8 == 0 && ((uintptr_t)dqkv & 15) == 0)
    return llc_attn_bwd_tc3(qkv, ld_qkv, o, ld_o, d_o, ld_do, lse, dqkv, ld_dqkv, N, L, H,
                            tok_stride_n, tok_stride_l, causal, (cudaStream_t)stream, 0);
  if (!legacy && llc_attn_long_eligible(L, causal) && ld_dqkv % 8 == 0 &&
      ((uintptr_t)dqkv & 15) == 0)
    return llc_attn_bwd_long(qkv, ld_qkv, o, ld_o, d_o, ld_do, lse, dqkv, ld_dqkv, N, L, H,
                             tok_stride_n, tok_stride_l, (cudaStream_t)stream);
  DISPATCH_LP(L, (launch_bwd<LP>((const __nv_bfloat16*)qkv, ld_qkv, (const __nv_bfloat16*)o, ld_o,
                                 (const __nv_bfloat16*)d_o, ld_do, lse, (__nv_bfloat16*)dqkv,
                                 ld_dqkv, N, L, H, tok_stride_n, tok_stride_l, causal,
                                 (cudaStream_t)stream)));
}

extern "C" int llc_attn_bwd(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o,
                            int ld_do, const float* lse, void* dqkv, int ld_dqkv, int N, int L,
                            int H, int tok_stride_n, int tok_stride_l, int causal, float* delta,
                            void* stream) {
  LLC_REQUIRE(delta, "llc_attn_bwd: delta scratch ([N*H*L] floats) is required");
  return llc_attn_bwd_ws(qkv, ld_qkv, o, ld_o, d_o, ld_do, lse, dqkv, ld_dqkv, N, L, H,
                         tok_stride_n, tok_stride_l, causal, delta, 0, stream);
}
