// Attention for sequences a little longer than the 256 keys the TMEM kernels hold per
// (sample, head): ViT-L/14's 257 tokens (BASELINE config 3). L = 256 + r, r <= kMaxSide.
//
// Softmax attention splits exactly over key subsets (log-sum-exp merge) and its backward is a sum
// over (query, key) blocks once the FULL row statistics lse_q and delta_q = dO_q . O_q are used:
//   forward   1. attn_fwd2_kernel on the first 256 tokens as a 256-token sequence (TMA maps with the
//                full sequence's row pitch): o_A, lse_A for queries < 256 over keys < 256;
//             2. attn_long_fwd_kernel (CUDA cores, one CTA per pair): merges the r side keys into
//                those rows (o = (w_A o_A + sum_j w_j v_j) / (w_A + sum_j w_j)) and computes the r
//                side queries over all L keys;
//   backward  1. attn_bwd3_kernel on the 256 x 256 block with the FULL lse and O (so its P and its
//                in-kernel delta are the true ones): exact dQ, dK, dV contributions of that block;
//             2. attn_long_bwd_kernel: the (query < 256, side key) and (side query, every key)
//                terms, added to the rows the block kernel wrote (each row is touched by one warp)
//                and written for the side tokens.
// The side work is 2 r / L of the score matrix: L2-bound row passes.
#include "common.cuh"

int llc_attn_fwd_tc2(const void* qkv, int ld_qkv, void* o, int ld_o, float* lse, int N, int L,
                     int H, int sn, int sl, int causal, cudaStream_t st, int lse_ld);
int llc_attn_bwd_tc3(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o,
                     int ld_do, const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H,
                     int sn, int sl, int causal, cudaStream_t st, int lse_ld);

namespace {

constexpr int HD = 64;
constexpr int L0 = 256;        // tokens handled by the TMEM kernels
constexpr int kMaxSide = 8;
constexpr int kThreads = 256;  // 8 warps
constexpr int kWarps = kThreads / 32;
// Row passes: EIGHT lanes own a token row (lane g of the group holds head dims 8g .. 8g+7: one
// 16-byte load), so one warp instruction moves four rows, a row dot is 8 local FMAs + 3 shuffles,
// and kU independent row quartets are in flight per warp. (The first version - a whole warp per
// row, 4 B per lane - was issue-bound: 41 M warp instructions for a 158 MB pass, IPC 0.5 with seven
// warps waiting on loads per issue, `profiles/r02_ncu_attn_long_v1.txt`.)
constexpr int kU = 2;
constexpr int kRowsPerIter = 4 * kU;

struct V8 {
  float v[8];
};
__device__ __forceinline__ V8 ld8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  V8 r;
  float2 f;
  f = unpack_bf16(u.x); r.v[0] = f.x; r.v[1] = f.y;
  f = unpack_bf16(u.y); r.v[2] = f.x; r.v[3] = f.y;
  f = unpack_bf16(u.z); r.v[4] = f.x; r.v[5] = f.y;
  f = unpack_bf16(u.w); r.v[6] = f.x; r.v[7] = f.y;
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const V8& a) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(a.v[0], a.v[1]), pack_bf16(a.v[2], a.v[3]),
                                            pack_bf16(a.v[4], a.v[5]), pack_bf16(a.v[6], a.v[7]));
}
__device__ __forceinline__ V8 lds8(const float* p) {   // 8 floats from shared memory
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  V8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ float dot8(const V8& a, const V8& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(a.v[i], b.v[i], s);
  return s;
}
__device__ __forceinline__ float group_sum(float v) {   // over the 8 lanes of a row group
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}
__device__ __forceinline__ float across_groups(float v) {   // over the 4 row groups of a warp
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : v + w;
  }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < kWarps; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// smem: ks[r][64] | vs[r][64] | sc[L (+pad)] | part[kWarps][64] | red[kWarps]
template <int R>   // R >= r: compile-time bound of the side-token count (1, 2, 4, 8)
__global__ void __launch_bounds__(kThreads)
attn_long_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, __nv_bfloat16* __restrict__ o,
                     int ld_o, float* __restrict__ lse, int L, int H, int sn, int sl) {
  extern __shared__ __align__(16) float sm[];
  const int r = L - L0;
  float* ks = sm;
  float* vs = ks + r * HD;
  float* sc = vs + r * HD;
  float* part = sc + ((L + 3) & ~3);
  float* red = part + kWarps * HD;
  pdl_wait();
  const int n = blockIdx.x / H, h = blockIdx.x % H, D = H * HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, g8 = (lane & 7) * 8;
  const size_t tok0 = (size_t)n * sn;
  const __nv_bfloat16* qbase = qkv + tok0 * ld_qkv + h * HD;
  __nv_bfloat16* obase = o + tok0 * ld_o + h * HD;
  const size_t qstep = (size_t)sl * ld_qkv, ostep = (size_t)sl * ld_o;
  float* lse_p = lse + (size_t)blockIdx.x * L;
  for (int i = tid; i < r * HD; i += kThreads) {
    const int j = i / HD, d = i % HD;
    ks[i] = __bfloat162float(qbase[(size_t)(L0 + j) * qstep + D + d]);
    vs[i] = __bfloat162float(qbase[(size_t)(L0 + j) * qstep + 2 * D + d]);
  }
  __syncthreads();
  // ---- queries < 256: merge the side keys into the block kernel's (o_A, lse_A)
  for (int q0 = warp * kRowsPerIter; q0 < L0; q0 += kWarps * kRowsPerIter) {
    V8 qv[kU], oa[kU];
    float la[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int q = q0 + u * 4 + grp;
      qv[u] = ld8(qbase + q * qstep + g8);
      oa[u] = ld8(obase + q * ostep + g8);
      la[u] = lse_p[q];
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int q = q0 + u * 4 + grp;
      float s[R];
      float m = la[u];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        s[j] = -INFINITY;
        if (j < r) {
          s[j] = 0.125f * group_sum(dot8(qv[u], lds8(ks + j * HD + g8)));
          m = fmaxf(m, s[j]);
        }
      }
      const float wa = __expf(la[u] - m);
      float den = wa;
      V8 acc;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] = wa * oa[u].v[i];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (j < r) {
          const float w = __expf(s[j] - m);
          den += w;
          const V8 vj = lds8(vs + j * HD + g8);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc.v[i] = fmaf(w, vj.v[i], acc.v[i]);
        }
      }
      const float inv = 1.0f / den;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] *= inv;
      st8(obase + q * ostep + g8, acc);
      if ((lane & 7) == 0) lse_p[q] = m + __logf(den);
    }
  }
  // ---- side queries: full rows over all L keys
  for (int i = 0; i < r; ++i) {
    const V8 qv = ld8(qbase + (size_t)(L0 + i) * qstep + g8);
    __syncthreads();
    for (int k0 = warp * kRowsPerIter; k0 < L; k0 += kWarps * kRowsPerIter) {
      V8 kv[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int k = k0 + u * 4 + grp;
        if (k < L) kv[u] = ld8(qbase + k * qstep + D + g8);
        else {
#pragma unroll
          for (int t = 0; t < 8; ++t) kv[u].v[t] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int k = k0 + u * 4 + grp;
        const float v = 0.125f * group_sum(dot8(qv, kv[u]));
        if ((lane & 7) == 0 && k < L) sc[k] = v;
      }
    }
    __syncthreads();
    float m = -INFINITY;
    for (int k = tid; k < L; k += kThreads) m = fmaxf(m, sc[k]);
    m = block_reduce(m, red, true);
    float z = 0.f;
    for (int k = tid; k < L; k += kThreads) {
      const float e = __expf(sc[k] - m);
      sc[k] = e;
      z += e;
    }
    z = block_reduce(z, red, false);
    V8 acc;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc.v[t] = 0.f;
    for (int k0 = warp * kRowsPerIter; k0 < L; k0 += kWarps * kRowsPerIter) {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int k = k0 + u * 4 + grp;
        if (k < L) {
          const V8 vv = ld8(qbase + k * qstep + 2 * D + g8);
          const float pk = sc[k];
#pragma unroll
          for (int t = 0; t < 8; ++t) acc.v[t] = fmaf(pk, vv.v[t], acc.v[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) acc.v[t] = across_groups(acc.v[t]);
    if (grp == 0) {
#pragma unroll
      for (int t = 0; t < 8; ++t) part[warp * HD + g8 + t] = acc.v[t];
    }
    __syncthreads();
    if (tid < HD) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) a += part[w * HD + tid];
      obase[(size_t)(L0 + i) * ostep + tid] = __float2bfloat16_rn(a / z);
    }
    if (tid == 0) lse_p[L0 + i] = m + __logf(z);
  }
}

// smem: ks | vs | qs[r][64] | gos[r][64] (dO of side queries) | dks[r][64] | dvs[r][64]
//       | dqs[r][64] | part[kWarps][3][64] | lse_s[8] | delta_s[8]
template <int R>
__global__ void __launch_bounds__(kThreads)
attn_long_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv,
                     const __nv_bfloat16* __restrict__ o, int ld_o,
                     const __nv_bfloat16* __restrict__ d_o, int ld_do, const float* __restrict__ lse,
                     __nv_bfloat16* __restrict__ dqkv, int ld_dqkv, int L, int H, int sn, int sl) {
  extern __shared__ __align__(16) float sm[];
  const int r = L - L0;
  float* ks = sm;
  float* vs = ks + r * HD;
  float* qs = vs + r * HD;
  float* gos = qs + r * HD;
  float* dks = gos + r * HD;
  float* dvs = dks + r * HD;
  float* dqs = dvs + r * HD;
  float* part = dqs + r * HD;
  float* lse_s = part + kWarps * 3 * HD;
  float* delta_s = lse_s + kMaxSide;
  pdl_wait();
  const int n = blockIdx.x / H, h = blockIdx.x % H, D = H * HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, g8 = (lane & 7) * 8;
  const size_t tok0 = (size_t)n * sn;
  const __nv_bfloat16* qbase = qkv + tok0 * ld_qkv + h * HD;
  const __nv_bfloat16* obase = o + tok0 * ld_o + h * HD;
  const __nv_bfloat16* gbase = d_o + tok0 * ld_do + h * HD;
  __nv_bfloat16* dbase = dqkv + tok0 * ld_dqkv + h * HD;
  const size_t qstep = (size_t)sl * ld_qkv, ostep = (size_t)sl * ld_o, gstep = (size_t)sl * ld_do,
               dstep = (size_t)sl * ld_dqkv;
  const float* lse_p = lse + (size_t)blockIdx.x * L;
  for (int i = tid; i < r * HD; i += kThreads) {
    const int j = i / HD, d = i % HD;
    ks[i] = __bfloat162float(qbase[(size_t)(L0 + j) * qstep + D + d]);
    vs[i] = __bfloat162float(qbase[(size_t)(L0 + j) * qstep + 2 * D + d]);
    qs[i] = __bfloat162float(qbase[(size_t)(L0 + j) * qstep + d]);
    gos[i] = __bfloat162float(gbase[(size_t)(L0 + j) * gstep + d]);
    dks[i] = dvs[i] = dqs[i] = 0.f;
  }
  __syncthreads();
  if (warp < r) {   // lse and delta = dO . O of the side queries
    const int i = warp;
    const float ov0 = __bfloat162float(obase[(size_t)(L0 + i) * ostep + 2 * lane]);
    const float ov1 = __bfloat162float(obase[(size_t)(L0 + i) * ostep + 2 * lane + 1]);
    float dl = ov0 * gos[i * HD + 2 * lane] + ov1 * gos[i * HD + 2 * lane + 1];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dl += __shfl_xor_sync(0xffffffffu, dl, off);
    if (lane == 0) { delta_s[i] = dl; lse_s[i] = lse_p[L0 + i]; }
  }
  __syncthreads();
  // ---- (query < 256) x (side key): dQ_q += dS k_s / 8 (row q: this row group only), dK_s, dV_s
  V8 ak[R], av[R];
#pragma unroll
  for (int j = 0; j < R; ++j)
#pragma unroll
    for (int t = 0; t < 8; ++t) ak[j].v[t] = av[j].v[t] = 0.f;
  for (int q0 = warp * kRowsPerIter; q0 < L0; q0 += kWarps * kRowsPerIter) {
    V8 qv[kU], gv[kU], ov[kU], dq[kU];
    float lq[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int q = q0 + u * 4 + grp;
      qv[u] = ld8(qbase + q * qstep + g8);
      gv[u] = ld8(gbase + q * gstep + g8);
      ov[u] = ld8(obase + q * ostep + g8);
      dq[u] = ld8(dbase + q * dstep + g8);
      lq[u] = lse_p[q];
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int q = q0 + u * 4 + grp;
      const float dl = group_sum(dot8(gv[u], ov[u]));
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (j < r) {
          const V8 kj = lds8(ks + j * HD + g8), vj = lds8(vs + j * HD + g8);
          const float s = 0.125f * group_sum(dot8(qv[u], kj));
          const float dp = group_sum(dot8(gv[u], vj));
          const float p = __expf(s - lq[u]);
          const float ds = p * (dp - dl) * 0.125f;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            dq[u].v[t] = fmaf(ds, kj.v[t], dq[u].v[t]);
            ak[j].v[t] = fmaf(ds, qv[u].v[t], ak[j].v[t]);
            av[j].v[t] = fmaf(p, gv[u].v[t], av[j].v[t]);
          }
        }
      }
      st8(dbase + q * dstep + g8, dq[u]);
    }
  }
#pragma unroll
  for (int j = 0; j < R; ++j) {   // sums over row groups, then over warps (fixed order)
    if (j >= r) break;            // uniform
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      ak[j].v[t] = across_groups(ak[j].v[t]);
      av[j].v[t] = across_groups(av[j].v[t]);
    }
    if (grp == 0) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        part[(warp * 3 + 0) * HD + g8 + t] = ak[j].v[t];
        part[(warp * 3 + 1) * HD + g8 + t] = av[j].v[t];
      }
    }
    __syncthreads();
    if (tid < 2 * HD) {
      const int which = tid / HD, d = tid % HD;
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) acc += part[(w * 3 + which) * HD + d];
      (which == 0 ? dks : dvs)[j * HD + d] += acc;
    }
    __syncthreads();
  }
  // ---- (side query) x (every key): dK_k += dS q_i / 8, dV_k += P dO_i (row k: this row group
  //      only), dq_i = sum_k dS k_k / 8
  for (int i = 0; i < r; ++i) {
    const V8 qi = lds8(qs + i * HD + g8), gi = lds8(gos + i * HD + g8);
    const float li = lse_s[i], dl = delta_s[i];
    V8 aq;
#pragma unroll
    for (int t = 0; t < 8; ++t) aq.v[t] = 0.f;
    for (int k0 = warp * kRowsPerIter; k0 < L; k0 += kWarps * kRowsPerIter) {
      V8 kv[kU], vv[kU], dk[kU], dv[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int k = k0 + u * 4 + grp;
        if (k < L) {
          kv[u] = ld8(qbase + k * qstep + D + g8);
          vv[u] = ld8(qbase + k * qstep + 2 * D + g8);
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t) kv[u].v[t] = vv[u].v[t] = 0.f;
        }
        if (k < L0) {
          dk[u] = ld8(dbase + k * dstep + D + g8);
          dv[u] = ld8(dbase + k * dstep + 2 * D + g8);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int k = k0 + u * 4 + grp;
        const float s = 0.125f * group_sum(dot8(qi, kv[u]));
        const float dp = group_sum(dot8(gi, vv[u]));
        if (k < L) {                                  // uniform within the row group
          const float p = __expf(s - li);
          const float ds = p * (dp - dl) * 0.125f;
#pragma unroll
          for (int t = 0; t < 8; ++t) aq.v[t] = fmaf(ds, kv[u].v[t], aq.v[t]);
          if (k < L0) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              dk[u].v[t] = fmaf(ds, qi.v[t], dk[u].v[t]);
              dv[u].v[t] = fmaf(p, gi.v[t], dv[u].v[t]);
            }
            st8(dbase + k * dstep + D + g8, dk[u]);
            st8(dbase + k * dstep + 2 * D + g8, dv[u]);
          } else {      // a side key: exactly one row group sees (i, k): accumulate in shared memory
            const int j = k - L0;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              dks[j * HD + g8 + t] += ds * qi.v[t];
              dvs[j * HD + g8 + t] += p * gi.v[t];
            }
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) aq.v[t] = across_groups(aq.v[t]);
    if (grp == 0) {
#pragma unroll
      for (int t = 0; t < 8; ++t) part[(warp * 3 + 2) * HD + g8 + t] = aq.v[t];
    }
    __syncthreads();
    if (tid < HD) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) acc += part[(w * 3 + 2) * HD + tid];
      dqs[i * HD + tid] = acc;
    }
    __syncthreads();
  }
  // ---- rows of the side tokens (nobody else writes them)
  for (int i = tid; i < r * HD; i += kThreads) {
    const int j = i / HD, d = i % HD;
    dbase[(size_t)(L0 + j) * dstep + d] = __float2bfloat16_rn(dqs[i]);
    dbase[(size_t)(L0 + j) * dstep + D + d] = __float2bfloat16_rn(dks[i]);
    dbase[(size_t)(L0 + j) * dstep + 2 * D + d] = __float2bfloat16_rn(dvs[i]);
  }
}

}  // namespace

bool llc_attn_long_eligible(int L, int causal) {
  return !causal && L > L0 && L <= L0 + kMaxSide;
}

int llc_attn_fwd_long(const void* qkv, int ld_qkv, void* o, int ld_o, float* lse, int N, int L,
                      int H, int sn, int sl, cudaStream_t st) {
  LLC_REQUIRE(lse != nullptr, "llc_attn_fwd: sequences of %d tokens need the lse buffer", L);
  if (int rc = llc_attn_fwd_tc2(qkv, ld_qkv, o, ld_o, lse, N, L0, H, sn, sl, 0, st, L)) return rc;
  const int r = L - L0;
  const size_t smem = (size_t)(2 * r * HD + ((L + 3) & ~3) + kWarps * HD + kWarps) * sizeof(float);
  LLC_PROF_BEGIN(LLC_K_ATTN_FWD, N * H, L, 2, 4.0 * N * H * (double)(2 * r) * L * HD,
                 2.0 * N * H * (double)L * HD * (3 + 2 * r), st);
#define LLC_LONG_FWD(RR)                                                                      \
  LLC_CUDA(llc_launch_pdl(attn_long_fwd_kernel<RR>, dim3(N * H), dim3(kThreads), smem, st,     \
                          reinterpret_cast<const __nv_bfloat16*>(qkv), ld_qkv,                 \
                          reinterpret_cast<__nv_bfloat16*>(o), ld_o, lse, L, H, sn, sl))
  if (r == 1) LLC_LONG_FWD(1); else if (r == 2) LLC_LONG_FWD(2);
  else if (r <= 4) LLC_LONG_FWD(4); else LLC_LONG_FWD(8);
#undef LLC_LONG_FWD
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_long_fwd_kernel");
  return 0;
}

int llc_attn_bwd_long(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o,
                      int ld_do, const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H,
                      int sn, int sl, cudaStream_t st) {
  if (int rc = llc_attn_bwd_tc3(qkv, ld_qkv, o, ld_o, d_o, ld_do, lse, dqkv, ld_dqkv, N, L0, H, sn,
                                sl, 0, st, L))
    return rc;
  const int r = L - L0;
  const size_t smem = (size_t)(7 * r * HD + kWarps * 3 * HD + 2 * kMaxSide) * sizeof(float);
  LLC_PROF_BEGIN(LLC_K_ATTN_BWD, N * H, L, 2, 8.0 * N * H * (double)(2 * r) * L * HD,
                 2.0 * N * H * (double)L * HD * (6 + 4 * r), st);
#define LLC_LONG_BWD(RR)                                                                      \
  LLC_CUDA(llc_launch_pdl(attn_long_bwd_kernel<RR>, dim3(N * H), dim3(kThreads), smem, st,     \
                          reinterpret_cast<const __nv_bfloat16*>(qkv), ld_qkv,                 \
                          reinterpret_cast<const __nv_bfloat16*>(o), ld_o,                     \
                          reinterpret_cast<const __nv_bfloat16*>(d_o), ld_do, lse,             \
                          reinterpret_cast<__nv_bfloat16*>(dqkv), ld_dqkv, L, H, sn, sl))
  if (r == 1) LLC_LONG_BWD(1); else if (r == 2) LLC_LONG_BWD(2);
  else if (r <= 4) LLC_LONG_BWD(4); else LLC_LONG_BWD(8);
#undef LLC_LONG_BWD
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_long_bwd_kernel");
  return 0;
}
