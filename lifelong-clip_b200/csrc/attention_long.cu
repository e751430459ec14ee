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
constexpr int kThreads = 256;  // 8 warps; a warp owns a token row, lane l the head dims 2l, 2l+1
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float2 ld2(const __nv_bfloat16* p) {
  return unpack_bf16(*reinterpret_cast<const uint32_t*>(p));
}
__device__ __forceinline__ float warp_sum2(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
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

// smem: ks[r][64] | vs[r][64] | sc[L] | part[kWarps][64] | red[kWarps]
template <int R>   // R >= r: compile-time bound of the side-token count (1, 2, 4, 8)
__global__ void __launch_bounds__(kThreads)
attn_long_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, __nv_bfloat16* __restrict__ o,
                     int ld_o, float* __restrict__ lse, int L, int H, int sn, int sl) {
  extern __shared__ float sm[];
  const int r = L - L0;
  float* ks = sm;
  float* vs = ks + r * HD;
  float* sc = vs + r * HD;
  float* part = sc + L;
  float* red = part + kWarps * HD;
  pdl_wait();
  const int n = blockIdx.x / H, h = blockIdx.x % H, D = H * HD;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t tok0 = (size_t)n * sn;
  auto qrow = [&](int l) { return qkv + (tok0 + (size_t)l * sl) * ld_qkv + h * HD; };
  auto orow = [&](int l) { return o + (tok0 + (size_t)l * sl) * ld_o + h * HD; };
  float* lse_p = lse + (size_t)blockIdx.x * L;
  for (int i = tid; i < r * HD; i += kThreads) {
    const int j = i / HD, d = i % HD;
    ks[i] = __bfloat162float(qrow(L0 + j)[D + d]);
    vs[i] = __bfloat162float(qrow(L0 + j)[2 * D + d]);
  }
  __syncthreads();
  // ---- queries < 256: merge the side keys into the block kernel's (o_A, lse_A). kU rows per
  // warp iteration: their loads are all in flight before the first shuffle chain starts
  constexpr int kU = 4;
  for (int q0 = warp * kU; q0 < L0; q0 += kWarps * kU) {
    float2 qv[kU], oa[kU];
    float la[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      qv[u] = ld2(qrow(q0 + u) + 2 * lane);
      oa[u] = ld2(orow(q0 + u) + 2 * lane);
      la[u] = lse_p[q0 + u];
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      float s[R];
      float m = la[u];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        s[j] = -INFINITY;
        if (j < r) {
          s[j] = 0.125f * warp_sum2(qv[u].x * ks[j * HD + 2 * lane] +
                                    qv[u].y * ks[j * HD + 2 * lane + 1]);
          m = fmaxf(m, s[j]);
        }
      }
      const float wa = __expf(la[u] - m);
      float den = wa, a0 = wa * oa[u].x, a1 = wa * oa[u].y;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (j < r) {
          const float w = __expf(s[j] - m);
          den += w;
          a0 = fmaf(w, vs[j * HD + 2 * lane], a0);
          a1 = fmaf(w, vs[j * HD + 2 * lane + 1], a1);
        }
      }
      const float inv = 1.0f / den;
      *reinterpret_cast<uint32_t*>(orow(q0 + u) + 2 * lane) = pack_bf16(a0 * inv, a1 * inv);
      if (lane == 0) lse_p[q0 + u] = m + __logf(den);
    }
  }
  // ---- side queries: full rows over all L keys
  for (int i = 0; i < r; ++i) {
    const float2 qv = ld2(qrow(L0 + i) + 2 * lane);
    __syncthreads();
    for (int k0 = warp * kU; k0 < L; k0 += kWarps * kU) {
      float2 kv[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u)
        kv[u] = k0 + u < L ? ld2(qrow(k0 + u) + D + 2 * lane) : make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const float v = 0.125f * warp_sum2(qv.x * kv[u].x + qv.y * kv[u].y);
        if (lane == 0 && k0 + u < L) sc[k0 + u] = v;
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
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
    for (int k = warp; k < L; k += kWarps) {
      const float2 vv = ld2(qrow(k) + 2 * D + 2 * lane);
      a0 = fmaf(sc[k], vv.x, a0);
      a1 = fmaf(sc[k], vv.y, a1);
    }
    part[warp * HD + 2 * lane] = a0;
    part[warp * HD + 2 * lane + 1] = a1;
    __syncthreads();
    if (tid < HD) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) acc += part[w * HD + tid];
      orow(L0 + i)[tid] = __float2bfloat16_rn(acc / z);
    }
    if (tid == 0) lse_p[L0 + i] = m + __logf(z);
  }
}

// smem: ks | vs | qs[r][64] | gos[r][64] (dO of side queries) | dks[r][64] | dvs[r][64]
//       | dqs[r][64] | part[kWarps][3][64] | lse_s[r] | delta_s[r]
template <int R>
__global__ void __launch_bounds__(kThreads)
attn_long_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv,
                     const __nv_bfloat16* __restrict__ o, int ld_o,
                     const __nv_bfloat16* __restrict__ d_o, int ld_do, const float* __restrict__ lse,
                     __nv_bfloat16* __restrict__ dqkv, int ld_dqkv, int L, int H, int sn, int sl) {
  extern __shared__ float sm[];
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
  const size_t tok0 = (size_t)n * sn;
  auto qrow = [&](int l) { return qkv + (tok0 + (size_t)l * sl) * ld_qkv + h * HD; };
  auto grow = [&](int l) { return dqkv + (tok0 + (size_t)l * sl) * ld_dqkv + h * HD; };
  const float* lse_p = lse + (size_t)blockIdx.x * L;
  for (int i = tid; i < r * HD; i += kThreads) {
    const int j = i / HD, d = i % HD;
    ks[i] = __bfloat162float(qrow(L0 + j)[D + d]);
    vs[i] = __bfloat162float(qrow(L0 + j)[2 * D + d]);
    qs[i] = __bfloat162float(qrow(L0 + j)[d]);
    gos[i] = __bfloat162float(d_o[(tok0 + (size_t)(L0 + j) * sl) * ld_do + h * HD + d]);
    dks[i] = dvs[i] = dqs[i] = 0.f;
  }
  __syncthreads();
  if (warp < r) {   // lse and delta = dO . O of the side queries
    const int i = warp;
    const float2 ov = ld2(o + (tok0 + (size_t)(L0 + i) * sl) * ld_o + h * HD + 2 * lane);
    const float dl = warp_sum2(ov.x * gos[i * HD + 2 * lane] + ov.y * gos[i * HD + 2 * lane + 1]);
    if (lane == 0) { delta_s[i] = dl; lse_s[i] = lse_p[L0 + i]; }
  }
  __syncthreads();
  // ---- (query < 256) x (side key): dQ_q += dS k_s / 8 (row q: this warp only), dK_s, dV_s sums
  float ak[R][2], av[R][2];
#pragma unroll
  for (int j = 0; j < R; ++j) ak[j][0] = ak[j][1] = av[j][0] = av[j][1] = 0.f;
  constexpr int kU = 4;
  for (int q0 = warp * kU; q0 < L0; q0 += kWarps * kU) {
    float2 qv[kU], gv[kU], ov[kU], dq[kU];
    float lq[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int q = q0 + u;
      qv[u] = ld2(qrow(q) + 2 * lane);
      gv[u] = ld2(d_o + (tok0 + (size_t)q * sl) * ld_do + h * HD + 2 * lane);
      ov[u] = ld2(o + (tok0 + (size_t)q * sl) * ld_o + h * HD + 2 * lane);
      dq[u] = ld2(grow(q) + 2 * lane);
      lq[u] = lse_p[q];
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const float dl = warp_sum2(gv[u].x * ov[u].x + gv[u].y * ov[u].y);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        if (j < r) {
          const float k0 = ks[j * HD + 2 * lane], k1 = ks[j * HD + 2 * lane + 1];
          const float v0 = vs[j * HD + 2 * lane], v1 = vs[j * HD + 2 * lane + 1];
          const float s = 0.125f * warp_sum2(qv[u].x * k0 + qv[u].y * k1);
          const float dp = warp_sum2(gv[u].x * v0 + gv[u].y * v1);
          const float p = __expf(s - lq[u]);
          const float ds = p * (dp - dl) * 0.125f;
          dq[u].x = fmaf(ds, k0, dq[u].x);
          dq[u].y = fmaf(ds, k1, dq[u].y);
          ak[j][0] = fmaf(ds, qv[u].x, ak[j][0]);
          ak[j][1] = fmaf(ds, qv[u].y, ak[j][1]);
          av[j][0] = fmaf(p, gv[u].x, av[j][0]);
          av[j][1] = fmaf(p, gv[u].y, av[j][1]);
        }
      }
      *reinterpret_cast<uint32_t*>(grow(q0 + u) + 2 * lane) = pack_bf16(dq[u].x, dq[u].y);
    }
  }
#pragma unroll
  for (int j = 0; j < R; ++j) {   // cross-warp sums of the side keys' dK, dV (fixed order)
    if (j >= r) break;            // uniform
    part[(warp * 3 + 0) * HD + 2 * lane] = ak[j][0];
    part[(warp * 3 + 0) * HD + 2 * lane + 1] = ak[j][1];
    part[(warp * 3 + 1) * HD + 2 * lane] = av[j][0];
    part[(warp * 3 + 1) * HD + 2 * lane + 1] = av[j][1];
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
  // ---- (side query) x (every key): dK_k += dS q_i / 8, dV_k += P dO_i (row k: this warp only),
  //      dq_i = sum_k dS k_k / 8
  for (int i = 0; i < r; ++i) {
    const float q0 = qs[i * HD + 2 * lane], q1 = qs[i * HD + 2 * lane + 1];
    const float g0 = gos[i * HD + 2 * lane], g1 = gos[i * HD + 2 * lane + 1];
    const float li = lse_s[i], dl = delta_s[i];
    float a0 = 0.f, a1 = 0.f;
    for (int k0 = warp * kU; k0 < L; k0 += kWarps * kU) {
      float2 kv[kU], vv[kU], dk[kU], dv[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int k = k0 + u;
        const bool in = k < L;
        kv[u] = in ? ld2(qrow(k) + D + 2 * lane) : make_float2(0.f, 0.f);
        vv[u] = in ? ld2(qrow(k) + 2 * D + 2 * lane) : make_float2(0.f, 0.f);
        dk[u] = k < L0 ? ld2(grow(k) + D + 2 * lane) : make_float2(0.f, 0.f);
        dv[u] = k < L0 ? ld2(grow(k) + 2 * D + 2 * lane) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int k = k0 + u;
        const float s = 0.125f * warp_sum2(q0 * kv[u].x + q1 * kv[u].y);
        const float dp = warp_sum2(g0 * vv[u].x + g1 * vv[u].y);
        if (k >= L) continue;                     // warp-uniform
        const float p = __expf(s - li);
        const float ds = p * (dp - dl) * 0.125f;
        a0 = fmaf(ds, kv[u].x, a0);
        a1 = fmaf(ds, kv[u].y, a1);
        if (k < L0) {
          dk[u].x = fmaf(ds, q0, dk[u].x); dk[u].y = fmaf(ds, q1, dk[u].y);
          dv[u].x = fmaf(p, g0, dv[u].x); dv[u].y = fmaf(p, g1, dv[u].y);
          *reinterpret_cast<uint32_t*>(grow(k) + D + 2 * lane) = pack_bf16(dk[u].x, dk[u].y);
          *reinterpret_cast<uint32_t*>(grow(k) + 2 * D + 2 * lane) = pack_bf16(dv[u].x, dv[u].y);
        } else {        // a side key: exactly one warp sees (i, k), accumulate in shared memory
          const int j = k - L0;
          dks[j * HD + 2 * lane] += ds * q0; dks[j * HD + 2 * lane + 1] += ds * q1;
          dvs[j * HD + 2 * lane] += p * g0; dvs[j * HD + 2 * lane + 1] += p * g1;
        }
      }
    }
    part[(warp * 3 + 2) * HD + 2 * lane] = a0;
    part[(warp * 3 + 2) * HD + 2 * lane + 1] = a1;
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
    grow(L0 + j)[d] = __float2bfloat16_rn(dqs[i]);
    grow(L0 + j)[D + d] = __float2bfloat16_rn(dks[i]);
    grow(L0 + j)[2 * D + d] = __float2bfloat16_rn(dvs[i]);
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
  const size_t smem = (size_t)(2 * r * HD + L + kWarps * HD + kWarps) * sizeof(float);
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
