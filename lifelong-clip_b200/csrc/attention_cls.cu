// Attention of the LAST block for the class token only (query = token 0 of every sample).
// The tower returns ln_post(x[:, 0, :]) @ proj (reference models/clip/model.py:782-785): of the
// last block's output only the CLS rows are ever read, so its attention needs one query row per
// (sample, head) - q_cls K^T, softmax, P V (models/clip/lora.py:950,1043,1063,1068) - while K and V
// still come from every token. 197 x 64 MACs per pair: CUDA cores, one CTA per (sample, head),
// fp32 math on the bf16 operands, HBM/L2-bound on reading K and V once.
// Backward: dP_k = dO . V_k, dS_k = P_k (dP_k - sum_j P_j dP_j), dq = sum_k dS_k K_k / 8 (row 0 of
// dQ, every other dQ row of the pair is zero), dK_k = dS_k q / 8, dV_k = P_k dO.
#include "common.cuh"

namespace {

constexpr int HD = 64;
constexpr int kThreads = 128;

__device__ __forceinline__ float block_sum_128(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}
__device__ __forceinline__ float block_max_128(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}
// out[k] = row_k . v for the keys of one (sample, head): one warp per key, lane l owns head dims
// 2l, 2l+1 (one coalesced 128 B row per warp load), four keys in flight per warp
__device__ __forceinline__ void rows_dot(const __nv_bfloat16* base, size_t row_stride, int L,
                                         float v0, float v1, float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int U = 8;     // keys in flight per warp
  for (int k0 = warp * U; k0 < L; k0 += 4 * U) {
    float acc[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      acc[i] = 0.f;
      if (k0 + i < L) {
        const float2 f = unpack_bf16(
            *reinterpret_cast<const uint32_t*>(base + (size_t)(k0 + i) * row_stride + 2 * lane));
        acc[i] = f.x * v0 + f.y * v1;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < U; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    // every lane holds all U sums: lane i stores the i-th
    float mine = acc[0];
#pragma unroll
    for (int i = 1; i < U; ++i) mine = lane == i ? acc[i] : mine;
    if (lane < U && k0 + lane < L) out[k0 + lane] = mine;
  }
}

// smem: q[64] | red[4] | p[L]
__global__ void __launch_bounds__(kThreads)
attn_cls_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, __nv_bfloat16* __restrict__ o_cls,
                    int ld_o, float* __restrict__ p_cls, int L, int H, int sn, int sl) {
  extern __shared__ float sm[];
  float* q = sm;
  float* red = q + HD;
  float* p = red + 4;
  pdl_wait();
  const int n = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * HD, tid = threadIdx.x;
  const size_t tok0 = (size_t)n * sn;
  if (tid < HD) q[tid] = 0.125f * __bfloat162float(qkv[tok0 * ld_qkv + h * HD + tid]);
  __syncthreads();
  rows_dot(qkv + tok0 * ld_qkv + D + h * HD, (size_t)sl * ld_qkv, L, q[2 * (tid & 31)],
           q[2 * (tid & 31) + 1], p);
  __syncthreads();
  float m = -INFINITY;
  for (int k = tid; k < L; k += kThreads) m = fmaxf(m, p[k]);
  m = block_max_128(m, red);
  float z = 0.f;
  for (int k = tid; k < L; k += kThreads) {
    const float e = __expf(p[k] - m);
    p[k] = e;
    z += e;
  }
  z = block_sum_128(z, red);
  const float inv = 1.0f / z;
  for (int k = tid; k < L; k += kThreads) {
    const float v = p[k] * inv;
    p[k] = v;
    p_cls[(size_t)blockIdx.x * L + k] = v;
  }
  __syncthreads();
  // o[d] = sum_k p_k V[k, d]: lanes own two head dims (one 128 B row per warp load), the four
  // warps own every fourth key
  const int d2 = (tid & 31) * 2, kq = tid >> 5;
  float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
  for (int k = kq; k < L; k += 4) {
    const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(
        qkv + (tok0 + (size_t)k * sl) * ld_qkv + 2 * D + h * HD + d2));
    a0 = fmaf(p[k], v.x, a0);
    a1 = fmaf(p[k], v.y, a1);
  }
  __syncthreads();
  float* part = p;   // reuse: [4][64]
  part[kq * HD + d2] = a0;
  part[kq * HD + d2 + 1] = a1;
  __syncthreads();
  if (tid < HD)
    o_cls[(size_t)n * ld_o + h * HD + tid] =
        __float2bfloat16_rn(part[tid] + part[HD + tid] + part[2 * HD + tid] + part[3 * HD + tid]);
}

// smem: q[64] | go[64] | red[4] | p[L] | ds[L]  (at least 256 floats behind red for the partials)
__global__ void __launch_bounds__(kThreads)
attn_cls_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, const float* __restrict__ p_cls,
                    const __nv_bfloat16* __restrict__ d_o_cls, int ld_do,
                    __nv_bfloat16* __restrict__ dqkv, int ld_dqkv, int L, int H, int sn, int sl) {
  extern __shared__ float sm[];
  float* q = sm;
  float* go = q + HD;
  float* red = go + HD;
  float* p = red + 4;
  float* ds = p + L;
  pdl_wait();
  const int n = blockIdx.x / H, h = blockIdx.x % H;
  const int D = H * HD, tid = threadIdx.x;
  const size_t tok0 = (size_t)n * sn;
  if (tid < HD) {
    q[tid] = 0.125f * __bfloat162float(qkv[tok0 * ld_qkv + h * HD + tid]);
    go[tid] = __bfloat162float(d_o_cls[(size_t)n * ld_do + h * HD + tid]);
  }
  __syncthreads();
  rows_dot(qkv + tok0 * ld_qkv + 2 * D + h * HD, (size_t)sl * ld_qkv, L, go[2 * (tid & 31)],
           go[2 * (tid & 31) + 1], ds);
  __syncthreads();
  float dl = 0.f;
  for (int k = tid; k < L; k += kThreads) {
    const float pk = p_cls[(size_t)blockIdx.x * L + k];
    p[k] = pk;
    dl = fmaf(pk, ds[k], dl);
  }
  dl = block_sum_128(dl, red);
  for (int k = tid; k < L; k += kThreads) ds[k] = p[k] * (ds[k] - dl);
  __syncthreads();
  // dK_k = dS_k q (q already carries 1/8), dV_k = P_k dO, dQ_k = 0 for k > 0: one warp per key,
  // lane l writes head dims 2l, 2l+1 (coalesced 128 B rows)
  {
    const int lane = tid & 31, warp = tid >> 5;
    const float q0 = q[2 * lane], q1 = q[2 * lane + 1], g0 = go[2 * lane], g1 = go[2 * lane + 1];
    for (int k = warp; k < L; k += 4) {
      __nv_bfloat16* row = dqkv + (tok0 + (size_t)k * sl) * ld_dqkv + h * HD + 2 * lane;
      const float dsk = ds[k], pk = p[k];
      *reinterpret_cast<uint32_t*>(row + D) = pack_bf16(dsk * q0, dsk * q1);
      *reinterpret_cast<uint32_t*>(row + 2 * D) = pack_bf16(pk * g0, pk * g1);
      if (k > 0) *reinterpret_cast<uint32_t*>(row) = 0u;
    }
  }
  // dq[d] = sum_k dS_k K[k, d] / 8
  const int d2 = (tid & 31) * 2, kq = tid >> 5;
  float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
  for (int k = kq; k < L; k += 4) {
    const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(
        qkv + (tok0 + (size_t)k * sl) * ld_qkv + D + h * HD + d2));
    a0 = fmaf(ds[k], v.x, a0);
    a1 = fmaf(ds[k], v.y, a1);
  }
  __syncthreads();
  float* part = p;   // reuse: [4][64] (L + L >= 256 is guaranteed by the launcher's smem size)
  part[kq * HD + d2] = a0;
  part[kq * HD + d2 + 1] = a1;
  __syncthreads();
  if (tid < HD)
    dqkv[tok0 * ld_dqkv + h * HD + tid] = __float2bfloat16_rn(
        0.125f * (part[tid] + part[HD + tid] + part[2 * HD + tid] + part[3 * HD + tid]));
}

}  // namespace

// o_cls [N, ld_o] bf16 (row n = sample n), p_cls [N*H, L] fp32 softmax rows of the CLS query
int llc_attn_cls_fwd(const void* qkv, int ld_qkv, void* o_cls, int ld_o, float* p_cls, int N, int L,
                     int H, int sn, int sl, cudaStream_t st) {
  const size_t smem = (size_t)(HD + 4 + (L > 256 ? L : 256)) * sizeof(float);
  LLC_PROF_BEGIN(LLC_K_ATTN_FWD, N * H, L, 1, 4.0 * N * H * (double)L * HD,
                 4.0 * N * H * (double)L * HD, st);
  LLC_CUDA(llc_launch_pdl(attn_cls_fwd_kernel, dim3(N * H), dim3(kThreads), smem, st,
                          reinterpret_cast<const __nv_bfloat16*>(qkv), ld_qkv,
                          reinterpret_cast<__nv_bfloat16*>(o_cls), ld_o, p_cls, L, H, sn, sl));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_cls_fwd_kernel");
  return 0;
}

// dqkv: every row of every pair is written (dQ rows other than the CLS row are zero)
int llc_attn_cls_bwd(const void* qkv, int ld_qkv, const float* p_cls, const void* d_o_cls, int ld_do,
                     void* dqkv, int ld_dqkv, int N, int L, int H, int sn, int sl, cudaStream_t st) {
  const int Lp = L > 128 ? L : 128;
  const size_t smem = (size_t)(2 * HD + 4 + 2 * Lp) * sizeof(float);
  LLC_PROF_BEGIN(LLC_K_ATTN_BWD, N * H, L, 1, 8.0 * N * H * (double)L * HD,
                 10.0 * N * H * (double)L * HD, st);
  LLC_CUDA(llc_launch_pdl(attn_cls_bwd_kernel, dim3(N * H), dim3(kThreads), smem, st,
                          reinterpret_cast<const __nv_bfloat16*>(qkv), ld_qkv, p_cls,
                          reinterpret_cast<const __nv_bfloat16*>(d_o_cls), ld_do,
                          reinterpret_cast<__nv_bfloat16*>(dqkv), ld_dqkv, L, H, sn, sl));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("attn_cls_bwd_kernel");
  return 0;
}
