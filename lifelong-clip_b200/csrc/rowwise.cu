// HBM-bound row-wise kernels of the ViT block: LayerNorm forward/backward (with the LoRA rank-r
// side products folded in), patch im2col, class-token/pos-emb/ln_pre assembly, LoRA side
// reductions, weight packing and the fused AdamW on the flat LoRA buffer.
// All are one-warp-per-row, 16-byte vectorised, shuffle-reduced; no shared-memory round trips
// except where a [C, r] factor is reused by every row.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr float kLnEps = 1e-5f;  // nn.LayerNorm default, reference models/clip/model.py:194-200
constexpr int kMaxR = 8;

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x-mean)*rstd*g + b -> bf16; optional u = y . A^T appended at column D.
// NV = D/128 float4 per lane.
template <int NV>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ gamma,
              const float* __restrict__ beta, int T, __nv_bfloat16* __restrict__ y, int ld_y,
              const float* __restrict__ lora_A, int r, int rev) {
  constexpr int D = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int row = blockIdx.x * 8 + warp;
  pdl_wait();
  if (row >= T) return;
  // rev: walk the rows from the END. The producer (a GEMM walking its m-tiles upwards) wrote the
  // last rows last: they are the ones still in L2 (x is larger than L2), and this kernel's own
  // output then has its FIRST rows freshest for the GEMM that consumes it upwards.
  if (rev) row = T - 1 - row;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ld_x);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[lane + 32 * i];
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + kLnEps);
  float u[kMaxR];
#pragma unroll
  for (int j = 0; j < kMaxR; ++j) u[j] = 0.f;
  uint2* yr = reinterpret_cast<uint2*>(y + (size_t)row * ld_y);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 o;
    o.x = v[i].x * rstd * g.x + b.x;
    o.y = v[i].y * rstd * g.y + b.y;
    o.z = v[i].z * rstd * g.z + b.z;
    o.w = v[i].w * rstd * g.w + b.w;
    yr[c4] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    if (lora_A != nullptr) {
#pragma unroll
      for (int j = 0; j < kMaxR; ++j) {
        if (j < r) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(lora_A + (size_t)j * D) + c4);
          u[j] += o.x * a.x + o.y * a.y + o.z * a.z + o.w * a.w;
        }
      }
    }
  }
  if (lora_A != nullptr) {
#pragma unroll
    for (int j = 0; j < kMaxR; ++j) u[j] = warp_sum(u[j]);
    if (lane < LLC_LORA_PAD / 2) {
      // lane l writes padded columns 2l, 2l+1
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxR; ++j) {
        if (j == 2 * lane && j < r) a = u[j];
        if (j == 2 * lane + 1 && j < r) b = u[j];
      }
      reinterpret_cast<uint32_t*>(y + (size_t)row * ld_y + D)[lane] = pack_bf16(a, b);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward w.r.t. the input only (affine parameters are frozen):
//   xh = (x-mean)*rstd, g = dy*gamma, dx = rstd*(g - mean(g) - xh*mean(g*xh))
// out = dx_in + dx -> fp32 (+ bf16 copy, + du = scale * out . B appended at column D of the copy)
template <int NV>
__global__ void __launch_bounds__(256, NV <= 6 ? 3 : 2)
ln_bwd_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ gamma,
              const __nv_bfloat16* __restrict__ dy, int ld_dy, const float* dx_in, float* dx_out,
              int T, __nv_bfloat16* __restrict__ dxb, int ld_dxb,
              const float* __restrict__ lora_B, int r, float scale, int rev) {
  constexpr int D = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int row = blockIdx.x * 8 + warp;
  pdl_wait();
  if (row >= T) return;
  if (rev) row = T - 1 - row;   // see ln_fwd_kernel
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ld_x);
  const uint2* dyr = reinterpret_cast<const uint2*>(dy + (size_t)row * ld_dy);
  const float4* dir = dx_in ? reinterpret_cast<const float4*>(dx_in + (size_t)row * D) : nullptr;
  // every global input of the row is requested before the first reduction: ~2.3 KB in flight
  // per warp instead of three dependent load phases
  float4 v[NV], g[NV], pin[NV];
  uint2 dq[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
#pragma unroll
  for (int i = 0; i < NV; ++i) dq[i] = dyr[lane + 32 * i];
#pragma unroll
  for (int i = 0; i < NV; ++i)
    pin[i] = dir ? dir[lane + 32 * i] : make_float4(0.f, 0.f, 0.f, 0.f);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + kLnEps);
  float c1 = 0.f, c2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float2 d01 = unpack_bf16(dq[i].x), d23 = unpack_bf16(dq[i].y);
    v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xh
    g[i].x = d01.x * gm.x; g[i].y = d01.y * gm.y; g[i].z = d23.x * gm.z; g[i].w = d23.y * gm.w;
    c1 += g[i].x + g[i].y + g[i].z + g[i].w;
    c2 += g[i].x * v[i].x + g[i].y * v[i].y + g[i].z * v[i].z + g[i].w * v[i].w;
  }
  c1 = warp_sum(c1) * (1.0f / D);
  c2 = warp_sum(c2) * (1.0f / D);
  float du[kMaxR];
#pragma unroll
  for (int j = 0; j < kMaxR; ++j) du[j] = 0.f;
  float4* dor = reinterpret_cast<float4*>(dx_out + (size_t)row * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    float4 o;
    o.x = rstd * (g[i].x - c1 - v[i].x * c2) + pin[i].x;
    o.y = rstd * (g[i].y - c1 - v[i].y * c2) + pin[i].y;
    o.z = rstd * (g[i].z - c1 - v[i].z * c2) + pin[i].z;
    o.w = rstd * (g[i].w - c1 - v[i].w * c2) + pin[i].w;
    dor[c4] = o;
    if (dxb != nullptr)
      reinterpret_cast<uint2*>(dxb + (size_t)row * ld_dxb)[c4] =
          make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    if (lora_B != nullptr) {
      const int c = c4 * 4;
      if (r == 4) {  // B rows of 4 floats: one 16-byte load per column
        const float4* bp = reinterpret_cast<const float4*>(lora_B) + c;
        const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1), b2 = __ldg(bp + 2), b3 = __ldg(bp + 3);
        du[0] += o.x * b0.x + o.y * b1.x + o.z * b2.x + o.w * b3.x;
        du[1] += o.x * b0.y + o.y * b1.y + o.z * b2.y + o.w * b3.y;
        du[2] += o.x * b0.z + o.y * b1.z + o.z * b2.z + o.w * b3.z;
        du[3] += o.x * b0.w + o.y * b1.w + o.z * b2.w + o.w * b3.w;
      } else {
#pragma unroll
        for (int j = 0; j < kMaxR; ++j) {
          if (j < r) {
            du[j] += o.x * __ldg(lora_B + (size_t)(c + 0) * r + j) +
                     o.y * __ldg(lora_B + (size_t)(c + 1) * r + j) +
                     o.z * __ldg(lora_B + (size_t)(c + 2) * r + j) +
                     o.w * __ldg(lora_B + (size_t)(c + 3) * r + j);
          }
        }
      }
    }
  }
  if (lora_B != nullptr && dxb != nullptr) {
#pragma unroll
    for (int j = 0; j < kMaxR; ++j) du[j] = warp_sum(du[j]) * scale;
    if (lane < LLC_LORA_PAD / 2) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxR; ++j) {
        if (j == 2 * lane && j < r) a = du[j];
        if (j == 2 * lane + 1 && j < r) b = du[j];
      }
      reinterpret_cast<uint32_t*>(dxb + (size_t)row * ld_dxb + D)[lane] = pack_bf16(a, b);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LoRA side reductions over X bf16 [T, ld_x] (C columns), see llc.h. Two streaming kernels, both
// one-warp-per-row with 16-byte loads (a warp reads 512 B of one row per instruction):
//   rowdot  u[t, j] = scale * sum_c X[t, c] * M[c, j]      -> bf16 into X's 16 pad columns
//   colsum  P[c, j] = sum_t X[t, c] * w[t, j]              -> per-CTA partials, reduced in fixed
//                                                             order by colsum_finish (deterministic)
constexpr int kSideThreads = 256;
}  // namespace
// tensor-core path for the column sums (lora_tc.cu)
bool llc_colsum_tc_eligible(const void* X, int ld_x, int T, int C, const void* w, int ld_w);
int llc_colsum_tc(const void* X, int ld_x, int T, int C, int R, const void* w, int ld_w,
                  float* partial, int* n_partials, cudaStream_t st);
namespace {
constexpr int kSideGroup = 256;  // columns per colsum CTA: one 16 B vector per lane

template <int R>
__global__ void __launch_bounds__(kSideThreads)
lora_rowdot_kernel(__nv_bfloat16* X, int ld_x, int T, int C, int r, const float* __restrict__ Mrd,
                   int rd_sc, int rd_sj, float rd_scale) {
  // factor staged as sM[(e * R + j) * nvec + v] (e = column within the 8-wide vector v): lanes
  // read consecutive words -> no bank conflicts
  extern __shared__ float sM[];
  const int nvec = C / 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < C * R; i += kSideThreads) {
    const int c = i / R, j = i % R;
    const float m = (j < r) ? Mrd[(size_t)c * rd_sc + (size_t)j * rd_sj] * rd_scale : 0.f;
    sM[((c & 7) * R + j) * nvec + (c >> 3)] = m;
  }
  __syncthreads();
  const int gw = blockIdx.x * (kSideThreads / 32) + warp;
  const int nw = gridDim.x * (kSideThreads / 32);
  constexpr int U = 4;  // rows per iteration: every staged factor word is used U times
  for (int t0 = gw * U; t0 < T; t0 += nw * U) {
    float d[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < R; ++j) d[u][j] = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      uint4 pk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        pk[u] = make_uint4(0, 0, 0, 0);
        if (t0 + u < T) pk[u] = reinterpret_cast<const uint4*>(X + (size_t)(t0 + u) * ld_x)[v];
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float m0[R], m1[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          m0[j] = sM[((2 * e) * R + j) * nvec + v];
          m1[j] = sM[((2 * e + 1) * R + j) * nvec + v];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t w = e == 0 ? pk[u].x : e == 1 ? pk[u].y : e == 2 ? pk[u].z : pk[u].w;
          const float2 f = unpack_bf16(w);
#pragma unroll
          for (int j = 0; j < R; ++j) d[u][j] += f.x * m0[j] + f.y * m1[j];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int j = 0; j < R; ++j) d[u][j] = warp_sum(d[u][j]);
      if (lane < LLC_LORA_PAD / 2 && t0 + u < T) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          if (j == 2 * lane) a = d[u][j];
          if (j == 2 * lane + 1) b = d[u][j];
        }
        reinterpret_cast<uint32_t*>(X + (size_t)(t0 + u) * ld_x + C)[lane] = pack_bf16(a, b);
      }
    }
  }
}

template <int R>
__global__ void __launch_bounds__(kSideThreads)
lora_colsum_kernel(const __nv_bfloat16* __restrict__ X, int ld_x, int T, int C, int r,
                   const __nv_bfloat16* __restrict__ w, int ld_w, int w_vec,
                   float* __restrict__ partial) {
  __shared__ float red[8 * R * 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int col = blockIdx.y * kSideGroup + lane * 8;  // this lane's 8 columns
  const bool live = col < C;
  float acc[8][R];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[e][j] = 0.f;
  const int gw = blockIdx.x * (kSideThreads / 32) + warp;
  const int nw = gridDim.x * (kSideThreads / 32);
  constexpr int U = 4;  // rows in flight per warp
  for (int t0 = gw; t0 < T; t0 += nw * U) {
    uint4 xv[U];
    float wj[U][R];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * nw;
      xv[u] = make_uint4(0, 0, 0, 0);
      if (t < T && live) xv[u] = *reinterpret_cast<const uint4*>(X + (size_t)t * ld_x + col);
      if (R == 4 && w_vec) {  // the 4 weights of a row in one 8-byte broadcast load
        uint2 q = make_uint2(0, 0);
        if (t < T) q = *reinterpret_cast<const uint2*>(w + (size_t)t * ld_w);
        const float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y);
        wj[u][0] = a.x; wj[u][1] = r > 1 ? a.y : 0.f;
        wj[u][2] = r > 2 ? b.x : 0.f; wj[u][3] = r > 3 ? b.y : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j)
          wj[u][j] = (t < T && j < r) ? __bfloat162float(w[(size_t)t * ld_w + j]) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t pw[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(pw[e]);
#pragma unroll
        for (int j = 0; j < R; ++j) {
          acc[2 * e][j] += f.x * wj[u][j];
          acc[2 * e + 1][j] += f.y * wj[u][j];
        }
      }
    }
  }
  // fixed-order cross-warp sum: red[(e * R + j) * 32 + lane]
  for (int wv = 0; wv < kSideThreads / 32; ++wv) {
    if (warp == wv) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          float* p = &red[(e * R + j) * 32 + lane];
          *p = (wv == 0) ? acc[e][j] : *p + acc[e][j];
        }
    }
    __syncthreads();
  }
  // partial[blockIdx.x][c][j], c = blockIdx.y * 256 + l * 8 + e
  float* out = partial + ((size_t)blockIdx.x * C + (size_t)blockIdx.y * kSideGroup) * R;
  for (int i = tid; i < kSideGroup * R; i += kSideThreads) {
    const int cl = i / R, j = i % R;
    if (blockIdx.y * kSideGroup + cl < C) out[i] = red[((cl & 7) * R + j) * 32 + (cl >> 3)];
  }
}

// out[c, j] = scale * sum_p partial[p][c][j]: 8 slices of the partials per output, summed in a
// fixed order (bit-deterministic)
__global__ void __launch_bounds__(256)
colsum_finish_kernel(const float* __restrict__ partial, int n_partials, int C, int R, int r,
                     float scale, float* __restrict__ out, int o_sc, int o_sj) {
  __shared__ float red[8][32];
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + o;
  const int n = C * R;
  float s = 0.f;
  if (i < n) {
#pragma unroll 4
    for (int p = sl; p < n_partials; p += 8) s += partial[(size_t)p * n + i];
  }
  red[sl][o] = s;
  __syncthreads();
  if (sl == 0 && i < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][o];
    const int c = i / R, j = i % R;
    if (j < r) out[(size_t)c * o_sc + (size_t)j * o_sj] = t * scale;
  }
}

constexpr int kMaxFinishJobs = 8;
struct FinishArgs {
  llc_finish_job j[kMaxFinishJobs];
};
// blockIdx.y = job; same fixed-order reduction as colsum_finish_kernel
__global__ void __launch_bounds__(1024)
colsum_finish_multi_kernel(const __grid_constant__ FinishArgs a, int R, int r) {
  // 32 slices of the partial list per output element (with one partial per SM from the fused
  // tensor-core pass every thread has <= 5 independent loads, one global round trip), four
  // consecutive elements per thread (16 B loads): a block reduces 128 elements
  __shared__ float4 red[32][32];
  const llc_finish_job& jb = a.j[blockIdx.y];
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int n = jb.C * R;                      // multiple of 4 (checked by the launcher)
  for (int base = blockIdx.x * 128; base < n; base += gridDim.x * 128) {
    const int i = base + 4 * o;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
#pragma unroll 8
      for (int p = sl; p < jb.n_partials; p += 32) {
        const float4 v = *reinterpret_cast<const float4*>(jb.partial + (size_t)p * n + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
    }
    red[sl][o] = s;
    __syncthreads();
    if (sl < 4 && i < n) {        // warp sl finishes component sl of every element group
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float4 v = red[k][o];
        t += sl == 0 ? v.x : sl == 1 ? v.y : sl == 2 ? v.z : v.w;
      }
      const int e = i + sl;
      const int c = e / R, j = e % R;
      if (j < r) jb.out[(size_t)c * jb.o_sc + (size_t)j * jb.o_sj] = t * jb.scale;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ src, int rows, int cols,
                                   int transpose, __nv_bfloat16* __restrict__ dst, int ld_dst) {
  // dst[i, j], i < rows, j < cols; src is [rows, cols] or (transpose) [cols, rows]
  __shared__ float tile[32][33];
  const int bi = blockIdx.y * 32, bj = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (!transpose) {
    for (int k = ty; k < 32; k += 8) {
      const int i = bi + k, j = bj + tx;
      if (i < rows && j < cols)
        dst[(size_t)i * ld_dst + j] = __float2bfloat16_rn(src[(size_t)i * cols + j]);
    }
  } else {
    for (int k = ty; k < 32; k += 8) {
      const int j = bj + k, i = bi + tx;  // read src[j, i] coalesced along i
      if (i < rows && j < cols) tile[k][tx] = src[(size_t)j * rows + i];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
      const int i = bi + k, j = bj + tx;
      if (i < rows && j < cols) dst[(size_t)i * ld_dst + j] = __float2bfloat16_rn(tile[tx][k]);
    }
  }
}

__global__ void pack_lora_cols_kernel(const float* __restrict__ src, int rows, int r, int s_i,
                                      int s_j, float scale, __nv_bfloat16* __restrict__ dst,
                                      int ld_dst, int col0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * LLC_LORA_PAD) return;
  const int i = idx / LLC_LORA_PAD, j = idx % LLC_LORA_PAD;
  const float v = (j < r) ? scale * src[(size_t)i * s_i + (size_t)j * s_j] : 0.f;
  dst[(size_t)i * ld_dst + col0 + j] = __float2bfloat16_rn(v);
}

// dst[j, c] = scale * src[j * s_j + c * s_c] for j < r, 0 for r <= j < 16 (bf16 [16, ld_dst])
__global__ void pack_factor_rows_kernel(const float* __restrict__ src, int r, int cols, int s_j,
                                        int s_c, float scale, __nv_bfloat16* __restrict__ dst,
                                        int ld_dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= LLC_LORA_PAD * cols) return;
  const int j = idx / cols, c = idx % cols;
  const float v = (j < r) ? scale * src[(size_t)j * s_j + (size_t)c * s_c] : 0.f;
  dst[(size_t)j * ld_dst + c] = __float2bfloat16_rn(v);
}

// Per-step refresh of every LoRA-dependent operand of the tower in ONE launch (it used to be six
// tiny kernels per layer): the 16 LoRA K-columns of the four augmented weights and the two
// [16, K] row-product factors, for all layers. blockIdx.y = layer.
struct RefreshLayer {
  const float *in_A, *in_B, *out_A, *out_B;
  __nv_bfloat16 *wqkv_aug, *wo_aug, *wqkvT_aug, *woT_aug, *f_out_A, *f_in_B, *f_out_B;
};
constexpr int kMaxRefreshLayers = 32;
struct RefreshArgs {
  RefreshLayer l[kMaxRefreshLayers];
};

__global__ void __launch_bounds__(256)
refresh_lora_kernel(const __grid_constant__ RefreshArgs args, int D, int r, float sc) {
  const RefreshLayer& y = args.l[blockIdx.y];
  const int DA = D + LLC_LORA_LD, QA = 3 * D + LLC_LORA_LD;
  const int P = LLC_LORA_PAD;
  // task sizes (elements): [W_in | s B_in] 3D*16, [W_o | s B_o] D*16, [W_in^T | A_in^T] D*16,
  // [W_o^T | A_o^T] D*16, F_oA 16*D, F_inB 16*3D
  const int n0 = 3 * D * P, n1 = D * P, n2 = D * P, n3 = D * P, n4 = P * D, n5 = P * 3 * D;
  const int n6 = P * D;
  const int total = n0 + n1 + n2 + n3 + n4 + n5 + n6;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int i = idx;
    if (i < n0) {            // wqkv_aug[c, D + j] = s * in_B[c, j]
      const int c = i / P, j = i % P;
      y.wqkv_aug[(size_t)c * DA + D + j] = __float2bfloat16_rn(j < r ? sc * y.in_B[c * r + j] : 0.f);
      continue;
    }
    i -= n0;
    if (i < n1) {            // wo_aug[c, D + j] = s * out_B[c, j]
      const int c = i / P, j = i % P;
      y.wo_aug[(size_t)c * DA + D + j] = __float2bfloat16_rn(j < r ? sc * y.out_B[c * r + j] : 0.f);
      continue;
    }
    i -= n1;
    if (i < n2) {            // wqkvT_aug[c, 3D + j] = in_A[j, c]
      const int c = i / P, j = i % P;
      y.wqkvT_aug[(size_t)c * QA + 3 * D + j] = __float2bfloat16_rn(j < r ? y.in_A[j * D + c] : 0.f);
      continue;
    }
    i -= n2;
    if (i < n3) {            // woT_aug[c, D + j] = out_A[j, c]
      const int c = i / P, j = i % P;
      y.woT_aug[(size_t)c * DA + D + j] = __float2bfloat16_rn(j < r ? y.out_A[j * D + c] : 0.f);
      continue;
    }
    i -= n3;
    if (i < n4) {            // f_out_A[j, c] = out_A[j, c]
      const int j = i / D, c = i % D;
      y.f_out_A[(size_t)j * D + c] = __float2bfloat16_rn(j < r ? y.out_A[j * D + c] : 0.f);
      continue;
    }
    i -= n4;
    if (i < n5) {            // f_in_B[j, c] = s * in_B[c, j]
      const int j = i / (3 * D), c = i % (3 * D);
      y.f_in_B[(size_t)j * 3 * D + c] = __float2bfloat16_rn(j < r ? sc * y.in_B[c * r + j] : 0.f);
      continue;
    }
    i -= n5;
    {                        // f_out_B[j, c] = s * out_B[c, j]
      const int j = i / D, c = i % D;
      y.f_out_B[(size_t)j * D + c] = __float2bfloat16_rn(j < r ? sc * y.out_B[c * r + j] : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// im2col of the stride-P patch conv: out[(n*G+py)*G+px, c*P*P + i*P + j] = img[n,c,py*P+i,px*P+j]
// (two pixels per thread: P is even, so a pair never straddles a patch). Columns
// [C*P*P, ld_out) are zero so that K can be padded to a multiple of 16 (ViT-L/14: 588 -> 592).
__global__ void patchify_kernel(const float* __restrict__ img, int N, int C, int HW, int P,
                                __nv_bfloat16* __restrict__ out, int ld_out) {
  const int G = HW / P;
  const size_t total = (size_t)N * C * HW * (HW / 2);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int x2 = (int)(idx % (HW / 2));
    size_t rest = idx / (HW / 2);
    const int yy = (int)(rest % HW); rest /= HW;
    const int c = (int)(rest % C);
    const int n = (int)(rest / C);
    const float2 v = reinterpret_cast<const float2*>(img)[idx];
    const int xx = x2 * 2;
    const int px = xx / P, j = xx % P;
    const int py = yy / P, i = yy % P;
    const size_t row = ((size_t)n * G + py) * G + px;
    const int col = c * P * P + i * P + j;
    *reinterpret_cast<uint32_t*>(out + row * ld_out + col) = pack_bf16(v.x, v.y);
  }
  const int K = C * P * P, pad = ld_out - K;
  if (pad > 0) {
    const size_t rows = (size_t)N * G * G;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < rows * pad;
         idx += (size_t)gridDim.x * blockDim.x)
      out[(idx / pad) * ld_out + K + (idx % pad)] = __float2bfloat16_rn(0.f);
  }
}

// Same im2col for patch sizes that are multiples of 8: eight pixels per thread (two 16 B loads,
// one 16 B store), blockIdx.y = image plane (n, c), 32-bit index arithmetic.
__global__ void __launch_bounds__(256)
patchify8_kernel(const float* __restrict__ img, int C, int HW, int P, __nv_bfloat16* __restrict__ out,
                 int ld_out) {
  const int G = HW / P, W8 = HW / 8;
  const int plane = blockIdx.y, n = plane / C, c = plane - n * C;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= HW * W8) return;
  const int yy = idx / W8, x8 = idx - yy * W8;
  const float4* src = reinterpret_cast<const float4*>(img + ((size_t)plane * HW + yy) * HW + x8 * 8);
  const float4 a = src[0], b = src[1];
  const int xx = x8 * 8;
  const int px = xx / P, j = xx - px * P;
  const int py = yy / P, i = yy - py * P;
  const size_t row = ((size_t)n * G + py) * G + px;
  const int col = c * P * P + i * P + j;
  *reinterpret_cast<uint4*>(out + row * ld_out + col) =
      make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
}

// class token + positional embedding + ln_pre (reference model.py:759-766)
template <int NV>
__global__ void __launch_bounds__(256)
embed_ln_pre_kernel(const float* __restrict__ patch_out, int ld_p, const float* __restrict__ cls,
                    const float* __restrict__ pos, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int N, int L, float* __restrict__ x0) {
  constexpr int D = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= N * L) return;
  const int n = row / L, l = row % L;
  const float4* src = (l == 0) ? reinterpret_cast<const float4*>(cls)
                               : reinterpret_cast<const float4*>(
                                     patch_out + ((size_t)n * (L - 1) + (l - 1)) * ld_p);
  const float4* pr = reinterpret_cast<const float4*>(pos + (size_t)l * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a = src[lane + 32 * i], p = __ldg(pr + lane + 32 * i);
    v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + kLnEps);
  float4* o = reinterpret_cast<float4*>(x0 + (size_t)row * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    o[c4] = make_float4(v[i].x * rstd * g.x + b.x, v[i].y * rstd * g.y + b.y,
                        v[i].z * rstd * g.z + b.z, v[i].w * rstd * g.w + b.w);
  }
}

// torch.optim.AdamW single-tensor update (decoupled weight decay), reference
// utils/train_utils.py:27-28; g is multiplied by grad_scale first (1/world, GradScaler unscale)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                             float* __restrict__ m, float* __restrict__ v, int n, float lr,
                             float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                             float grad_scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale;
  float pi = p[i] * (1.0f - lr * wd);
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi -= (lr / bc1) * (mi / denom);
  p[i] = pi;
}

#define DISPATCH_NV(D, CALL)                                                      \
  switch ((D) / 128) {                                                            \
    case 1: { constexpr int NV = 1; CALL; } break;                                \
    case 2: { constexpr int NV = 2; CALL; } break;                                \
    case 3: { constexpr int NV = 3; CALL; } break;                                \
    case 4: { constexpr int NV = 4; CALL; } break;                                \
    case 5: { constexpr int NV = 5; CALL; } break;                                \
    case 6: { constexpr int NV = 6; CALL; } break;                                \
    case 8: { constexpr int NV = 8; CALL; } break;                                \
    case 10: { constexpr int NV = 10; CALL; } break;                              \
    default:                                                                      \
      llc_set_error("LayerNorm width %d unsupported (need a multiple of 128 up to 1280)", (D)); \
      return LLC_ERR_ARG;                                                         \
  }

}  // namespace

extern "C" int llc_ln_fwd(const float* x, int ld_x, const float* gamma, const float* beta, int T,
                          int D, void* y, int ld_y, const float* lora_A, int r, void* stream) {
  LLC_REQUIRE(x && gamma && beta && y, "llc_ln_fwd: null pointer");
  LLC_REQUIRE(T >= 0 && D % 128 == 0 && ld_x % 4 == 0 && ld_y % 8 == 0, "llc_ln_fwd: bad shape");
  LLC_REQUIRE(lora_A == nullptr || (r >= 1 && r <= kMaxR && ld_y >= D + LLC_LORA_PAD),
              "llc_ln_fwd: LoRA rank %d unsupported or ld_y too small", r);
  if (T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  LLC_PROF_BEGIN(LLC_K_LN_FWD, T, D, 0, 0.0, 6.0 * T * D, st);
  DISPATCH_NV(D, (llc_launch_pdl(ln_fwd_kernel<NV>, dim3((T + 7) / 8), dim3(256), 0, st, x, ld_x,
                                 gamma, beta, T, (__nv_bfloat16*)y, ld_y, lora_A, r,
                                 g_llc_traversal & 1)));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("ln_fwd_kernel");
  return 0;
}

extern "C" int llc_ln_bwd(const float* x, int ld_x, const float* gamma, const void* dy, int ld_dy,
                          const float* dx_in, float* dx_out, int T, int D, void* dxb, int ld_dxb,
                          const float* lora_B, int r, float scale, void* stream) {
  LLC_REQUIRE(x && gamma && dy && dx_out, "llc_ln_bwd: null pointer");
  LLC_REQUIRE(T >= 0 && D % 128 == 0 && ld_x % 4 == 0 && ld_dy % 4 == 0, "llc_ln_bwd: bad shape");
  LLC_REQUIRE(dxb == nullptr || ld_dxb % 8 == 0, "llc_ln_bwd: ld_dxb %% 8");
  LLC_REQUIRE(lora_B == nullptr || (r >= 1 && r <= kMaxR && dxb && ld_dxb >= D + LLC_LORA_PAD),
              "llc_ln_bwd: LoRA rank %d unsupported or no room for du", r);
  LLC_REQUIRE(lora_B == nullptr || r != 4 || ((uintptr_t)lora_B & 15) == 0,
              "llc_ln_bwd: lora_B must be 16-byte aligned");
  if (T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  LLC_PROF_BEGIN(LLC_K_LN_BWD, T, D, 0, 0.0,
                 (double)T * D * (4 + 2 + (dx_in ? 4 : 0) + 4 + (dxb ? 2 : 0)), st);
  DISPATCH_NV(D, (llc_launch_pdl(ln_bwd_kernel<NV>, dim3((T + 7) / 8), dim3(256), 0, st, x, ld_x,
                                 gamma, (const __nv_bfloat16*)dy, ld_dy, dx_in, dx_out, T,
                                 (__nv_bfloat16*)dxb, ld_dxb, lora_B, r, scale,
                                 (g_llc_traversal >> 1) & 1)));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("ln_bwd_kernel");
  return 0;
}

extern "C" int llc_lora_side_max_partials(void) { return 4 * llc_num_sms(); }

extern "C" int llc_lora_side(void* X, int ld_x, int T, int C, int r, const float* Mrd, int rd_sc,
                             int rd_sj, float rd_scale, const void* w, int ld_w, float* partial,
                             int* n_partials, void* stream) {
  LLC_REQUIRE(X && T > 0 && C > 0, "llc_lora_side: empty input");
  LLC_REQUIRE(C % 8 == 0 && ld_x % 8 == 0 && ((uintptr_t)X & 15) == 0,
              "llc_lora_side: C=%d / ld_x=%d must be multiples of 8, X 16-byte aligned", C, ld_x);
  LLC_REQUIRE(r >= 1 && r <= kMaxR, "llc_lora_side: rank %d unsupported", r);
  LLC_REQUIRE(Mrd == nullptr || ld_x >= C + LLC_LORA_PAD, "llc_lora_side: no room for rowdot");
  LLC_REQUIRE(w == nullptr || (partial && n_partials), "llc_lora_side: colsum needs partial");
  const int R = r <= 4 ? 4 : 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (Mrd != nullptr) {
    const size_t smem = (size_t)C * R * sizeof(float);
    LLC_REQUIRE(smem <= 160 * 1024, "llc_lora_side: C=%d too wide for the staged factor", C);
    const int grid = min((T + 31) / 32, 4 * llc_num_sms());
    LLC_PROF_BEGIN(LLC_K_LORA_SIDE, T, C, 0, 2.0 * T * C * r, 2.0 * T * C, st);
    if (R == 4) {
      LLC_CONFIGURE_SMEM(lora_rowdot_kernel<4>, 160 * 1024);
      lora_rowdot_kernel<4><<<grid, kSideThreads, smem, st>>>((__nv_bfloat16*)X, ld_x, T, C, r, Mrd,
                                                             rd_sc, rd_sj, rd_scale);
    } else {
      LLC_CONFIGURE_SMEM(lora_rowdot_kernel<8>, 160 * 1024);
      lora_rowdot_kernel<8><<<grid, kSideThreads, smem, st>>>((__nv_bfloat16*)X, ld_x, T, C, r, Mrd,
                                                             rd_sc, rd_sj, rd_scale);
    }
    LLC_PROF_END(st);
    LLC_COUNT_LAUNCH();
    LLC_LAUNCH_CHECK("lora_rowdot_kernel");
  }
  static const bool cc_colsum = llc_dev_env("LLC_COLSUM_LEGACY") != nullptr;
  if (w != nullptr && !cc_colsum && llc_colsum_tc_eligible(X, ld_x, T, C, w, ld_w))
    return llc_colsum_tc(X, ld_x, T, C, R, w, ld_w, partial, n_partials, st);
  if (w != nullptr) {
    const int groups = (C + kSideGroup - 1) / kSideGroup;
    int gx = llc_lora_side_max_partials() / groups;   // ~4 CTAs per SM over all column groups
    const int by_rows = (T + 31) / 32;  // at least ~4 rows per warp
    if (gx > by_rows) gx = by_rows;
    if (gx < 1) gx = 1;
    LLC_PROF_BEGIN(LLC_K_LORA_SIDE, T, C, 2, 2.0 * T * C * r, 2.0 * T * C, st);
    const int w_vec = (((uintptr_t)w & 7) == 0 && ld_w % 4 == 0) ? 1 : 0;
    if (R == 4)
      lora_colsum_kernel<4><<<dim3(gx, groups), kSideThreads, 0, st>>>(
          (const __nv_bfloat16*)X, ld_x, T, C, r, (const __nv_bfloat16*)w, ld_w, w_vec, partial);
    else
      lora_colsum_kernel<8><<<dim3(gx, groups), kSideThreads, 0, st>>>(
          (const __nv_bfloat16*)X, ld_x, T, C, r, (const __nv_bfloat16*)w, ld_w, 0, partial);
    LLC_PROF_END(st);
    LLC_COUNT_LAUNCH();
    LLC_LAUNCH_CHECK("lora_colsum_kernel");
    *n_partials = gx;
  } else if (n_partials) {
    *n_partials = 0;
  }
  return 0;
}

extern "C" int llc_lora_colsum_finish(const float* partial, int n_partials, int C, int r,
                                      float cs_scale, float* out, int o_sc, int o_sj,
                                      void* stream) {
  LLC_REQUIRE(partial && out && n_partials > 0 && r >= 1 && r <= kMaxR,
              "llc_lora_colsum_finish: bad args");
  const int R = r <= 4 ? 4 : 8;
  LLC_PROF_BEGIN(LLC_K_LORA_SIDE, n_partials, C, 1, 0.0, 4.0 * n_partials * C * R,
                 (cudaStream_t)stream);
  colsum_finish_kernel<<<(C * R + 31) / 32, 256, 0, (cudaStream_t)stream>>>(
      partial, n_partials, C, R, r, cs_scale, out, o_sc, o_sj);
  LLC_PROF_END((cudaStream_t)stream);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("colsum_finish_kernel");
  return 0;
}

extern "C" int llc_lora_colsum_finish_multi(const llc_finish_job* jobs, int n_jobs, int r,
                                            void* stream) {
  LLC_REQUIRE(jobs && n_jobs >= 1 && n_jobs <= kMaxFinishJobs && r >= 1 && r <= kMaxR,
              "llc_lora_colsum_finish_multi: bad args");
  const int R = r <= 4 ? 4 : 8;
  FinishArgs a;
  int maxn = 0;
  for (int i = 0; i < n_jobs; ++i) {
    LLC_REQUIRE(jobs[i].partial && jobs[i].out && jobs[i].n_partials > 0 && jobs[i].C > 0,
                "llc_lora_colsum_finish_multi: job %d incomplete", i);
    a.j[i] = jobs[i];
    if (jobs[i].C * R > maxn) maxn = jobs[i].C * R;
  }
  LLC_PROF_BEGIN(LLC_K_LORA_SIDE, n_jobs, maxn, 1, 0.0, 0.0, (cudaStream_t)stream);
  for (int i = 0; i < n_jobs; ++i)
    LLC_REQUIRE(((uintptr_t)jobs[i].partial & 15) == 0,
                "llc_lora_colsum_finish_multi: job %d partials must be 16-byte aligned", i);
  colsum_finish_multi_kernel<<<dim3((maxn + 127) / 128, n_jobs), 1024, 0, (cudaStream_t)stream>>>(
      a, R, r);
  LLC_PROF_END((cudaStream_t)stream);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("colsum_finish_multi_kernel");
  return 0;
}

extern "C" int llc_pack_weight(const float* src, int rows, int cols, int transpose, void* dst,
                               int ld_dst, void* stream) {
  LLC_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_dst >= cols, "llc_pack_weight: bad args");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  pack_weight_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, rows, cols, transpose,
                                                               (__nv_bfloat16*)dst, ld_dst);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("pack_weight_kernel");
  return 0;
}

extern "C" int llc_pack_lora_cols(const float* src, int rows, int r, int s_i, int s_j, float scale,
                                  void* dst, int ld_dst, int col0, void* stream) {
  LLC_REQUIRE(src && dst && rows > 0 && r >= 1 && r <= LLC_LORA_PAD &&
                  ld_dst >= col0 + LLC_LORA_PAD,
              "llc_pack_lora_cols: bad args");
  const int n = rows * LLC_LORA_PAD;
  pack_lora_cols_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      src, rows, r, s_i, s_j, scale, (__nv_bfloat16*)dst, ld_dst, col0);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("pack_lora_cols_kernel");
  return 0;
}

int llc_refresh_lora_all(const llc_vit_layer* layers, int n_layers, int D, int r, float sc,
                         cudaStream_t st) {
  LLC_REQUIRE(n_layers >= 1 && n_layers <= kMaxRefreshLayers, "llc_vit_refresh_lora: %d layers (max %d)",
              n_layers, kMaxRefreshLayers);
  RefreshArgs a;
  for (int i = 0; i < n_layers; ++i) {
    const llc_vit_layer& y = layers[i];
    LLC_REQUIRE(y.in_A && y.in_B && y.out_A && y.out_B && y.wqkv_aug && y.wo_aug && y.wqkvT_aug &&
                    y.woT_aug && y.f_out_A && y.f_in_B && y.f_out_B,
                "llc_vit_refresh_lora: layer %d has a null operand", i);
    a.l[i] = RefreshLayer{y.in_A, y.in_B, y.out_A, y.out_B,
                          (__nv_bfloat16*)y.wqkv_aug, (__nv_bfloat16*)y.wo_aug,
                          (__nv_bfloat16*)y.wqkvT_aug, (__nv_bfloat16*)y.woT_aug,
                          (__nv_bfloat16*)y.f_out_A, (__nv_bfloat16*)y.f_in_B,
                          (__nv_bfloat16*)y.f_out_B};
  }
  refresh_lora_kernel<<<dim3(48, n_layers), 256, 0, st>>>(a, D, r, sc);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("refresh_lora_kernel");
  return 0;
}

extern "C" int llc_pack_factor_rows(const float* src, int r, int cols, int s_j, int s_c, float scale,
                                    void* dst, int ld_dst, void* stream) {
  LLC_REQUIRE(src && dst && r >= 1 && r <= LLC_LORA_PAD && cols > 0 && ld_dst >= cols,
              "llc_pack_factor_rows: bad args");
  const int n = LLC_LORA_PAD * cols;
  pack_factor_rows_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      src, r, cols, s_j, s_c, scale, (__nv_bfloat16*)dst, ld_dst);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("pack_factor_rows_kernel");
  return 0;
}

extern "C" int llc_patchify(const float* img, int N, int C, int HW, int P, void* out, int ld_out,
                            void* stream) {
  LLC_REQUIRE(img && out && N > 0, "llc_patchify: bad args");
  LLC_REQUIRE(P % 2 == 0 && HW % P == 0 && ld_out % 2 == 0 && ld_out >= C * P * P,
              "llc_patchify: image %d / patch %d unsupported", HW, P);
  const size_t total = (size_t)N * C * HW * (HW / 2);
  const int grid = (int)((total + 255) / 256 < (size_t)(llc_num_sms() * 16)
                             ? (total + 255) / 256
                             : (size_t)(llc_num_sms() * 16));
  LLC_PROF_BEGIN(LLC_K_EMBED, N, C * HW * HW, 0, 0.0, 6.0 * N * C * HW * HW,
                 (cudaStream_t)stream);
  // fast path: 8 pixels per thread; needs 16 B aligned rows on both sides and no K padding
  if (P % 8 == 0 && ld_out == C * P * P && ld_out % 8 == 0 && ((uintptr_t)img & 15) == 0 &&
      ((uintptr_t)out & 15) == 0 && (size_t)N * C <= 65535) {
    const dim3 g8((unsigned)((HW * (HW / 8) + 255) / 256), (unsigned)(N * C));
    patchify8_kernel<<<g8, 256, 0, (cudaStream_t)stream>>>(img, C, HW, P, (__nv_bfloat16*)out,
                                                           ld_out);
  } else {
    patchify_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, N, C, HW, P, (__nv_bfloat16*)out,
                                                            ld_out);
  }
  LLC_PROF_END((cudaStream_t)stream);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("patchify_kernel");
  return 0;
}

extern "C" int llc_embed_ln_pre(const float* patch_out, int ld_p, const float* class_emb,
                                const float* pos, const float* gamma, const float* beta, int N,
                                int L, int D, float* x0, void* stream) {
  LLC_REQUIRE(patch_out && class_emb && pos && gamma && beta && x0 && N > 0 && L > 1,
              "llc_embed_ln_pre: bad args");
  LLC_REQUIRE(D % 128 == 0 && ld_p % 4 == 0, "llc_embed_ln_pre: bad shape");
  const int T = N * L;
  cudaStream_t st = (cudaStream_t)stream;
  LLC_PROF_BEGIN(LLC_K_EMBED, T, D, 1, 0.0, 8.0 * T * D, st);
  DISPATCH_NV(D, (embed_ln_pre_kernel<NV><<<(T + 7) / 8, 256, 0, st>>>(
                     patch_out, ld_p, class_emb, pos, gamma, beta, N, L, x0)));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("embed_ln_pre_kernel");
  return 0;
}

extern "C" int llc_adamw(float* p, const float* g, float* m, float* v, int n, float lr, float beta1,
                         float beta2, float eps, float wd, int step, float grad_scale,
                         void* stream) {
  LLC_REQUIRE(p && g && m && v && n > 0 && step >= 1, "llc_adamw: bad args");
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
  LLC_PROF_BEGIN(LLC_K_OTHER, n, 0, 0, 0.0, 28.0 * n, (cudaStream_t)stream);
  adamw_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2,
                                                                  eps, wd, bc1, bc2_sqrt,
                                                                  grad_scale);
  LLC_PROF_END((cudaStream_t)stream);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("adamw_kernel");
  return 0;
}
