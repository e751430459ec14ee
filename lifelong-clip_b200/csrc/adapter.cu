// Bottleneck adapter of the adapter-clip method (reference models/clip/adapter.py:11-73, used by
// ResidualAttentionBlock_Adapter, models/clip/model.py:418-442):
//     adapter(y) = y + scale * (drop(relu(y W_d^T + b_d)) W_u^T + b_u)        W_d [64, D], W_u [D, 64]
// The two projections run on the tcgen05 GEMM (llc_gemm_bf16_tn: N = 64 and K = 64 problems); this
// file holds what is specific to the adapter:
//   * the activation pass on the [T, 64] bottleneck (ReLU + inverted dropout, forward / backward,
//     the backward also forming the column sums that are d b_d),
//   * the WEIGHT gradients, a token-reduction GEMM  P[c, j] = sum_t X[t, c] * w[t, j]  (d W_u =
//     scale dx^T a, d W_d = dz^T y). X [T, D] and w [T, 64] are consumed by tcgen05.mma AS THEY LIE
//     (both MN-major operands, K = tokens), X streams through shared memory once; a constant
//     "ones" operand tile adds the plain column sums of X (d b_u) to the same pass,
//   * the fixed-order finish kernels (bit-deterministic gradients) and the operand refresh.
#include "common.cuh"

namespace {

constexpr int kDim = LLC_ADAPTER_DIM;        // bottleneck width (adapter.py:39 hard-codes 64)
constexpr int kPW = 80;                      // floats per column in a partial: 64 + ones + pad
constexpr int kKB = 64;                      // tokens per k-block
constexpr int kATile = 2 * kKB * 128;        // two 64-column atoms of X
constexpr int kWTile = kKB * 128;            // 64 tokens x 64 bottleneck columns
constexpr int kStage = kATile + kWTile;
constexpr int kStages = 6;
constexpr int kSmem = 1024 + kStages * kStage + kWTile + 256;
constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads, 1)
tokgemm_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  int T, int C, int tok_per_split, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* sOnes = smem + kStages * kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + kWTile);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* done_bar = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 128;
  const int t_begin = blockIdx.y * tok_per_split;
  const int t_end = min(T, t_begin + tok_per_split);
  const int num_kb = t_end > t_begin ? (t_end - t_begin + kKB - 1) / kKB : 0;

  // constant B operand [64 tokens x 16]: column 0 = 1 -> accumulator column 64 = sum_t X[t, c].
  // MN-major rows of 128 B under TMA's 128 B swizzle: logical 16 B chunk 0 of row k sits at
  // physical chunk k & 7.
  for (int i = threadIdx.x; i < kWTile / 16; i += kThreads) {
    const int row = i >> 3, chunk = i & 7;
    reinterpret_cast<uint4*>(sOnes)[i] =
        make_uint4(chunk == (row & 7) ? 0x00003F80u : 0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(done_bar), 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<128>(smem_u32(tmem_slot));
  fence_proxy_async_smem();
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
        const int t0 = t_begin + kb * kKB;
        mbar_expect_tx(fb, kStage);
        tma_load_2d(sa, &tmX, fb, c0, t0);
        tma_load_2d(sa + kKB * 128, &tmX, fb, c0 + 64, t0);
        tma_load_2d(sa + kATile, &tmW, fb, 0, t0);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc_w = umma_idesc_bf16(128, kDim, 1, 1);   // A and B both MN-major
    const uint32_t idesc_1 = umma_idesc_bf16(128, 16, 1, 1);
    const uint32_t so = smem_u32(sOnes);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(smem_u32(&full_bar[stage]), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * kStage);
#pragma unroll
        for (int ks = 0; ks < kKB / 16; ++ks) {
          const uint64_t da = umma_desc_mn_sw128(sa + ks * 2048, kKB * 128, 1024);
          umma_bf16(tmem_base, da, umma_desc_mn_sw128(sa + kATile + ks * 2048, 8192, 1024),
                    idesc_w, (kb | ks) != 0);
          umma_bf16(tmem_base + kDim, da, umma_desc_mn_sw128(so + ks * 2048, 8192, 1024), idesc_1,
                    (kb | ks) != 0);
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        if (kb == num_kb - 1) umma_commit(smem_u32(done_bar));
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  }
  __syncwarp();
  // epilogue (all four warps): accumulator lane = column c0 + 32 warp + lane of X
  const int col = warp * 32 + lane;
  float4* out = reinterpret_cast<float4*>(partial + ((size_t)blockIdx.y * C + c0 + col) * kPW);
  if (num_kb > 0) {
    mbar_wait(smem_u32(done_bar), 0);
    tc_fence_after();
    const uint32_t tb = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int g = 0; g < kPW / 16; ++g) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
            "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]),
            "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(tb + g * 16)
          : "memory");
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q)
        out[g * 4 + q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                     __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
    }
  } else {
    for (int q = 0; q < kPW / 4; ++q) out[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// out_w[c * o_sc + j * o_sj] (+)= scale * sum_p P[p][c][j], j < 64; out_b[c] (+)= scale * sum_p
// P[p][c][64]. One thread per output, partials added in index order (bit-deterministic).
__global__ void __launch_bounds__(256)
tokgemm_finish_kernel(const float* __restrict__ partial, int n_partials, int C, float scale,
                      float* __restrict__ out_w, int o_sc, int o_sj, float* __restrict__ out_b,
                      int accumulate) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int per = kDim + 1;
  if (i >= C * per) return;
  const int c = i / per, j = i % per;
  if (j == kDim && out_b == nullptr) return;
  float s = 0.f;
  for (int p = 0; p < n_partials; ++p) s += partial[((size_t)p * C + c) * kPW + j];
  s *= scale;
  float* dst = j < kDim ? out_w + (size_t)c * o_sc + (size_t)j * o_sj : out_b + c;
  *dst = accumulate ? *dst + s : s;
}

// keep decision of element i of dropout stream (seed, use): a 64-bit mix (splitmix64 finaliser)
__device__ __forceinline__ bool keep_elem(unsigned long long seed, unsigned use, size_t i, float p) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)i + 1) +
                         0xD1B54A32D192ED03ull * (use + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f) >= p;
}

// a <- relu(a) * keep / (1 - p), in place on bf16 [T, 64]; mask (uint8, 1 = keep) overrides the
// generated stream
__global__ void __launch_bounds__(256)
adapter_act_kernel(__nv_bfloat16* __restrict__ a, size_t n8, const unsigned char* __restrict__ mask,
                   unsigned long long seed, unsigned use, float p) {
  const float inv = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8;
       i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = reinterpret_cast<uint4*>(a)[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    unsigned long long mk = 0x0101010101010101ull;
    if (mask) mk = *reinterpret_cast<const unsigned long long*>(mask + i * 8);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = unpack_bf16(w[e]);
      bool k0, k1;
      if (mask) {
        k0 = (mk >> (16 * e)) & 0xff;
        k1 = (mk >> (16 * e + 8)) & 0xff;
      } else if (p > 0.f) {
        k0 = keep_elem(seed, use, i * 8 + 2 * e, p);
        k1 = keep_elem(seed, use, i * 8 + 2 * e + 1, p);
      } else {
        k0 = k1 = true;
      }
      f.x = (f.x > 0.f && k0) ? f.x * inv : 0.f;
      f.y = (f.y > 0.f && k1) ? f.y * inv : 0.f;
      w[e] = pack_bf16(f.x, f.y);
    }
    reinterpret_cast<uint4*>(a)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// da <- da * [a > 0] / (1 - p) in place (the gradient of the pre-activation), and per-CTA column
// sums of the result -> part_b[blockIdx.x][64] (d b_d, finished in block order)
constexpr int kActRows = 32;   // rows per CTA step: 256 threads = 32 rows x 8 chunks
__global__ void __launch_bounds__(256)
adapter_act_bwd_kernel(__nv_bfloat16* __restrict__ da, const __nv_bfloat16* __restrict__ a, int T,
                       float p, float* __restrict__ part_b) {
  __shared__ float red[kActRows][kDim + 1];
  const float inv = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  const int r = threadIdx.x >> 3, ch = threadIdx.x & 7;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int t = blockIdx.x * kActRows + r; t < T; t += gridDim.x * kActRows) {
    const size_t i = (size_t)t * 8 + ch;
    const uint4 g = reinterpret_cast<const uint4*>(da)[i];
    const uint4 v = reinterpret_cast<const uint4*>(a)[i];
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, vw[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 gf = unpack_bf16(gw[e]), af = unpack_bf16(vw[e]);
      const float x = af.x > 0.f ? gf.x * inv : 0.f, y = af.y > 0.f ? gf.y * inv : 0.f;
      o[e] = pack_bf16(x, y);
      const float2 rf = unpack_bf16(o[e]);     // sum what the weight-gradient GEMM will read
      acc[2 * e] += rf.x;
      acc[2 * e + 1] += rf.y;
    }
    reinterpret_cast<uint4*>(da)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[r][ch * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < kDim) {
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < kActRows; ++k) s += red[k][threadIdx.x];
    part_b[(size_t)blockIdx.x * kDim + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(kDim)
adapter_bias_finish_kernel(const float* __restrict__ part_b, int n, float* __restrict__ out,
                           int accumulate) {
  float s = 0.f;
  for (int p = 0; p < n; ++p) s += part_b[(size_t)p * kDim + threadIdx.x];
  out[threadIdx.x] = accumulate ? out[threadIdx.x] + s : s;
}

// x[t, c] += y[t, c] (fp32 stream += bf16 branch)
__global__ void __launch_bounds__(256)
add_bf16_rows_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ y, int ld_y, int T,
                     int D) {
  const int per_row = D / 8;
  const size_t total = (size_t)T * per_row;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / per_row;
    const int c8 = (int)(i % per_row);
    const uint4 v = *reinterpret_cast<const uint4*>(y + row * ld_y + c8 * 8);
    float4* px = reinterpret_cast<float4*>(x + row * D + c8 * 8);
    float4 a = px[0], b = px[1];
    const float2 f0 = unpack_bf16(v.x), f1 = unpack_bf16(v.y), f2 = unpack_bf16(v.z),
                 f3 = unpack_bf16(v.w);
    a.x += f0.x; a.y += f0.y; a.z += f1.x; a.w += f1.y;
    b.x += f2.x; b.y += f2.y; b.z += f3.x; b.w += f3.y;
    px[0] = a;
    px[1] = b;
  }
}

// bf16 operands of the four projections from the live fp32 parameters (one launch):
//   wd [64, D] = W_d          wu [D, 64] = s W_u          wdT [D, 64] = W_d^T       wuT [64, D] = s W_u^T
//   bu_s [D] = s b_u
__global__ void __launch_bounds__(256)
adapter_refresh_kernel(const float* __restrict__ down_w, const float* __restrict__ up_w,
                       const float* __restrict__ up_b, float s, int D, __nv_bfloat16* wd,
                       __nv_bfloat16* wu, __nv_bfloat16* wdT, __nv_bfloat16* wuT, float* bu_s) {
  const int n = kDim * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = i / D, c = i % D;              // i indexes [64, D]
    const float d = down_w[i], u = s * up_w[(size_t)c * kDim + j];
    wd[i] = __float2bfloat16(d);
    wdT[(size_t)c * kDim + j] = __float2bfloat16(d);
    wu[(size_t)c * kDim + j] = __float2bfloat16(u);
    wuT[i] = __float2bfloat16(u);
    if (j == 0) bu_s[c] = s * up_b[c];
  }
}

// dst[t, col0 .. col0 + 64) = a[t, :]: the [T, 64] bottleneck gradient into the pad columns of a
// K-augmented operand
__global__ void __launch_bounds__(256)
adapter_scatter_kernel(const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ dst,
                       int ld_dst, int col0, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t t = i >> 3;
    const int ch = (int)(i & 7);
    *reinterpret_cast<uint4*>(dst + t * ld_dst + col0 + ch * 8) =
        reinterpret_cast<const uint4*>(a)[i];
  }
}

int grid_for(size_t items, int threads) {
  const size_t want = (items + threads - 1) / threads;
  const size_t cap = (size_t)llc_num_sms() * 8;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

int act_bwd_blocks(int T) {
  const int want = (T + kActRows - 1) / kActRows;
  const int cap = llc_num_sms() * 4;
  return want < cap ? want : cap;
}

// partial buffer: [slices <= SMs / (C / 128)][C][kPW] floats of the token-reduction GEMM, then
// [CTAs <= 4 SMs][64] of the bias column sums
size_t tok_region_floats() { return (size_t)llc_num_sms() * 128 * kPW; }

int tok_splits(int T, int C, int* per_out) {
  const int mtiles = C / 128;
  int splits = llc_num_sms() / mtiles;
  if (splits < 1) splits = 1;
  int per = (T + splits - 1) / splits;
  per = (per + kKB - 1) / kKB * kKB;          // k-blocks never straddle two slices
  *per_out = per;
  return (T + per - 1) / per;
}

// P[p][c][0..63] = sum over the tokens of slice p of X[t, c] * w[t, j]; P[p][c][64] = sum X[t, c]
int tokgemm(const void* X, int ld_x, const void* w, int T, int C, float* partial, int* n_partials,
            cudaStream_t st) {
  CUtensorMap tx, tw;
  if (int rc = llc_encode_tmap_2d(&tx, X, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)C,
                                  (uint64_t)T, (uint64_t)ld_x * 2, 64, kKB,
                                  CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  if (int rc = llc_encode_tmap_2d(&tw, w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)kDim,
                                  (uint64_t)T, (uint64_t)kDim * 2, kDim, kKB,
                                  CU_TENSOR_MAP_SWIZZLE_128B))
    return rc;
  int per = 0;
  const int splits = tok_splits(T, C, &per);
  LLC_CONFIGURE_SMEM(tokgemm_tc_kernel, kSmem);
  LLC_PROF_BEGIN(LLC_K_OTHER, T, C, kDim, 2.0 * T * C * (kDim + 16), 2.0 * T * (C + kDim), st);
  LLC_CUDA(llc_launch_pdl(tokgemm_tc_kernel, dim3(C / 128, splits), dim3(kThreads), (size_t)kSmem,
                          st, tx, tw, T, C, per, partial));
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("tokgemm_tc_kernel");
  *n_partials = splits;
  return 0;
}

int tok_finish(const float* partial, int n_partials, int C, float scale, float* out_w, int o_sc,
               int o_sj, float* out_b, int accumulate, cudaStream_t st) {
  const int n = C * (kDim + 1);
  tokgemm_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(partial, n_partials, C, scale, out_w, o_sc,
                                                        o_sj, out_b, accumulate);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("tokgemm_finish_kernel");
  return 0;
}

#define RUN(call)            \
  do {                       \
    int _rc = (call);        \
    if (_rc != 0) return _rc; \
  } while (0)

int check_adapter(const llc_adapter* ad, int T, int D, const char* who) {
  LLC_REQUIRE(ad && ad->wd && ad->wu && ad->wdT && ad->wuT && ad->bu_s && ad->down_b,
              "%s: adapter operands missing (llc_adapter_refresh)", who);
  LLC_REQUIRE(T > 0 && D % 128 == 0, "%s: T=%d D=%d unsupported (D must be a multiple of 128)", who,
              T, D);
  LLC_REQUIRE(ad->dropout >= 0.f && ad->dropout < 1.f, "%s: dropout %f", who, ad->dropout);
  return 0;
}

}  // namespace

int llc_adapter_scatter(const void* a, void* dst, int ld_dst, int col0, int T, void* stream) {
  LLC_REQUIRE(a && dst && ld_dst % 8 == 0 && col0 % 8 == 0 && T > 0, "llc_adapter_scatter: bad args");
  const size_t n8 = (size_t)T * kDim / 8;
  cudaStream_t st = (cudaStream_t)stream;
  LLC_PROF_BEGIN(LLC_K_OTHER, T, kDim, 2, 0.0, 4.0 * T * kDim, st);
  adapter_scatter_kernel<<<grid_for(n8, 256), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<__nv_bfloat16*>(dst), ld_dst, col0,
      n8);
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("adapter_scatter_kernel");
  return 0;
}

extern "C" size_t llc_adapter_partial_floats(int D) {
  if (D <= 0 || D % 128 != 0) return 0;
  return tok_region_floats() + (size_t)llc_num_sms() * 4 * kDim;
}

extern "C" int llc_adapter_refresh(const llc_adapter* ad, int D, void* stream) {
  LLC_REQUIRE(ad && ad->down_w && ad->up_w && ad->up_b && ad->wd && ad->wu && ad->wdT && ad->wuT &&
                  ad->bu_s && D > 0,
              "llc_adapter_refresh: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  adapter_refresh_kernel<<<grid_for((size_t)kDim * D, 256), 256, 0, st>>>(
      ad->down_w, ad->up_w, ad->up_b, ad->scale, D, reinterpret_cast<__nv_bfloat16*>(ad->wd),
      reinterpret_cast<__nv_bfloat16*>(ad->wu), reinterpret_cast<__nv_bfloat16*>(ad->wdT),
      reinterpret_cast<__nv_bfloat16*>(ad->wuT), ad->bu_s);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("adapter_refresh_kernel");
  // block use: the composed columns (W_d W_o)^T [D, 64] and (W_d W_proj)^T [mlp, 64] behind the
  // transposed frozen weights, so that the backward's d_o / dz GEMMs take the bottleneck gradient
  // as 64 more K columns (llc_adapter_block_backward)
  if (ad->woT_ad != nullptr) {
    LLC_REQUIRE(ad->wprojT_ad && ad->mlp_dim > 0, "llc_adapter_refresh: incomplete block operands");
    const int KAD = D + kDim;
    llc_gemm_epi e{};
    e.out = reinterpret_cast<__nv_bfloat16*>(ad->woT_ad) + D; e.ld_out = D + LLC_LORA_LD;
    RUN(llc_gemm_bf16_tn(ad->woT_ad, D + LLC_LORA_LD, ad->wd, D, D, kDim, D, &e, stream));
    e = llc_gemm_epi{};
    e.out = reinterpret_cast<__nv_bfloat16*>(ad->wprojT_ad) + D; e.ld_out = KAD;
    RUN(llc_gemm_bf16_tn(ad->wprojT_ad, KAD, ad->wd, D, ad->mlp_dim, kDim, D, &e, stream));
  }
  return 0;
}

// out = [resid +] [y +] scale * (drop(relu(y W_d^T + b_d)) W_u^T + b_u); a <- the bottleneck after
// ReLU and dropout (bf16 [T, 64], saved for backward)
extern "C" int llc_adapter_forward(const llc_adapter* ad, const void* y, int ld_y,
                                   const float* resid, int add_y, void* a,
                                   const unsigned char* mask, unsigned use, int training,
                                   float* out, int T, int D, void* stream) {
  RUN(check_adapter(ad, T, D, "llc_adapter_forward"));
  LLC_REQUIRE(y && a && out && ld_y % 8 == 0 && ld_y >= D, "llc_adapter_forward: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  llc_gemm_epi e{};
  e.bias = ad->down_b; e.out = a; e.ld_out = kDim;
  RUN(llc_gemm_bf16_tn(y, ld_y, ad->wd, D, T, kDim, D, &e, stream));
  const float p = training ? ad->dropout : 0.f;
  const size_t n8 = (size_t)T * kDim / 8;
  LLC_PROF_BEGIN(LLC_K_OTHER, T, kDim, 0, 0.0, 4.0 * T * kDim, st);
  adapter_act_kernel<<<grid_for(n8, 256), 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(a), n8,
                                                       training ? mask : nullptr, ad->seed, use, p);
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("adapter_act_kernel");
  e = llc_gemm_epi{};
  e.bias = ad->bu_s; e.resid = resid; e.ld_resid = D; e.out = out; e.ld_out = D; e.out_fp32 = 1;
  RUN(llc_gemm_bf16_tn(a, kDim, ad->wu, kDim, T, D, kDim, &e, stream));
  if (add_y) {
    LLC_PROF_BEGIN(LLC_K_OTHER, T, D, 0, 0.0, 10.0 * T * D, st);
    add_bf16_rows_kernel<<<grid_for((size_t)T * D / 8, 256), 256, 0, st>>>(
        out, reinterpret_cast<const __nv_bfloat16*>(y), ld_y, T, D);
    LLC_PROF_END(st);
    LLC_COUNT_LAUNCH();
    LLC_LAUNCH_CHECK("add_bf16_rows_kernel");
  }
  return 0;
}

// dx fp32 [T, D] = gradient of `out` (dxb: its bf16 copy). Accumulates (accumulate = 1) or writes
// the four parameter gradients; d_y fp32 [T, D] = [dx +] dz W_d, the gradient of y (pass_dx: y
// also reached `out` directly). da: scratch bf16 [T, 64]; partial: llc_adapter_partial_floats.
extern "C" int llc_adapter_backward(const llc_adapter* ad, const void* y, int ld_y, const void* a,
                                    const float* dx, const void* dxb, int ld_dxb, float* d_y,
                                    int pass_dx, void* da, float* partial, int accumulate,
                                    int training, int T, int D, void* stream) {
  RUN(check_adapter(ad, T, D, "llc_adapter_backward"));
  LLC_REQUIRE(y && a && dxb && da && partial && ld_y % 8 == 0 && ld_dxb % 8 == 0,
              "llc_adapter_backward: bad args");
  LLC_REQUIRE(ad->g_down_w && ad->g_down_b && ad->g_up_w && ad->g_up_b,
              "llc_adapter_backward: gradient slots missing");
  cudaStream_t st = (cudaStream_t)stream;
  const float p = training ? ad->dropout : 0.f;
  // da = dx (s W_u)
  llc_gemm_epi e{};
  e.out = da; e.ld_out = kDim;
  RUN(llc_gemm_bf16_tn(dxb, ld_dxb, ad->wuT, D, T, kDim, D, &e, stream));
  // dz = da o [a > 0] / (1 - p); d b_d = column sums
  float* part_b = partial + tok_region_floats();
  const int nb = act_bwd_blocks(T);
  LLC_PROF_BEGIN(LLC_K_OTHER, T, kDim, 1, 0.0, 6.0 * T * kDim, st);
  adapter_act_bwd_kernel<<<nb, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(da),
                                            reinterpret_cast<const __nv_bfloat16*>(a), T, p, part_b);
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("adapter_act_bwd_kernel");
  adapter_bias_finish_kernel<<<1, kDim, 0, st>>>(part_b, nb, ad->g_down_b, accumulate);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("adapter_bias_finish_kernel");
  // d W_u [D, 64] = s dx^T a, d b_u = s colsum(dx)   (s is folded here: `a` fed an s-scaled W_u)
  int np = 0;
  RUN(tokgemm(dxb, ld_dxb, a, T, D, partial, &np, st));
  RUN(tok_finish(partial, np, D, ad->scale, ad->g_up_w, kDim, 1, ad->g_up_b, accumulate, st));
  // d W_d [64, D] = dz^T y
  RUN(tokgemm(y, ld_y, da, T, D, partial, &np, st));
  RUN(tok_finish(partial, np, D, 1.0f, ad->g_down_w, 1, D, nullptr, accumulate, st));
  // d_y = [dx +] dz W_d
  if (d_y) {
    LLC_REQUIRE(!pass_dx || dx, "llc_adapter_backward: pass_dx needs dx");
    e = llc_gemm_epi{};
    e.resid = pass_dx ? dx : nullptr; e.ld_resid = D; e.out = d_y; e.ld_out = D; e.out_fp32 = 1;
    RUN(llc_gemm_bf16_tn(da, kDim, ad->wdT, kDim, T, D, kDim, &e, stream));
  }
  return 0;
}
