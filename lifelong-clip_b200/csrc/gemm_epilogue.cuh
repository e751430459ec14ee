// Epilogue of the tcgen05 GEMMs: TMEM accumulator -> registers -> fused element-wise work of the
// ViT block -> per-warp smem transpose -> coalesced 16-byte global I/O.
//   EPI_BF16   out = bf16(acc + bias)                                  (qkv, d_o, dh)
//   EPI_GELU   z = bf16(acc + bias) (optional), out2 = bf16(QuickGELU(z))   model.py:203-206,219-222
//   EPI_DGELU  out = bf16(acc * QuickGELU'(aux))                         backward of the above
//   EPI_F32    out = fp32(acc + bias + resid)                            residual stream model.py:234-235
// One warp owns 32 accumulator rows (its TMEM lane quarter; thread = row after tcgen05.ld 32x32b)
// and walks NCH chunks of 32 columns.
//
// bf16 modes do all math in the row-per-thread layout straight out of TMEM (32 independent
// elements per thread: no shared-memory latency in the dependency chain), then transpose the
// PACKED bf16 result through a 2 KB XOR-swizzled staging tile so that global stores are 16 B per
// lane with 4 lanes per 64 B row segment. EPI_DGELU brings its aux operand in through the same
// tile in the opposite direction (coalesced load -> row-per-thread). EPI_F32 transposes the fp32
// accumulator (4 KB tile) and adds bias/residual in the coalesced layout, where the residual
// loads are 16 B per lane. Global inputs of chunk c+1 are requested before chunk c is processed.
#pragma once
#include "common.cuh"

enum { EPI_BF16 = 0, EPI_GELU = 1, EPI_DGELU = 2, EPI_F32 = 3 };

struct EpiParams {
  const float* bias;
  const float* resid;
  int ld_resid;
  int act;
  const __nv_bfloat16* aux;
  int ld_aux;
  void* out;
  int ld_out;
  int out_fp32;
  __nv_bfloat16* out2;
  int ld_out2;
  int dbg;  // development only (LLC_GEMM_DBG): 256 = no global stores, 512 = no TMEM loads
  int keep_out;  // the bf16 output is small enough to stay in L2 for its consumer: no evict-first
  void* ws;      // stream-K workspace (llc_gemm_ws_bytes, flags zero) or nullptr
  size_t ws_bytes;
};

constexpr int kEpiWarpBytes = 8192;  // TMA-store tiles (see epi_warp_tile_tma) / one 4 KB fp32 tile

// QuickGELU x*sigmoid(1.702x) with sigmoid(y) = 0.5 + 0.5 tanh(y/2): one MUFU op per element
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float quick_gelu_fast(float x) {
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(0.851f * x), hx);
}
__device__ __forceinline__ float quick_gelu_grad_fast(float x) {
  const float s = fmaf(0.5f, tanh_approx(0.851f * x), 0.5f);
  return s * fmaf(1.702f * x, 1.0f - s, 1.0f);
}

// ---- staging tiles ------------------------------------------------------------------------------
// bf16 tile: 32 rows x 64 B, 16 B chunk c of row r stored at chunk c ^ ((r >> 1) & 3): both the
// row-per-thread side (8 consecutive rows, one chunk) and the coalesced side (2 rows x 4 chunks)
// touch 32 distinct banks per 8-lane phase.
__device__ __forceinline__ uint32_t bf_tile_off(int row, int chunk) {
  return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
}
// fp32 tile: 32 rows x 128 B, chunk c of row r at c ^ (r & 7)
__device__ __forceinline__ uint32_t f32_tile_off(int row, int chunk) {
  return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

template <int MODE>
struct EpiPre {
  float4 bias[8];  // bf16 modes: the chunk's 32 bias values (uniform); F32: [0] = this lane's 4
  uint4 aux[4];    // DGELU: this lane's 16 B of 4 rows (coalesced mapping)
  float4 rs[8];    // F32: this lane's 16 B of 8 rows
};

template <int MODE>
__device__ __forceinline__ void epi_prefetch(EpiPre<MODE>& p, const EpiParams& ep, int row0,
                                             int col, int M, int N, int lane) {
  if (MODE == EPI_F32) {
    const int cc = col + (lane & 7) * 4;
    p.bias[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias != nullptr && cc < N)
      p.bias[0] = __ldg(reinterpret_cast<const float4*>(ep.bias + cc));
    if (ep.resid != nullptr) {
#pragma unroll
      for (int pass = 0; pass < 8; ++pass) {
        const int row = row0 + pass * 4 + (lane >> 3);
        p.rs[pass] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < M && cc < N)
          p.rs[pass] = *reinterpret_cast<const float4*>(ep.resid + (size_t)row * ep.ld_resid + cc);
      }
    }
  } else {
    if (MODE != EPI_DGELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        p.bias[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias != nullptr && col + 4 * j < N)  // same address in every lane: one wavefront
          p.bias[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + col + 4 * j));
      }
    } else {
      const int cc = col + (lane & 3) * 8;
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        const int row = row0 + pass * 8 + (lane >> 2);
        p.aux[pass] = make_uint4(0, 0, 0, 0);
        if (row < M && cc < N)
          p.aux[pass] = *reinterpret_cast<const uint4*>(ep.aux + (size_t)row * ep.ld_aux + cc);
      }
    }
  }
}

#define g_dbg_nostore (ep_dbg_flags & 256)
// packed bf16 row (32 columns = 4 x 16 B) of this thread -> staging tile
__device__ __forceinline__ void stage_row_bf16(uint8_t* tile, int lane, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<uint4*>(tile + bf_tile_off(lane, c)) =
        make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
}
// staging tile -> global, 4 lanes per row, 8 rows per instruction
__device__ __forceinline__ void store_tile_bf16(const uint8_t* tile, __nv_bfloat16* out, int ld,
                                                int row0, int col, int M, int N, int lane,
                                                int ep_dbg_flags) {
  const int cc = col + (lane & 3) * 8;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int rr = pass * 8 + (lane >> 2), row = row0 + rr;
    const uint4 v = *reinterpret_cast<const uint4*>(tile + bf_tile_off(rr, lane & 3));
    if (row < M && cc < N && !(g_dbg_nostore)) *reinterpret_cast<uint4*>(out + (size_t)row * ld + cc) = v;
  }
}

template <int MODE>
__device__ __forceinline__ void epi_chunk(const EpiPre<MODE>& p, const EpiParams& ep,
                                          const uint32_t (&acc)[32], uint8_t* tile, int row0,
                                          int col, int M, int N, int lane) {
  if (MODE == EPI_F32) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4*>(tile + f32_tile_off(lane, j)) =
          make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    __syncwarp();
    const int c4 = lane & 7, cc = col + c4 * 4;
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
      const int rr = pass * 4 + (lane >> 3), row = row0 + rr;
      float4 v = *reinterpret_cast<const float4*>(tile + f32_tile_off(rr, c4));
      v.x += p.bias[0].x; v.y += p.bias[0].y; v.z += p.bias[0].z; v.w += p.bias[0].w;
      if (ep.resid != nullptr) {
        v.x += p.rs[pass].x; v.y += p.rs[pass].y; v.z += p.rs[pass].z; v.w += p.rs[pass].w;
      }
      if (row < M && cc < N)
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (size_t)row * ep.ld_out + cc) = v;
    }
    __syncwarp();
    return;
  }
  uint8_t* t0 = tile;
  uint8_t* t1 = tile + 2048;
  float v[32];
  if (MODE == EPI_DGELU) {
    // aux arrives in the coalesced mapping; turn it into this thread's row through tile 1
#pragma unroll
    for (int pass = 0; pass < 4; ++pass)
      *reinterpret_cast<uint4*>(t1 + bf_tile_off(pass * 8 + (lane >> 2), lane & 3)) = p.aux[pass];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 zq = *reinterpret_cast<const uint4*>(t1 + bf_tile_off(lane, c));
      const uint32_t zw[4] = {zq.x, zq.y, zq.z, zq.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 z = unpack_bf16(zw[e]);
        v[8 * c + 2 * e] = __uint_as_float(acc[8 * c + 2 * e]) * quick_gelu_grad_fast(z.x);
        v[8 * c + 2 * e + 1] = __uint_as_float(acc[8 * c + 2 * e + 1]) * quick_gelu_grad_fast(z.y);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] = __uint_as_float(acc[4 * j]) + p.bias[j].x;
      v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + p.bias[j].y;
      v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + p.bias[j].z;
      v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + p.bias[j].w;
    }
  }
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(ep.out);
  if (MODE == EPI_GELU) {
    if (out != nullptr) stage_row_bf16(t0, lane, pk);
    // the activation is applied to the bf16-rounded pre-activation, so that backward (which only
    // sees the saved bf16 z) differentiates exactly the function forward evaluated
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 z = unpack_bf16(pk[i]);
      pk[i] = pack_bf16(quick_gelu_fast(z.x), quick_gelu_fast(z.y));
    }
    stage_row_bf16(t1, lane, pk);
    __syncwarp();
    if (out != nullptr) store_tile_bf16(t0, out, ep.ld_out, row0, col, M, N, lane, ep.dbg);
    store_tile_bf16(t1, ep.out2, ep.ld_out2, row0, col, M, N, lane, ep.dbg);
  } else {
    stage_row_bf16(t0, lane, pk);
    __syncwarp();
    store_tile_bf16(t0, out, ep.ld_out, row0, col, M, N, lane, ep.dbg);
  }
  __syncwarp();  // tiles are rewritten by the next chunk
}

// DRAM -> L2 request for the rows a warp will read as aux / residual one tile later
// (row0/col0 of that tile; the warp covers 32 rows x NCH*32 columns)
template <int MODE, int NCH>
__device__ __forceinline__ void epi_l2_prefetch(const EpiParams& ep, int row0, int col0, int M,
                                                int lane) {
  const char* base = nullptr;
  size_t pitch = 0;
  int row_bytes = 0;
  if (MODE == EPI_F32 && ep.resid != nullptr) {
    base = reinterpret_cast<const char*>(ep.resid + col0);
    pitch = (size_t)ep.ld_resid * 4;
    row_bytes = NCH * 32 * 4;
  } else if (MODE == EPI_DGELU) {
    base = reinterpret_cast<const char*>(ep.aux + col0);
    pitch = (size_t)ep.ld_aux * 2;
    row_bytes = NCH * 32 * 2;
  } else {
    return;
  }
  const int lines = row_bytes / 128;  // per row
  for (int i = lane; i < 32 * lines; i += 32) {
    const int r = i / lines, l = i % lines;
    if (row0 + r < M)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)(row0 + r) * pitch + l * 128));
  }
}

// One warp drains NCH chunks (32 rows x 32 columns each) starting at TMEM address t_addr.
template <int MODE, int NCH>
__device__ __forceinline__ void epi_warp_tile(const EpiParams& ep, uint32_t t_addr, uint8_t* tile,
                                              int row0, int col0, int M, int N, int lane) {
  if (MODE == EPI_F32) {
    // 8 x 16 B of residual per lane and chunk: not double-buffered (registers); the loads were
    // requested into L2 one tile ahead by epi_l2_prefetch
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      EpiPre<MODE> cur;
      epi_prefetch<MODE>(cur, ep, row0, col0 + c * 32, M, N, lane);
      uint32_t acc[32];
      tmem_ld_32x32(t_addr + c * 32, acc);
      tmem_ld_wait();
      epi_chunk<MODE>(cur, ep, acc, tile, row0, col0 + c * 32, M, N, lane);
    }
    return;
  }
  EpiPre<MODE> cur, nxt;
  epi_prefetch<MODE>(cur, ep, row0, col0, M, N, lane);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c + 1 < NCH) epi_prefetch<MODE>(nxt, ep, row0, col0 + (c + 1) * 32, M, N, lane);
    uint32_t acc[32];
    if (!(ep.dbg & 512)) {
      tmem_ld_32x32(t_addr + c * 32, acc);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0x3f800000u + j + lane;
    }
    epi_chunk<MODE>(cur, ep, acc, tile, row0, col0 + c * 32, M, N, lane);
    if (c + 1 < NCH) cur = nxt;
  }
}


// ---- bf16 modes with TMA stores (CTA-pair kernel) ------------------------------------------------
// Register-side math as above, but the packed rows of TWO chunks (64 columns = one full 128 B
// line per row) are staged in a [32 x 128 B] tile in TMA's SWIZZLE_128B pattern and written with
// one cp.async.bulk.tensor store per warp: no LDS, no per-lane global stores (the 16 B STG path
// cost 30-70 us per GEMM in LSU wavefronts and write transactions, profiles/), full-line writes,
// row tail clipped by the tensor map. Per warp 8 KB: BF16 double-buffers the tile, GELU holds the
// z and g tiles, DGELU one tile + a 2 KB scratch that turns the coalesced aux load into rows.
__device__ __forceinline__ uint32_t line_tile_off(int row, int c16) {
  return (uint32_t)(row * 128 + ((c16 ^ (row & 7)) << 4));
}

template <int MODE>
__device__ __forceinline__ void epi_rows_bf16(const EpiPre<MODE>& p, const uint32_t (&acc)[32],
                                              uint8_t* scratch, int lane, uint32_t (&pk)[16],
                                              uint32_t (&pg)[16]) {
  float v[32];
  if (MODE == EPI_DGELU) {
#pragma unroll
    for (int pass = 0; pass < 4; ++pass)
      *reinterpret_cast<uint4*>(scratch + bf_tile_off(pass * 8 + (lane >> 2), lane & 3)) =
          p.aux[pass];
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint4 zq = *reinterpret_cast<const uint4*>(scratch + bf_tile_off(lane, c));
      const uint32_t zw[4] = {zq.x, zq.y, zq.z, zq.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 z = unpack_bf16(zw[e]);
        v[8 * c + 2 * e] = __uint_as_float(acc[8 * c + 2 * e]) * quick_gelu_grad_fast(z.x);
        v[8 * c + 2 * e + 1] = __uint_as_float(acc[8 * c + 2 * e + 1]) * quick_gelu_grad_fast(z.y);
      }
    }
    __syncwarp();  // scratch is rewritten by the next chunk
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] = __uint_as_float(acc[4 * j]) + p.bias[j].x;
      v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + p.bias[j].y;
      v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + p.bias[j].z;
      v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + p.bias[j].w;
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
  if (MODE == EPI_GELU) {
    // activation of the bf16-ROUNDED pre-activation (what backward will see)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 z = unpack_bf16(pk[i]);
      pg[i] = pack_bf16(quick_gelu_fast(z.x), quick_gelu_fast(z.y));
    }
  }
}

// one warp, NCH (even) chunks of 32 columns, 32 rows; wtile = this warp's 8 KB
template <int MODE, int NCH>
__device__ __forceinline__ void epi_warp_tile_tma(const EpiParams& ep, const CUtensorMap* tmO,
                                                  const CUtensorMap* tmO2, uint32_t t_addr,
                                                  uint8_t* wtile, int row0, int col0, int M, int N,
                                                  int lane) {
  EpiPre<MODE> cur, nxt;
  epi_prefetch<MODE>(cur, ep, row0, col0, M, N, lane);
#pragma unroll
  for (int g = 0; g < NCH / 2; ++g) {
    uint8_t* tA = (MODE == EPI_BF16) ? wtile + (g & 1) * 4096 : wtile;
    uint8_t* tB = wtile + 4096;   // GELU: g tile; DGELU: aux scratch
    // the tile(s) about to be rewritten must have been read by their previous TMA store
    if (lane == 0) {
      if (MODE == EPI_BF16) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
    }
    __syncwarp();
#pragma unroll
    for (int hc = 0; hc < 2; ++hc) {
      const int c = 2 * g + hc;
      if (c + 1 < NCH) epi_prefetch<MODE>(nxt, ep, row0, col0 + (c + 1) * 32, M, N, lane);
      uint32_t acc[32], pk[16], pg[16];
      tmem_ld_32x32(t_addr + c * 32, acc);
      tmem_ld_wait();
      epi_rows_bf16<MODE>(cur, acc, tB, lane, pk, pg);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE != EPI_GELU || ep.out != nullptr)
          *reinterpret_cast<uint4*>(tA + line_tile_off(lane, hc * 4 + j)) =
              make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        if (MODE == EPI_GELU)
          *reinterpret_cast<uint4*>(tB + line_tile_off(lane, hc * 4 + j)) =
              make_uint4(pg[4 * j], pg[4 * j + 1], pg[4 * j + 2], pg[4 * j + 3]);
      }
      if (c + 1 < NCH) cur = nxt;
    }
    fence_proxy_async_smem();   // generic-proxy tile writes -> visible to the TMA engine
    __syncwarp();
    if (lane == 0) {
      if (MODE != EPI_GELU || ep.out != nullptr)
        tma_store_2d_hint(tmO, smem_u32(tA), col0 + g * 64, row0,
                          ep.keep_out ? kEvictNormal : kEvictFirst);
      if (MODE == EPI_GELU) tma_store_2d_hint(tmO2, smem_u32(tB), col0 + g * 64, row0, kEvictFirst);
      tma_store_commit();
    }
  }
}


// ---- fp32 residual epilogue with TMA in both directions (CTA-pair kernel) -------------------------
// out = fp32(acc + bias + resid). Per 32-column chunk the residual rows arrive by TMA in a
// [32 x 128 B] swizzled tile, each thread adds its accumulator row IN PLACE (row-per-thread, 8 x
// 16 B, conflict-free), and the tile leaves by TMA store: full 128 B lines both ways, no per-lane
// global access (the LSU version cost ~35 us per GEMM, profiles/). Two tiles per warp alternate;
// the loads of a tile's first two chunks are issued one tile ahead, the other two as soon as the
// stores of the first two have read their tiles (their rows were requested into L2 a tile ago).
struct EpiF32State {
  uint32_t ph[2];
};
__device__ __forceinline__ void epi_f32_load(const CUtensorMap* tmR, uint8_t* tile, uint64_t* mb,
                                             int col, int row) {
  mbar_expect_tx(smem_u32(mb), 4096);
  tma_load_2d_hint(smem_u32(tile), tmR, smem_u32(mb), col, row, kEvictFirst);
}
// issue the residual loads of chunks 0 and 1 of the tile at (row0, col0); lane 0 only
__device__ __forceinline__ void epi_f32_prime(const CUtensorMap* tmR, uint8_t* wtile, uint64_t* mb,
                                              int row0, int col0) {
  tma_store_wait_read<0>();   // both tiles have been read by their stores
  epi_f32_load(tmR, wtile, mb, col0, row0);
  epi_f32_load(tmR, wtile + 4096, mb + 1, col0 + 32, row0);
}

template <int NCH>
__device__ __forceinline__ void epi_warp_tile_f32_tma(const EpiParams& ep, const CUtensorMap* tmO,
                                                      const CUtensorMap* tmR, uint32_t t_addr,
                                                      uint8_t* wtile, uint64_t* mb, EpiF32State& st,
                                                      int row0, int col0, int N, int lane) {
  const bool has_resid = ep.resid != nullptr;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int b = c & 1;
    uint8_t* T = wtile + b * 4096;
    float4 bias[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bias[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ep.bias != nullptr && col0 + c * 32 + 4 * j < N)
        bias[j] = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + c * 32 + 4 * j));
    }
    uint32_t acc[32];
    tmem_ld_32x32(t_addr + c * 32, acc);
    if (has_resid) {
      mbar_wait(smem_u32(mb + b), st.ph[b]);
      st.ph[b] ^= 1;
    } else {
      if (lane == 0) tma_store_wait_read<1>();   // the store two chunks ago has read this tile
      __syncwarp();
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4* p = reinterpret_cast<float4*>(T + f32_tile_off(lane, j));
      float4 v = has_resid ? *p : make_float4(0.f, 0.f, 0.f, 0.f);
      v.x += __uint_as_float(acc[4 * j]) + bias[j].x;
      v.y += __uint_as_float(acc[4 * j + 1]) + bias[j].y;
      v.z += __uint_as_float(acc[4 * j + 2]) + bias[j].z;
      v.w += __uint_as_float(acc[4 * j + 3]) + bias[j].w;
      *p = v;
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d_hint(tmO, smem_u32(T), col0 + c * 32, row0, kEvictFirst);
      tma_store_commit();
      if (has_resid && c + 2 < NCH) {
        tma_store_wait_read<0>();
        epi_f32_load(tmR, T, mb + b, col0 + (c + 2) * 32, row0);
      }
    }
  }
}

inline int epi_mode_of(const EpiParams& ep) {
  if (ep.act == 1) return EPI_GELU;
  if (ep.act == 2) return EPI_DGELU;
  return ep.out_fp32 ? EPI_F32 : EPI_BF16;
}
