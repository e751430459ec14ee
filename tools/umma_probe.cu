// Dev probe (not part of libllc): validates tcgen05 pieces the attention kernel relies on, in
// isolation on one CTA:  D[128 x 64] = A[128 x K] . B[K x 64]  with
//   B given as [K rows][64 cols] (the natural layout of V: "MN-major" B operand, 128B swizzle)
//   variant 0: A from shared memory (K-major, 128B swizzle)           (SS)
//   variant 1: A written to TMEM with tcgen05.st as packed bf16 pairs  (TS), as P in attention
// build: nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tools/_probe.so tools/umma_probe.cu
#include "../lifelong-clip_b200/csrc/common.cuh"

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
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// K must be a multiple of 16, <= 256
__global__ void __launch_bounds__(128, 1)
probe_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K,
             int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // K-major: 4 atoms [128 rows x 128 B] (64 k each) = 64 KB
  uint8_t* sB = smem + 65536;         // MN-major: [K rows][128 B], 8-row groups of 1024 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 256 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // A -> smem K-major SW128: element (r, k): atom = k / 64, within: row r, byte (k % 64) * 2,
  // 16 B chunk index XOR (r & 7)
  for (int i = tid; i < 128 * 256; i += 128) {
    const int r = i / 256, k = i % 256;
    const float v = (k < K) ? A[r * K + k] : 0.f;
    const int atom = k / 64, kk = k % 64;
    const int chunk = (kk / 8) ^ (r & 7);
    *reinterpret_cast<__nv_bfloat16*>(sA + atom * 16384 + r * 128 + chunk * 16 + (kk % 8) * 2) =
        __float2bfloat16_rn(v);
  }
  // B [K][64] -> smem rows of 128 B, chunk XOR (row & 7)
  for (int i = tid; i < 256 * 64; i += 128) {
    const int k = i / 64, n = i % 64;
    const float v = (k < K) ? B[k * 64 + n] : 0.f;
    const int chunk = (n / 8) ^ (k & 7);
    *reinterpret_cast<__nv_bfloat16*>(sB + k * 128 + chunk * 16 + (n % 8) * 2) =
        __float2bfloat16_rn(v);
  }
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(smem_u32(slot));
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *slot;
  const uint32_t tD = tbase;          // 64 columns
  const uint32_t tA = tbase + 64;     // K/2 columns of packed bf16 pairs

  if (variant == 1) {
    // thread = row (TMEM lane = warp*32 + lane): packed pairs (k, k+1) -> 32-bit column k/2
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = 2 * (c0 + j);
        w[j] = pack_bf16(A[r * K + k], A[r * K + k + 1]);
      }
      tmem_st_32x32b_x8(tA + ((uint32_t)(warp * 32) << 16) + c0, w);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (warp == 1) {
    // idesc: M=128, N=64, A K-major, B MN-major (bit 16)
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
    if (elect_one()) {
      for (int ks = 0; ks < K / 16; ++ks) {
        // B: 16 k rows = 2 groups of 8 rows (1024 B each) -> advance 2048 B per step
        const uint64_t bdesc = umma_desc_mn_sw128(smem_u32(sB) + ks * 2048, 8192 /*unused: N=64*/, 1024);
        if (variant == 0) {
          const uint64_t adesc = umma_desc_k_sw128(smem_u32(sA) + (ks / 4) * 16384) + 2 * (ks % 4);
          umma_bf16(tD, adesc, bdesc, idesc, ks != 0);
        } else {
          umma_bf16_ts(tD, tA + ks * 8, bdesc, idesc, ks != 0);
        }
      }
      umma_commit(smem_u32(bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(bar), 0);
  tc_fence_after();
  {
    const int r = warp * 32 + lane;
    uint32_t v[32];
    for (int c = 0; c < 64; c += 32) {
      tmem_ld_32x32(tD + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[r * 64 + c + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tbase);
  }
}

void llc_set_error(const char*, ...) {}
int llc_check_cuda(cudaError_t e, const char*) { return (int)e; }

extern "C" int probe_run(const float* A, const float* B, float* D, int K, int variant) {
  const int smem = 1024 + 65536 + 256 * 128 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(A, B, D, K, variant);
  return (int)cudaDeviceSynchronize();
}
