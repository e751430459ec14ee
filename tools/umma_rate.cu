// Dev probe (not part of libllc): tcgen05.mma issue/execute rate for the small shapes the
// attention kernels use. One CTA, one issuing thread, R back-to-back MMAs into one accumulator,
// timed with clock64 from the first issue to the commit's mbarrier completion.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -I include -o tools/_umma_rate tools/umma_rate.cu
#include <cstdio>
#include "../lifelong-clip_b200/csrc/common.cuh"

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// variant: bit0 = A from TMEM, bit1 = A MN-major (SS only), bit2 = B MN-major
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int variant, int R, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1) {
    const bool ts = variant & 1, amn = variant & 2, bmn = variant & 4;
    const uint32_t idesc = umma_idesc_bf16(128, N, amn ? 1 : 0, bmn ? 1 : 0);
    const uint32_t sA = smem_u32(smem), sB = sA + 64 * 1024;
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < R; ++i) {
          const int ks = i & 3;
          const uint64_t ad = amn ? umma_desc_mn_sw128(sA + ks * 2048, 16384, 1024)
                                  : umma_desc_k_sw128(sA) + 2 * ks;
          const uint64_t bd = bmn ? umma_desc_mn_sw128(sB + ks * 2048, 16384, 1024)
                                  : umma_desc_k_sw128(sB) + 2 * ks;
          if (ts) umma_bf16_ts(tmem, tmem + 256 + ks * 8, bd, idesc, 1);
          else umma_bf16(tmem, ad, bd, idesc, 1);
        }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      const long long ti = clock64();
      mbar_wait(smem_u32(&bar), rep & 1);
      t1 = clock64();
      if (rep == 2 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const char* names[8] = {"SS A-K  B-K ", "TS      B-K ", "SS A-MN B-K ", "-", "SS A-K  B-MN", "TS      B-MN", "SS A-MN B-MN", "-"};
  const int R = 256;
  for (int N : {16, 64, 128, 256})
    for (int v : {0, 1, 4, 5, 6}) {
      rate_kernel<<<1, 128, 160 * 1024>>>(N, v, R, d);
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("N=%3d %s: %6.1f clk/MMA total, %6.1f clk/MMA issue  (%s)\n", N, names[v],
             (double)h[0] / R, (double)h[1] / R, cudaGetErrorString(e));
    }
  return 0;
}
