// Shared device/host helpers for the LifeLong-CLIP B200 hot path (sm_100a only).
// PTX wrappers for mbarrier / TMA / tcgen05 / TMEM, error plumbing for the C-ABI.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/llc.h"

// Development A/B switches (LLC_ATTN_LEGACY, LLC_GEMM_DBG, ...) exist only in -DLLC_DEV builds; the
// shipped library reads no environment variables.
#ifdef LLC_DEV
inline const char* llc_dev_env(const char* name) { return getenv(name); }
#else
inline const char* llc_dev_env(const char*) { return nullptr; }
#endif

// ---------------------------------------------------------------- errors
void llc_set_error(const char* fmt, ...);
int llc_check_cuda(cudaError_t e, const char* what);

#define LLC_CUDA(call)                                         \
  do {                                                         \
    cudaError_t _e = (call);                                   \
    if (_e != cudaSuccess) return llc_check_cuda(_e, #call);   \
  } while (0)

#define LLC_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      llc_set_error(__VA_ARGS__);       \
      return LLC_ERR_ARG;               \
    }                                   \
  } while (0)

#define LLC_LAUNCH_CHECK(name)                                   \
  do {                                                           \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return llc_check_cuda(_e, name);      \
  } while (0)

// Opt a kernel into more than 48 KB of dynamic shared memory, once per DEVICE (the attribute is
// per device; a process-wide flag would leave the second GPU of a process unconfigured).
#define LLC_CONFIGURE_SMEM(kernel, bytes)                                                       \
  do {                                                                                          \
    static int cfg_done_[64];                                                                   \
    int dev_ = 0;                                                                               \
    LLC_CUDA(cudaGetDevice(&dev_));                                                             \
    if (dev_ < 0 || dev_ >= 64 || cfg_done_[dev_] < (int)(bytes)) {                             \
      LLC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                    (int)(bytes)));                                             \
      if (dev_ >= 0 && dev_ < 64) cfg_done_[dev_] = (int)(bytes);                               \
    }                                                                                           \
  } while (0)

// launch counter (bench.py reports it as gpu_launches)
extern unsigned long long g_llc_launches;
#define LLC_COUNT_LAUNCH() (++g_llc_launches)

// per-launch event timing (llc_prof_enable / llc_prof_read)
extern int g_llc_prof_on;
void llc_prof_begin(int kind, int m, int n, int k, double flops, double bytes, cudaStream_t st);
void llc_prof_end(cudaStream_t st);
#define LLC_PROF_BEGIN(kind, m, n, k, flops, bytes, st)                       \
  do {                                                                        \
    if (g_llc_prof_on) llc_prof_begin(kind, m, n, k, flops, bytes, st);       \
  } while (0)
#define LLC_PROF_END(st)                  \
  do {                                    \
    if (g_llc_prof_on) llc_prof_end(st);  \
  } while (0)

// Programmatic dependent launch: kernels launched through llc_launch_pdl may be scheduled while
// the previous kernel of the stream is still draining (its CTAs exit one by one); they run their
// local prologue (barrier init, TMEM allocation, descriptor prefetch) and then block in
// pdl_wait() until the predecessor has completed and its writes are visible. Saves the launch
// latency + tail + prologue (~8 us) between the ~260 kernels of a step. LLC_NO_PDL=1 disables it.
extern int g_llc_pdl;
extern int g_llc_pdl_trigger;
extern int g_llc_traversal;     // llc_set_traversal: bit 0 ln_fwd, bit 1 ln_bwd walk rows from the end   // llc_set_pdl_trigger: persistent kernels trigger dependents early
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
cudaError_t llc_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                           cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_llc_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
// Early trigger: a persistent kernel calls this when a CTA starts its LAST work item. Once every
// CTA of the grid has (or has exited), the next kernel of the stream may be scheduled: its CTAs
// take SMs as ours retire, run their prologue (barriers, TMEM, descriptor prefetch) and park in
// pdl_wait() until this grid has completed and flushed. Hides the launch + prologue latency in our
// tail. (Triggering at kernel START was measured slower in round 1: dependents parked for the
// whole kernel compete with multi-wave kernels' own CTAs.)
__device__ __forceinline__ void pdl_trigger() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

// TMA descriptor encode (driver entry point fetched through the runtime, no -lcuda)
int llc_encode_tmap_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int elem_bytes,
                       uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz);
int llc_encode_tmap_3d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int elem_bytes,
                       uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                       uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                       CUtensorMapSwizzle swz);
int llc_num_sms();

#ifdef __CUDACC__
// ---------------------------------------------------------------- small math
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}
// QuickGELU x*sigmoid(1.702x) (reference models/clip/model.py:203-206) and its derivative
__device__ __forceinline__ float quick_gelu(float x) {
  return x / (1.0f + __expf(-1.702f * x));
}
__device__ __forceinline__ float quick_gelu_grad(float x) {
  float s = 1.0f / (1.0f + __expf(-1.702f * x));
  return s * (1.0f + 1.702f * x * (1.0f - s));
}

// One lane of a fully converged warp. Single-thread instructions (TMA, tcgen05.mma/commit) are
// issued as `if (elect_one()) ...` from warp-uniform code so that their operands stay in uniform
// registers; issuing them from inside `if (lane == 0)` makes ptxas wrap every one in an
// R2UR "waterfall" loop, which measured ~230 cycles per tcgen05.mma (profiles/).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// warp index as a provably warp-uniform value
__device__ __forceinline__ int uniform_warp_idx() {
  return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
}

// ---------------------------------------------------------------- smem / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ unsigned long long llc_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a pipeline bug traps after ~4 s instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3fffu) == 0) {
      unsigned long long t = llc_globaltimer();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) {
        printf("llc: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x,
               threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
// Wait for warps that expect to wait LONG (epilogue / element-wise warps parked until a whole
// tile's MMAs retire): back off with nanosleep between polls instead of spinning - the spinning
// warps compete for issue slots with the single-thread MMA/TMA issuers and burn power on a
// power-capped part.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(100);
    if ((++spins & 0x3ffu) == 0) {
      unsigned long long t = llc_globaltimer();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) {
        printf("llc: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x,
               threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// DRAM -> L2 only (no smem destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(m), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                                 int c0, int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// createpolicy-encoded L2 hints (same constants CUTLASS uses for SM90+ TMA)
static constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
static constexpr uint64_t kEvictLast = 0x14F0000000000000ull;
static constexpr uint64_t kEvictNormal = 0x1000000000000000ull;

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// make the mbarrier track completion of all MMAs issued so far by this thread
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------- CTA pair (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same smem offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default .release.cta semantics as CUTLASS's ClusterBarrier::arrive: a cluster-scope release
  // here costs a full fence (+ L1 invalidate) per call and serialised the pipeline (profiles/)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; completes bytes on an mbarrier that may live in the
// peer CTA (bar is a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* m,
                                                uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(m), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2_hint(uint32_t dst, const CUtensorMap* m,
                                                     uint32_t bar_cluster, int c0, int c1,
                                                     uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      ".L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(m), "r"(bar_cluster), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* m, uint32_t src, int c0, int c1,
                                                  uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(m),
      "r"(src), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
               : "memory");
}
// one MMA over the CTA pair: M = 256 (128 rows per CTA), issued by the leader CTA's thread
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far retire) on the barrier at this smem offset in every CTA
// of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}

// K-major, 128B-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart).
// Field layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor (start>>4 | LBO<<16 | SBO<<32 |
// version=1<<46 | layout_type<<61, SWIZZLE_128B = 2).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;            // LBO: unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO: 8 rows * 128 B
  d |= (uint64_t)1 << 46;            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
// MN-major, 128B-swizzled operand tile: 64 contiguous MN elements (128 B) per K row, 8 K rows
// per 1024 B atom; lbo = byte stride between 64-wide MN atoms, sbo = between 8-row K groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                       uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor, kind::f16, A/B = bf16, D = fp32 (InstrDescriptor in the same header):
// c_format[4,6)=1(F32) a_format[7,10)=1(BF16) b_format[10,13)=1 a_major[15] b_major[16]
// n_dim[17,23)=N>>3 m_dim[24,29)=M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major = 0,
                                                       int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#endif  // __CUDACC__
