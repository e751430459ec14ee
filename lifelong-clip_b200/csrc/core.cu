// Library plumbing: error slot, device check, TMA descriptor encoding, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_llc_launches = 0;
int g_llc_pdl = llc_dev_env("LLC_NO_PDL") == nullptr;
int g_llc_pdl_trigger = 0;
int g_llc_traversal = 0;
extern "C" int llc_set_traversal(int mask) {
  const int old = g_llc_traversal;
  g_llc_traversal = mask;
  return old;
}
extern "C" int llc_set_pdl_trigger(int on) {
  const int old = g_llc_pdl_trigger;
  g_llc_pdl_trigger = on != 0;
  return old;
}

void llc_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int llc_check_cuda(cudaError_t e, const char* what) {
  llc_set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

extern "C" int llc_version(void) { return LLC_VERSION; }
extern "C" const char* llc_last_error(void) { return g_err; }
extern "C" unsigned long long llc_launch_count(void) { return g_llc_launches; }

// ---------------------------------------------------------------- per-launch event timing
int g_llc_prof_on = 0;
namespace {
struct ProfSlot {
  llc_prof_rec rec;
  cudaEvent_t e0, e1;
};
std::vector<ProfSlot> g_prof;
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t take_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void release_all() {
  for (auto& s : g_prof) {
    g_event_pool.push_back(s.e0);
    g_event_pool.push_back(s.e1);
  }
  g_prof.clear();
}
}  // namespace

void llc_prof_begin(int kind, int m, int n, int k, double flops, double bytes, cudaStream_t st) {
  ProfSlot s;
  s.rec.kind = kind; s.rec.m = m; s.rec.n = n; s.rec.k = k;
  s.rec.ms = 0.f; s.rec.flops = flops; s.rec.bytes = bytes;
  s.e0 = take_event();
  s.e1 = take_event();
  cudaEventRecord(s.e0, st);
  g_prof.push_back(s);
}
void llc_prof_end(cudaStream_t st) {
  if (!g_prof.empty()) cudaEventRecord(g_prof.back().e1, st);
}

extern "C" int llc_prof_enable(int on) {
  if (on) release_all();
  g_llc_prof_on = on ? 1 : 0;
  return 0;
}

extern "C" int llc_prof_read(llc_prof_rec* out, int max) {
  for (auto& s : g_prof) {
    cudaError_t e = cudaEventSynchronize(s.e1);
    if (e != cudaSuccess) return -llc_check_cuda(e, "llc_prof_read: cudaEventSynchronize");
    e = cudaEventElapsedTime(&s.rec.ms, s.e0, s.e1);
    if (e != cudaSuccess) return -llc_check_cuda(e, "llc_prof_read: cudaEventElapsedTime");
  }
  const int n = (int)g_prof.size();
  if (out != nullptr)
    for (int i = 0; i < n && i < max; ++i) out[i] = g_prof[i].rec;
  return n;
}

extern "C" int llc_check_device(int dev) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return llc_check_cuda(e, "cudaGetDeviceCount");
  LLC_REQUIRE(dev >= 0 && dev < n, "llc_check_device: device %d of %d", dev, n);
  int major = 0, minor = 0;
  LLC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  LLC_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    llc_set_error("llc: device %d is sm_%d%d; this library holds sm_100a code only (no fallback)",
                  dev, major, minor);
    return LLC_ERR_ARCH;
  }
  return 0;
}

int llc_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return sms;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
  }
  return fn;
}

static int encode(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int rank,
                  const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                  CUtensorMapSwizzle swz) {
  PFN_encodeTiled fn = get_encode();
  if (fn == nullptr) {
    llc_set_error("llc: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return LLC_ERR_ARCH;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    llc_set_error("llc: cuTensorMapEncodeTiled failed (CUresult %d) base=%p rank=%d dims=%llu,%llu "
                  "stride=%llu box=%u,%u",
                  (int)r, base, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)strides[0], box[0], box[1]);
    return LLC_ERR_ARG;
  }
  return 0;
}

int llc_encode_tmap_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int elem_bytes,
                       uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swz) {
  (void)elem_bytes;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  return encode(map, base, dt, 2, dims, strides, box, swz);
}

int llc_encode_tmap_3d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, int elem_bytes,
                       uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                       uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2,
                       CUtensorMapSwizzle swz) {
  (void)elem_bytes;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  return encode(map, base, dt, 3, dims, strides, box, swz);
}
