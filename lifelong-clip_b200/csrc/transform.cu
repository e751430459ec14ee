// GPU input transform feeding the patch embedding. The reference builds every
// training batch on the host with torchvision (methods/_trainer.py:236-242):
//     Resize((S, S)) -> RandomCrop(S, padding=4) -> RandomHorizontalFlip() -> Normalize(mean, std)
// applied to the whole [B, 3, h, w] float batch at once (one crop offset and one flip decision per
// batch), and ships 224x224 fp32 images over PCIe. Here the raw batch (uint8 0..255 as the dataset
// stores it, or float 0..1) is copied as is and ONE pass produces either the fp32 NCHW tensor
// (drop-in output of `train_transform`) or directly the bf16 im2col rows of the stride-P patch
// convolution (model.py:756-758) that the patch-embedding GEMM reads.
//   bilinear resize: align_corners=False, source index clamped at 0 - the arithmetic of
//   aten/native/UpSample.h (area_pixel_compute_source_index, compute_source_index_and_lambda);
//   crop offsets / flip are read from device memory so a captured CUDA graph can replay with new
//   random draws; the draws themselves stay on the host (torch's global CPU generator, in
//   torchvision's order: i, j, then the flip coin).
#include "common.cuh"

namespace {

struct TxK {
  const void* src;
  int src_u8, h, w, S, pad;
  int crop_i, crop_j, flip;
  const int* dyn;   // device {crop_i, crop_j, flip} or nullptr
  float mean[3], std[3];
  float scale_h, scale_w;
};

struct Tap {
  int i0, i1;
  float l0, l1;
};

__device__ __forceinline__ Tap make_tap(float scale, int dst, int in_size) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  Tap t;
  t.i0 = (int)src;
  if (t.i0 > in_size - 1) t.i0 = in_size - 1;
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.l1 = fminf(fmaxf(src - (float)t.i0, 0.f), 1.f);
  t.l0 = 1.f - t.l1;
  return t;
}

__device__ __forceinline__ float load_px(const TxK& a, size_t plane, int y, int x) {
  const size_t idx = (plane * a.h + y) * a.w + x;
  if (a.src_u8) return __fdiv_rn((float)reinterpret_cast<const uint8_t*>(a.src)[idx], 255.0f);
  return reinterpret_cast<const float*>(a.src)[idx];
}

// value of the transformed image at output pixel (y, x) of plane (n, c)
__device__ __forceinline__ float tx_pixel(const TxK& a, size_t plane, int c, int y, int x, int ci,
                                          int cj, int flip) {
  const int ry = y + ci - a.pad;
  const int rx = (flip ? a.S - 1 - x : x) + cj - a.pad;
  float v = 0.f;   // RandomCrop pads with zeros BEFORE Normalize
  if (ry >= 0 && ry < a.S && rx >= 0 && rx < a.S) {
    const Tap ty = make_tap(a.scale_h, ry, a.h), tx = make_tap(a.scale_w, rx, a.w);
    const float t0 = __fadd_rn(__fmul_rn(tx.l0, load_px(a, plane, ty.i0, tx.i0)),
                               __fmul_rn(tx.l1, load_px(a, plane, ty.i0, tx.i1)));
    const float t1 = __fadd_rn(__fmul_rn(tx.l0, load_px(a, plane, ty.i1, tx.i0)),
                               __fmul_rn(tx.l1, load_px(a, plane, ty.i1, tx.i1)));
    v = __fadd_rn(__fmul_rn(ty.l0, t0), __fmul_rn(ty.l1, t1));
  }
  return __fdiv_rn(__fsub_rn(v, a.mean[c]), a.std[c]);
}

// PATCH = 0: fp32 NCHW out [N, 3, S, S]; PATCH = P: bf16 im2col rows [N*G*G, ld_out]
template <bool PATCHES>
__global__ void __launch_bounds__(256)
transform_kernel(TxK a, int N, int P, void* __restrict__ out, int ld_out) {
  const int ci = a.dyn ? a.dyn[0] : a.crop_i, cj = a.dyn ? a.dyn[1] : a.crop_j;
  const int flip = a.dyn ? a.dyn[2] : a.flip;
  const int S = a.S, W2 = S / 2;
  const int plane = blockIdx.y, n = plane / 3, c = plane - 3 * n;
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= S * W2) return;
  const int y = idx / W2, x = (idx - y * W2) * 2;
  const float v0 = tx_pixel(a, (size_t)plane, c, y, x, ci, cj, flip);
  const float v1 = tx_pixel(a, (size_t)plane, c, y, x + 1, ci, cj, flip);
  if (PATCHES) {
    const int G = S / P;
    const int px = x / P, j = x - px * P, py = y / P, i = y - py * P;
    const size_t row = ((size_t)n * G + py) * G + px;
    const int col = c * P * P + i * P + j;
    *reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(out) + row * ld_out + col) =
        pack_bf16(v0, v1);
    // K padding columns (ViT-L/14: 588 -> 592) are zeroed by the thread that owns pixel (0, 0)
    // of the patch's first plane
    if (c == 0 && i == 0 && j == 0)
      for (int k = 3 * P * P; k < ld_out; ++k)
        reinterpret_cast<__nv_bfloat16*>(out)[row * ld_out + k] = __float2bfloat16_rn(0.f);
  } else {
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + ((size_t)plane * S + y) * S + x) =
        make_float2(v0, v1);
  }
}

int to_k(const llc_img_transform* t, TxK* k, const char* who) {
  LLC_REQUIRE(t && t->src, "%s: null input", who);
  LLC_REQUIRE(t->h > 0 && t->w > 0 && t->out_size > 0 && t->out_size % 2 == 0,
              "%s: bad geometry (%d x %d -> %d)", who, t->h, t->w, t->out_size);
  LLC_REQUIRE(t->h <= t->out_size && t->w <= t->out_size,
              "%s: down-sampling (%d x %d -> %d) needs torchvision's anti-aliased filter, which "
              "this kernel does not implement", who, t->h, t->w, t->out_size);
  LLC_REQUIRE(t->pad >= 0 && t->crop_i >= 0 && t->crop_j >= 0 && t->crop_i <= 2 * t->pad &&
              t->crop_j <= 2 * t->pad, "%s: crop offset outside the padded frame", who);
  k->src = t->src; k->src_u8 = t->src_u8; k->h = t->h; k->w = t->w; k->S = t->out_size;
  k->pad = t->pad; k->crop_i = t->crop_i; k->crop_j = t->crop_j; k->flip = t->flip;
  k->dyn = t->dyn_params;
  for (int c = 0; c < 3; ++c) {
    LLC_REQUIRE(t->std[c] != 0.f, "%s: zero std", who);
    k->mean[c] = t->mean[c]; k->std[c] = t->std[c];
  }
  // area_pixel_compute_scale<float>(input, output, align_corners=false, scale=nullopt)
  k->scale_h = (float)t->h / (float)t->out_size;
  k->scale_w = (float)t->w / (float)t->out_size;
  return 0;
}

}  // namespace

extern "C" int llc_transform_images(const llc_img_transform* t, int N, float* out, void* stream) {
  TxK k;
  if (int rc = to_k(t, &k, "llc_transform_images")) return rc;
  LLC_REQUIRE(out && N > 0 && (size_t)N * 3 <= 65535, "llc_transform_images: bad batch");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = k.S;
  LLC_PROF_BEGIN(LLC_K_EMBED, N, 3 * S * S, 2, 0.0, 4.0 * N * 3 * S * S, st);
  transform_kernel<false><<<dim3((S * (S / 2) + 255) / 256, N * 3), 256, 0, st>>>(k, N, 0, out, 0);
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("transform_kernel");
  return 0;
}

extern "C" int llc_transform_patchify(const llc_img_transform* t, int N, int P, void* out,
                                      int ld_out, void* stream) {
  TxK k;
  if (int rc = to_k(t, &k, "llc_transform_patchify")) return rc;
  LLC_REQUIRE(out && N > 0 && (size_t)N * 3 <= 65535, "llc_transform_patchify: bad batch");
  LLC_REQUIRE(P > 0 && P % 2 == 0 && k.S % P == 0 && ld_out >= 3 * P * P && ld_out % 2 == 0,
              "llc_transform_patchify: image %d / patch %d unsupported", k.S, P);
  cudaStream_t st = (cudaStream_t)stream;
  const int S = k.S;
  LLC_PROF_BEGIN(LLC_K_EMBED, N, 3 * S * S, 3, 0.0, 2.0 * N * 3 * S * S, st);
  transform_kernel<true><<<dim3((S * (S / 2) + 255) / 256, N * 3), 256, 0, st>>>(k, N, P, out,
                                                                               ld_out);
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("transform_kernel");
  return 0;
}
