// Evaluation tail of the online protocol, on the device:
//   per-task counters of methods/_trainer.py:519-534 (_interpret_pred: bin = y // n_tasks, ten
//   bins) and the confusion matrix of methods/adapter_clip.py:157-166 accumulated with integer
//   atomics instead of per-batch .tolist() + host sklearn;
//   row L2 normalisation (model.py:966-969) emitting the bf16 operand of the tensor-core logit
//   GEMM, and the row softmax + arg-max over [N, C] logits (models/adapter_clip.py:99,
//   methods/adapter_clip.py:149) for evaluation-sized heads (BASELINE config 5: 4096 x 1000).
#include "common.cuh"

namespace {

__global__ void eval_accum_kernel(const int64_t* __restrict__ y, const int64_t* __restrict__ pred,
                                  int n, int n_tasks, int n_classes,
                                  unsigned long long* __restrict__ cm,
                                  unsigned long long* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t yi = y[i], pi = pred[i];
  // bin = y // n_tasks (labels are non-negative); bins past the reference's ten land in slot 10
  int64_t bin = yi >= 0 ? yi / n_tasks : 10;
  if (bin > 10) bin = 10;
  atomicAdd(counts + bin, 1ull);
  if (yi == pi) atomicAdd(counts + 11 + bin, 1ull);
  if (cm && yi >= 0 && yi < n_classes && pi >= 0 && pi < n_classes)
    atomicAdd(cm + yi * n_classes + pi, 1ull);
}

// one warp per row
__global__ void __launch_bounds__(256)
l2norm_rows_kernel(const float* __restrict__ x, int ld_x, int N, int E, float scale,
                   float* __restrict__ y, int ld_y, __nv_bfloat16* __restrict__ yb, int ld_yb) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* xr = x + (size_t)row * ld_x;
  float s = 0.f;
  for (int e = lane; e < E; e += 32) { const float v = xr[e]; s += v * v; }
  const float inv = scale / sqrtf(warp_sum(s));
  for (int e = lane; e < E; e += 32) {
    const float v = xr[e] * inv;
    if (y) y[(size_t)row * ld_y + e] = v;
    if (yb) yb[(size_t)row * ld_yb + e] = __float2bfloat16_rn(v);
  }
}

// one CTA (128 threads) per row: softmax over C logits (+ additive mask), arg-max = lowest index
// among the maxima (torch.argmax)
__global__ void __launch_bounds__(128)
softmax_argmax_kernel(const float* __restrict__ logits, int ld, int C,
                      const float* __restrict__ add_mask, float* __restrict__ probs, int ld_p,
                      int64_t* __restrict__ pred) {
  __shared__ float red[4];
  __shared__ int redi[4];
  const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* lr = logits + (size_t)row * ld;
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int c = tid; c < C; c += 128) {
    const float v = lr[c] + (add_mask ? add_mask[c] : 0.f);
    if (v > m) { m = v; am = c; }
  }
  // (max, lowest index) reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const int a2 = __shfl_xor_sync(0xffffffffu, am, o);
    if (m2 > m || (m2 == m && a2 < am)) { m = m2; am = a2; }
  }
  if (lane == 0) { red[warp] = m; redi[warp] = am; }
  __syncthreads();
  m = red[0]; am = redi[0];
#pragma unroll
  for (int w = 1; w < 4; ++w)
    if (red[w] > m || (red[w] == m && redi[w] < am)) { m = red[w]; am = redi[w]; }
  __syncthreads();
  float z = 0.f;
  for (int c = tid; c < C; c += 128) z += __expf(lr[c] + (add_mask ? add_mask[c] : 0.f) - m);
  z = warp_sum(z);
  if (lane == 0) red[warp] = z;
  __syncthreads();
  z = red[0] + red[1] + red[2] + red[3];
  if (probs) {
    const float inv = 1.0f / z;
    for (int c = tid; c < C; c += 128)
      probs[(size_t)row * ld_p + c] = __expf(lr[c] + (add_mask ? add_mask[c] : 0.f) - m) * inv;
  }
  if (tid == 0 && pred) pred[row] = (int64_t)am;
}

}  // namespace

extern "C" int llc_eval_accum(const int64_t* y, const int64_t* pred, int n, int n_tasks,
                              int n_classes, unsigned long long* cm, unsigned long long* counts,
                              void* stream) {
  LLC_REQUIRE(n >= 0 && n_tasks > 0 && counts, "llc_eval_accum: bad args");
  if (n == 0) return 0;
  LLC_REQUIRE(y && pred && (cm == nullptr || n_classes > 0), "llc_eval_accum: null input");
  eval_accum_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(y, pred, n, n_tasks,
                                                                       n_classes, cm, counts);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("eval_accum_kernel");
  return 0;
}

extern "C" int llc_l2norm_rows(const float* x, int ld_x, int N, int E, float scale, float* y,
                               int ld_y, void* y_bf16, int ld_yb, void* stream) {
  LLC_REQUIRE(x && N > 0 && E > 0 && (y || y_bf16), "llc_l2norm_rows: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  LLC_PROF_BEGIN(LLC_K_HEAD, N, E, 2, 3.0 * N * E, 4.0 * N * E + (y ? 4.0 : 0.0) * N * E +
                 (y_bf16 ? 2.0 : 0.0) * N * E, st);
  l2norm_rows_kernel<<<(N + 7) / 8, 256, 0, st>>>(x, ld_x, N, E, scale, y, ld_y,
                                                  reinterpret_cast<__nv_bfloat16*>(y_bf16), ld_yb);
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("l2norm_rows_kernel");
  return 0;
}

extern "C" int llc_softmax_argmax(const float* logits, int ld, int N, int C, const float* add_mask,
                                  float* probs, int ld_p, int64_t* pred, void* stream) {
  LLC_REQUIRE(logits && N > 0 && C > 0 && ld >= C && (probs == nullptr || ld_p >= C),
              "llc_softmax_argmax: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  LLC_PROF_BEGIN(LLC_K_HEAD, N, C, 3, 4.0 * N * C, (probs ? 8.0 : 4.0) * N * C, st);
  softmax_argmax_kernel<<<N, 128, 0, st>>>(logits, ld, C, add_mask, probs, ld_p, pred);
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("softmax_argmax_kernel");
  return 0;
}
