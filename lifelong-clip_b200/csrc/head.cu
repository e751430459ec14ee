// Image-text head of the online step, one CTA per sample:
//   ln_post(CLS) @ proj                      reference models/clip/model.py:782-785
//   f = z/|z|, logits = exp(logit_scale) f T^T                    model.py:966-973
//   probs = softmax(logits)                              models/adapter_clip.py:99
//   loss = CrossEntropy(probs, y)  (CE applied to probabilities)  methods/adapter_clip.py:89,
//          criterion methods/_trainer.py:164;  pred = argmax      methods/adapter_clip.py:90
//   class restriction: gather of the visible class rows (methods/adapter_clip.py:53-61,84) or the
//   additive seen-class mask (methods/mvp_clip.py:113-118)
// and the analytic backward down to the CLS rows of the residual stream (proj, ln_post frozen).
// Also the integer label remap (methods/adapter_clip.py:75-76) and the per-step loss/acc scalars.
#include <stdlib.h>

#include "common.cuh"

namespace {

// 1024 threads per sample: the two fat loops (y @ proj and dlogits @ T) are chains of L2-latency
// bound loads; with 256 threads a sample took ~190 us whatever the batch (a third of a 32-image
// step's critical path). Every loop below is split so that all 32 warps hold loads in flight.
constexpr int kThreads = 1024;
constexpr int kMaxSplit = 8;

// number of slices the reduction dimension of an [n_out]-wide loop is cut into so that
// n_out * split covers the CTA (power of two, <= kMaxSplit)
__device__ __forceinline__ int split_for(int n_out) {
  int s = 1;
  while (s < kMaxSplit && n_out * s * 2 <= kThreads) s *= 2;
  return s;
}
constexpr float kLnEps = 1e-5f;

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) s += red[i];
  return s;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = -INFINITY;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) s = fmaxf(s, red[i]);
  return s;
}

struct HeadK {
  const float* x; int cls_stride, ld_x;
  const float *ln_g, *ln_b, *proj, *text;
  const int64_t* cls_idx;
  const float* add_mask;
  float logit_scale;
  int N, D, E, C;
  const int64_t* labels;
  int double_softmax;
  float inv_batch;
  float *feat, *fnorm, *logits, *probs, *loss_rows;
  int64_t* pred;
  const float* d_feat;
  int skip_logit_grad;
  const int64_t* row_idx;   // row of sample n in x / dx (text tower: the EOT token), else n*cls_stride
  const float* d_fnorm;     // backward: gradient w.r.t. the NORMALISED features (text side)
  float* dlogits;           // backward: dL/dlogits [N, C] written out (for llc_head_dtext)
  int d_is_logits;          // backward: the incoming gradient is w.r.t. the LOGITS, not the probs
  int rot;   // bit 0: rotate the proj row order per CTA, bit 1: the text row order (LLC_HEAD_ROT)
};

// One sample per CTA (kS = 1; more samples per CTA measured slower at N = 256: the loops are
// latency-, not L2-bandwidth-bound, and fewer CTAs cover less of the machine).
constexpr int kS = 1;

// smem: y[kS][D] | f[kS][E] | p[kS][C] | red[32]
__global__ void __launch_bounds__(kThreads) head_fwd_kernel(HeadK a) {
  extern __shared__ float sm[];
  float* sy = sm;
  float* sf = sy + kS * a.D;
  float* sp = sf + kS * a.E;
  float* red = sp + kS * a.C;
  __shared__ int s_arg;
  const int n0 = blockIdx.x * kS, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ln_post of each sample's CLS row
  for (int sI = 0; sI < kS; ++sI) {
    const int n = min(n0 + sI, a.N - 1);     // tail samples recompute the last one (not stored)
    const float* xr = a.x + (a.row_idx ? (size_t)a.row_idx[n] : (size_t)n * a.cls_stride) * a.ld_x;
    float* y = sy + sI * a.D;
    float s = 0.f;
    for (int k = tid; k < a.D; k += kThreads) { y[k] = xr[k]; s += y[k]; }
    const float mean = block_sum(s, red) / a.D;
    float q = 0.f;
    for (int k = tid; k < a.D; k += kThreads) { const float d = y[k] - mean; q += d * d; }
    const float rstd = rsqrtf(block_sum(q, red) / a.D + kLnEps);
    for (int k = tid; k < a.D; k += kThreads) y[k] = (y[k] - mean) * rstd * a.ln_g[k] + a.ln_b[k];
  }
  __syncthreads();

  // z = y @ proj: thread (e, slice) accumulates its slice of the D rows for output e; the
  // slices are summed through shared memory in a fixed order
  {
    float* part = red + 32;                         // [split][E]
    const int split = split_for(a.E);
    const int per = kThreads / split;               // threads per slice
    const int sl = tid / per, e0 = tid - sl * per;
    const int kb = (int)((long long)a.D * sl / split), ke = (int)((long long)a.D * (sl + 1) / split);
    // every CTA walks the proj rows from its own starting row: in lock step all CTAs would ask
    // the same L2 lines at the same time (one slice serving every requester per line)
    const int rot = (a.rot & 1) ? (int)((blockIdx.x * 6u) % (unsigned)(ke - kb)) : 0;
    for (int e = e0; e < a.E; e += per) {
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      int i = 0;
      const int len = ke - kb;
#pragma unroll 2
      for (; i + 4 <= len; i += 4) {
        int k0 = i + rot; if (k0 >= len) k0 -= len;
        int k1 = k0 + 1; if (k1 >= len) k1 -= len;
        int k2 = k1 + 1; if (k2 >= len) k2 -= len;
        int k3 = k2 + 1; if (k3 >= len) k3 -= len;
        const float p0 = __ldg(a.proj + (size_t)(kb + k0) * a.E + e);
        const float p1 = __ldg(a.proj + (size_t)(kb + k1) * a.E + e);
        const float p2 = __ldg(a.proj + (size_t)(kb + k2) * a.E + e);
        const float p3 = __ldg(a.proj + (size_t)(kb + k3) * a.E + e);
        acc0 = fmaf(sy[kb + k0], p0, acc0);
        acc1 = fmaf(sy[kb + k1], p1, acc1);
        acc2 = fmaf(sy[kb + k2], p2, acc2);
        acc3 = fmaf(sy[kb + k3], p3, acc3);
      }
      for (; i < len; ++i) {
        int k0 = i + rot; if (k0 >= len) k0 -= len;
        acc0 = fmaf(sy[kb + k0], __ldg(a.proj + (size_t)(kb + k0) * a.E + e), acc0);
      }
      part[sl * a.E + e] = (acc0 + acc1) + (acc2 + acc3);
    }
    __syncthreads();
    float nrm = 0.f;
    for (int e = tid; e < a.E; e += kThreads) {
      float v = 0.f;
      for (int j = 0; j < split; ++j) v += part[j * a.E + e];
      sf[e] = v;
      a.feat[(size_t)n0 * a.E + e] = v;
      nrm += v * v;
    }
    const float inv_norm = 1.0f / sqrtf(block_sum(nrm, red));
    for (int e = tid; e < a.E; e += kThreads) {
      const float v = sf[e] * inv_norm;
      sf[e] = v;
      a.fnorm[(size_t)n0 * a.E + e] = v;
    }
  }
  __syncthreads();

  // logits: one warp per class, lanes stride the embedding, kS samples per text row read
  const int c0 = (a.rot & 2) ? (int)((blockIdx.x * 3u) % (unsigned)a.C) : 0;   // as for proj
  for (int ci = warp; ci < a.C; ci += kThreads / 32) {
    int c = ci + c0;
    if (c >= a.C) c -= a.C;
    const int64_t row = a.cls_idx ? a.cls_idx[c] : (int64_t)c;
    const float* tr = a.text + (size_t)row * a.E;
    float acc[kS];
#pragma unroll
    for (int sI = 0; sI < kS; ++sI) acc[sI] = 0.f;
    for (int e = lane; e < a.E; e += 32) {
      const float t = __ldg(tr + e);
#pragma unroll
      for (int sI = 0; sI < kS; ++sI) acc[sI] = fmaf(sf[sI * a.E + e], t, acc[sI]);
    }
#pragma unroll
    for (int sI = 0; sI < kS; ++sI) {
      float v = warp_sum(acc[sI]) * a.logit_scale;
      if (a.add_mask) v += a.add_mask[c];
      if (lane == 0) {
        sp[sI * a.C + c] = v;
        if (n0 + sI < a.N) a.logits[(size_t)(n0 + sI) * a.C + c] = v;
      }
    }
  }
  __syncthreads();

  for (int sI = 0; sI < kS; ++sI) {
    const int n = n0 + sI;
    if (n >= a.N) break;               // uniform across the block
    float* p_ = sp + sI * a.C;
    // softmax
    float m = -INFINITY;
    for (int c = tid; c < a.C; c += kThreads) m = fmaxf(m, p_[c]);
    m = block_max(m, red);
    float z = 0.f;
    for (int c = tid; c < a.C; c += kThreads) z += __expf(p_[c] - m);
    z = block_sum(z, red);
    const float logz = m + logf(z);
    float pm = -1.f;
    for (int c = tid; c < a.C; c += kThreads) {
      const float p = __expf(p_[c] - m) / z;
      p_[c] = p;
      a.probs[(size_t)n * a.C + c] = p;
      pm = fmaxf(pm, p);
    }
    pm = block_max(pm, red);
    // argmax: lowest index among the maxima
    if (tid == 0) s_arg = 0x7fffffff;
    __syncthreads();
    for (int c = tid; c < a.C; c += kThreads)
      if (p_[c] == pm) atomicMin(&s_arg, c);
    __syncthreads();
    if (tid == 0 && a.pred) a.pred[n] = (int64_t)s_arg;

    if (a.labels && a.loss_rows) {
      const int64_t yv = a.labels[n];
      float loss;
      if (a.double_softmax) {
        // CE on probabilities: -p_y + log sum_j exp(p_j)
        float z2 = 0.f;
        for (int c = tid; c < a.C; c += kThreads) z2 += __expf(p_[c]);
        z2 = block_sum(z2, red);
        const float py = (yv >= 0 && yv < a.C) ? p_[yv] : 0.f;
        loss = -py + logf(z2);
      } else {
        const float ly = (yv >= 0 && yv < a.C) ? a.logits[(size_t)n * a.C + yv] : 0.f;
        loss = -(ly - logz);
      }
      if (tid == 0) a.loss_rows[n] = loss * a.inv_batch;
    }
    __syncthreads();
  }
}

// smem: g[kS][C] | f[kS][E] | z[kS][E](df then dz) | dy[kS][D] | xh[kS][D] | red[32]
__global__ void __launch_bounds__(kThreads)
head_bwd_kernel(HeadK a, const float* __restrict__ d_probs, float loss_scale,
                float* __restrict__ dx, int ld_dx) {
  extern __shared__ float sm[];
  float* sg = sm;
  float* sf = sg + kS * a.C;
  float* sz = sf + kS * a.E;
  float* sdy = sz + kS * a.E;
  float* sxh = sdy + kS * a.D;
  float* red = sxh + kS * a.D;
  const int n0 = blockIdx.x * kS, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (!a.skip_logit_grad) {
    // dL/dlogits of each sample
    for (int sI = 0; sI < kS; ++sI) {
      const int n = min(n0 + sI, a.N - 1);
      const float* pr = a.probs + (size_t)n * a.C;
      float* g = sg + sI * a.C;
      if (d_probs != nullptr && a.d_is_logits) {
        for (int c = tid; c < a.C; c += kThreads) g[c] = d_probs[(size_t)n * a.C + c] * loss_scale;
        __syncthreads();
      } else if (d_probs == nullptr && !a.double_softmax) {
        const int64_t yv = a.labels[n];
        for (int c = tid; c < a.C; c += kThreads)
          g[c] = (pr[c] - (c == yv ? 1.f : 0.f)) * a.inv_batch * loss_scale;
      } else {
        if (d_probs != nullptr) {
          for (int c = tid; c < a.C; c += kThreads) g[c] = d_probs[(size_t)n * a.C + c] * loss_scale;
        } else {
          const int64_t yv = a.labels[n];
          float z2 = 0.f;
          for (int c = tid; c < a.C; c += kThreads) z2 += __expf(pr[c]);
          z2 = block_sum(z2, red);
          for (int c = tid; c < a.C; c += kThreads)
            g[c] = (__expf(pr[c]) / z2 - (c == yv ? 1.f : 0.f)) * a.inv_batch * loss_scale;
        }
        __syncthreads();
        float dot = 0.f;
        for (int c = tid; c < a.C; c += kThreads) dot += g[c] * pr[c];
        dot = block_sum(dot, red);
        for (int c = tid; c < a.C; c += kThreads) g[c] = pr[c] * (g[c] - dot);
      }
      if (a.dlogits && n0 + sI < a.N) {
        for (int c = tid; c < a.C; c += kThreads) a.dlogits[(size_t)n * a.C + c] = g[c];
      }
      for (int e = tid; e < a.E; e += kThreads) sf[sI * a.E + e] = a.fnorm[(size_t)n * a.E + e];
    }
    __syncthreads();

    // df = scale * dlogits @ T: thread (e, slice) sums its slice of the classes
    float fd[kS], zz[kS];
    fd[0] = zz[0] = 0.f;
    {
      float* part = red + 32;                         // [split][E]
      const int split = split_for(a.E);
      const int per = kThreads / split;
      const int sl = tid / per, e0 = tid - sl * per;
      const int cb = (int)((long long)a.C * sl / split), ce = (int)((long long)a.C * (sl + 1) / split);
      for (int e = e0; e < a.E; e += per) {
        float acc0 = 0.f, acc1 = 0.f;
        int c = cb;
#pragma unroll 4
        for (; c + 2 <= ce; c += 2) {
          const int64_t r0 = a.cls_idx ? a.cls_idx[c] : (int64_t)c;
          const int64_t r1 = a.cls_idx ? a.cls_idx[c + 1] : (int64_t)(c + 1);
          acc0 = fmaf(sg[c], __ldg(a.text + (size_t)r0 * a.E + e), acc0);
          acc1 = fmaf(sg[c + 1], __ldg(a.text + (size_t)r1 * a.E + e), acc1);
        }
        if (c < ce) {
          const int64_t r0 = a.cls_idx ? a.cls_idx[c] : (int64_t)c;
          acc0 = fmaf(sg[c], __ldg(a.text + (size_t)r0 * a.E + e), acc0);
        }
        part[sl * a.E + e] = acc0 + acc1;
      }
      __syncthreads();
      const int n = n0;
      for (int e = tid; e < a.E; e += kThreads) {
        float v = 0.f;
        for (int j = 0; j < split; ++j) v += part[j * a.E + e];
        v *= a.logit_scale;
        sz[e] = v;
        fd[0] += v * sf[e];
        const float zf = a.feat[(size_t)n * a.E + e];
        zz[0] += zf * zf;
      }
    }
    // dz = (df - f (f.df)) / |z|
#pragma unroll
    for (int sI = 0; sI < kS; ++sI) {
      const int n = min(n0 + sI, a.N - 1);
      const float fds = block_sum(fd[sI], red);
      const float inv_norm = 1.0f / sqrtf(block_sum(zz[sI], red));
      for (int e = tid; e < a.E; e += kThreads) {
        float v = (sz[sI * a.E + e] - sf[sI * a.E + e] * fds) * inv_norm;
        if (a.d_feat) v += a.d_feat[(size_t)n * a.E + e] * loss_scale;
        sz[sI * a.E + e] = v;
      }
    }
  } else if (a.d_fnorm) {
    // gradient arrives w.r.t. f = z/|z| (the text side of the logit product):
    // dz = (df - f (f.df)) / |z|  [+ d_feat]
    for (int sI = 0; sI < kS; ++sI) {
      const int n = min(n0 + sI, a.N - 1);
      float fd = 0.f, zz = 0.f;
      for (int e = tid; e < a.E; e += kThreads) {
        const float f = a.fnorm[(size_t)n * a.E + e];
        const float v = a.d_fnorm[(size_t)n * a.E + e] * loss_scale;
        const float zf = a.feat[(size_t)n * a.E + e];
        sf[sI * a.E + e] = f;
        sz[sI * a.E + e] = v;
        fd += v * f;
        zz += zf * zf;
      }
      fd = block_sum(fd, red);
      const float inv_norm = 1.0f / sqrtf(block_sum(zz, red));
      for (int e = tid; e < a.E; e += kThreads) {
        float v = (sz[sI * a.E + e] - sf[sI * a.E + e] * fd) * inv_norm;
        if (a.d_feat) v += a.d_feat[(size_t)n * a.E + e] * loss_scale;
        sz[sI * a.E + e] = v;
      }
    }
  } else {
    for (int sI = 0; sI < kS; ++sI) {
      const int n = min(n0 + sI, a.N - 1);
      for (int e = tid; e < a.E; e += kThreads)
        sz[sI * a.E + e] = a.d_feat[(size_t)n * a.E + e] * loss_scale;
    }
  }
  __syncthreads();

  // dy = dz @ proj^T : one warp per k, every proj row read once for the kS samples (rows
  // rotated per CTA, see head_fwd_kernel)
  const int kr0 = (a.rot & 1) ? (int)((blockIdx.x * 6u) % (unsigned)a.D) : 0;
  for (int ki = warp; ki < a.D; ki += kThreads / 32) {
    int k = ki + kr0;
    if (k >= a.D) k -= a.D;
    const float* prow = a.proj + (size_t)k * a.E;
    float acc[kS];
#pragma unroll
    for (int sI = 0; sI < kS; ++sI) acc[sI] = 0.f;
    for (int e = lane; e < a.E; e += 32) {
      const float pv = __ldg(prow + e);
#pragma unroll
      for (int sI = 0; sI < kS; ++sI) acc[sI] = fmaf(sz[sI * a.E + e], pv, acc[sI]);
    }
#pragma unroll
    for (int sI = 0; sI < kS; ++sI) {
      const float v = warp_sum(acc[sI]);
      if (lane == 0) sdy[sI * a.D + k] = v;
    }
  }
  __syncthreads();
  // ln_post backward (input gradient only)
  for (int sI = 0; sI < kS; ++sI) {
    const int n = n0 + sI;
    if (n >= a.N) break;
    const size_t xrow = a.row_idx ? (size_t)a.row_idx[n] : (size_t)n * a.cls_stride;
    const float* xr = a.x + xrow * a.ld_x;
    float* xh_ = sxh + sI * a.D;
    float* dy_ = sdy + sI * a.D;
    float s = 0.f;
    for (int k = tid; k < a.D; k += kThreads) { xh_[k] = xr[k]; s += xh_[k]; }
    const float mean = block_sum(s, red) / a.D;
    float q = 0.f;
    for (int k = tid; k < a.D; k += kThreads) { const float d = xh_[k] - mean; q += d * d; }
    const float rstd = rsqrtf(block_sum(q, red) / a.D + kLnEps);
    float c1 = 0.f, c2 = 0.f;
    for (int k = tid; k < a.D; k += kThreads) {
      const float xh = (xh_[k] - mean) * rstd;
      const float g = dy_[k] * a.ln_g[k];
      xh_[k] = xh;
      dy_[k] = g;
      c1 += g;
      c2 += g * xh;
    }
    c1 = block_sum(c1, red) / a.D;
    c2 = block_sum(c2, red) / a.D;
    float* dr = dx + xrow * ld_dx;
    for (int k = tid; k < a.D; k += kThreads) dr[k] = rstd * (dy_[k] - c1 - xh_[k] * c2);
  }
}

__global__ void label_remap_kernel(const int64_t* __restrict__ y, const int64_t* __restrict__ lut,
                                   int lut_size, int64_t* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = y[i];
  out[i] = (v >= 0 && v < lut_size) ? lut[v] : (int64_t)-1;
}

__global__ void loss_acc_kernel(const float* __restrict__ loss_rows,
                                const int64_t* __restrict__ pred,
                                const int64_t* __restrict__ labels, int n, float* __restrict__ out) {
  __shared__ float red[32];
  float l = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    l += loss_rows[i];
    c += (pred[i] == labels[i]) ? 1.f : 0.f;
  }
  l = block_sum(l, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) { out[0] = l; out[1] = c; }
}

// d_text[c, e] = scale * sum_n dlogits[n, c] * fnorm[n, e]: the gradient of the logit product
// (model.py:972) w.r.t. the normalised text row of visible class c. One CTA per class, samples in
// a fixed order: deterministic.
__global__ void __launch_bounds__(kThreads)
head_dtext_kernel(const float* __restrict__ dlogits, const float* __restrict__ fnorm, int N, int C,
                  int E, float scale, float* __restrict__ d_text) {
  const int c = blockIdx.x;
  for (int e = threadIdx.x; e < E; e += kThreads) {
    float acc = 0.f;
    for (int n = 0; n < N; ++n)
      acc = fmaf(__ldg(dlogits + (size_t)n * C + c), __ldg(fnorm + (size_t)n * E + e), acc);
    d_text[(size_t)c * E + e] = acc * scale;
  }
}

int to_k(const llc_head_args* a, HeadK* k, const char* who) {
  LLC_REQUIRE(a && a->x && a->ln_g && a->ln_b && a->proj && a->text, "%s: null input", who);
  LLC_REQUIRE(a->N > 0 && a->D > 0 && a->E > 0 && a->C > 0, "%s: empty problem", who);
  LLC_REQUIRE(a->feat && a->fnorm && a->logits && a->probs, "%s: null output", who);
  k->x = a->x; k->cls_stride = a->cls_stride; k->ld_x = a->ld_x;
  k->ln_g = a->ln_g; k->ln_b = a->ln_b; k->proj = a->proj; k->text = a->text;
  k->cls_idx = a->cls_idx; k->add_mask = a->add_mask; k->logit_scale = a->logit_scale;
  k->N = a->N; k->D = a->D; k->E = a->E; k->C = a->C; k->labels = a->labels;
  k->double_softmax = a->double_softmax; k->inv_batch = a->inv_batch;
  k->feat = a->feat; k->fnorm = a->fnorm; k->logits = a->logits; k->probs = a->probs;
  k->loss_rows = a->loss_rows; k->pred = a->pred;
  k->d_feat = a->d_feat; k->skip_logit_grad = a->skip_logit_grad;
  k->row_idx = a->row_idx; k->d_fnorm = a->d_fnorm; k->dlogits = a->dlogits;
  k->d_is_logits = a->d_is_logits;
  static const int rot = llc_dev_env("LLC_HEAD_ROT") ? atoi(llc_dev_env("LLC_HEAD_ROT")) : 1;
  k->rot = rot;
  return 0;
}

}  // namespace

extern "C" int llc_head_fwd(const llc_head_args* a, void* stream) {
  HeadK k;
  if (int rc = to_k(a, &k, "llc_head_fwd")) return rc;
  const size_t smem = (size_t)(kS * (a->D + a->E + a->C) + 32 + kMaxSplit * a->E) * sizeof(float);
  LLC_REQUIRE(smem <= 200 * 1024, "llc_head_fwd: D+E+C too large for one CTA");
  if (smem > 48 * 1024) LLC_CONFIGURE_SMEM(head_fwd_kernel, smem);
  LLC_PROF_BEGIN(LLC_K_HEAD, a->N, a->C, 0, 2.0 * a->N * ((double)a->D * a->E + (double)a->E * a->C),
                 4.0 * a->N * (a->D + 2 * a->E + 2 * a->C), (cudaStream_t)stream);
  head_fwd_kernel<<<(a->N + kS - 1) / kS, kThreads, smem, (cudaStream_t)stream>>>(k);
  LLC_PROF_END((cudaStream_t)stream);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("head_fwd_kernel");
  return 0;
}

extern "C" int llc_head_bwd(const llc_head_args* a, const float* d_probs, float loss_scale,
                            float* dx, int ld_dx, void* stream) {
  HeadK k;
  if (int rc = to_k(a, &k, "llc_head_bwd")) return rc;
  LLC_REQUIRE(dx && ld_dx >= a->D, "llc_head_bwd: bad dx");
  LLC_REQUIRE(d_probs || a->labels || (a->skip_logit_grad && (a->d_feat || a->d_fnorm)),
              "llc_head_bwd: need d_probs, labels or d_feat / d_fnorm");
  LLC_REQUIRE(!a->skip_logit_grad || a->d_feat || a->d_fnorm,
              "llc_head_bwd: skip_logit_grad needs d_feat or d_fnorm");
  const size_t smem =
      (size_t)(kS * (a->C + 2 * a->E + 2 * a->D) + 32 + kMaxSplit * a->E) * sizeof(float);
  LLC_REQUIRE(smem <= 200 * 1024, "llc_head_bwd: sizes too large for one CTA");
  if (smem > 48 * 1024) LLC_CONFIGURE_SMEM(head_bwd_kernel, smem);
  LLC_PROF_BEGIN(LLC_K_HEAD, a->N, a->C, 1, 2.0 * a->N * ((double)a->D * a->E + (double)a->E * a->C),
                 4.0 * a->N * (2 * a->D + 2 * a->E + a->C), (cudaStream_t)stream);
  head_bwd_kernel<<<(a->N + kS - 1) / kS, kThreads, smem, (cudaStream_t)stream>>>(
      k, d_probs, loss_scale, dx, ld_dx);
  LLC_PROF_END((cudaStream_t)stream);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("head_bwd_kernel");
  return 0;
}

extern "C" int llc_head_dtext(const float* dlogits, const float* fnorm, int N, int C, int E,
                              float scale, float* d_text, void* stream) {
  LLC_REQUIRE(dlogits && fnorm && d_text && N > 0 && C > 0 && E > 0, "llc_head_dtext: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  LLC_PROF_BEGIN(LLC_K_HEAD, N, C, 4, 2.0 * N * C * E, 4.0 * ((double)N * C + (double)N * E + (double)C * E), st);
  head_dtext_kernel<<<C, kThreads, 0, st>>>(dlogits, fnorm, N, C, E, scale, d_text);
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("head_dtext_kernel");
  return 0;
}

extern "C" int llc_label_remap(const int64_t* y_global, const int64_t* lut, int lut_size,
                               int64_t* y_local, int n, void* stream) {
  LLC_REQUIRE(n >= 0, "llc_label_remap: negative count");
  if (n == 0) return 0;  // empty batch: nothing to do, pointers may be NULL
  LLC_REQUIRE(y_global && lut && y_local && lut_size > 0, "llc_label_remap: bad args");
  label_remap_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(y_global, lut, lut_size,
                                                                        y_local, n);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("label_remap_kernel");
  return 0;
}

extern "C" int llc_loss_acc(const float* loss_rows, const int64_t* pred, const int64_t* labels,
                            int n, float* out2, void* stream) {
  LLC_REQUIRE(loss_rows && pred && labels && out2 && n > 0, "llc_loss_acc: bad args");
  loss_acc_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(loss_rows, pred, labels, n, out2);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("loss_acc_kernel");
  return 0;
}
