// Host-side orchestration of the image tower: the launch sequence of one
// ResidualAttentionBlock_LoRA (reference models/clip/model.py:233-236,400-415 ->
// models/clip/lora.py:825-840,950,1002-1074) forward and backward, and of the whole
// VisualTransformer.forward (model.py:755-787). Pure launch code: every tensor lives in a caller
// owned arena whose layout is defined here, so one C call replaces ~100 Python-dispatched ops and
// the sequence is CUDA-graph capturable.
#include "common.cuh"

bool llc_lora_fused_eligible(const void* X, int ld_x, int T, int C, int R, const void* w, int ld_w,
                             const void* F, int ld_f, const void* U, int ld_u);
int llc_lora_fused_tc(const void* X, int ld_x, int T, int C, int R, const void* w, int ld_w,
                      const void* F, int ld_f, void* U, int ld_u, float* partial, int* n_partials,
                      cudaStream_t st);
int llc_attn_cls_fwd(const void* qkv, int ld_qkv, void* o_cls, int ld_o, float* p_cls, int N, int L,
                     int H, int sn, int sl, cudaStream_t st);
int llc_attn_cls_bwd(const void* qkv, int ld_qkv, const float* p_cls, const void* d_o_cls, int ld_do,
                     void* dqkv, int ld_dqkv, int N, int L, int H, int sn, int sl, cudaStream_t st);
int llc_attn_bwd_ws(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o,
                    int ld_do, const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H,
                    int tok_stride_n, int tok_stride_l, int causal, float* delta_ws, int delta_ready,
                    void* stream);
bool llc_colsum_tc_eligible(const void* X, int ld_x, int T, int C, const void* w, int ld_w);
int llc_colsum_tc_delta(const void* X, int ld_x, int T, int C, int R, const void* w, int ld_w,
                        float* partial, int* n_partials, const void* d_o, int ld_do, float* delta,
                        int delta_ld, cudaStream_t st);
bool llc_attn_bwd_uses_delta(int L);
int llc_adapter_scatter(const void* a, void* dst, int ld_dst, int col0, int T, void* stream);

namespace {

}  // namespace
int llc_refresh_lora_all(const llc_vit_layer* layers, int n_layers, int D, int r, float sc,
                         cudaStream_t st);
namespace {
inline size_t align_up(size_t x, size_t a = 1024) { return (x + a - 1) / a * a; }

// stream-K workspace of the GEMMs launched by the tower calls: lives at the head of the arena
// (caller-zeroed once, see llc_vit_arena_bytes); block-level calls have none and keep the
// whole-tile schedule
thread_local void* t_gemm_ws = nullptr;
struct WsScope {
  explicit WsScope(void* ws) { t_gemm_ws = ws; }
  ~WsScope() { t_gemm_ws = nullptr; }
};
inline int GEMM(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                llc_gemm_epi* e, void* stream) {
  e->ws = t_gemm_ws;
  e->ws_bytes = t_gemm_ws ? llc_gemm_ws_bytes() : 0;
  return llc_gemm_bf16_tn(A, lda, B, ldb, M, N, K, e, stream);
}

struct Dims {
  int N, L, T, D, H, M, E, G, P, PK /* padded 3*P*P */, r, layers;
};

// context > 0: a text tower (sequence length = context, no patch front end)
Dims make_dims(const llc_vit_cfg* c, int N, int context = 0) {
  Dims d;
  d.N = N;
  d.G = context > 0 ? 0 : c->image_size / c->patch;
  d.L = context > 0 ? context : d.G * d.G + 1;
  d.T = N * d.L;
  d.D = c->width;
  d.H = c->heads;
  d.M = c->mlp_dim;
  d.E = c->embed_dim;
  d.P = c->patch;
  d.PK = (3 * c->patch * c->patch + 15) / 16 * 16;
  d.r = c->lora_r;
  d.layers = c->layers;
  return d;
}

// Arena layout. Training keeps one activation set per layer (saved for backward); inference
// reuses a single set.
struct Arena {
  size_t gemm_ws;   // first: its flag words must be zero before the first launch
  size_t patches, patch_out, x /*[layers+1]*/, x_stride;
  size_t h1, qkv, lse, o, x_mid, z, layer_stride;  // per-layer block (training) or shared
  size_t h2, g;
  size_t dxb, dz, dh, d_o, dqkv, partial, delta;  // backward scratch
  // compact buffers of the class-token-only last block (rows = samples), llc_vit_*_cls
  size_t c_o, c_p, c_xmid, c_h2, c_z, c_g, c_dx, c_dxb, c_dz, c_dh, c_do;
  size_t total;
};

Arena plan(const Dims& d, int training) {
  Arena a;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  const size_t T = d.T;
  a.gemm_ws = take(llc_gemm_ws_bytes());
  a.patches = take((size_t)d.N * d.G * d.G * d.PK * 2);
  a.patch_out = take((size_t)d.N * d.G * d.G * d.D * 4);
  a.x_stride = align_up(T * d.D * 4);
  a.x = take(a.x_stride * (training ? d.layers + 1 : 2));
  const size_t l0 = off;
  a.h1 = take(T * (d.D + LLC_LORA_LD) * 2);
  a.qkv = take(T * (3 * d.D + LLC_LORA_LD) * 2);
  a.lse = take((size_t)d.N * d.H * d.L * 4);
  a.o = take(T * (d.D + LLC_LORA_LD) * 2);
  a.x_mid = take(T * d.D * 4);
  a.z = take(T * d.M * 2);
  a.layer_stride = off - l0;
  if (training) off = l0 + a.layer_stride * d.layers;
  a.h2 = take(T * d.D * 2);
  a.g = take(T * d.M * 2);
  if (training) {
    a.dxb = take(T * (d.D + LLC_LORA_LD) * 2);
    a.dz = take(T * d.M * 2);
    a.dh = take(T * d.D * 2);
    a.d_o = take(T * d.D * 2);
    a.dqkv = take(T * (3 * d.D + LLC_LORA_LD) * 2);
    a.partial = take((size_t)llc_lora_side_max_partials() * 3 * d.D * 2 * 4 * 4);  // 4 regions
    a.delta = take((size_t)d.N * d.H * d.L * 4);   // rowsum(dO o O) for the attention backward
  } else {
    a.dxb = a.dz = a.dh = a.d_o = a.dqkv = a.partial = a.delta = 0;
  }
  a.c_o = take((size_t)d.N * (d.D + LLC_LORA_LD) * 2);
  a.c_p = take((size_t)d.N * d.H * d.L * 4);
  a.c_xmid = take((size_t)d.N * d.D * 4);
  a.c_h2 = take((size_t)d.N * d.D * 2);
  a.c_z = take((size_t)d.N * d.M * 2);
  a.c_g = take((size_t)d.N * d.M * 2);
  if (training) {
    a.c_dx = take((size_t)d.N * d.D * 4);
    a.c_dxb = take((size_t)d.N * (d.D + LLC_LORA_LD) * 2);
    a.c_dz = take((size_t)d.N * d.M * 2);
    a.c_dh = take((size_t)d.N * d.D * 2);
    a.c_do = take((size_t)d.N * d.D * 2);
  } else {
    a.c_dx = a.c_dxb = a.c_dz = a.c_dh = a.c_do = 0;
  }
  a.total = off;
  return a;
}

int check_cfg(const llc_vit_cfg* c, const char* who) {
  LLC_REQUIRE(c, "%s: null cfg", who);
  LLC_REQUIRE(c->width % 128 == 0 && c->heads * 64 == c->width,
              "%s: width %d / heads %d unsupported (head dim must be 64)", who, c->width, c->heads);
  LLC_REQUIRE(c->patch > 0 && c->patch % 2 == 0 && c->image_size % c->patch == 0,
              "%s: bad patch geometry", who);
  LLC_REQUIRE(c->mlp_dim % 8 == 0 && c->layers > 0 && c->embed_dim > 0, "%s: bad dims", who);
  LLC_REQUIRE(c->lora_r >= 1 && c->lora_r <= 8, "%s: LoRA rank %d unsupported (1..8)", who,
              c->lora_r);
  return 0;
}

void fill_bufs(const Dims& d, const Arena& a, uint8_t* base, int layer, int training,
               llc_block_bufs* b) {
  const size_t lo = training ? a.layer_stride * layer : 0;
  const int xi = training ? layer : (layer & 1);
  const int xo = training ? layer + 1 : ((layer + 1) & 1);
  b->x_in = reinterpret_cast<float*>(base + a.x + a.x_stride * xi);
  b->x_out = reinterpret_cast<float*>(base + a.x + a.x_stride * xo);
  b->h1 = base + a.h1 + lo;
  b->qkv = base + a.qkv + lo;
  b->lse = reinterpret_cast<float*>(base + a.lse + lo);
  b->o = base + a.o + lo;
  b->x_mid = reinterpret_cast<float*>(base + a.x_mid + lo);
  b->z = training ? base + a.z + lo : nullptr;
  b->h2 = base + a.h2;
  b->g = base + a.g;
  (void)d;
}

#define RUN(call)            \
  do {                       \
    int _rc = (call);        \
    if (_rc != 0) return _rc; \
  } while (0)

}  // namespace

__global__ void cast_rows_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                      int T, int D, int ld_dst) {
  const int per_row = D / 4;
  const size_t total = (size_t)T * per_row;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t row = i / per_row;
    const int c4 = (int)(i % per_row);
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    *reinterpret_cast<uint2*>(dst + row * ld_dst + c4 * 4) =
        make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}

extern "C" int llc_cast_bf16(const float* src, void* dst, int T, int D, int ld_dst, void* stream) {
  LLC_REQUIRE(src && dst && T > 0 && D % 4 == 0 && ld_dst % 4 == 0 && ld_dst >= D,
              "llc_cast_bf16: bad args");
  const size_t total = (size_t)T * D / 4;
  const int grid = (int)((total + 255) / 256 < (size_t)llc_num_sms() * 8
                             ? (total + 255) / 256
                             : (size_t)llc_num_sms() * 8);
  LLC_PROF_BEGIN(LLC_K_OTHER, T, D, 1, 0.0, 6.0 * T * D, (cudaStream_t)stream);
  cast_rows_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst), T, D, ld_dst);
  LLC_PROF_END((cudaStream_t)stream);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("cast_rows_bf16_kernel");
  return 0;
}

// attention half of a block, from the normalised (or raw, for llc_mha_forward) input in b->h1
// (bf16 | u = h A_in^T in the pad columns) to the out-projection, whose epilogue is given by `eo`
static int attn_half_forward(const llc_vit_cfg* cfg, const llc_vit_layer* w, const llc_block_bufs* b,
                             llc_gemm_epi eo, int N, int L, int sn, int sl, int causal,
                             void* stream) {
  const int D = cfg->width, H = cfg->heads;
  const int T = N * L, DA = D + LLC_LORA_LD, QA = 3 * D + LLC_LORA_LD;  // row pitches
  const int KA = D + LLC_LORA_PAD;                                       // K extent
  llc_gemm_epi e;
  // qkv = h1 W_in^T + b_in + s (h1 A^T) B^T   (one accumulator, K = D + 16)
  e = llc_gemm_epi{};
  e.bias = w->bqkv; e.out = b->qkv; e.ld_out = QA;
  RUN(GEMM(b->h1, DA, w->wqkv_aug, DA, T, 3 * D, KA, &e, stream));
  RUN(llc_attn_fwd(b->qkv, QA, b->o, DA, b->lse, N, L, H, sn, sl, causal, stream));
  // u_o = o A_o^T into o's pad columns
  e = llc_gemm_epi{};
  e.out = reinterpret_cast<__nv_bfloat16*>(b->o) + D; e.ld_out = DA;
  RUN(GEMM(b->o, DA, w->f_out_A, D, T, LLC_LORA_PAD, D, &e, stream));
  // out = [resid +] o W_o^T + b_o + s (o A_o^T) B_o^T
  eo.bias = w->bo;
  RUN(GEMM(b->o, DA, w->wo_aug, DA, T, D, KA, &eo, stream));
  return 0;
}

extern "C" int llc_block_forward(const llc_vit_cfg* cfg, const llc_vit_layer* w,
                                 const llc_block_bufs* b, int N, int L, int sn, int sl, int causal,
                                 void* stream) {
  RUN(check_cfg(cfg, "llc_block_forward"));
  LLC_REQUIRE(w && b && N > 0 && L > 0, "llc_block_forward: bad args");
  const int D = cfg->width, M = cfg->mlp_dim, r = cfg->lora_r;
  const int T = N * L, DA = D + LLC_LORA_LD;
  llc_gemm_epi e;
  // x -> ln_1 -> h1 | u = h1 A_in^T
  RUN(llc_ln_fwd(b->x_in, D, w->ln1_g, w->ln1_b, T, D, b->h1, DA, w->in_A, r, stream));
  // x_mid = x + attention(h1)
  e = llc_gemm_epi{};
  e.resid = b->x_in; e.ld_resid = D; e.out = b->x_mid; e.ld_out = D; e.out_fp32 = 1;
  RUN(attn_half_forward(cfg, w, b, e, N, L, sn, sl, causal, stream));
  // mlp
  RUN(llc_ln_fwd(b->x_mid, D, w->ln2_g, w->ln2_b, T, D, b->h2, D, nullptr, 0, stream));
  e = llc_gemm_epi{};
  e.bias = w->bfc; e.act = 1; e.out = b->z; e.ld_out = M; e.out2 = b->g; e.ld_out2 = M;
  RUN(GEMM(b->h2, D, w->wfc, D, T, M, D, &e, stream));
  e = llc_gemm_epi{};
  e.bias = w->bproj; e.resid = b->x_mid; e.ld_resid = D; e.out = b->x_out; e.ld_out = D;
  e.out_fp32 = 1;
  RUN(GEMM(b->g, M, w->wproj, M, T, D, M, &e, stream));
  return 0;
}

// lora.MultiheadAttention.forward for self-attention (reference models/clip/lora.py:454-702 ->
// multi_head_attention_forward :732-1082): x fp32 [T, D] -> out fp32 [T, D]. b->h1 / qkv / lse / o
// are written (saved for llc_mha_backward); b->x_in = x, b->x_out = out; the other members are
// unused.
extern "C" int llc_mha_forward(const llc_vit_cfg* cfg, const llc_vit_layer* w,
                               const llc_block_bufs* b, int N, int L, int sn, int sl, int causal,
                               void* stream) {
  RUN(check_cfg(cfg, "llc_mha_forward"));
  LLC_REQUIRE(w && b && b->x_in && b->x_out && b->h1 && b->qkv && b->o && b->lse && N > 0 && L > 0,
              "llc_mha_forward: bad args");
  const int D = cfg->width, r = cfg->lora_r, T = N * L, DA = D + LLC_LORA_LD;
  RUN(llc_cast_bf16(b->x_in, b->h1, T, D, DA, stream));
  // u = x A_in^T (fp32 factor [r, D]: element (j, c) at j*D + c) into the pad columns of h1
  RUN(llc_lora_side(b->h1, DA, T, D, r, w->in_A, 1, D, 1.0f, nullptr, 0, nullptr, nullptr, stream));
  llc_gemm_epi e{};
  e.out = b->x_out; e.ld_out = D; e.out_fp32 = 1;
  return attn_half_forward(cfg, w, b, e, N, L, sn, sl, causal, stream);
}

// attention half of the backward: from s->dxb = bf16 gradient of the out-projection's output
// (pad columns free) to the LoRA gradients and, if need_dh1, s->dh = gradient of the attention
// input h1 (bf16 [T, D])
// k_do: K extent of the d_o GEMM when the caller has put more than the LoRA row product into the
// pad columns of dxb / woT_aug (the adapter block: 64 columns, see llc_adapter_block_backward)
static int attn_half_backward(const llc_vit_cfg* cfg, const llc_vit_layer* w,
                              const llc_block_bufs* b, const llc_block_bwd_bufs* s, int N, int L,
                              int sn, int sl, int causal, int need_dh1, void* stream,
                              int k_do = 0) {
  const int D = cfg->width, H = cfg->heads, r = cfg->lora_r;
  const float sc = cfg->lora_scale;
  const int T = N * L, DA = D + LLC_LORA_LD, QA = 3 * D + LLC_LORA_LD;  // row pitches
  const int KA = D + LLC_LORA_PAD, KQ = 3 * D + LLC_LORA_PAD;              // K extents
  __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(s->dxb);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(b->o);
  __nv_bfloat16* h1 = reinterpret_cast<__nv_bfloat16*>(b->h1);
  __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(s->dqkv);
  llc_gemm_epi e;
  // LoRA weight gradients: four column sums, each into its own partial region, reduced by ONE
  // finish launch at the end of the layer.
  const size_t preg = (size_t)llc_lora_side_max_partials() * 3 * D * 2;   // floats per region
  float* pr[4] = {s->partial, s->partial + preg, s->partial + 2 * preg, s->partial + 3 * preg};
  int np4[4] = {0, 0, 0, 0};
  cudaStream_t cst = (cudaStream_t)stream;
  // frozen attention (no gradient slots: the vanilla block under the adapter-clip method, whose
  // LoRA factors are zero buffers): nothing to reduce. The pad columns of dxb / dqkv are then
  // never written; the caller provides them zeroed (they meet zero weight columns).
  const bool frozen = !w->g_in_A && !w->g_in_B && !w->g_out_A && !w->g_out_B;
  // out-proj: du_o = s dx_mid B_o (-> dxb pad cols) and dB_o = s dx_mid^T u_o read the same
  // dx_mid: one fused pass when the shape allows, else a skinny GEMM plus a column-sum launch
  if (frozen) {
  } else if (llc_lora_fused_eligible(s->dxb, DA, T, D, r, o + D, DA, w->f_out_B, D, dxb + D, DA)) {
    RUN(llc_lora_fused_tc(s->dxb, DA, T, D, r <= 4 ? 4 : 8, o + D, DA, w->f_out_B, D, dxb + D, DA, pr[0],
                          &np4[0], cst));
  } else {
    e = llc_gemm_epi{};
    e.out = dxb + D; e.ld_out = DA;
    RUN(GEMM(s->dxb, DA, w->f_out_B, D, T, LLC_LORA_PAD, D, &e, stream));
    RUN(llc_lora_side(s->dxb, DA, T, D, r, nullptr, 0, 0, 0.f, o + D, DA, pr[0], &np4[0], stream));
  }
  // d_o = dx_mid W_o + du_o A_o
  e = llc_gemm_epi{};
  e.out = s->d_o; e.ld_out = D;
  RUN(GEMM(s->dxb, DA, w->woT_aug, DA, T, D, k_do > 0 ? k_do : KA, &e, stream));
  // dA_o = du_o^T o. The same pass over O also forms delta = rowsum(dO o O) for the attention
  // backward (from the O tiles it streams anyway and the L2-hot dO), when the shapes allow
  int delta_ready = 0;
  if (frozen) {
  } else if (s->delta && D == H * 64 && llc_attn_bwd_uses_delta(L) &&
      llc_colsum_tc_eligible(b->o, DA, T, D, dxb + D, DA)) {
    RUN(llc_colsum_tc_delta(b->o, DA, T, D, r <= 4 ? 4 : 8, dxb + D, DA, pr[1], &np4[1], s->d_o, D, s->delta, H,
                            cst));
    delta_ready = 1;
  } else {
    RUN(llc_lora_side(b->o, DA, T, D, r, nullptr, 0, 0, 0.f, dxb + D, DA, pr[1], &np4[1], stream));
  }
  RUN(llc_attn_bwd_ws(b->qkv, QA, b->o, DA, s->d_o, D, b->lse, s->dqkv, QA, N, L, H, sn, sl, causal,
                      s->delta, delta_ready, stream));
  // in-proj: du = s dqkv B_in (-> dqkv pad cols), dB_in = s dqkv^T u, dA_in = du^T h1
  if (frozen) {
  } else if (llc_lora_fused_eligible(s->dqkv, QA, T, 3 * D, r, h1 + D, DA, w->f_in_B, 3 * D,
                                     dqkv + 3 * D, QA)) {
    RUN(llc_lora_fused_tc(s->dqkv, QA, T, 3 * D, r <= 4 ? 4 : 8, h1 + D, DA, w->f_in_B, 3 * D, dqkv + 3 * D, QA,
                          pr[2], &np4[2], cst));
  } else {
    e = llc_gemm_epi{};
    e.out = dqkv + 3 * D; e.ld_out = QA;
    RUN(GEMM(s->dqkv, QA, w->f_in_B, 3 * D, T, LLC_LORA_PAD, 3 * D, &e, stream));
    RUN(llc_lora_side(s->dqkv, QA, T, 3 * D, r, nullptr, 0, 0, 0.f, h1 + D, DA, pr[2], &np4[2],
                      stream));
  }
  if (!frozen)
    RUN(llc_lora_side(b->h1, DA, T, D, r, nullptr, 0, 0, 0.f, dqkv + 3 * D, QA, pr[3], &np4[3],
                      stream));
  if (!frozen) {
    llc_finish_job jobs[4] = {
        {pr[0], np4[0], D, sc, w->g_out_B, r, 1},
        {pr[1], np4[1], D, 1.0f, w->g_out_A, 1, D},
        {pr[2], np4[2], 3 * D, sc, w->g_in_B, r, 1},
        {pr[3], np4[3], D, 1.0f, w->g_in_A, 1, D},
    };
    RUN(llc_lora_colsum_finish_multi(jobs, 4, r, stream));
  }
  if (need_dh1) {
    // dh1 = dqkv W_in + du A_in
    e = llc_gemm_epi{};
    e.out = s->dh; e.ld_out = D;
    RUN(GEMM(s->dqkv, QA, w->wqkvT_aug, QA, T, D, KQ, &e, stream));
  }
  return 0;
}

extern "C" int llc_block_backward(const llc_vit_cfg* cfg, const llc_vit_layer* w,
                                  const llc_block_bufs* b, const llc_block_bwd_bufs* s, int N,
                                  int L, int sn, int sl, int causal, int need_dx_in,
                                  void* stream) {
  RUN(check_cfg(cfg, "llc_block_backward"));
  LLC_REQUIRE(w && b && s && N > 0 && L > 0, "llc_block_backward: bad args");
  LLC_REQUIRE(b->z, "llc_block_backward: forward was not run in training mode");
  const int D = cfg->width, M = cfg->mlp_dim;
  const int T = N * L, DA = D + LLC_LORA_LD;
  llc_gemm_epi e;
  // dz = (dx W_proj) o QuickGELU'(z)
  e = llc_gemm_epi{};
  e.act = 2; e.aux = b->z; e.ld_aux = M; e.out = s->dz; e.ld_out = M;
  RUN(GEMM(s->dxb, DA, w->wprojT, D, T, M, D, &e, stream));
  // dh2 = dz W_fc
  e = llc_gemm_epi{};
  e.out = s->dh; e.ld_out = D;
  RUN(GEMM(s->dz, M, w->wfcT, M, T, D, M, &e, stream));
  // dx_mid = dx + LN2'(dh2); bf16 copy (the row product du_o = s dx_mid B_o runs on the tensor
  // cores inside attn_half_backward: fused into the LayerNorm kernel it cost 52 us per launch in
  // L1 traffic for the factor, profiles/)
  RUN(llc_ln_bwd(b->x_mid, D, w->ln2_g, s->dh, D, s->dy ? s->dy : s->dx, s->dx, T, D, s->dxb, DA,
                 nullptr, 0, 0.f, stream));
  RUN(attn_half_backward(cfg, w, b, s, N, L, sn, sl, causal, need_dx_in, stream));
  if (need_dx_in)   // dx_in = dx_mid + LN1'(dh1)
    RUN(llc_ln_bwd(b->x_in, D, w->ln1_g, s->dh, D, s->dx, s->dx, T, D, s->dxb, DA, nullptr, 0, 0.f,
                   stream));
  return 0;
}

// backward of llc_mha_forward: s->dx = fp32 gradient of the output [T, D] (read only);
// LoRA gradients -> w->g_*; if need_dx_in, s->dh = bf16 gradient of the input x [T, D].
// s->dz is unused.
extern "C" int llc_mha_backward(const llc_vit_cfg* cfg, const llc_vit_layer* w,
                                const llc_block_bufs* b, const llc_block_bwd_bufs* s, int N, int L,
                                int sn, int sl, int causal, int need_dx_in, void* stream) {
  RUN(check_cfg(cfg, "llc_mha_backward"));
  LLC_REQUIRE(w && b && s && s->dx && s->dxb && s->dh && s->d_o && s->dqkv && s->partial &&
              s->delta && N > 0 && L > 0, "llc_mha_backward: bad args");
  const int D = cfg->width, T = N * L, DA = D + LLC_LORA_LD;
  RUN(llc_cast_bf16(s->dx, s->dxb, T, D, DA, stream));
  return attn_half_backward(cfg, w, b, s, N, L, sn, sl, causal, need_dx_in, stream);
}

// ResidualAttentionBlock_Adapter (reference models/clip/model.py:418-442): the frozen block with
// the bottleneck adapter (adapter.cu) applied to both branches. The branches leave their GEMMs in
// bf16 (they are the adapter's A operand), the adapter's up-projection carries the fp32 residual.
extern "C" int llc_adapter_block_forward(const llc_vit_cfg* cfg, const llc_vit_layer* w,
                                         const llc_adapter* ad, const llc_block_bufs* b,
                                         const llc_adapter_bufs* ab, int N, int L, int sn, int sl,
                                         int causal, int training, void* stream) {
  RUN(check_cfg(cfg, "llc_adapter_block_forward"));
  LLC_REQUIRE(w && ad && b && ab && ab->ya && ab->a1 && ab->m && ab->a2 && N > 0 && L > 0,
              "llc_adapter_block_forward: bad args");
  const int D = cfg->width, M = cfg->mlp_dim, r = cfg->lora_r;
  const int T = N * L, DA = D + LLC_LORA_LD;
  llc_gemm_epi e;
  RUN(llc_ln_fwd(b->x_in, D, w->ln1_g, w->ln1_b, T, D, b->h1, DA, w->in_A, r, stream));
  e = llc_gemm_epi{};
  e.out = ab->ya; e.ld_out = D;
  RUN(attn_half_forward(cfg, w, b, e, N, L, sn, sl, causal, stream));
  RUN(llc_adapter_forward(ad, ab->ya, D, b->x_in, 1, ab->a1, ab->mask1, 0, training, b->x_mid, T,
                          D, stream));
  RUN(llc_ln_fwd(b->x_mid, D, w->ln2_g, w->ln2_b, T, D, b->h2, D, nullptr, 0, stream));
  e = llc_gemm_epi{};
  e.bias = w->bfc; e.act = 1; e.out = b->z; e.ld_out = M; e.out2 = b->g; e.ld_out2 = M;
  RUN(GEMM(b->h2, D, w->wfc, D, T, M, D, &e, stream));
  e = llc_gemm_epi{};
  e.bias = w->bproj; e.out = ab->m; e.ld_out = D;
  RUN(GEMM(b->g, M, w->wproj, M, T, D, M, &e, stream));
  RUN(llc_adapter_forward(ad, ab->m, D, b->x_mid, 1, ab->a2, ab->mask2, 1, training, b->x_out, T,
                          D, stream));
  return 0;
}

extern "C" int llc_adapter_block_backward(const llc_vit_cfg* cfg, const llc_vit_layer* w,
                                          const llc_adapter* ad, const llc_block_bufs* b,
                                          const llc_adapter_bufs* ab, const llc_block_bwd_bufs* s,
                                          int N, int L, int sn, int sl, int causal, int need_dx_in,
                                          int training, void* stream) {
  RUN(check_cfg(cfg, "llc_adapter_block_backward"));
  LLC_REQUIRE(w && ad && b && ab && s && ab->da && ab->partial && N > 0 && L > 0,
              "llc_adapter_block_backward: bad args");
  LLC_REQUIRE(b->z, "llc_adapter_block_backward: forward was not run in training mode");
  LLC_REQUIRE(ad->wprojT_ad, "llc_adapter_block_backward: composed operands missing "
                             "(llc_adapter_refresh with the block's weights)");
  const int D = cfg->width, M = cfg->mlp_dim;
  const int T = N * L, DA = D + LLC_LORA_LD, KAD = D + LLC_ADAPTER_DIM;
  llc_gemm_epi e;
  // x_out = x_mid + m + s up(a2): adapter gradients from dx; the gradient of the branch,
  // d_m = dx + dz2 W_d, is never formed - it only feeds d_m W_proj = dx W_proj + dz2 (W_d W_proj):
  // dz2 rides in the 64 pad columns of dxb against the composed columns of wprojT_ad
  RUN(llc_adapter_backward(ad, ab->m, D, ab->a2, s->dx, s->dxb, DA, nullptr, 0, ab->da,
                           ab->partial, 0, training, T, D, stream));
  RUN(llc_adapter_scatter(ab->da, s->dxb, DA, D, T, stream));
  // dz = ([dx | dz2] [W_proj^T | (W_d W_proj)^T]^T) o QuickGELU'(z); dh2 = dz W_fc
  e = llc_gemm_epi{};
  e.act = 2; e.aux = b->z; e.ld_aux = M; e.out = s->dz; e.ld_out = M;
  RUN(GEMM(s->dxb, DA, ad->wprojT_ad, KAD, T, M, KAD, &e, stream));
  e = llc_gemm_epi{};
  e.out = s->dh; e.ld_out = D;
  RUN(GEMM(s->dz, M, w->wfcT, M, T, D, M, &e, stream));
  RUN(llc_ln_bwd(b->x_mid, D, w->ln2_g, s->dh, D, s->dy ? s->dy : s->dx, s->dx, T, D, s->dxb, DA,
                 nullptr, 0, 0.f, stream));
  // x_mid = x + ya + s up(a1): the same with d_o = dx W_o + dz1 (W_d W_o) (pad columns of woT_aug)
  RUN(llc_adapter_backward(ad, ab->ya, D, ab->a1, s->dx, s->dxb, DA, nullptr, 0, ab->da,
                           ab->partial, 1, training, T, D, stream));
  RUN(llc_adapter_scatter(ab->da, s->dxb, DA, D, T, stream));
  RUN(attn_half_backward(cfg, w, b, s, N, L, sn, sl, causal, need_dx_in, stream, KAD));
  if (need_dx_in)
    RUN(llc_ln_bwd(b->x_in, D, w->ln1_g, s->dh, D, s->dx, s->dx, T, D, s->dxb, DA, nullptr, 0, 0.f,
                   stream));
  return 0;
}

extern "C" size_t llc_vit_arena_bytes(const llc_vit_cfg* cfg, int N, int training) {
  if (check_cfg(cfg, "llc_vit_arena_bytes") != 0 || N <= 0) return 0;
  return plan(make_dims(cfg, N), training).total;
}

namespace {
// ------------------------------------------------------------------------------------------------
// Class-token-only last block. The tower returns ln_post(x[:, 0, :]) @ proj (reference
// models/clip/model.py:782-785): of the last block's output only the CLS rows are ever read, and
// on the way back only they carry a gradient. LN1, the qkv GEMM (K and V of every token) and the
// in-projection's backward stay full size; attention runs for one query per (sample, head), and
// out-proj, LN2 and the MLP shrink from N*L rows to N rows. Identical results on the CLS rows.
struct ClsBufs {
  __nv_bfloat16 *o, *h2, *z, *g, *dxb, *dz, *dh, *d_o;
  float *p, *x_mid, *dx;
};

ClsBufs fill_cls(const Arena& a, uint8_t* base) {
  ClsBufs c;
  c.o = reinterpret_cast<__nv_bfloat16*>(base + a.c_o);
  c.p = reinterpret_cast<float*>(base + a.c_p);
  c.x_mid = reinterpret_cast<float*>(base + a.c_xmid);
  c.h2 = reinterpret_cast<__nv_bfloat16*>(base + a.c_h2);
  c.z = reinterpret_cast<__nv_bfloat16*>(base + a.c_z);
  c.g = reinterpret_cast<__nv_bfloat16*>(base + a.c_g);
  c.dx = reinterpret_cast<float*>(base + a.c_dx);
  c.dxb = reinterpret_cast<__nv_bfloat16*>(base + a.c_dxb);
  c.dz = reinterpret_cast<__nv_bfloat16*>(base + a.c_dz);
  c.dh = reinterpret_cast<__nv_bfloat16*>(base + a.c_dh);
  c.d_o = reinterpret_cast<__nv_bfloat16*>(base + a.c_do);
  return c;
}

// dst[n, :] = src[n * row_stride, :] (fp32) and its bf16 copy (pitch ld_b)
__global__ void gather_cls_rows_kernel(const float* __restrict__ src, size_t row_stride, int N, int D,
                                       float* __restrict__ dst, __nv_bfloat16* __restrict__ dstb,
                                       int ld_b) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = src[(size_t)n * row_stride + c];
    dst[(size_t)n * D + c] = v;
    dstb[(size_t)n * ld_b + c] = __float2bfloat16_rn(v);
  }
}
// dx[n * row_stride, :] += add[n, :]; bf16 copy refreshed
__global__ void add_cls_rows_kernel(float* __restrict__ dx, size_t row_stride, int N, int D,
                                    const float* __restrict__ add, __nv_bfloat16* __restrict__ dxb,
                                    size_t row_stride_b) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const float v = dx[(size_t)n * row_stride + c] + add[(size_t)n * D + c];
    dx[(size_t)n * row_stride + c] = v;
    dxb[(size_t)n * row_stride_b + c] = __float2bfloat16_rn(v);
  }
}

int block_forward_cls(const llc_vit_cfg* cfg, const llc_vit_layer* w, const llc_block_bufs* b,
                      const ClsBufs& c, int N, int L, void* stream) {
  const int D = cfg->width, M = cfg->mlp_dim, H = cfg->heads, r = cfg->lora_r;
  const int T = N * L, DA = D + LLC_LORA_LD, QA = 3 * D + LLC_LORA_LD;
  const int KA = D + LLC_LORA_PAD;
  cudaStream_t st = (cudaStream_t)stream;
  llc_gemm_epi e;
  // full size: x -> ln_1 -> h1 | u ; K and V of every token (output features D .. 3D), Q of the
  // class tokens only (rows L apart)
  RUN(llc_ln_fwd(b->x_in, D, w->ln1_g, w->ln1_b, T, D, b->h1, DA, w->in_A, r, stream));
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(b->qkv);
  const __nv_bfloat16* wqkv = reinterpret_cast<const __nv_bfloat16*>(w->wqkv_aug);
  e = llc_gemm_epi{};
  e.bias = w->bqkv + D; e.out = qkv + D; e.ld_out = QA;
  RUN(GEMM(b->h1, DA, wqkv + (size_t)D * DA, DA, T, 2 * D, KA, &e, stream));
  e = llc_gemm_epi{};
  e.bias = w->bqkv; e.out = qkv; e.ld_out = L * QA;
  RUN(GEMM(b->h1, L * DA, wqkv, DA, N, D, KA, &e, stream));
  // one query per (sample, head)
  RUN(llc_attn_cls_fwd(b->qkv, QA, c.o, DA, c.p, N, L, H, L, 1, st));
  e = llc_gemm_epi{};
  e.out = c.o + D; e.ld_out = DA;
  RUN(GEMM(c.o, DA, w->f_out_A, D, N, LLC_LORA_PAD, D, &e, stream));
  // x_mid[cls] = x[cls] + o W_o^T + b_o + s (o A_o^T) B_o^T : N rows, residual rows L*D apart
  e = llc_gemm_epi{};
  e.bias = w->bo; e.resid = b->x_in; e.ld_resid = L * D; e.out = c.x_mid; e.ld_out = D;
  e.out_fp32 = 1;
  RUN(GEMM(c.o, DA, w->wo_aug, DA, N, D, KA, &e, stream));
  RUN(llc_ln_fwd(c.x_mid, D, w->ln2_g, w->ln2_b, N, D, c.h2, D, nullptr, 0, stream));
  e = llc_gemm_epi{};
  e.bias = w->bfc; e.act = 1; e.out = c.z; e.ld_out = M; e.out2 = c.g; e.ld_out2 = M;
  RUN(GEMM(c.h2, D, w->wfc, D, N, M, D, &e, stream));
  // x_out[cls rows of the full buffer] = x_mid + g W_proj^T + b
  e = llc_gemm_epi{};
  e.bias = w->bproj; e.resid = c.x_mid; e.ld_resid = D; e.out = b->x_out; e.ld_out = L * D;
  e.out_fp32 = 1;
  RUN(GEMM(c.g, M, w->wproj, M, N, D, M, &e, stream));
  return 0;
}

int block_backward_cls(const llc_vit_cfg* cfg, const llc_vit_layer* w, const llc_block_bufs* b,
                       const ClsBufs& c, const llc_block_bwd_bufs* s, int N, int L, int need_dx_in,
                       void* stream) {
  const int D = cfg->width, M = cfg->mlp_dim, H = cfg->heads, r = cfg->lora_r;
  const float sc = cfg->lora_scale;
  const int T = N * L, DA = D + LLC_LORA_LD, QA = 3 * D + LLC_LORA_LD;
  const int KA = D + LLC_LORA_PAD, KQ = 3 * D + LLC_LORA_PAD;
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* h1 = reinterpret_cast<__nv_bfloat16*>(b->h1);
  __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(s->dqkv);
  __nv_bfloat16* dxb_full = reinterpret_cast<__nv_bfloat16*>(s->dxb);
  llc_gemm_epi e;
  // the head's gradient sits in the CLS rows of the full buffer
  gather_cls_rows_kernel<<<N, 256, 0, st>>>(s->dx, (size_t)L * D, N, D, c.dx, c.dxb, DA);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("gather_cls_rows_kernel");
  // MLP backward on N rows
  e = llc_gemm_epi{};
  e.act = 2; e.aux = c.z; e.ld_aux = M; e.out = c.dz; e.ld_out = M;
  RUN(GEMM(c.dxb, DA, w->wprojT, D, N, M, D, &e, stream));
  e = llc_gemm_epi{};
  e.out = c.dh; e.ld_out = D;
  RUN(GEMM(c.dz, M, w->wfcT, M, N, D, M, &e, stream));
  RUN(llc_ln_bwd(c.x_mid, D, w->ln2_g, c.dh, D, c.dx, c.dx, N, D, c.dxb, DA, nullptr, 0, 0.f,
                 stream));
  // out-proj LoRA: du_o = s dx_mid B_o, dB_o = s dx_mid^T u_o, dA_o = du_o^T o   (N rows)
  const size_t preg = (size_t)llc_lora_side_max_partials() * 3 * D * 2;   // floats per region
  float* pr[4] = {s->partial, s->partial + preg, s->partial + 2 * preg, s->partial + 3 * preg};
  int np4[4] = {0, 0, 0, 0};
  e = llc_gemm_epi{};
  e.out = c.dxb + D; e.ld_out = DA;
  RUN(GEMM(c.dxb, DA, w->f_out_B, D, N, LLC_LORA_PAD, D, &e, stream));
  RUN(llc_lora_side(c.dxb, DA, N, D, r, nullptr, 0, 0, 0.f, c.o + D, DA, pr[0], &np4[0], stream));
  RUN(llc_lora_side(c.o, DA, N, D, r, nullptr, 0, 0, 0.f, c.dxb + D, DA, pr[1], &np4[1], stream));
  // d_o = dx_mid W_o + du_o A_o ; attention backward of the CLS query -> dqkv of every token
  e = llc_gemm_epi{};
  e.out = c.d_o; e.ld_out = D;
  RUN(GEMM(c.dxb, DA, w->woT_aug, DA, N, D, KA, &e, stream));
  RUN(llc_attn_cls_bwd(b->qkv, QA, c.p, c.d_o, D, s->dqkv, QA, N, L, H, L, 1, st));
  // in-projection: full size again
  if (llc_lora_fused_eligible(s->dqkv, QA, T, 3 * D, r, h1 + D, DA, w->f_in_B, 3 * D, dqkv + 3 * D,
                              QA)) {
    RUN(llc_lora_fused_tc(s->dqkv, QA, T, 3 * D, r <= 4 ? 4 : 8, h1 + D, DA, w->f_in_B, 3 * D, dqkv + 3 * D, QA,
                          pr[2], &np4[2], st));
  } else {
    e = llc_gemm_epi{};
    e.out = dqkv + 3 * D; e.ld_out = QA;
    RUN(GEMM(s->dqkv, QA, w->f_in_B, 3 * D, T, LLC_LORA_PAD, 3 * D, &e, stream));
    RUN(llc_lora_side(s->dqkv, QA, T, 3 * D, r, nullptr, 0, 0, 0.f, h1 + D, DA, pr[2], &np4[2],
                      stream));
  }
  RUN(llc_lora_side(b->h1, DA, T, D, r, nullptr, 0, 0, 0.f, dqkv + 3 * D, QA, pr[3], &np4[3],
                    stream));
  {
    llc_finish_job jobs[4] = {
        {pr[0], np4[0], D, sc, w->g_out_B, r, 1},
        {pr[1], np4[1], D, 1.0f, w->g_out_A, 1, D},
        {pr[2], np4[2], 3 * D, sc, w->g_in_B, r, 1},
        {pr[3], np4[3], D, 1.0f, w->g_in_A, 1, D},
    };
    RUN(llc_lora_colsum_finish_multi(jobs, 4, r, stream));
  }
  if (need_dx_in) {
    // dx_in = LN1'(dh1) for every token; the residual path adds dx_mid on the CLS rows only.
    // dQ is zero outside the class-token rows: the full-size GEMM contracts over K, V and the
    // LoRA columns only (K range D .. 3D+16), the class-token rows are redone over all of K
    const __nv_bfloat16* wT = reinterpret_cast<const __nv_bfloat16*>(w->wqkvT_aug);
    e = llc_gemm_epi{};
    e.out = s->dh; e.ld_out = D;
    RUN(GEMM(dqkv + D, QA, wT + D, QA, T, D, KQ - D, &e, stream));
    e = llc_gemm_epi{};
    e.out = s->dh; e.ld_out = L * D;
    RUN(GEMM(s->dqkv, L * QA, w->wqkvT_aug, QA, N, D, KQ, &e, stream));
    RUN(llc_ln_bwd(b->x_in, D, w->ln1_g, s->dh, D, nullptr, s->dx, T, D, s->dxb, DA, nullptr, 0, 0.f,
                   stream));
    add_cls_rows_kernel<<<N, 256, 0, st>>>(s->dx, (size_t)L * D, N, D, c.dx, dxb_full,
                                           (size_t)L * DA);
    LLC_COUNT_LAUNCH();
    LLC_LAUNCH_CHECK("add_cls_rows_kernel");
  }
  return 0;
}

}  // namespace

extern "C" int llc_vit_refresh_lora(const llc_vit_cfg* cfg, const llc_vit_weights* w,
                                    void* stream) {
  RUN(check_cfg(cfg, "llc_vit_refresh_lora"));
  LLC_REQUIRE(w && w->layers, "llc_vit_refresh_lora: null weights");
  return llc_refresh_lora_all(w->layers, cfg->layers, cfg->width, cfg->lora_r, cfg->lora_scale,
                              (cudaStream_t)stream);
}

static int vit_forward_impl(const llc_vit_cfg* cfg, const llc_vit_weights* w, const float* images,
                            const llc_img_transform* tx, int N, void* arena, int training,
                            float** x_final, void* stream, bool cls_only) {
  RUN(check_cfg(cfg, "llc_vit_forward"));
  LLC_REQUIRE(w && w->layers && (images || tx) && arena && N > 0, "llc_vit_forward: bad args");
  const Dims d = make_dims(cfg, N);
  const Arena a = plan(d, training);
  uint8_t* base = reinterpret_cast<uint8_t*>(arena);
  WsScope ws_scope(base + a.gemm_ws);
  // patch embedding: (input transform +) im2col -> GEMM -> class token, positional embedding,
  // ln_pre
  if (tx) {
    LLC_REQUIRE(tx->out_size == cfg->image_size, "llc_vit_forward_tx: transform output %d != %d",
                tx->out_size, cfg->image_size);
    RUN(llc_transform_patchify(tx, N, cfg->patch, base + a.patches, d.PK, stream));
  } else {
    RUN(llc_patchify(images, N, 3, cfg->image_size, cfg->patch, base + a.patches, d.PK, stream));
  }
  llc_gemm_epi e{};
  e.out = base + a.patch_out; e.ld_out = d.D; e.out_fp32 = 1;
  RUN(GEMM(base + a.patches, d.PK, w->wpatch, d.PK, N * d.G * d.G, d.D, d.PK, &e,
                       stream));
  float* x0 = reinterpret_cast<float*>(base + a.x);
  RUN(llc_embed_ln_pre(reinterpret_cast<float*>(base + a.patch_out), d.D, w->class_emb, w->pos_emb,
                       w->ln_pre_g, w->ln_pre_b, N, d.L, d.D, x0, stream));
  llc_block_bufs b;
  for (int l = 0; l < d.layers; ++l) {
    fill_bufs(d, a, base, l, training, &b);
    if (cls_only && l == d.layers - 1)
      RUN(block_forward_cls(cfg, &w->layers[l], &b, fill_cls(a, base), N, d.L, stream));
    else
      RUN(llc_block_forward(cfg, &w->layers[l], &b, N, d.L, d.L, 1, 0, stream));
  }
  if (x_final) *x_final = b.x_out;
  return 0;
}

extern "C" int llc_vit_forward(const llc_vit_cfg* cfg, const llc_vit_weights* w,
                               const float* images, int N, void* arena, int training,
                               float** x_final, void* stream) {
  return vit_forward_impl(cfg, w, images, nullptr, N, arena, training, x_final, stream, false);
}

extern "C" int llc_vit_forward_cls(const llc_vit_cfg* cfg, const llc_vit_weights* w,
                                   const float* images, int N, void* arena, int training,
                                   float** x_final, void* stream) {
  return vit_forward_impl(cfg, w, images, nullptr, N, arena, training, x_final, stream, true);
}

extern "C" int llc_vit_forward_tx(const llc_vit_cfg* cfg, const llc_vit_weights* w,
                                  const llc_img_transform* tx, int N, void* arena, int training,
                                  int cls_only, float** x_final, void* stream) {
  LLC_REQUIRE(tx, "llc_vit_forward_tx: null transform");
  return vit_forward_impl(cfg, w, nullptr, tx, N, arena, training, x_final, stream, cls_only != 0);
}

static int vit_backward_impl(const llc_vit_cfg* cfg, const llc_vit_weights* w, int N, void* arena,
                             float* dx_final, void* stream, bool cls_only) {
  RUN(check_cfg(cfg, "llc_vit_backward"));
  LLC_REQUIRE(w && w->layers && arena && dx_final && N > 0, "llc_vit_backward: bad args");
  const Dims d = make_dims(cfg, N);
  const Arena a = plan(d, 1);
  uint8_t* base = reinterpret_cast<uint8_t*>(arena);
  WsScope ws_scope(base + a.gemm_ws);
  llc_block_bwd_bufs s{};
  s.dx = dx_final;
  s.dxb = base + a.dxb;
  s.dz = base + a.dz;
  s.dh = base + a.dh;
  s.d_o = base + a.d_o;
  s.dqkv = base + a.dqkv;
  s.partial = reinterpret_cast<float*>(base + a.partial);
  s.delta = reinterpret_cast<float*>(base + a.delta);
  if (!cls_only) RUN(llc_cast_bf16(dx_final, s.dxb, d.T, d.D, d.D + LLC_LORA_LD, stream));
  llc_block_bufs b;
  for (int l = d.layers - 1; l >= 0; --l) {
    fill_bufs(d, a, base, l, 1, &b);
    if (cls_only && l == d.layers - 1)
      RUN(block_backward_cls(cfg, &w->layers[l], &b, fill_cls(a, base), &s, N, d.L, l > 0, stream));
    else
      RUN(llc_block_backward(cfg, &w->layers[l], &b, &s, N, d.L, d.L, 1, 0, l > 0, stream));
  }
  return 0;
}

extern "C" int llc_vit_backward(const llc_vit_cfg* cfg, const llc_vit_weights* w, int N,
                                void* arena, float* dx_final, void* stream) {
  return vit_backward_impl(cfg, w, N, arena, dx_final, stream, false);
}

extern "C" int llc_vit_backward_cls(const llc_vit_cfg* cfg, const llc_vit_weights* w, int N,
                                    void* arena, float* dx_final, void* stream) {
  return vit_backward_impl(cfg, w, N, arena, dx_final, stream, true);
}


// ------------------------------------------------------------------------------------------------
// Text tower: CLIP.encode_text, reference models/clip/model.py:941-956, for
// peft_encoder='both' (what scripts/lora_clip.sh sets): token embedding + positional embedding ->
// LoRA blocks under the causal mask (:926-932) -> [C*ctx, D] fp32. ln_final, the EOT gather
// (:953-954) and text_projection run in the head kernels (llc_head_fwd with row_idx).
namespace {
template <int NV>
__global__ void __launch_bounds__(256)
text_embed_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ emb, int vocab,
                  const float* __restrict__ pos, int T, int ctx, float* __restrict__ x0) {
  constexpr int D = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= T) return;
  int64_t tok = tokens[row];
  if (tok < 0 || tok >= vocab) tok = 0;   // checked on the host side of the binding
  const float4* e = reinterpret_cast<const float4*>(emb + (size_t)tok * D);
  const float4* p = reinterpret_cast<const float4*>(pos + (size_t)(row % ctx) * D);
  float4* o = reinterpret_cast<float4*>(x0 + (size_t)row * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a = __ldg(e + lane + 32 * i), b = __ldg(p + lane + 32 * i);
    o[lane + 32 * i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}
}  // namespace

extern "C" size_t llc_text_arena_bytes(const llc_vit_cfg* cfg, int context, int C, int training) {
  if (check_cfg(cfg, "llc_text_arena_bytes") != 0 || C <= 0 || context <= 0) return 0;
  return plan(make_dims(cfg, C, context), training).total;
}

extern "C" int llc_text_forward(const llc_vit_cfg* cfg, const llc_text_weights* w,
                                const int64_t* tokens, int C, void* arena, int training,
                                float** x_final, void* stream) {
  RUN(check_cfg(cfg, "llc_text_forward"));
  LLC_REQUIRE(w && w->layers && w->tok_emb && w->pos_emb && tokens && arena && C > 0 &&
              w->context > 0 && w->vocab > 0, "llc_text_forward: bad args");
  const Dims d = make_dims(cfg, C, w->context);
  const Arena a = plan(d, training);
  uint8_t* base = reinterpret_cast<uint8_t*>(arena);
  WsScope ws_scope(base + a.gemm_ws);
  float* x0 = reinterpret_cast<float*>(base + a.x);
  cudaStream_t st = (cudaStream_t)stream;
  LLC_PROF_BEGIN(LLC_K_EMBED, d.T, d.D, 4, 0.0, 8.0 * d.T * d.D, st);
  switch (d.D / 128) {
    case 4: text_embed_kernel<4><<<(d.T + 7) / 8, 256, 0, st>>>(tokens, w->tok_emb, w->vocab, w->pos_emb, d.T, d.L, x0); break;
    case 5: text_embed_kernel<5><<<(d.T + 7) / 8, 256, 0, st>>>(tokens, w->tok_emb, w->vocab, w->pos_emb, d.T, d.L, x0); break;
    case 6: text_embed_kernel<6><<<(d.T + 7) / 8, 256, 0, st>>>(tokens, w->tok_emb, w->vocab, w->pos_emb, d.T, d.L, x0); break;
    case 8: text_embed_kernel<8><<<(d.T + 7) / 8, 256, 0, st>>>(tokens, w->tok_emb, w->vocab, w->pos_emb, d.T, d.L, x0); break;
    case 3: text_embed_kernel<3><<<(d.T + 7) / 8, 256, 0, st>>>(tokens, w->tok_emb, w->vocab, w->pos_emb, d.T, d.L, x0); break;
    case 1: text_embed_kernel<1><<<(d.T + 7) / 8, 256, 0, st>>>(tokens, w->tok_emb, w->vocab, w->pos_emb, d.T, d.L, x0); break;
    case 2: text_embed_kernel<2><<<(d.T + 7) / 8, 256, 0, st>>>(tokens, w->tok_emb, w->vocab, w->pos_emb, d.T, d.L, x0); break;
    default:
      llc_set_error("llc_text_forward: width %d unsupported", d.D);
      return LLC_ERR_ARG;
  }
  LLC_PROF_END(st);
  LLC_COUNT_LAUNCH();
  LLC_LAUNCH_CHECK("text_embed_kernel");
  llc_block_bufs b;
  for (int l = 0; l < d.layers; ++l) {
    fill_bufs(d, a, base, l, training, &b);
    RUN(llc_block_forward(cfg, &w->layers[l], &b, C, d.L, d.L, 1, 1, stream));
  }
  if (x_final) *x_final = b.x_out;
  return 0;
}

extern "C" int llc_text_backward(const llc_vit_cfg* cfg, const llc_text_weights* w, int C,
                                 void* arena, float* dx_final, void* stream) {
  RUN(check_cfg(cfg, "llc_text_backward"));
  LLC_REQUIRE(w && w->layers && arena && dx_final && C > 0 && w->context > 0,
              "llc_text_backward: bad args");
  const Dims d = make_dims(cfg, C, w->context);
  const Arena a = plan(d, 1);
  uint8_t* base = reinterpret_cast<uint8_t*>(arena);
  WsScope ws_scope(base + a.gemm_ws);
  llc_block_bwd_bufs s{};
  s.dx = dx_final;
  s.dxb = base + a.dxb;
  s.dz = base + a.dz;
  s.dh = base + a.dh;
  s.d_o = base + a.d_o;
  s.dqkv = base + a.dqkv;
  s.partial = reinterpret_cast<float*>(base + a.partial);
  s.delta = reinterpret_cast<float*>(base + a.delta);
  RUN(llc_cast_bf16(dx_final, s.dxb, d.T, d.D, d.D + LLC_LORA_LD, stream));
  llc_block_bufs b;
  for (int l = d.layers - 1; l >= 0; --l) {
    fill_bufs(d, a, base, l, 1, &b);
    // token / positional embeddings are frozen: the first block needs no input gradient
    RUN(llc_block_backward(cfg, &w->layers[l], &b, &s, C, d.L, d.L, 1, 1, l > 0, stream));
  }
  return 0;
}
