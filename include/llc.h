/*
 * llc.h — C ABI of libllc.so: the B200 (sm_100a) implementation of LifeLong-CLIP's online-step hot
 * path (CLIP ViT image tower forward/backward with LoRA on the attention projections + the cosine
 * logit / softmax / cross-entropy head).
 *
 * The reference (qcNPU/LifeLong-CLIP) is pure Python/PyTorch and has NO FFI layer; its operator
 * API is a set of nn.Module classes. Each entry point below names the reference lines it replaces
 * (paths relative to the reference root); the ctypes binding a maintainer would add is shown in
 * INTEGRATION.md and implemented in lifelong-clip_b200/_capi.py.
 *
 * Conventions
 *  - extern "C", plain pointers + sizes; every device pointer is owned by the caller (PyTorch's
 *    caching allocator); the library retains nothing past the call and allocates no device
 *    memory of its own (scratch is passed in).
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it and never
 *    synchronise the device; all are CUDA-graph capturable.
 *  - return 0 on success; <0 = argument/shape error found before launch; >0 = cudaError_t.
 *    Text via llc_last_error() (thread-local). No CPU fallback exists: a missing GPU or a
 *    non-sm_100 device is an error.
 *  - token layout: activations are row-major [T, ld] with T = N*L tokens; token (n, l) is row
 *    n*tok_stride_n + l*tok_stride_l (sample-major [N,L,D]: (L,1); the reference's
 *    sequence-first [L,N,D], models/clip/model.py:767: (1,N)).
 */
#ifndef LLC_H_
#define LLC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLC_VERSION 100
#define LLC_ERR_ARG (-1)
#define LLC_ERR_ARCH (-2)
#define LLC_LORA_PAD 16 /* rank-r LoRA columns are padded to one UMMA K step (16 bf16) */
#define LLC_LORA_LD 64  /* ...and the ROW PITCH of an augmented buffer grows by 64 elements (128 B),
                           so every row still starts on a cache-line boundary for TMA */

int llc_version(void);
const char* llc_last_error(void);
/* 0 when device `dev` is sm_100 (B200); LLC_ERR_ARCH otherwise. */
int llc_check_device(int dev);
/* process-wide switch (default off; returns the previous setting): the persistent kernels
 * (GEMM, attention) release their programmatic dependents when every CTA has started its last
 * work item, so the next kernel's prologue overlaps this kernel's tail */
int llc_set_pdl_trigger(int on);
/* process-wide traversal order of the row kernels (returns the previous mask): bit 0 - llc_ln_fwd,
 * bit 1 - llc_ln_bwd walk their rows from the END, so that they start on the rows their producer
 * (a GEMM walking m-tiles upwards) wrote last and that are still in L2 */
int llc_set_traversal(int mask);
/* number of kernels this process has launched through the library (bench.py: gpu_launches) */
unsigned long long llc_launch_count(void);

/* ---- per-launch device timing (bench.py roofline: CUDA events on the launching stream) --------
 * When enabled every llc_* launch is bracketed by two events; llc_prof_read synchronises on them
 * and returns one record per launch. Off by default (events cost ~1 us each). Not graph-capturable
 * while enabled. */
enum {
  LLC_K_GEMM = 0, LLC_K_ATTN_FWD = 1, LLC_K_ATTN_BWD = 2, LLC_K_LN_FWD = 3, LLC_K_LN_BWD = 4,
  LLC_K_LORA_SIDE = 5, LLC_K_HEAD = 6, LLC_K_EMBED = 7, LLC_K_OTHER = 8, LLC_K_COUNT = 9
};
typedef struct llc_prof_rec {
  int kind, m, n, k;   /* m,n,k: GEMM shape, or rows/cols of the pass (k = 0) */
  float ms;            /* device time of this launch */
  double flops, bytes; /* algorithmic work the caller attributes to it */
} llc_prof_rec;
int llc_prof_enable(int on); /* 1: start recording (clears old records); 0: stop */
/* copies up to max records into out (may be NULL to just count); returns the record count, <0 on
 * error */
int llc_prof_read(llc_prof_rec* out, int max);

/* ---- dense contraction: out[M,N] = epi(A[M,K] . B[N,K]^T), bf16 operands, fp32 accumulate ----
 * TMA-fed tcgen05/TMEM GEMM. Replaces every F.linear on the path: in-proj lora.py:837-839,
 * out-proj lora.py:1072-1074, mlp model.py:219-222, and their activation-gradient transposes.
 * A and B are K-major (row-major [rows, ld]); K % 16 == 0, ld % 8 == 0, N % 8 == 0. */
typedef struct llc_gemm_epi {
  const float* bias;  /* [N] fp32 or NULL */
  const float* resid; /* fp32 [M, ld_resid] or NULL: added to the result (residual stream) */
  int ld_resid;
  int act;            /* 0 none | 1 QuickGELU: out=z (may be NULL), out2=z*sigmoid(1.702z)
                         | 2 multiply by QuickGELU'(aux) (backward of model.py:203-206) */
  const void* aux;    /* bf16 [M, ld_aux], pre-activation z for act==2 */
  int ld_aux;
  void* out;          /* [M, ld_out] */
  int ld_out;
  int out_fp32;       /* 0: bf16, 1: fp32 */
  void* out2;         /* bf16 [M, ld_out2], only act==1 */
  int ld_out2;
  void* ws;           /* optional stream-K workspace: llc_gemm_ws_bytes() bytes of device memory,
                         16 B aligned, whose first 8 KB (flags) were zero before the first use;
                         the kernel leaves them zero. NULL: whole-tile schedule only */
  size_t ws_bytes;
} llc_gemm_epi;
/* Workspace size of the stream-K schedule: when the 256 x 256 tile count leaves a poorly filled
 * last wave on the 74 CTA pairs (T = 197 x images rarely divides), the (tile, k-block) units are
 * cut into equal contiguous ranges and split tiles are summed through this workspace. */
size_t llc_gemm_ws_bytes(void);
/* process-wide switch of the stream-K schedule (default on; A/B measurements); returns the
 * previous setting */
int llc_gemm_set_stream_k(int on);
int llc_gemm_bf16_tn(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                     const llc_gemm_epi* epi, void* stream);

/* ---- LayerNorm (model.py:194-200: fp32, eps 1e-5, affine) --------------------------------- */
/* y = LN(x)*g+b -> bf16 [T, ld_y]; if lora_A != NULL also writes u = LN(x) . lora_A^T (r cols,
 * bf16, zero padded to LLC_LORA_PAD) at column D of y: the K-augmentation that folds the rank-r
 * update of lora.py:838-839 into the following GEMM. */
int llc_ln_fwd(const float* x, int ld_x, const float* gamma, const float* beta, int T, int D,
               void* y, int ld_y, const float* lora_A, int r, void* stream);
/* dx_out = dx_in + LNbackward(dy) (gamma/beta frozen: no weight grads, methods/adapter_clip.py:
 * 115-119). Writes fp32 dx_out [T, D] (may alias dx_in), and if dxb != NULL a bf16 copy
 * [T, ld_dxb]; if lora_B != NULL appends du = scale * dx_out . lora_B (r cols) at column D of dxb. */
int llc_ln_bwd(const float* x, int ld_x, const float* gamma, const void* dy, int ld_dy,
               const float* dx_in, float* dx_out, int T, int D, void* dxb, int ld_dxb,
               const float* lora_B, int r, float scale, void* stream);

/* ---- attention core (lora.py:950,1002-1006,1043,1063-1071): softmax(q k^T / sqrt(hd)) v ------
 * qkv bf16 [T, ld_qkv] holds q | k | v at columns [0,D) [D,2D) [2D,3D); head h of sample n uses
 * columns h*hd..; no mask (vision tower) or causal (text tower). o bf16 [T, ld_o]; lse fp32
 * [N*H*L] saved for backward. hd must be 64. */
int llc_attn_fwd(const void* qkv, int ld_qkv, void* o, int ld_o, float* lse, int N, int L, int H,
                 int tok_stride_n, int tok_stride_l, int causal, void* stream);
/* delta: caller-owned scratch of N*H*L floats (rowsum(dO o O), formed by a pre-pass); the library
 * keeps no buffer of its own */
int llc_attn_bwd(const void* qkv, int ld_qkv, const void* o, int ld_o, const void* d_o, int ld_do,
                 const float* lse, void* dqkv, int ld_dqkv, int N, int L, int H, int tok_stride_n,
                 int tok_stride_l, int causal, float* delta, void* stream);

/* ---- LoRA side reductions (lora.py:838-839,1072-1074 and their autograd) ---------------------
 * One pass over X bf16 [T, ld_x] (C columns):
 *   rowdot: if Mrd != NULL, w_out[t, j] = rd_scale * sum_c X[t,c] * Mrd[c*rd_sc + j*rd_sj], j<r,
 *           written bf16 (zero padded to LLC_LORA_PAD) at X[t, C..C+16)   (u = x A^T, du = s g B)
 *   colsum: if w != NULL (bf16 [T, ld_w], r cols), partial[c, j] += sum_t X[t,c] * w[t,j]
 * followed by llc_lora_colsum_finish which reduces the per-CTA partials deterministically into
 * out[c*o_sc + j*o_sj] = cs_scale * sum (dB = s g^T u, dA = du^T h). */
int llc_lora_side(void* X, int ld_x, int T, int C, int r, const float* Mrd, int rd_sc, int rd_sj,
                  float rd_scale, const void* w, int ld_w, float* partial, int* n_partials,
                  void* stream);
int llc_lora_colsum_finish(const float* partial, int n_partials, int C, int r, float cs_scale,
                           float* out, int o_sc, int o_sj, void* stream);
int llc_lora_side_max_partials(void);
/* Fused form for the two reductions of a LoRA projection's backward that read the same X
 * (bf16 [T, ld_x], C columns), ONE pass over X on the tensor cores:
 *   partial[p][c][j] = sum over CTA p's tokens of X[t,c] * w[t,j], j < r   (-> colsum_finish)
 *   U[t, 0..16)      = sum_c X[t,c] * F[j,c]   bf16, F = packed factors bf16 [16, ld_f]
 * (dB = s g^T u and du = s g B of lora.py:838-839,1072-1074 under autograd). Needs
 * C % 128 == 0, C <= 2304, T >= 1024, 16 B aligned pointers and pitches % 8 == 0; returns
 * LLC_ERR_ARG otherwise (the caller then uses llc_lora_side + a skinny GEMM). */
int llc_lora_side_fused(const void* X, int ld_x, int T, int C, int r, const void* w, int ld_w,
                        const void* F, int ld_f, void* U, int ld_u, float* partial,
                        int* n_partials, void* stream);
/* several finishes in one launch (one per LoRA tensor of a layer) */
typedef struct llc_finish_job {
  const float* partial;
  int n_partials, C;
  float scale;
  float* out;
  int o_sc, o_sj;
} llc_finish_job;
int llc_lora_colsum_finish_multi(const llc_finish_job* jobs, int n_jobs, int r, void* stream);

/* ---- weight preparation (frozen backbone: done once; LoRA columns refreshed every step) ------ */
/* dst bf16 [rows, ld_dst] <- src fp32 [rows, cols] (row-major) or its transpose */
int llc_pack_weight(const float* src, int rows, int cols, int transpose, void* dst, int ld_dst,
                    void* stream);
/* dst[i, col0 + j] = scale * src[i*s_i + j*s_j], j < r; columns r..LLC_LORA_PAD zeroed */
int llc_pack_lora_cols(const float* src, int rows, int r, int s_i, int s_j, float scale, void* dst,
                       int ld_dst, int col0, void* stream);

/* dst bf16 [16, ld_dst]: row j < r = scale * src[j*s_j + c*s_c], rows r..15 zero: the [16, K]
 * factor of a rank-r row product computed as a skinny GEMM (u = X . F^T) */
int llc_pack_factor_rows(const float* src, int r, int cols, int s_j, int s_c, float scale, void* dst,
                         int ld_dst, void* stream);

/* ---- patch embedding front end (model.py:756-767) -------------------------------------------- */
/* NCHW fp32 image -> bf16 patch rows [N*G*G, ld_out] (im2col of the stride-P conv, no bias);
 * columns >= 3*P*P are zero-filled (K padding to a multiple of 16) */
int llc_patchify(const float* img, int N, int C, int HW, int P, void* out, int ld_out,
                 void* stream);
/* x0[n,0,:] = LN_pre(class_emb + pos[0]); x0[n,1+g,:] = LN_pre(patch_out[n*G*G+g] + pos[1+g]) */
int llc_embed_ln_pre(const float* patch_out, int ld_p, const float* class_emb, const float* pos,
                     const float* gamma, const float* beta, int N, int L, int D, float* x0,
                     void* stream);

/* ---- head (model.py:782-785, 966-973; models/adapter_clip.py:99; methods/adapter_clip.py:89-90;
 *           mask variant methods/mvp_clip.py:113-118) ---------------------------------------- */
typedef struct llc_head_args {
  const float* x;        /* residual stream fp32; CLS token of sample n at row n*cls_stride */
  int cls_stride, ld_x;
  const float* ln_g;     /* ln_post */
  const float* ln_b;
  const float* proj;     /* [D, E] fp32 (visual.proj) */
  const float* text;     /* [C_all, E] fp32, L2-normalised cached class text features */
  const int64_t* cls_idx;/* [C] rows of `text` visible this step (gather-then-compute), or NULL */
  const float* add_mask; /* [C] additive logit mask (0 / -inf), or NULL */
  float logit_scale;     /* exp(logit_scale) */
  int N, D, E, C;
  const int64_t* labels; /* [N] local (remapped) labels, or NULL (inference) */
  const float* d_feat;   /* backward only: [N, E] extra gradient w.r.t. feat (pre-normalisation),
                            added to the one coming from the logits; NULL if none */
  int skip_logit_grad;   /* backward only: 1 = use d_feat alone (VisualTransformer.forward used
                            without the head) */
  int double_softmax;    /* 1: CE applied to the probabilities (the reference's loss) */
  float inv_batch;       /* 1 / global batch (mean reduction) */
  /* outputs */
  float* feat;           /* [N, E] image features before normalisation */
  float* fnorm;          /* [N, E] normalised */
  float* logits;         /* [N, C] */
  float* probs;          /* [N, C] */
  float* loss_rows;      /* [N] per-sample loss (already * inv_batch) */
  int64_t* pred;         /* [N] argmax (lowest index among ties) */
  /* text side (peft_encoder='both', model.py:941-956) */
  const int64_t* row_idx;/* [N] row of sample n in x / dx, or NULL (then n*cls_stride): the EOT
                            token of every prompt, text.argmax(-1) of model.py:953-954 */
  const float* d_fnorm;  /* backward only: [N, E] gradient w.r.t. the NORMALISED features (used
                            with skip_logit_grad; chained through f = z/|z| in the kernel) */
  float* dlogits;        /* backward only: [N, C] dL/dlogits written out, or NULL */
  int d_is_logits;       /* backward only: llc_head_bwd's d_probs is the gradient w.r.t. the LOGITS
                            (models/maple.py:250-253 returns logits; methods/maple.py:96) */
} llc_head_args;
int llc_head_fwd(const llc_head_args* a, void* stream);
/* d_probs (may be NULL: then the analytic gradient of the fused loss is used, scaled by
 * loss_scale) -> d_x rows of the CLS tokens: fp32 [T, ld_dx] (other rows untouched) */
int llc_head_bwd(const llc_head_args* a, const float* d_probs, float loss_scale, float* dx,
                 int ld_dx, void* stream);
/* d_text[c, e] = scale * sum_n dlogits[n, c] * fnorm[n, e] ([C, E] fp32): gradient of the logit
 * product (model.py:972) w.r.t. the normalised text features when the text tower is trainable */
int llc_head_dtext(const float* dlogits, const float* fnorm, int N, int C, int E, float scale,
                   float* d_text, void* stream);
/* y_local[i] = lut[y_global[i]] (methods/adapter_clip.py:75-76; -1 when the class is unseen) */
int llc_label_remap(const int64_t* y_global, const int64_t* lut, int lut_size, int64_t* y_local,
                    int n, void* stream);
/* per-step scalars: out[0] = sum(loss_rows), out[1] = #(pred == labels) */
int llc_loss_acc(const float* loss_rows, const int64_t* pred, const int64_t* labels, int n,
                 float* out2, void* stream);

/* ---- evaluation tail (methods/adapter_clip.py:132-176, methods/_trainer.py:519-534) ---------- */
/* counts [22] (uint64): counts[b] += #(y // n_tasks == b), counts[11 + b] += #(... and y == pred)
 * for the reference's ten bins b < 10; bins past ten (where the reference raises IndexError) land
 * in slot 10. cm [n_classes * n_classes] (uint64, may be NULL): cm[y * n_classes + pred] += 1. */
int llc_eval_accum(const int64_t* y, const int64_t* pred, int n, int n_tasks, int n_classes,
                   unsigned long long* cm, unsigned long long* counts, void* stream);
/* y = scale * x / |x| per row (model.py:966-969): fp32 [N, ld_y] and/or bf16 [N, ld_yb] (the
 * operand of the tensor-core logit GEMM) */
int llc_l2norm_rows(const float* x, int ld_x, int N, int E, float scale, float* y, int ld_y,
                    void* y_bf16, int ld_yb, void* stream);
/* probs = softmax(logits + add_mask) per row, pred = argmax (lowest index among ties);
 * models/adapter_clip.py:99, methods/adapter_clip.py:149; probs / pred / add_mask may be NULL */
int llc_softmax_argmax(const float* logits, int ld, int N, int C, const float* add_mask,
                       float* probs, int ld_p, int64_t* pred, void* stream);

/* ---- GPU input transform (methods/_trainer.py:236-247): Resize((S,S)) -> RandomCrop(S, padding)
 *      -> RandomHorizontalFlip -> Normalize(mean, std) on the raw batch, one pass ------------- */
typedef struct llc_img_transform {
  const void* src;        /* raw batch [N, 3, h, w]: uint8 0..255 (src_u8 = 1) or fp32 0..1 */
  int src_u8, h, w;
  int out_size;           /* S (even, >= h and w: bilinear up-sampling, align_corners=False) */
  int pad;                /* RandomCrop padding (zeros); 0 = no crop */
  int crop_i, crop_j;     /* crop offset inside the padded frame, 0..2*pad */
  int flip;               /* horizontal flip of the whole batch */
  const int* dyn_params;  /* device int[3] {crop_i, crop_j, flip} overriding the three fields
                             above (so a captured CUDA graph replays with new draws), or NULL */
  float mean[3], std[3];
} llc_img_transform;
/* fp32 NCHW [N, 3, S, S] (what train_transform / test_transform return) */
int llc_transform_images(const llc_img_transform* t, int N, float* out, void* stream);
/* bf16 im2col rows of the stride-P patch conv [N*G*G, ld_out] (as llc_patchify) */
int llc_transform_patchify(const llc_img_transform* t, int N, int P, void* out, int ld_out,
                           void* stream);

/* ---- fused AdamW on the flat LoRA buffer (utils/train_utils.py:27-28, torch.optim.AdamW) ----- */
int llc_adamw(float* p, const float* g, float* m, float* v, int n, float lr, float beta1,
              float beta2, float eps, float wd, int step, float grad_scale, void* stream);

/* ---- whole image tower (VisualTransformer.forward model.py:755-787 on LoRA blocks :400-415) - */
typedef struct llc_vit_cfg {
  int image_size, patch, width, layers, heads, mlp_dim, embed_dim;
  int lora_r;
  float lora_scale; /* alpha / r */
} llc_vit_cfg;

typedef struct llc_vit_layer {
  /* frozen, prepared bf16 K-major (llc_pack_weight); *_aug carry LLC_LORA_PAD extra K columns */
  /* "+16" = LLC_LORA_PAD K columns; the row pitch of these buffers is +LLC_LORA_LD (64) */
  void* wqkv_aug;  /* [3D, D+16]: W_in | s*B_in        */
  void* wo_aug;    /* [D, D+16]:  W_o  | s*B_o         */
  void* wfc;       /* [mlp, D]                          */
  void* wproj;     /* [D, mlp]                          */
  void* wqkvT_aug; /* [D, 3D+16]: W_in^T | A_in^T       */
  void* woT_aug;   /* [D, D+16]:  W_o^T  | A_o^T        */
  void* wfcT;      /* [D, mlp]                          */
  void* wprojT;    /* [mlp, D]                          */
  /* rank-r row products as skinny GEMMs on the tensor cores: [16, K] bf16 factors (rows >= r 0) */
  void* f_out_A;   /* [16, D]:  A_o            u_o  = o . A_o^T              (forward)  */
  void* f_in_B;    /* [16, 3D]: s * B_in^T     du_in = s dqkv . B_in         (backward) */
  void* f_out_B;   /* [16, D]:  s * B_o^T      du_o  = s dx_mid . B_o        (backward) */
  const float *bqkv, *bo, *bfc, *bproj, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  /* live fp32 LoRA parameters and their gradient slots (views of the flat buffers). All four
   * slots NULL = frozen attention (vanilla block with zero factors): the backward skips the LoRA
   * reductions; the pad columns of the dxb / dqkv scratch must then be zero on entry */
  const float *in_A, *in_B, *out_A, *out_B; /* [r,D] [3D,r] [r,D] [D,r] */
  float *g_in_A, *g_in_B, *g_out_A, *g_out_B;
} llc_vit_layer;

typedef struct llc_vit_weights {
  void* wpatch;            /* [D, 3*P*P] bf16 (conv1.weight viewed 2-D) */
  const float *class_emb, *pos_emb, *ln_pre_g, *ln_pre_b;
  const llc_vit_layer* layers;
} llc_vit_weights;

/* activation arena: sizes from llc_vit_arena_bytes; saved-for-backward tensors live here. Its
 * first 8 KB (the flag words of the GEMMs' stream-K workspace) must be ZERO before the first
 * call that uses the arena; every call leaves them zero. */
size_t llc_vit_arena_bytes(const llc_vit_cfg* cfg, int N, int training);
/* images fp32 NCHW [N,3,S,S] -> x_final fp32 [N*L, D] (pointer returned inside the arena) */
int llc_vit_forward(const llc_vit_cfg* cfg, const llc_vit_weights* w, const float* images, int N,
                    void* arena, int training, float** x_final, void* stream);
/* dx_final fp32 [N*L, D] (the head's gradient; consumed/overwritten) -> LoRA grads in w->layers */
int llc_vit_backward(const llc_vit_cfg* cfg, const llc_vit_weights* w, int N, void* arena,
                     float* dx_final, void* stream);
/* The same two calls when the consumer reads only the class-token rows of x_final (the tower's
 * own output, ln_post(x[:, 0, :]) @ proj, model.py:782-785): the last block then runs its
 * attention for one query per (sample, head) and its out-proj / LN2 / MLP on N rows instead of
 * N*L. llc_vit_forward_cls leaves the other rows of x_final undefined; llc_vit_backward_cls reads
 * only rows n*L of dx_final (the other rows need not be initialised). Results on the CLS rows and
 * all LoRA gradients are identical to the full calls. */
int llc_vit_forward_cls(const llc_vit_cfg* cfg, const llc_vit_weights* w, const float* images,
                        int N, void* arena, int training, float** x_final, void* stream);
int llc_vit_backward_cls(const llc_vit_cfg* cfg, const llc_vit_weights* w, int N, void* arena,
                         float* dx_final, void* stream);
/* llc_vit_forward[_cls] with the input transform fused in front of the patch embedding: the raw
 * batch described by `tx` replaces the fp32 images */
int llc_vit_forward_tx(const llc_vit_cfg* cfg, const llc_vit_weights* w,
                       const llc_img_transform* tx, int N, void* arena, int training, int cls_only,
                       float** x_final, void* stream);
/* dst bf16 [T, ld_dst] <- src fp32 [T, D] (contiguous rows) */
int llc_cast_bf16(const float* src, void* dst, int T, int D, int ld_dst, void* stream);
/* refresh the LoRA columns of the augmented weights from the live parameters */
int llc_vit_refresh_lora(const llc_vit_cfg* cfg, const llc_vit_weights* w, void* stream);

/* one transformer block on [T, D] fp32 (ResidualAttentionBlock.forward model.py:233-236) */
/* buffers written "[T, X+16]" have row pitch X + LLC_LORA_LD */
typedef struct llc_block_bufs {
  float* x_in;   /* [T, D] fp32 input (saved)                     */
  void* h1;      /* [T, D+16] bf16  LN1 out | u_in (saved)        */
  void* qkv;     /* [T, 3D+16] bf16 (saved; bwd reuses pad cols)  */
  float* lse;    /* [N*H*L]                                        */
  void* o;       /* [T, D+16] bf16 attn out | u_o (saved)         */
  float* x_mid;  /* [T, D] fp32 (saved)                           */
  void* h2;      /* [T, D] bf16 transient                         */
  void* z;       /* [T, mlp] bf16 (saved)                         */
  void* g;       /* [T, mlp] bf16 transient                       */
  float* x_out;  /* [T, D] fp32                                    */
} llc_block_bufs;
int llc_block_forward(const llc_vit_cfg* cfg, const llc_vit_layer* w, const llc_block_bufs* b,
                      int N, int L, int tok_stride_n, int tok_stride_l, int causal, void* stream);
typedef struct llc_block_bwd_bufs {
  float* dx;     /* [T, D] fp32 in: grad wrt x_out; out: grad wrt x_in (in place) */
  void* dxb;     /* [T, D+16] bf16: in: bf16 copy of dx; out: copy of the new dx  */
  void* dz;      /* [T, mlp] bf16 scratch                                          */
  void* dh;      /* [T, D] bf16 scratch                                            */
  void* d_o;     /* [T, D] bf16 scratch                                            */
  void* dqkv;    /* [T, 3D+16] bf16 scratch                                        */
  float* partial;/* llc_lora_side partials: max_partials * 3D * r floats           */
  float* delta;  /* [N * heads * L] fp32 scratch: rowsum(dO o O) of the attention backward */
  const float* dy; /* optional: the gradient of x_out in a buffer that must NOT be modified (an
                      autograd input). dx is then output only; NULL = dx holds it on entry */
} llc_block_bwd_bufs;
int llc_block_backward(const llc_vit_cfg* cfg, const llc_vit_layer* w, const llc_block_bufs* b,
                       const llc_block_bwd_bufs* s, int N, int L, int tok_stride_n,
                       int tok_stride_l, int causal, int need_dx_in, void* stream);

/* lora.MultiheadAttention.forward(x, x, x, need_weights=False) (lora.py:454-702 -> :732-1082) as
 * its own call: b->x_in = x fp32 [T, D] -> b->x_out fp32 [T, D]; h1 / qkv / lse / o are saved for
 * the backward. llc_mha_backward: s->dx = gradient of the output (fp32, read only) -> LoRA
 * gradients and, if need_dx_in, s->dh = gradient of x (bf16 [T, D]). */
int llc_mha_forward(const llc_vit_cfg* cfg, const llc_vit_layer* w, const llc_block_bufs* b, int N,
                    int L, int tok_stride_n, int tok_stride_l, int causal, void* stream);
int llc_mha_backward(const llc_vit_cfg* cfg, const llc_vit_layer* w, const llc_block_bufs* b,
                     const llc_block_bwd_bufs* s, int N, int L, int tok_stride_n, int tok_stride_l,
                     int causal, int need_dx_in, void* stream);

/* ---- bottleneck adapter of the adapter-clip method (models/clip/adapter.py:11-73) --------------
 * adapter(y) = y + scale * (dropout(relu(y W_d^T + b_d)) W_u^T + b_u), W_d [64, D], W_u [D, 64]
 * (down_proj is hard-coded 64 wide, adapter.py:39; scale = 0.1, dropout = 0.1, init "lora",
 * model.py:432-439). The live fp32 parameters are the module's; llc_adapter_refresh prepares the
 * bf16 operands from them (after every optimizer step). */
#define LLC_ADAPTER_DIM 64
typedef struct llc_adapter {
  float scale;             /* adapter_scalar */
  float dropout;           /* p, applied when the call says training */
  unsigned long long seed; /* dropout stream of this call (ignored with an explicit mask) */
  const float *down_w, *down_b, *up_w, *up_b;   /* [64, D] [64] [D, 64] [D] */
  float *g_down_w, *g_down_b, *g_up_w, *g_up_b; /* gradient slots (backward only) */
  void *wd, *wu, *wdT, *wuT;                    /* bf16 [64,D] [D,64] [D,64] [64,D] (wu, wuT scaled) */
  float* bu_s;                                  /* [D] scale * b_u */
  /* block use only (NULL for a stand-alone module): the frozen block's transposed weights with 64
   * spare K columns, which llc_adapter_refresh fills with the COMPOSED factors so that the
   * bottleneck gradient rides in the pad columns of the backward's big GEMMs:
   *   woT_ad    [D, D+64]   = the layer's woT_aug (W_o^T | pad): pad <- (W_d W_o)^T
   *   wprojT_ad [mlp, D+64] = W_proj^T | (W_d W_proj)^T        (left part filled by the caller) */
  void *woT_ad, *wprojT_ad;
  int mlp_dim;
} llc_adapter;
int llc_adapter_refresh(const llc_adapter* ad, int D, void* stream);
/* floats of the `partial` scratch of llc_adapter_backward */
size_t llc_adapter_partial_floats(int D);
/* out fp32 [T, D] = [resid +] [y +] scale * up(drop(relu(down(y)))): y bf16 [T, ld_y]; resid fp32
 * [T, D] or NULL; a bf16 [T, 64] receives the bottleneck after ReLU and dropout (saved for the
 * backward). mask: optional uint8 [T, 64] keep flags replacing the generated dropout stream
 * (seed, use). Adapter.forward(x) with add_residual: resid = x, y = bf16(x), add_y = 0; the
 * block's x + adaptmlp(branch): resid = x, y = branch, add_y = 1. */
int llc_adapter_forward(const llc_adapter* ad, const void* y, int ld_y, const float* resid,
                        int add_y, void* a, const unsigned char* mask, unsigned use, int training,
                        float* out, int T, int D, void* stream);
/* dx fp32 [T, D] = gradient of `out`, dxb its bf16 copy [T, ld_dxb]. Writes (accumulate = 0) or
 * adds (1: the block applies the same module twice) the four parameter gradients; d_y fp32 [T, D]
 * (may be NULL) = [dx +] dz W_d, the gradient of y. da: scratch bf16 [T, 64]. */
int llc_adapter_backward(const llc_adapter* ad, const void* y, int ld_y, const void* a,
                         const float* dx, const void* dxb, int ld_dxb, float* d_y, int pass_dx,
                         void* da, float* partial, int accumulate, int training, int T, int D,
                         void* stream);
/* ResidualAttentionBlock_Adapter.forward (model.py:440-442):
 *   x_mid = x + adaptmlp(attention(ln_1(x)));  x_out = x_mid + adaptmlp(mlp(ln_2(x_mid)))
 * on the frozen block `w` (zero LoRA factors) with ONE adapter shared by both branches. */
typedef struct llc_adapter_bufs {
  void* ya;  /* [T, D] bf16 attention branch (saved)   */
  void* a1;  /* [T, 64] bf16 (saved)                   */
  void* m;   /* [T, D] bf16 MLP branch (saved)         */
  void* a2;  /* [T, 64] bf16 (saved)                   */
  const unsigned char *mask1, *mask2; /* optional explicit dropout masks [T, 64] */
  void* da;        /* backward scratch [T, 64] bf16                        */
  float* d_branch; /* unused (kept for layout stability)                   */
  float* partial;  /* backward scratch, llc_adapter_partial_floats(D)      */
} llc_adapter_bufs;
int llc_adapter_block_forward(const llc_vit_cfg* cfg, const llc_vit_layer* w, const llc_adapter* ad,
                              const llc_block_bufs* b, const llc_adapter_bufs* ab, int N, int L,
                              int tok_stride_n, int tok_stride_l, int causal, int training,
                              void* stream);
/* s->dx = gradient of x_out (fp32, becomes the gradient of x_in when need_dx_in), s->dxb its bf16
 * copy on entry; adapter gradients -> ad->g_* (written, both branches summed) */
int llc_adapter_block_backward(const llc_vit_cfg* cfg, const llc_vit_layer* w, const llc_adapter* ad,
                               const llc_block_bufs* b, const llc_adapter_bufs* ab,
                               const llc_block_bwd_bufs* s, int N, int L, int tok_stride_n,
                               int tok_stride_l, int causal, int need_dx_in, int training,
                               void* stream);

/* ---- text tower with LoRA (CLIP.encode_text model.py:941-956, causal mask :926-932;
 *      peft_encoder='both' of scripts/lora_clip.sh) ------------------------------------------- */
typedef struct llc_text_weights {
  const float* tok_emb;  /* [vocab, D] fp32 (token_embedding.weight, frozen) */
  const float* pos_emb;  /* [context, D] fp32 (positional_embedding, frozen) */
  const llc_vit_layer* layers;
  int vocab, context;
} llc_text_weights;
/* cfg: width / layers / heads / mlp_dim / lora_* of the text transformer (patch geometry unused
 * but must be valid). tokens int64 [C, context] -> x_final fp32 [C*context, D] inside the arena;
 * ln_final + EOT gather + text_projection are llc_head_fwd with row_idx. */
size_t llc_text_arena_bytes(const llc_vit_cfg* cfg, int context, int C, int training);
int llc_text_forward(const llc_vit_cfg* cfg, const llc_text_weights* w, const int64_t* tokens,
                     int C, void* arena, int training, float** x_final, void* stream);
int llc_text_backward(const llc_vit_cfg* cfg, const llc_text_weights* w, int C, void* arena,
                      float* dx_final, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LLC_H_ */
