"""Thin torch-tensor wrappers over the C-ABI entry points (one Python function per llc_* symbol).
Every function launches on torch's current CUDA stream and raises RuntimeError on failure.
No function here computes anything on the host or through torch ops.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi as K

PAD = K.LORA_LD   # row-pitch pad of augmented buffers (the K extension itself is LORA_PAD)


def _lib():
    return K.load()


def _s():
    return K.stream_ptr()


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (libllc has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
    if t.stride(-1) != 1:
        raise RuntimeError(f"{name}: innermost dimension must be contiguous")


def check_device(dev: int = 0) -> None:
    K.check(_lib().llc_check_device(dev), "llc_check_device")


def launch_count() -> int:
    return int(_lib().llc_launch_count())


def prof_enable(on: bool) -> None:
    K.check(_lib().llc_prof_enable(int(on)), "llc_prof_enable")


def prof_read():
    """Per-launch records [(kind, m, n, k, ms, flops, bytes)] since prof_enable(True); blocks on
    the recorded events."""
    lib = _lib()
    n = lib.llc_prof_read(None, 0)
    if n < 0:
        K.check(n, "llc_prof_read")
    buf = (K.ProfRec * max(n, 1))()
    n = lib.llc_prof_read(buf, n)
    return [(K.PROF_KINDS[r.kind], r.m, r.n, r.k, r.ms, r.flops, r.bytes) for r in buf[:n]]


def gemm_tn(A, B, M, N, Kdim, out, *, bias=None, resid=None, act=0, aux=None, out2=None):
    """out[M,N] = epi(A[M,K] @ B[N,K]^T). A, B bf16 2-D (row stride = leading dim)."""
    _req(A, torch.bfloat16, "A"); _req(B, torch.bfloat16, "B")
    e = K.GemmEpi()
    e.bias = K.ptr(bias)
    e.resid = K.ptr(resid); e.ld_resid = resid.stride(0) if resid is not None else 0
    e.act = act
    e.aux = K.ptr(aux); e.ld_aux = aux.stride(0) if aux is not None else 0
    e.out = K.ptr(out); e.ld_out = out.stride(0) if out is not None else 0
    e.out_fp32 = int(out is not None and out.dtype == torch.float32)
    e.out2 = K.ptr(out2); e.ld_out2 = out2.stride(0) if out2 is not None else 0
    K.check(_lib().llc_gemm_bf16_tn(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N,
                                    Kdim, C.byref(e), _s()), "llc_gemm_bf16_tn")
    return out


def ln_fwd(x, gamma, beta, y, lora_A=None, r=0):
    _req(x, torch.float32, "x"); _req(y, torch.bfloat16, "y")
    T, D = x.shape[0], gamma.numel()
    K.check(_lib().llc_ln_fwd(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), T, D,
                              y.data_ptr(), y.stride(0), K.ptr(lora_A), r, _s()), "llc_ln_fwd")
    return y


def ln_bwd(x, gamma, dy, dx_in, dx_out, dxb=None, lora_B=None, r=0, scale=0.0):
    _req(x, torch.float32, "x"); _req(dy, torch.bfloat16, "dy")
    T, D = x.shape[0], gamma.numel()
    K.check(_lib().llc_ln_bwd(x.data_ptr(), x.stride(0), gamma.data_ptr(), dy.data_ptr(),
                              dy.stride(0), K.ptr(dx_in), dx_out.data_ptr(), T, D, K.ptr(dxb),
                              dxb.stride(0) if dxb is not None else 0, K.ptr(lora_B), r,
                              float(scale), _s()), "llc_ln_bwd")
    return dx_out


def attn_fwd(qkv, o, lse, N, L, H, sn, sl, causal=False):
    _req(qkv, torch.bfloat16, "qkv"); _req(o, torch.bfloat16, "o")
    K.check(_lib().llc_attn_fwd(qkv.data_ptr(), qkv.stride(0), o.data_ptr(), o.stride(0),
                                K.ptr(lse), N, L, H, sn, sl, int(causal), _s()), "llc_attn_fwd")
    return o


def attn_bwd(qkv, o, d_o, lse, dqkv, N, L, H, sn, sl, causal=False, delta=None):
    if delta is None:   # rowsum(dO o O) scratch: caller-owned, the library keeps no buffer
        delta = torch.empty(N * H * L, dtype=torch.float32, device=qkv.device)
    K.check(_lib().llc_attn_bwd(qkv.data_ptr(), qkv.stride(0), o.data_ptr(), o.stride(0),
                                d_o.data_ptr(), d_o.stride(0), lse.data_ptr(), dqkv.data_ptr(),
                                dqkv.stride(0), N, L, H, sn, sl, int(causal), delta.data_ptr(),
                                _s()), "llc_attn_bwd")
    return dqkv


def lora_side(X, T, Cc, r, Mrd=None, rd_sc=0, rd_sj=0, rd_scale=1.0, w=None, ld_w=0,
              partial=None):
    n = C.c_int(0)
    K.check(_lib().llc_lora_side(X.data_ptr(), X.stride(0), T, Cc, r, K.ptr(Mrd), rd_sc, rd_sj,
                                 float(rd_scale), K.ptr(w), ld_w, K.ptr(partial), C.byref(n),
                                 _s()), "llc_lora_side")
    return n.value


def lora_side_fused(X, T, Cc, r, w, ld_w, F, U, partial):
    """One pass over X: column sums with w into `partial`, row products with the packed factors
    F [16, C] into U [T, 16]. Returns the number of partial slices."""
    n = C.c_int(0)
    K.check(_lib().llc_lora_side_fused(X.data_ptr(), X.stride(0), T, Cc, r, w.data_ptr(), ld_w,
                                       F.data_ptr(), F.stride(0), U.data_ptr(), U.stride(0),
                                       partial.data_ptr(), C.byref(n), _s()),
            "llc_lora_side_fused")
    return n.value


def lora_colsum_finish(partial, n_partials, Cc, r, scale, out, o_sc, o_sj):
    K.check(_lib().llc_lora_colsum_finish(partial.data_ptr(), n_partials, Cc, r, float(scale),
                                          out.data_ptr(), o_sc, o_sj, _s()),
            "llc_lora_colsum_finish")
    return out


def lora_side_max_partials() -> int:
    return int(_lib().llc_lora_side_max_partials())


def pack_weight(src, dst, transpose=False):
    """dst bf16 [rows, ld] <- src fp32 (row-major [rows, cols], or [cols, rows] if transpose)."""
    _req(src, torch.float32, "src"); _req(dst, torch.bfloat16, "dst")
    rows, cols = (src.shape[1], src.shape[0]) if transpose else (src.shape[0], src.shape[1])
    K.check(_lib().llc_pack_weight(src.data_ptr(), rows, cols, int(transpose), dst.data_ptr(),
                                   dst.stride(0), _s()), "llc_pack_weight")
    return dst


def pack_lora_cols(src, rows, r, s_i, s_j, scale, dst, col0):
    K.check(_lib().llc_pack_lora_cols(src.data_ptr(), rows, r, s_i, s_j, float(scale),
                                      dst.data_ptr(), dst.stride(0), col0, _s()),
            "llc_pack_lora_cols")
    return dst


def patchify(img, P, out):
    _req(img, torch.float32, "img")
    N, Cc, HW, _ = img.shape
    K.check(_lib().llc_patchify(img.data_ptr(), N, Cc, HW, P, out.data_ptr(), out.stride(0), _s()),
            "llc_patchify")
    return out


def embed_ln_pre(patch_out, class_emb, pos, gamma, beta, N, L, D, x0):
    K.check(_lib().llc_embed_ln_pre(patch_out.data_ptr(), patch_out.stride(0), class_emb.data_ptr(),
                                    pos.data_ptr(), gamma.data_ptr(), beta.data_ptr(), N, L, D,
                                    x0.data_ptr(), _s()), "llc_embed_ln_pre")
    return x0


def label_remap(y_global, lut, y_local=None):
    if y_local is None:
        y_local = torch.empty_like(y_global)
    _req(y_global, torch.int64, "y_global"); _req(lut, torch.int64, "lut")
    K.check(_lib().llc_label_remap(y_global.data_ptr(), lut.data_ptr(), lut.numel(),
                                   y_local.data_ptr(), y_global.numel(), _s()), "llc_label_remap")
    return y_local


def loss_acc(loss_rows, pred, labels, out2):
    K.check(_lib().llc_loss_acc(loss_rows.data_ptr(), pred.data_ptr(), labels.data_ptr(),
                                loss_rows.numel(), out2.data_ptr(), _s()), "llc_loss_acc")
    return out2


def adamw(p, g, m, v, lr, beta1, beta2, eps, wd, step, grad_scale=1.0):
    K.check(_lib().llc_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                             lr, beta1, beta2, eps, wd, step, grad_scale, _s()), "llc_adamw")


def head_dtext(dlogits, fnorm, scale, d_text=None):
    """d_text [C, E] = scale * dlogits^T [C, N] @ fnorm [N, E] (gradient of the logit product w.r.t.
    the normalised text features; model.py:972 with a trainable text tower)."""
    N, Cn = dlogits.shape
    E = fnorm.shape[1]
    if d_text is None:
        d_text = torch.empty(Cn, E, device=dlogits.device)
    K.check(_lib().llc_head_dtext(dlogits.data_ptr(), fnorm.data_ptr(), N, Cn, E, float(scale),
                                  d_text.data_ptr(), _s()), "llc_head_dtext")
    return d_text


def eval_accum(y, pred, n_tasks, n_classes, cm, counts):
    """counts [22] int64: per-task totals / corrects of methods/_trainer.py:519-534 (slot 10 and
    21: bins past the reference's ten); cm [n_classes, n_classes] int64 confusion counts."""
    _req(y, torch.int64, "y"); _req(pred, torch.int64, "pred")
    K.check(_lib().llc_eval_accum(y.data_ptr(), pred.data_ptr(), y.numel(), int(n_tasks),
                                  int(n_classes), K.ptr(cm), counts.data_ptr(), _s()),
            "llc_eval_accum")


def l2norm_rows(x, scale=1.0, y=None, y_bf16=None):
    _req(x, torch.float32, "x")
    N, E = x.shape
    K.check(_lib().llc_l2norm_rows(x.data_ptr(), x.stride(0), N, E, float(scale), K.ptr(y),
                                   y.stride(0) if y is not None else 0, K.ptr(y_bf16),
                                   y_bf16.stride(0) if y_bf16 is not None else 0, _s()),
            "llc_l2norm_rows")


def softmax_argmax(logits, C_, add_mask=None, probs=None, pred=None):
    _req(logits, torch.float32, "logits")
    K.check(_lib().llc_softmax_argmax(logits.data_ptr(), logits.stride(0), logits.shape[0], C_,
                                      K.ptr(add_mask), K.ptr(probs),
                                      probs.stride(0) if probs is not None else 0, K.ptr(pred),
                                      _s()), "llc_softmax_argmax")


def cast_bf16(src, dst):
    """dst bf16 [T, ld] <- src fp32 [T, D] contiguous."""
    _req(src, torch.float32, "src"); _req(dst, torch.bfloat16, "dst")
    T, D = src.shape
    K.check(_lib().llc_cast_bf16(src.data_ptr(), dst.data_ptr(), T, D, dst.stride(0), _s()),
            "llc_cast_bf16")
    return dst


def make_transform(src, out_size, mean, std, pad=0, crop=(0, 0), flip=False, dyn=None):
    """llc_img_transform for a raw batch [N, 3, h, w] (uint8 0..255 or float32 0..1)."""
    if not src.is_cuda or src.dim() != 4 or src.shape[1] != 3 or not src.is_contiguous():
        raise RuntimeError("transform: expected a contiguous CUDA batch [N, 3, h, w]")
    if src.dtype not in (torch.uint8, torch.float32):
        raise RuntimeError(f"transform: uint8 or float32 input, got {src.dtype}")
    t = K.ImgTransform()
    t.src = src.data_ptr(); t.src_u8 = int(src.dtype == torch.uint8)
    t.h, t.w, t.out_size, t.pad = src.shape[2], src.shape[3], int(out_size), int(pad)
    t.crop_i, t.crop_j, t.flip = int(crop[0]), int(crop[1]), int(bool(flip))
    t.dyn_params = K.ptr(dyn)
    for i in range(3):
        t.mean[i] = float(mean[i]); t.std[i] = float(std[i])
    return t


def transform_images(t, N, out):
    K.check(_lib().llc_transform_images(C.byref(t), N, out.data_ptr(), _s()),
            "llc_transform_images")
    return out


def transform_patchify(t, N, P, out):
    K.check(_lib().llc_transform_patchify(C.byref(t), N, P, out.data_ptr(), out.stride(0), _s()),
            "llc_transform_patchify")
    return out


class Head:
    """Argument block of llc_head_fwd / llc_head_bwd; owns the small output tensors."""

    def __init__(self, x, cls_stride, ln_g, ln_b, proj, text, logit_scale_exp, N, *, cls_idx=None,
                 add_mask=None, labels=None, double_softmax=True, inv_batch=None, row_idx=None,
                 want_dlogits=False):
        D, E = proj.shape
        Cn = cls_idx.numel() if cls_idx is not None else text.shape[0]
        dev = x.device
        self.N, self.D, self.E, self.C = N, D, E, Cn
        self.keep = (x, ln_g, ln_b, proj, text, cls_idx, add_mask, labels, row_idx)
        self.dlogits = torch.empty(N, Cn, device=dev) if want_dlogits else None
        self.feat = torch.empty(N, E, device=dev)
        self.fnorm = torch.empty(N, E, device=dev)
        self.logits = torch.empty(N, Cn, device=dev)
        self.probs = torch.empty(N, Cn, device=dev)
        self.loss_rows = torch.zeros(N, device=dev)
        self.pred = torch.empty(N, dtype=torch.int64, device=dev)
        a = K.HeadArgs()
        a.x = x.data_ptr(); a.cls_stride = cls_stride; a.ld_x = x.stride(0)
        a.ln_g = ln_g.data_ptr(); a.ln_b = ln_b.data_ptr(); a.proj = proj.data_ptr()
        a.text = text.data_ptr(); a.cls_idx = K.ptr(cls_idx); a.add_mask = K.ptr(add_mask)
        a.logit_scale = float(logit_scale_exp)
        a.N, a.D, a.E, a.C = N, D, E, Cn
        a.labels = K.ptr(labels); a.double_softmax = int(double_softmax)
        a.inv_batch = float(inv_batch if inv_batch is not None else 1.0 / N)
        a.feat = self.feat.data_ptr(); a.fnorm = self.fnorm.data_ptr()
        a.logits = self.logits.data_ptr(); a.probs = self.probs.data_ptr()
        a.loss_rows = self.loss_rows.data_ptr(); a.pred = self.pred.data_ptr()
        a.row_idx = K.ptr(row_idx); a.dlogits = K.ptr(self.dlogits)
        self.args = a

    def forward(self):
        K.check(_lib().llc_head_fwd(C.byref(self.args), _s()), "llc_head_fwd")
        return self

    def backward(self, dx, d_probs=None, loss_scale=1.0):
        K.check(_lib().llc_head_bwd(C.byref(self.args), K.ptr(d_probs), float(loss_scale),
                                    dx.data_ptr(), dx.stride(0), _s()), "llc_head_bwd")
        return dx
