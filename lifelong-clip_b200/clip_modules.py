"""Drop-in mirrors of the reference's vision-tower modules (models/clip/model.py, models/clip/lora.py)
backed by libllc.so.

Same constructor signatures, parameter names and state_dict keys as the reference classes, so
OpenAI-CLIP checkpoints load and the `"lora" in name` freeze filter of
methods/adapter_clip.py:115-119 keeps working:

    LayerNorm, QuickGELU                         model.py:194-206
    Linear (LoRA)                                lora.py:100-173
    MultiheadAttention (LoRA)                    lora.py:371-452 (parameters), :732-1082 (math)
    ResidualAttentionBlock[_LoRA]                model.py:209-236, :400-415
    Transformer                                  model.py:639-686
    VisualTransformer                            model.py:689-787

Compute: ResidualAttentionBlock_LoRA.forward -> llc_block_forward/backward on the reference's
[L, N, D] layout (token strides (1, N), no copy); VisualTransformer.forward -> one llc_vit_forward
call for the whole tower on [N, L, D]. The frozen backbone gets no weight gradients (they are
never computed); only the LoRA factors do.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _capi as K
from . import ops

PAD = K.LORA_LD   # row-pitch pad of augmented buffers (the K extension itself is LORA_PAD)


class LayerNorm(nn.LayerNorm):
    """model.py:194-200. On the hot path LayerNorm is fused into the tower kernels
    (llc_ln_fwd/llc_ln_bwd); this forward only serves callers that reach into the module directly
    (e.g. models/mvp_clip.py:259-261) and keeps the reference semantics (fp32 compute)."""

    def forward(self, x: torch.Tensor):
        orig_type = x.dtype
        ret = super().forward(x.type(torch.float32))
        return ret.type(orig_type)


class QuickGELU(nn.Module):
    """model.py:203-206 (fused into the c_fc GEMM epilogue on the hot path)."""

    def forward(self, x: torch.Tensor):
        return x * torch.sigmoid(1.702 * x)


class _LoraLinearFn(torch.autograd.Function):
    """y = x W^T + b + s (x A^T) B^T (lora.py:162-173) on libllc: the rank-r update rides as 16
    extra K columns of the same tcgen05 GEMM; backward gives dx and the two LoRA gradients, the
    frozen weight/bias get none."""

    @staticmethod
    def forward(ctx, lin, x, A, B):
        dev = x.device
        if dev.type != "cuda":
            raise RuntimeError("lifelong_clip_b200 modules compute on CUDA (sm_100a) only; "
                               "move the module and its input to the GPU (no CPU fallback)")
        I, O, r, s = lin.in_features, lin.out_features, lin.r, lin.scaling
        x2 = x.detach().float().contiguous().view(-1, I)
        T = x2.shape[0]
        xb = _bf16(T, I + PAD, dev)
        ops.cast_bf16(x2, xb)
        f32 = lambda p: p.detach().float().contiguous()
        A32, B32 = f32(A), f32(B)
        ops.lora_side(xb, T, I, r, Mrd=A32, rd_sc=1, rd_sj=I, rd_scale=1.0)      # u = x A^T
        w_aug = ops.pack_weight(f32(lin.weight), _bf16(O, I + PAD, dev))
        ops.pack_lora_cols(B32, O, r, r, 1, s, w_aug, I)                          # | s B
        y = torch.empty(T, O, device=dev)
        ops.gemm_tn(xb, w_aug, T, O, I + K.LORA_PAD, y,
                    bias=f32(lin.bias) if lin.bias is not None else None)
        ctx.lin, ctx.xb, ctx.A32, ctx.B32, ctx.shape = lin, xb, A32, B32, x.shape
        return y.view(*x.shape[:-1], O)

    @staticmethod
    def backward(ctx, dy):
        lin, xb, A32, B32 = ctx.lin, ctx.xb, ctx.A32, ctx.B32
        I, O, r, s = lin.in_features, lin.out_features, lin.r, lin.scaling
        dev = dy.device
        d2 = dy.detach().float().contiguous().view(-1, O)
        T = d2.shape[0]
        db = _bf16(T, O + PAD, dev)
        ops.cast_bf16(d2, db)
        part = torch.empty(ops.lora_side_max_partials() * max(I, O) * 8, device=dev)
        # du = s dy B (-> pad columns of db); dB = s dy^T u
        n1 = ops.lora_side(db, T, O, r, Mrd=B32, rd_sc=r, rd_sj=1, rd_scale=s,
                           w=xb[:, I:], ld_w=xb.stride(0), partial=part)
        gB = torch.empty(O, r, device=dev)
        ops.lora_colsum_finish(part, n1, O, r, s, gB, r, 1)
        # dA = du^T x
        n2 = ops.lora_side(xb, T, I, r, w=db[:, O:], ld_w=db.stride(0), partial=part)
        gA = torch.empty(r, I, device=dev)
        ops.lora_colsum_finish(part, n2, I, r, 1.0, gA, 1, I)
        gx = None
        if ctx.needs_input_grad[1]:
            # dx = dy W + du A : one GEMM over K = O + 16 against [W^T | A^T]
            wT = ops.pack_weight(lin.weight.detach().float().contiguous(),
                                 _bf16(I, O + PAD, dev), True)
            ops.pack_lora_cols(A32, I, r, 1, I, 1.0, wT, O)
            gx = torch.empty(T, I, device=dev)
            ops.gemm_tn(db, wT, T, I, O + K.LORA_PAD, gx)
            gx = gx.view(ctx.shape)
        return None, gx, gA, gB


class Linear(nn.Linear):
    """lora.py:100-173 LoRA dense layer (out_proj of the LoRA attention): weight, bias, lora_A
    kaiming-uniform(a=sqrt 5), lora_B zeros, scaling = alpha / r; forward adds the rank-r update.
    Inside a block the layer is only a parameter container (the projection is fused into
    llc_block_forward); called on its own it runs _LoraLinearFn."""

    def __init__(self, in_features, out_features, r=0, lora_alpha=1, lora_dropout=0.,
                 fan_in_fan_out=False, merge_weights=True, **kwargs):
        super().__init__(in_features, out_features, **kwargs)
        if lora_dropout != 0.:
            raise NotImplementedError("LoRA dropout is 0 on the reference path (lora.py:376)")
        if fan_in_fan_out:
            raise NotImplementedError("fan_in_fan_out is never set on the reference path")
        self.r, self.lora_alpha, self.merged, self.merge_weights = r, lora_alpha, False, merge_weights
        self.fan_in_fan_out = fan_in_fan_out
        if r > 0:
            self.lora_A = nn.Parameter(self.weight.new_zeros((r, in_features)))
            self.lora_B = nn.Parameter(self.weight.new_zeros((out_features, r)))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = False
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)

    def train(self, mode: bool = True):
        """lora.py:141-160 merges W += s B A in eval mode when merge_weights is set; the reference
        constructs out_proj with merge_weights=False (lora.py:430-435), the only supported use."""
        if self.merge_weights and self.r > 0 and not mode:
            raise NotImplementedError("merge_weights=True (weight merging in eval mode) is not "
                                      "used on the reference path")
        return super().train(mode)

    def forward(self, x: torch.Tensor):
        if self.r > 0 and not self.merged:
            return _LoraLinearFn.apply(self, x, self.lora_A, self.lora_B)
        raise NotImplementedError("LoRA rank 0 / merged weights: use nn.Linear")


class _MhaFn(torch.autograd.Function):
    """lora.MultiheadAttention.forward(x, x, x) through llc_mha_forward / llc_mha_backward on the
    reference's [L, N, D] layout (token strides (1, N))."""

    @staticmethod
    def forward(ctx, attn, x, causal, *lora):
        L, N, D = x.shape
        T, dev = L * N, x.device
        if dev.type != "cuda":
            raise RuntimeError("lifelong_clip_b200 modules compute on CUDA (sm_100a) only; "
                               "move the module and its input to the GPU (no CPU fallback)")
        bufs = dict(x_in=x.detach().float().contiguous().view(T, D), h1=_bf16p(T, D, dev),
                    qkv=_bf16p(T, 3 * D, dev),
                    lse=torch.empty(N * attn.num_heads * L, device=dev), o=_bf16p(T, D, dev),
                    x_out=torch.empty(T, D, device=dev))
        b = K.BlockBufs()
        for k, v in bufs.items():
            setattr(b, k, K.ptr(v))
        layer = attn._layer_struct(lora, [None] * 4)
        attn._refresh_lora(layer)
        K.check(K.load().llc_mha_forward(C.byref(attn._cfg), C.byref(layer), C.byref(b), N, L, 1,
                                         N, int(causal), K.stream_ptr()), "llc_mha_forward")
        ctx.attn, ctx.bufs, ctx.shape, ctx.causal, ctx.lora = attn, bufs, (L, N, D), causal, lora
        return bufs["x_out"].view(L, N, D)

    @staticmethod
    def backward(ctx, dy):
        attn, bufs, (L, N, D) = ctx.attn, ctx.bufs, ctx.shape
        T, dev = L * N, dy.device
        train_lora = any(p.requires_grad for p in ctx.lora)
        grads = [torch.zeros_like(p, dtype=torch.float32) if train_lora else None
                 for p in ctx.lora]
        scratch = dict(
            dx=dy.detach().float().contiguous().view(T, D), dxb=_bf16p(T, D, dev),
            dh=_bf16(T, D, dev), d_o=_bf16(T, D, dev), dqkv=_bf16p(T, 3 * D, dev),
            partial=torch.empty(ops.lora_side_max_partials() * 3 * D * 8, device=dev),
            delta=torch.empty(N * attn.num_heads * L, device=dev))
        s = K.BlockBwdBufs()
        for k, v in scratch.items():
            setattr(s, k, v.data_ptr())
        b = K.BlockBufs()
        for k, v in bufs.items():
            setattr(b, k, K.ptr(v))
        layer = attn._layer_struct(ctx.lora, grads)
        need_dx = bool(ctx.needs_input_grad[1])
        K.check(K.load().llc_mha_backward(C.byref(attn._cfg), C.byref(layer), C.byref(b),
                                          C.byref(s), N, L, 1, N, int(ctx.causal), int(need_dx),
                                          K.stream_ptr()), "llc_mha_backward")
        gx = scratch["dh"].float().view(L, N, D) if need_dx else None
        return (None, gx, None) + tuple(
            g.to(p.dtype) if (g is not None and p.requires_grad) else None
            for g, p in zip(grads, ctx.lora))


_CAUSAL_CHECKED = {}


def _causal_flag(attn_mask, L: int) -> int:
    """None -> 0; the text tower's additive -inf upper triangle (model.py:926-932) -> 1. The
    comparison (a device-to-host read when the mask lives on the GPU) runs once per mask object
    and length, not once per block call."""
    if attn_mask is None:
        return 0
    key = (id(attn_mask), attn_mask._version, attn_mask.data_ptr(), L)
    if _CAUSAL_CHECKED.get(key):
        return 1
    want = torch.full((L, L), float("-inf")).triu(1)
    m = attn_mask.detach().float().cpu()
    ok = (tuple(m.shape) == (L, L) and torch.equal(m, want)) or \
        (m.dim() == 2 and m.shape[0] >= L and torch.equal(m[:L, :L], want))
    if not ok:   # (second form: a mask built for the full context, applied to a shorter sequence)
        raise NotImplementedError("only attn_mask=None or the causal mask is supported")
    if len(_CAUSAL_CHECKED) > 256:
        _CAUSAL_CHECKED.clear()
    _CAUSAL_CHECKED[key] = True
    return 1


class MultiheadAttention(nn.Module):
    """lora.py:371-452: in_proj_weight/bias + one rank-r (A [r,D], B [3D,r]) pair shared by q,k,v
    (both xavier-uniform), out_proj = LoRA Linear. forward (lora.py:454-702) supports what the
    reference's blocks call: self-attention on [L, N, D], need_weights=False, attn_mask None or
    causal. Inside ResidualAttentionBlock_LoRA.forward the same math runs fused in
    llc_block_forward."""

    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False,
                 add_zero_attn=False, kdim=None, vdim=None, lora_alpha: int = 1, r: int = 0):
        super().__init__()
        assert r > 0
        if dropout != 0. or add_bias_kv or add_zero_attn or not bias or \
                (kdim not in (None, embed_dim)) or (vdim not in (None, embed_dim)):
            raise NotImplementedError("only the configuration the reference instantiates "
                                      "(model.py:412-415) is supported")
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.kdim = self.vdim = embed_dim
        self._qkv_same_embed_dim = True
        self.dropout, self.batch_first = 0., False
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim
        self.lora_alpha, self.r = lora_alpha, r
        self.scaling = lora_alpha / r
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.in_proj_weight_lora_A = nn.Parameter(torch.empty(r, embed_dim))
        self.in_proj_weight_lora_B = nn.Parameter(torch.empty(3 * embed_dim, r))
        self.out_proj = Linear(embed_dim, embed_dim, bias=True, merge_weights=False,
                               lora_alpha=lora_alpha, r=r)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.)
        nn.init.xavier_uniform_(self.in_proj_weight_lora_A)
        nn.init.xavier_uniform_(self.in_proj_weight_lora_B)
        self._packed = None
        self._cfg = _cfg_struct(embed_dim, num_heads, embed_dim * 4, r, lora_alpha / r)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def lora_params(self):
        return (self.in_proj_weight_lora_A, self.in_proj_weight_lora_B, self.out_proj.lora_A,
                self.out_proj.lora_B)

    def _layer_struct(self, lora, grads) -> K.VitLayer:
        if self._packed is None:
            self._packed = PackedLayer(self, None)
        s = K.VitLayer()
        self._packed.fill(s, lora, grads)
        return s

    def _refresh_lora(self, s: K.VitLayer):
        w = K.VitWeights()
        w.layers = (K.VitLayer * 1)(s)
        K.check(K.load().llc_vit_refresh_lora(C.byref(self._cfg), C.byref(w), K.stream_ptr()),
                "llc_vit_refresh_lora")

    def forward(self, query, key=None, value=None, key_padding_mask=None, need_weights=True,
                attn_mask=None, average_attn_weights=True, is_causal=False):
        if key is None:
            key = query
        if value is None:
            value = query
        if key is not query or value is not query:
            raise NotImplementedError("only self-attention (q is k is v), the call of "
                                      "model.py:226-231, is implemented")
        if need_weights or key_padding_mask is not None:
            raise NotImplementedError("need_weights=True / key_padding_mask: the reference's "
                                      "blocks call with need_weights=False (model.py:230)")
        if query.dim() != 3 or query.shape[-1] != self.embed_dim:
            raise RuntimeError(f"expected [L, N, {self.embed_dim}], got {tuple(query.shape)}")
        causal = 1 if (is_causal and attn_mask is None) else _causal_flag(attn_mask,
                                                                          query.shape[0])
        out = _MhaFn.apply(self, query, causal, *self.lora_params())
        return out, None


LORA_NAMES = ("attn.in_proj_weight_lora_A", "attn.in_proj_weight_lora_B",
              "attn.out_proj.lora_A", "attn.out_proj.lora_B")


def _bf16(rows, cols, device):
    """Zero-filled bf16 buffer: the K-augmented operands (pad columns must hold zeros where no
    kernel writes them)."""
    return torch.zeros(rows, cols, dtype=torch.bfloat16, device=device)


def _bf16p(rows, cols, device):
    """bf16 [rows, cols + PAD] whose PAD columns are zero and whose first `cols` columns are
    uninitialised: the activation-sized K-augmented buffers (h1, qkv, o, dxb, dqkv). Their body is
    always written by a kernel; zero-filling all of it cost 0.7 GB of memset per block call."""
    t = torch.empty(rows, cols + PAD, dtype=torch.bfloat16, device=device)
    t[:, cols:].zero_()
    return t


def _bf16e(rows, cols, device):
    """Uninitialised bf16 buffer (every element is written before it is read)."""
    return torch.empty(rows, cols, dtype=torch.bfloat16, device=device)


class PackedLayer:
    """Prepared operands of one block: frozen weights as bf16 K-major matrices (forward and
    transposed for the activation-gradient GEMMs), 16 spare K columns for the LoRA factors."""

    def __init__(self, attn: "MultiheadAttention", blk=None):
        """attn: the block's attention module; blk: the block (None: attention-only packing for
        MultiheadAttention.forward, the MLP / LayerNorm slots then stay empty)."""
        a = attn
        D, dev = a.embed_dim, a.in_proj_weight.device
        if dev.type != "cuda":
            raise RuntimeError("lifelong_clip_b200 modules compute on CUDA (sm_100a) only; "
                               "move the module to the GPU first (no CPU fallback)")
        f32 = lambda p: p.detach().float().contiguous()
        self.wqkv_aug = ops.pack_weight(f32(a.in_proj_weight), _bf16(3 * D, D + PAD, dev))
        self.wo_aug = ops.pack_weight(f32(a.out_proj.weight), _bf16(D, D + PAD, dev))
        self.wqkvT_aug = ops.pack_weight(f32(a.in_proj_weight), _bf16(D, 3 * D + PAD, dev), True)
        self.woT_aug = ops.pack_weight(f32(a.out_proj.weight), _bf16(D, D + PAD, dev), True)
        self.wfc = self.wproj = self.wfcT = self.wprojT = None
        if blk is not None:
            M = blk.mlp.c_fc.out_features
            self.wfc = ops.pack_weight(f32(blk.mlp.c_fc.weight), _bf16(M, D, dev))
            self.wproj = ops.pack_weight(f32(blk.mlp.c_proj.weight), _bf16(D, M, dev))
            self.wfcT = ops.pack_weight(f32(blk.mlp.c_fc.weight), _bf16(D, M, dev), True)
            self.wprojT = ops.pack_weight(f32(blk.mlp.c_proj.weight), _bf16(M, D, dev), True)
        self.f_out_A = _bf16(16, D, dev)       # refreshed from the live LoRA factors every step
        self.f_in_B = _bf16(16, 3 * D, dev)
        self.f_out_B = _bf16(16, D, dev)
        if blk is not None:
            self.small = [f32(p) for p in (a.in_proj_bias, a.out_proj.bias, blk.mlp.c_fc.bias,
                                           blk.mlp.c_proj.bias, blk.ln_1.weight, blk.ln_1.bias,
                                           blk.ln_2.weight, blk.ln_2.bias)]
        else:
            self.small = [f32(a.in_proj_bias), f32(a.out_proj.bias)] + [None] * 6

    def fill(self, s: K.VitLayer, lora, grads):
        for n in ("wqkv_aug", "wo_aug", "wfc", "wproj", "wqkvT_aug", "woT_aug", "wfcT", "wprojT",
                  "f_out_A", "f_in_B", "f_out_B"):
            setattr(s, n, K.ptr(getattr(self, n)))
        for n, t in zip(("bqkv", "bo", "bfc", "bproj", "ln1_g", "ln1_b", "ln2_g", "ln2_b"),
                        self.small):
            setattr(s, n, K.ptr(t))
        for n, t in zip(("in_A", "in_B", "out_A", "out_B"), lora):
            setattr(s, n, t.data_ptr())
        for n, t in zip(("g_in_A", "g_in_B", "g_out_A", "g_out_B"), grads):
            setattr(s, n, t.data_ptr() if t is not None else None)


def _cfg_struct(width, heads, mlp_dim, r, scale, layers=1, image_size=32, patch=16, embed_dim=1):
    c = K.VitCfg()
    c.image_size, c.patch, c.width, c.layers = image_size, patch, width, layers
    c.heads, c.mlp_dim, c.embed_dim, c.lora_r, c.lora_scale = heads, mlp_dim, embed_dim, r, scale
    return c


class _BlockFn(torch.autograd.Function):
    """One block on the reference's [L, N, D] layout through llc_block_forward/backward."""

    @staticmethod
    def forward(ctx, blk, x, *lora):
        L, N, D = x.shape
        T, M, dev = L * N, blk.mlp.c_fc.out_features, x.device
        x2 = x.detach().float().contiguous().view(T, D)
        # (grad mode is off inside Function.forward: ask autograd which inputs need gradients)
        need_grad = any(ctx.needs_input_grad)
        bufs = dict(
            x_in=x2, h1=_bf16p(T, D, dev), qkv=_bf16p(T, 3 * D, dev),
            lse=torch.empty(N * blk.n_head * L, device=dev), o=_bf16p(T, D, dev),
            x_mid=torch.empty(T, D, device=dev), h2=_bf16e(T, D, dev),
            z=_bf16e(T, M, dev) if need_grad else None, g=_bf16e(T, M, dev),
            x_out=torch.empty(T, D, device=dev))
        b = K.BlockBufs()
        for k, v in bufs.items():
            setattr(b, k, K.ptr(v))
        layer = blk._layer_struct(lora, [None] * 4)
        blk._refresh_lora(layer)
        causal = blk._causal_flag(L)
        K.check(K.load().llc_block_forward(C.byref(blk._cfg), C.byref(layer), C.byref(b), N, L,
                                           1, N, causal, K.stream_ptr()), "llc_block_forward")
        out = bufs.pop("x_out")
        bufs["h2"] = bufs["g"] = None      # not read by the backward: freed with the call
        ctx.blk, ctx.bufs, ctx.shape, ctx.causal = blk, bufs, (L, N, D), causal
        ctx.x_needs_grad = x.requires_grad
        ctx.lora = lora
        return out.view(L, N, D)

    @staticmethod
    def backward(ctx, dy):
        blk, bufs, (L, N, D) = ctx.blk, ctx.bufs, ctx.shape
        T, M, dev = L * N, blk.mlp.c_fc.out_features, dy.device
        if bufs["z"] is None:
            raise RuntimeError("block forward ran without grad; cannot backpropagate")
        # dy belongs to autograd and is only read (llc_block_bwd_bufs.dy); dx receives the
        # gradient of the block's input
        dyc = dy.detach().float().contiguous().view(T, D)
        dx = torch.empty(T, D, device=dev)
        # vanilla (frozen) block: no gradient slots -> the backward skips the LoRA reductions
        # (include/llc.h llc_vit_layer); the scratch below is zero-filled as that path requires
        train_lora = any(p.requires_grad for p in ctx.lora)
        grads = [torch.zeros_like(p, dtype=torch.float32) if train_lora else None
                 for p in ctx.lora]
        s = K.BlockBwdBufs()
        scratch = dict(
            dx=dx, dxb=_bf16p(T, D, dev), dz=_bf16e(T, M, dev), dh=_bf16e(T, D, dev),
            d_o=_bf16e(T, D, dev), dqkv=_bf16p(T, 3 * D, dev),
            partial=torch.empty(ops.lora_side_max_partials() * 3 * D * 8, device=dev),
            delta=torch.empty(N * blk.n_head * L, device=dev))
        for k, v in scratch.items():
            setattr(s, k, v.data_ptr())
        s.dy = dyc.data_ptr()
        lib = K.load()
        K.check(lib.llc_cast_bf16(dyc.data_ptr(), scratch["dxb"].data_ptr(), T, D, D + PAD,
                                  K.stream_ptr()), "llc_cast_bf16")
        b = K.BlockBufs()
        for k, v in bufs.items():
            setattr(b, k, K.ptr(v))
        layer = blk._layer_struct(ctx.lora, grads)
        K.check(lib.llc_block_backward(C.byref(blk._cfg), C.byref(layer), C.byref(b), C.byref(s),
                                       N, L, 1, N, ctx.causal, int(ctx.x_needs_grad),
                                       K.stream_ptr()), "llc_block_backward")
        gx = dx.view(L, N, D) if ctx.x_needs_grad else None
        return (None, gx) + tuple(g.to(p.dtype) if (g is not None and p.requires_grad) else None
                                  for g, p in zip(grads, ctx.lora))


class FrozenMultiheadAttention(nn.Module):
    """Parameter container with nn.MultiheadAttention's state_dict keys (in_proj_weight,
    in_proj_bias, out_proj.weight, out_proj.bias) for the vanilla block (model.py:217). The LoRA
    slots are zero buffers (not Parameters, not in the state_dict), so the same kernels compute
    exactly x W^T + b."""

    def __init__(self, embed_dim, num_heads, r=4):
        super().__init__()
        self.embed_dim, self.num_heads, self.r = embed_dim, num_heads, r
        self.head_dim = embed_dim // num_heads
        self.batch_first, self.dropout = False, 0.
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)       # keys out_proj.weight / .bias
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.)
        for n, shape in (("_z_in_A", (r, embed_dim)), ("_z_in_B", (3 * embed_dim, r)),
                         ("_z_out_A", (r, embed_dim)), ("_z_out_B", (embed_dim, r))):
            self.register_buffer(n, torch.zeros(shape), persistent=False)

    def lora_params(self):
        return (self._z_in_A, self._z_in_B, self._z_out_A, self._z_out_B)


class ResidualAttentionBlock(nn.Module):
    """model.py:209-236: x + attn(ln_1(x)); x + mlp(ln_2(x)) on [L, N, D], frozen (no weight
    gradients are ever computed on this path; the input gradient is)."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.d_model, self.n_head = d_model, n_head
        self.lora_alpha, self.lora_r = 1, 4      # zero factors: see FrozenMultiheadAttention
        self.attn = FrozenMultiheadAttention(d_model, n_head, r=self.lora_r)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)),
                                              ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask
        self._packed = None
        self._cfg = _cfg_struct(d_model, n_head, d_model * 4, self.lora_r,
                                self.lora_alpha / self.lora_r)

    # -- packed-weight cache ------------------------------------------------------------------
    def invalidate_packed(self):
        """Call after modifying frozen weights in place (load_state_dict and .to()/.cuda() do it
        automatically)."""
        self._packed = None
        if hasattr(self.attn, "_packed"):
            self.attn._packed = None

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def packed(self) -> PackedLayer:
        if self._packed is None:
            self._packed = PackedLayer(self.attn, self)
        return self._packed

    def lora_params(self):
        return self.attn.lora_params()

    def _layer_struct(self, lora, grads) -> K.VitLayer:
        s = K.VitLayer()
        for p in lora:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("LoRA parameters must be contiguous fp32")
        self.packed().fill(s, lora, grads)
        return s

    def _refresh_lora(self, s: K.VitLayer):
        cfg = self._cfg
        w = K.VitWeights()
        arr = (K.VitLayer * 1)(s)
        w.layers = arr
        K.check(K.load().llc_vit_refresh_lora(C.byref(cfg), C.byref(w), K.stream_ptr()),
                "llc_vit_refresh_lora")

    def _causal_flag(self, L: int) -> int:
        return _causal_flag(self.attn_mask, L)

    def attention(self, x: torch.Tensor):
        """model.py:226-231: self.attn(x, x, x, need_weights=False, attn_mask=self.attn_mask)[0]."""
        if isinstance(self.attn, MultiheadAttention):
            return self.attn(x, x, x, need_weights=False, attn_mask=self.attn_mask)[0]
        # vanilla block: the same kernels with zero LoRA factors
        if getattr(self, "_plain_mha", None) is None:
            self._plain_mha = _PlainMha(self.attn)
        return _MhaFn.apply(self._plain_mha, x, self._causal_flag(x.shape[0]),
                            *self.attn.lora_params())

    def forward(self, x: torch.Tensor):
        if x.dim() != 3 or x.shape[-1] != self.d_model:
            raise RuntimeError(f"expected [L, N, {self.d_model}], got {tuple(x.shape)}")
        return _BlockFn.apply(self, x, *self.lora_params())


class _PlainMha:
    """Adapter giving a FrozenMultiheadAttention the packing interface _MhaFn expects."""

    def __init__(self, attn: FrozenMultiheadAttention):
        self.attn, self.num_heads = attn, attn.num_heads
        self._cfg = _cfg_struct(attn.embed_dim, attn.num_heads, attn.embed_dim * 4, attn.r,
                                1.0 / attn.r)
        self._packed = None
        self._key = None

    def _layer_struct(self, lora, grads):
        key = (self.attn.in_proj_weight.data_ptr(), self.attn.in_proj_weight._version)
        if self._packed is None or key != self._key:
            self._packed, self._key = PackedLayer(self.attn, None), key
        s = K.VitLayer()
        self._packed.fill(s, lora, grads)
        return s

    _refresh_lora = MultiheadAttention._refresh_lora


class ResidualAttentionBlock_LoRA(ResidualAttentionBlock):
    """model.py:400-415: the vanilla block with lora.MultiheadAttention as `attn`."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None,
                 design_details: dict = {}):
        super().__init__(d_model, n_head, attn_mask)
        self.lora_alpha = design_details.get('lora_alpha', 1)
        self.lora_r = design_details.get('lora_r', 4)
        self.attn = MultiheadAttention(d_model, n_head, lora_alpha=self.lora_alpha, r=self.lora_r)
        self._cfg = _cfg_struct(d_model, n_head, d_model * 4, self.lora_r,
                                self.lora_alpha / self.lora_r)


class Transformer(nn.Module):
    """model.py:639-686: LoRA blocks when method='lora' (adapter blocks when method='adapter')
    and peft_encoder covers this modality, vanilla (frozen) blocks otherwise. The MoE / prefix
    flavours belong to other methods of the reference (out of scope)."""

    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None,
                 design_details: dict = {}, modal='text'):
        super().__init__()
        self.width, self.layers = width, layers
        res_type = design_details.get('method', 'vanilla')
        peft_flag = design_details.get('peft_encoder', 'none') in ['both', modal]
        if res_type in ('moe', 'prefix_prompt') and peft_flag:
            raise NotImplementedError(
                f"method={res_type!r}: only the LoRA, adapter and vanilla blocks are built by "
                "lifelong_clip_b200 (scripts/lora_clip.sh, scripts/adapter_clip.sh)")
        if res_type == 'adapter' and peft_flag:
            from .adapter_modules import ResidualAttentionBlock_Adapter
            self.resblocks = nn.Sequential(*[
                ResidualAttentionBlock_Adapter(width, heads, attn_mask, design_details)
                for _ in range(layers)])
        elif res_type == 'lora' and peft_flag:
            self.resblocks = nn.Sequential(*[
                ResidualAttentionBlock_LoRA(width, heads, attn_mask, design_details)
                for _ in range(layers)])
        else:
            self.resblocks = nn.Sequential(*[
                ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: torch.Tensor):
        return self.resblocks(x)


class _TowerFn(torch.autograd.Function):
    """images -> feat [N, E] = ln_post(x[:, 0]) @ proj through llc_vit_forward + llc_head_fwd."""

    @staticmethod
    def forward(ctx, vit, images, *lora):
        need_grad = any(ctx.needs_input_grad)
        eng = vit.engine()
        eng.forward(images, training=need_grad)
        head = eng.features_only()
        ctx.vit, ctx.lora, ctx.need_grad = vit, lora, need_grad
        return head.feat.clone()

    @staticmethod
    def backward(ctx, d_feat):
        if not ctx.need_grad:
            raise RuntimeError("tower forward ran without grad")
        eng = ctx.vit.engine()
        eng.backward_from_feat(d_feat.detach().float().contiguous())
        return (None, None) + tuple(g.clone() if p.requires_grad else None
                                    for g, p in zip(eng.lora_grad_views, ctx.lora))


class VisualTransformer(nn.Module):
    """model.py:689-787. forward(x) keeps the reference signature; the prompt arguments of the
    proto-CLIP refactor (model.py:755,769-780) are accepted but must be unused on the LoRA path."""

    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int,
                 output_dim: int, modal=None, design_details: dict = {}):
        super().__init__()
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.patch_size, self.width, self.layers, self.heads = patch_size, width, layers, heads
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(
            scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads, modal=modal,
                                       design_details=design_details)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._engine = None
        return super()._load_from_state_dict(*a, **k)

    def invalidate_packed(self):
        self._engine = None
        for b in self.transformer.resblocks:
            b.invalidate_packed()

    def engine(self):
        from .engine import VitEngine
        if self.block_by_block:
            raise RuntimeError("adapter blocks run block by block (forward_tokens); the fused "
                               "tower engine only knows the LoRA / vanilla block")
        if self._engine is None:
            self._engine = VitEngine(self)
        return self._engine

    def lora_params(self):
        return tuple(p for b in self.transformer.resblocks for p in b.lora_params())

    def forward(self, x, prompt_module=None, register_blk=-1, q=None, patch_tokens=None,
                train=False, task_id=None):
        if prompt_module is not None:
            raise NotImplementedError("prompt_module belongs to the proto-CLIP method "
                                      "(out of scope; SURVEY.md §2)")
        if self.block_by_block:
            f32 = lambda t: t.detach().float().contiguous()
            t = self.forward_tokens(x)
            rows = torch.arange(t.shape[1], device=t.device, dtype=torch.int64)
            return _RowFeatFn.apply(t, rows, f32(self.ln_post.weight), f32(self.ln_post.bias),
                                    f32(self.proj), False)
        return _TowerFn.apply(self, x, *self.lora_params())

    @property
    def block_by_block(self) -> bool:
        """Adapter blocks carry their own trainable tensors and run one llc_adapter_block_* call
        per block under autograd; LoRA / vanilla towers run as ONE llc_vit_forward call."""
        from .adapter_modules import ResidualAttentionBlock_Adapter
        return isinstance(self.transformer.resblocks[0], ResidualAttentionBlock_Adapter)

    def forward_tokens(self, x: torch.Tensor) -> torch.Tensor:
        """model.py:757-781 block by block: images -> the transformer's output [L, N, D]."""
        if x.device.type != "cuda":
            raise RuntimeError("lifelong_clip_b200 modules compute on CUDA (sm_100a) only; move "
                               "the module and its input to the GPU (no CPU fallback)")
        t = embed_images(self, x).permute(1, 0, 2).contiguous()     # NLD -> LND
        return self.transformer(t)

    def get_patch_feature(self, x: torch.Tensor):
        """model.py:731-753: ln_post(CLS) without the projection, returned twice."""
        with torch.no_grad():
            if self.block_by_block:
                cls = self.forward_tokens(x)[0]
            else:
                eng = self.engine()
                eng.forward(x, training=False)
                cls = eng.cls_rows()
            y = F.layer_norm(cls, (self.width,), self.ln_post.weight.float(),
                             self.ln_post.bias.float(), self.ln_post.eps)
        return y, y


class _RowFeatFn(torch.autograd.Function):
    """x [L, N, D] -> features of one row per sample, LN(x[row_n]) @ proj, L2-normalised unless
    normalise=False: the tail of both encoders (model.py:782-785 / :951-956 + :966-969) through
    llc_head_fwd / llc_head_bwd."""

    @staticmethod
    def forward(ctx, x, rows, ln_g, ln_b, proj, normalise=True):
        L, N, D = x.shape
        x2 = x.detach().float().contiguous().view(L * N, D)
        dummy = torch.zeros(1, proj.shape[1], device=x.device)
        dummy.narrow(1, 0, 1).fill_(1.0)       # (a device-side fill: CUDA-graph capturable)
        head = ops.Head(x2, 1, ln_g, ln_b, proj, dummy, 1.0, N, row_idx=rows).forward()
        ctx.head, ctx.shape, ctx.normalise = head, (L, N, D), normalise
        return (head.fnorm if normalise else head.feat).clone()

    @staticmethod
    def backward(ctx, d_out):
        head, (L, N, D) = ctx.head, ctx.shape
        d = d_out.detach().float().contiguous()
        head.keep = head.keep + (d,)
        head.args.d_fnorm = d.data_ptr() if ctx.normalise else None
        head.args.d_feat = None if ctx.normalise else d.data_ptr()
        head.args.skip_logit_grad = 1
        dx = torch.zeros(L * N, D, device=d.device)
        head.backward(dx)
        return dx.view(L, N, D), None, None, None, None, None


class _CosineLogitFn(torch.autograd.Function):
    """x_img [L, N, D], t_hat [C, E] -> logits [N, C] = s * normalise(LN(x[0, n]) @ proj) @ t_hat^T
    (models/maple.py:244-251), gradients to x_img (class-token rows) and t_hat."""

    @staticmethod
    def forward(ctx, x, text, ln_g, ln_b, proj, scale):
        L, N, D = x.shape
        x2 = x.detach().float().contiguous().view(L * N, D)
        rows = torch.arange(N, device=x.device, dtype=torch.int64)     # token (0, n) = row n
        t = text.detach().float().contiguous()
        head = ops.Head(x2, 1, ln_g, ln_b, proj, t, float(scale), N, row_idx=rows,
                        want_dlogits=True).forward()
        ctx.head, ctx.shape, ctx.scale = head, (L, N, D), float(scale)
        return head.logits.clone()

    @staticmethod
    def backward(ctx, d_logits):
        head, (L, N, D) = ctx.head, ctx.shape
        d = d_logits.detach().float().contiguous()
        head.args.d_is_logits = 1
        head.args.skip_logit_grad = 0
        dx = torch.zeros(L * N, D, device=d.device)
        head.backward(dx, d)
        d_text = ops.head_dtext(head.dlogits, head.fnorm, ctx.scale)
        return dx.view(L, N, D), d_text, None, None, None, None

class _HeadProbsFn(torch.autograd.Function):
    """x [L, N, D] (class-token rows = rows 0..N-1 of the flattened [L*N] axis), normalised text
    features [C, E] -> (probs, normalised image features, pred, loss_sum): the cosine-logit head
    of models/adapter_clip.py:94-100 on a tower that ran block by block. With labels the loss
    (CE on the probabilities, methods/adapter_clip.py:89) and its gradient are formed by the head
    kernels; without, backward takes the caller's d_probs."""

    @staticmethod
    def forward(ctx, x, text, ln_g, ln_b, proj, scale, add_mask, labels, inv_batch,
                double_softmax):
        L, N, D = x.shape
        x2 = x.detach().float().contiguous().view(L * N, D)
        rows = torch.arange(N, device=x.device, dtype=torch.int64)
        t = text.detach().float().contiguous()
        head = ops.Head(x2, 1, ln_g, ln_b, proj, t, float(scale), N, row_idx=rows,
                        add_mask=add_mask, labels=labels, inv_batch=inv_batch,
                        double_softmax=double_softmax, want_dlogits=True).forward()
        ctx.head, ctx.shape, ctx.scale, ctx.fused = head, (L, N, D), float(scale), labels is not None
        loss = head.loss_rows.sum() if labels is not None else head.loss_rows.new_zeros(())
        # clones: an output that is also reachable from ctx (ctx.head) would close a reference
        # cycle output -> grad_fn -> ctx -> output that Python's GC cannot see through, and the
        # whole graph (every block's saved activations) would outlive the step
        probs, fnorm, pred = head.probs.clone(), head.fnorm.clone(), head.pred.clone()
        ctx.mark_non_differentiable(pred)
        return probs, fnorm, pred, loss

    @staticmethod
    def backward(ctx, d_probs, d_fnorm, _, d_loss):
        head, (L, N, D) = ctx.head, ctx.shape
        dev = head.probs.device
        dx = torch.zeros(L * N, D, device=dev)
        head.args.skip_logit_grad = 0
        if ctx.fused:
            head.backward(dx, None, float(d_loss) if d_loss is not None else 1.0)
        else:
            dp = d_probs.detach().float().contiguous() if d_probs is not None else \
                torch.zeros_like(head.probs)
            head.backward(dx, dp)
        d_text = None
        if ctx.needs_input_grad[1]:
            d_text = ops.head_dtext(head.dlogits, head.fnorm, ctx.scale)
        return (dx.view(L, N, D), d_text) + (None,) * 8


def embed_images(v: "VisualTransformer", image: torch.Tensor) -> torch.Tensor:
    """model.py:757-767: stride-P conv + class token + positional embedding + ln_pre -> fp32
    [N, L, D] (llc_patchify + tcgen05 GEMM + llc_embed_ln_pre; frozen, no gradient)."""
    N, P = image.shape[0], v.patch_size
    G, D = v.input_resolution // P, v.width
    dev = image.device
    kp = (3 * P * P + 15) // 16 * 16
    key = (v.conv1.weight.data_ptr(), v.conv1.weight._version)
    if getattr(v, "_wpatch_key", None) != key:
        v._wpatch = ops.pack_weight(
            v.conv1.weight.detach().float().reshape(D, 3 * P * P).contiguous(),
            torch.zeros(D, kp, dtype=torch.bfloat16, device=dev))
        v._wpatch_key = key
    patches = torch.empty(N * G * G, kp, dtype=torch.bfloat16, device=dev)
    ops.patchify(image.float().contiguous(), P, patches)
    po = torch.empty(N * G * G, D, device=dev)
    ops.gemm_tn(patches, v._wpatch, N * G * G, D, kp, po)
    x0 = torch.empty(N * (G * G + 1), D, device=dev)
    f32 = lambda t: t.detach().float().contiguous()
    ops.embed_ln_pre(po, f32(v.class_embedding), f32(v.positional_embedding),
                     f32(v.ln_pre.weight), f32(v.ln_pre.bias), N, G * G + 1, D, x0)
    return x0.view(N, G * G + 1, D)
