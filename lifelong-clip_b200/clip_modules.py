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


class Linear(nn.Linear):
    """lora.py:100-139 LoRA dense layer: parameter container for out_proj (weight, bias, lora_A
    kaiming-uniform(a=sqrt 5), lora_B zeros, scaling = alpha / r)."""

    def __init__(self, in_features, out_features, r=0, lora_alpha=1, lora_dropout=0.,
                 fan_in_fan_out=False, merge_weights=True, **kwargs):
        super().__init__(in_features, out_features, **kwargs)
        if lora_dropout != 0.:
            raise NotImplementedError("LoRA dropout is 0 on the reference path (lora.py:376)")
        self.r, self.lora_alpha, self.merged, self.merge_weights = r, lora_alpha, False, merge_weights
        self.fan_in_fan_out = fan_in_fan_out
        if r > 0:
            self.lora_A = nn.Parameter(self.weight.new_zeros((r, in_features)))
            self.lora_B = nn.Parameter(self.weight.new_zeros((out_features, r)))
            self.scaling = self.lora_alpha / self.r
            self.weight.requires_grad = False
            nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_B)


class MultiheadAttention(nn.Module):
    """lora.py:371-452: in_proj_weight/bias + one rank-r (A [r,D], B [3D,r]) pair shared by q,k,v
    (both xavier-uniform), out_proj = LoRA Linear. Parameter container; the math
    (lora.py:825-840,950,1002-1074) runs inside llc_block_forward."""

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


LORA_NAMES = ("attn.in_proj_weight_lora_A", "attn.in_proj_weight_lora_B",
              "attn.out_proj.lora_A", "attn.out_proj.lora_B")


def _bf16(rows, cols, device):
    return torch.zeros(rows, cols, dtype=torch.bfloat16, device=device)


class PackedLayer:
    """Prepared operands of one block: frozen weights as bf16 K-major matrices (forward and
    transposed for the activation-gradient GEMMs), 16 spare K columns for the LoRA factors."""

    def __init__(self, blk: "ResidualAttentionBlock_LoRA"):
        a = blk.attn
        D, M, dev = blk.d_model, blk.mlp.c_fc.out_features, a.in_proj_weight.device
        if dev.type != "cuda":
            raise RuntimeError("lifelong_clip_b200 modules compute on CUDA (sm_100a) only; "
                               "move the module to the GPU first (no CPU fallback)")
        f32 = lambda p: p.detach().float().contiguous()
        self.wqkv_aug = ops.pack_weight(f32(a.in_proj_weight), _bf16(3 * D, D + PAD, dev))
        self.wo_aug = ops.pack_weight(f32(a.out_proj.weight), _bf16(D, D + PAD, dev))
        self.wfc = ops.pack_weight(f32(blk.mlp.c_fc.weight), _bf16(M, D, dev))
        self.wproj = ops.pack_weight(f32(blk.mlp.c_proj.weight), _bf16(D, M, dev))
        self.wqkvT_aug = ops.pack_weight(f32(a.in_proj_weight), _bf16(D, 3 * D + PAD, dev), True)
        self.woT_aug = ops.pack_weight(f32(a.out_proj.weight), _bf16(D, D + PAD, dev), True)
        self.wfcT = ops.pack_weight(f32(blk.mlp.c_fc.weight), _bf16(D, M, dev), True)
        self.wprojT = ops.pack_weight(f32(blk.mlp.c_proj.weight), _bf16(M, D, dev), True)
        self.f_out_A = _bf16(16, D, dev)       # refreshed from the live LoRA factors every step
        self.f_in_B = _bf16(16, 3 * D, dev)
        self.f_out_B = _bf16(16, D, dev)
        self.small = [f32(p) for p in (a.in_proj_bias, a.out_proj.bias, blk.mlp.c_fc.bias,
                                       blk.mlp.c_proj.bias, blk.ln_1.weight, blk.ln_1.bias,
                                       blk.ln_2.weight, blk.ln_2.bias)]

    def fill(self, s: K.VitLayer, lora, grads):
        for n in ("wqkv_aug", "wo_aug", "wfc", "wproj", "wqkvT_aug", "woT_aug", "wfcT", "wprojT",
                  "f_out_A", "f_in_B", "f_out_B"):
            setattr(s, n, getattr(self, n).data_ptr())
        for n, t in zip(("bqkv", "bo", "bfc", "bproj", "ln1_g", "ln1_b", "ln2_g", "ln2_b"),
                        self.small):
            setattr(s, n, t.data_ptr())
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
            x_in=x2, h1=_bf16(T, D + PAD, dev), qkv=_bf16(T, 3 * D + PAD, dev),
            lse=torch.empty(N * blk.n_head * L, device=dev), o=_bf16(T, D + PAD, dev),
            x_mid=torch.empty(T, D, device=dev), h2=_bf16(T, D, dev),
            z=_bf16(T, M, dev) if need_grad else None, g=_bf16(T, M, dev),
            x_out=torch.empty(T, D, device=dev))
        b = K.BlockBufs()
        for k, v in bufs.items():
            setattr(b, k, K.ptr(v))
        layer = blk._layer_struct(lora, [None] * 4)
        blk._refresh_lora(layer)
        causal = blk._causal_flag(L)
        K.check(K.load().llc_block_forward(C.byref(blk._cfg), C.byref(layer), C.byref(b), N, L,
                                           1, N, causal, K.stream_ptr()), "llc_block_forward")
        ctx.blk, ctx.bufs, ctx.shape, ctx.causal = blk, bufs, (L, N, D), causal
        ctx.x_needs_grad = x.requires_grad
        ctx.lora = lora
        return bufs["x_out"].view(L, N, D)

    @staticmethod
    def backward(ctx, dy):
        blk, bufs, (L, N, D) = ctx.blk, ctx.bufs, ctx.shape
        T, M, dev = L * N, blk.mlp.c_fc.out_features, dy.device
        if bufs["z"] is None:
            raise RuntimeError("block forward ran without grad; cannot backpropagate")
        dx = dy.detach().float().contiguous().view(T, D).clone()
        grads = [torch.zeros_like(p, dtype=torch.float32) for p in ctx.lora]
        s = K.BlockBwdBufs()
        scratch = dict(
            dx=dx, dxb=_bf16(T, D + PAD, dev), dz=_bf16(T, M, dev), dh=_bf16(T, D, dev),
            d_o=_bf16(T, D, dev), dqkv=_bf16(T, 3 * D + PAD, dev),
            partial=torch.empty(ops.lora_side_max_partials() * 3 * D * 8, device=dev),
            delta=torch.empty(N * blk.n_head * L, device=dev))
        for k, v in scratch.items():
            setattr(s, k, v.data_ptr())
        lib = K.load()
        K.check(lib.llc_cast_bf16(dx.data_ptr(), scratch["dxb"].data_ptr(), T, D, D + PAD,
                                  K.stream_ptr()), "llc_cast_bf16")
        b = K.BlockBufs()
        for k, v in bufs.items():
            setattr(b, k, K.ptr(v))
        layer = blk._layer_struct(ctx.lora, grads)
        K.check(lib.llc_block_backward(C.byref(blk._cfg), C.byref(layer), C.byref(b), C.byref(s),
                                       N, L, 1, N, ctx.causal, int(ctx.x_needs_grad),
                                       K.stream_ptr()), "llc_block_backward")
        gx = dx.view(L, N, D) if ctx.x_needs_grad else None
        return (None, gx) + tuple(g.to(p.dtype) if p.requires_grad else None
                                  for g, p in zip(grads, ctx.lora))


class ResidualAttentionBlock_LoRA(nn.Module):
    """model.py:400-415 (on top of :209-236): x + attn(ln_1(x)); x + mlp(ln_2(x)) on [L, N, D]."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None,
                 design_details: dict = {}):
        super().__init__()
        self.d_model, self.n_head = d_model, n_head
        self.lora_alpha = design_details.get('lora_alpha', 1)
        self.lora_r = design_details.get('lora_r', 4)
        self.attn = MultiheadAttention(d_model, n_head, lora_alpha=self.lora_alpha, r=self.lora_r)
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

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def packed(self) -> PackedLayer:
        if self._packed is None:
            self._packed = PackedLayer(self)
        return self._packed

    def lora_params(self):
        a = self.attn
        return (a.in_proj_weight_lora_A, a.in_proj_weight_lora_B, a.out_proj.lora_A,
                a.out_proj.lora_B)

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
        m = self.attn_mask
        if m is None:
            return 0
        want = torch.full((L, L), float("-inf")).triu(1)
        if tuple(m.shape) == (L, L) and torch.equal(m.detach().float().cpu(), want):
            return 1  # the text tower's mask (model.py:926-932)
        raise NotImplementedError("only attn_mask=None or the causal mask is supported")

    def attention(self, x: torch.Tensor):
        raise NotImplementedError("attention() is fused into forward(); call the block")

    def forward(self, x: torch.Tensor):
        if x.dim() != 3 or x.shape[-1] != self.d_model:
            raise RuntimeError(f"expected [L, N, {self.d_model}], got {tuple(x.shape)}")
        return _BlockFn.apply(self, x, *self.lora_params())


class Transformer(nn.Module):
    """model.py:639-686. Only the 'lora' (both/this modality) flavour holds trainable blocks on
    this path; other methods of the reference are out of scope."""

    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor = None,
                 design_details: dict = {}, modal='text'):
        super().__init__()
        self.width, self.layers = width, layers
        res_type = design_details.get('method', 'vanilla')
        peft_flag = design_details.get('peft_encoder', 'none') in ['both', modal]
        if not (res_type == 'lora' and peft_flag):
            raise NotImplementedError(
                f"method={res_type!r} on modal={modal!r}: only LoRA blocks are built by "
                "lifelong_clip_b200 (the image tower of scripts/lora_clip.sh)")
        self.resblocks = nn.Sequential(*[
            ResidualAttentionBlock_LoRA(width, heads, attn_mask, design_details)
            for _ in range(layers)])

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
        return _TowerFn.apply(self, x, *self.lora_params())

    def get_patch_feature(self, x: torch.Tensor):
        """model.py:731-753: ln_post(CLS) without the projection, returned twice."""
        eng = self.engine()
        with torch.no_grad():
            eng.forward(x, training=False)
            y = F.layer_norm(eng.cls_rows(), (self.width,), self.ln_post.weight.float(),
                             self.ln_post.bias.float(), self.ln_post.eps)
        return y, y
