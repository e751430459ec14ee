"""The adapter-clip method's modules (`--method adapter-clip`, scripts/adapter_clip.sh) on libllc:

    Adapter                              models/clip/adapter.py:11-73
    ResidualAttentionBlock_Adapter       models/clip/model.py:418-442

Same constructor signatures, parameter names (`adaptmlp.down_proj.*`, `adaptmlp.up_proj.*`: the
`"adaptmlp" in name` freeze filter of methods/adapter_clip.py:115-119 keeps working) and init as
the reference. The block runs as ONE llc_adapter_block_forward / _backward pair on the reference's
[L, N, D] layout: frozen attention and MLP on the tcgen05 GEMM / attention kernels, the adapter's
two projections on the same GEMM, its weight gradients on a token-reduction tcgen05 kernel
(csrc/adapter.cu). Dropout (p = 0.1 on the bottleneck in training mode) draws from a counter-based
stream seeded from torch's generator - the reference's own draws come from torch's Philox stream
and are not reproducible outside torch; `Adapter.push_masks` injects explicit keep masks (tests).
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _capi as K
from . import ops
from .clip_modules import PAD, ResidualAttentionBlock, _bf16, _bf16e, _bf16p

DIM = K.ADAPTER_DIM


class Adapter(nn.Module):
    """models/clip/adapter.py:11-73 with adapter_layernorm_option='none' and a float scalar (what
    ResidualAttentionBlock_Adapter constructs, model.py:432-439)."""

    def __init__(self, d_model=None, bottleneck=None, dropout=0.0, init_option="lora",
                 adapter_scalar="1.0", adapter_layernorm_option="in"):
        super().__init__()
        if adapter_layernorm_option not in ("none", None):
            raise NotImplementedError("adapter_layernorm_option 'in'/'out' is not used by the "
                                      "reference's adapter block (model.py:438)")
        if adapter_scalar == "learnable_scalar":
            raise NotImplementedError("learnable adapter scalar is not used by the reference's "
                                      "adapter block (model.py:437)")
        if bottleneck != DIM:
            raise NotImplementedError(f"the reference hard-codes down_proj to {DIM} outputs "
                                      f"(adapter.py:39); bottleneck must be {DIM}")
        self.n_embd, self.down_size = d_model, bottleneck
        self.adapter_layernorm_option = adapter_layernorm_option
        self.adapter_layer_norm_before = None
        self.scale = float(adapter_scalar)
        self.down_proj = nn.Linear(self.n_embd, DIM)
        self.non_linear_func = nn.ReLU()
        self.up_proj = nn.Linear(self.down_size, self.n_embd)
        self.dropout = dropout
        if init_option == "bert":
            raise NotImplementedError
        elif init_option == "lora":
            with torch.no_grad():
                nn.init.kaiming_uniform_(self.down_proj.weight, a=math.sqrt(5))
                nn.init.zeros_(self.up_proj.weight)
                nn.init.zeros_(self.down_proj.bias)
                nn.init.zeros_(self.up_proj.bias)
        self._ops = None
        self._ops_key = None
        self._calls = 0
        self._masks = []
        # keep_bottleneck: retain the bf16 [T, 64] bottleneck(s) after ReLU and dropout of the
        # latest call (one per application) - which gates were open (diagnostics / tests)
        self.keep_bottleneck = False
        self.last_bottleneck = ()

    # -- operands ------------------------------------------------------------------------------
    def params(self):
        return (self.down_proj.weight, self.down_proj.bias, self.up_proj.weight,
                self.up_proj.bias)

    def push_masks(self, *masks):
        """Explicit dropout keep masks (uint8/bool [T, 64], token-major as the kernels see them)
        consumed by the next applications of this module in training mode, first in first out."""
        self._masks += [m.to(torch.uint8).contiguous() for m in masks]

    def pop_mask(self, T, device):
        if not self._masks:
            return None
        m = self._masks.pop(0).to(device)
        if m.numel() != T * DIM:
            raise RuntimeError(f"dropout mask has {m.numel()} elements, expected {T * DIM}")
        return m

    def next_seed(self) -> int:
        self._calls += 1
        return (torch.initial_seed() * 0x9E3779B1 + self._calls * 0x85EBCA77) & (2 ** 63 - 1)

    def struct(self, params, grads=None, seed=0, block=None) -> K.Adapter:
        """llc_adapter for the given live tensors; the bf16 operands are re-derived whenever a
        parameter changed (its version counter or storage). block: the owning
        ResidualAttentionBlock_Adapter - its transposed frozen weights then get the composed
        columns (W_d W_o)^T / (W_d W_proj)^T the block's backward folds into its big GEMMs."""
        D = self.n_embd
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device.type != "cuda":
                raise RuntimeError("adapter parameters must be contiguous fp32 CUDA tensors "
                                   "(lifelong_clip_b200 has no CPU fallback)")
        dev = params[0].device
        if self._ops is None or self._ops[0].device != dev:
            self._ops = (_bf16(DIM, D, dev), _bf16(D, DIM, dev), _bf16(D, DIM, dev),
                         _bf16(DIM, D, dev), torch.zeros(D, device=dev))
            self._ops_key = None
        s = K.Adapter()
        s.scale, s.dropout, s.seed = self.scale, float(self.dropout), seed
        for n, p in zip(("down_w", "down_b", "up_w", "up_b"), params):
            setattr(s, n, p.data_ptr())
        for n, t in zip(("wd", "wu", "wdT", "wuT", "bu_s"), self._ops):
            setattr(s, n, t.data_ptr())
        if grads is not None:
            for n, g in zip(("g_down_w", "g_down_b", "g_up_w", "g_up_b"), grads):
                setattr(s, n, g.data_ptr())
        packed = None
        if block is not None:
            packed = block.packed()
            M = block.mlp.c_fc.out_features
            if getattr(self, "_wprojT_src", None) is not packed:
                self._wprojT_ad = _bf16(M, D + DIM, dev)
                self._wprojT_ad[:, :D].copy_(packed.wprojT)
                self._wprojT_src = packed
            s.woT_ad = packed.woT_aug.data_ptr()        # [D, D + 64]: pad columns are free
            s.wprojT_ad = self._wprojT_ad.data_ptr()
            s.mlp_dim = M
        key = tuple((p.data_ptr(), p._version) for p in params) + (id(packed),)
        if key != self._ops_key:
            K.check(K.load().llc_adapter_refresh(C.byref(s), D, K.stream_ptr()),
                    "llc_adapter_refresh")
            self._ops_key = key
        return s

    def _apply(self, fn, *a, **k):
        self._ops = self._ops_key = None
        self._wprojT_src = None
        return super()._apply(fn, *a, **k)

    def forward(self, x, add_residual=True, residual=None):
        return _AdapterFn.apply(self, x, residual, bool(add_residual), *self.params())


class _AdapterFn(torch.autograd.Function):
    """Adapter.forward(x, add_residual, residual) (adapter.py:53-73) on [.., D] through
    llc_adapter_forward / llc_adapter_backward."""

    @staticmethod
    def forward(ctx, ad, x, residual, add_residual, *params):
        if x.device.type != "cuda":
            raise RuntimeError("lifelong_clip_b200 modules compute on CUDA (sm_100a) only; move "
                               "the module and its input to the GPU (no CPU fallback)")
        D, dev = ad.n_embd, x.device
        x2 = x.detach().float().contiguous().view(-1, D)
        T = x2.shape[0]
        res = None
        if add_residual:
            res = x2 if residual is None else residual.detach().float().contiguous().view(-1, D)
        yb = _bf16(T, D, dev)
        ops.cast_bf16(x2, yb)
        a = _bf16(T, DIM, dev)
        out = torch.empty(T, D, device=dev)
        training = ad.training and ad.dropout > 0
        mask = ad.pop_mask(T, dev) if training else None
        s = ad.struct(params, seed=ad.next_seed())
        K.check(K.load().llc_adapter_forward(C.byref(s), yb.data_ptr(), D, K.ptr(res), 0,
                                             a.data_ptr(), K.ptr(mask), 0, int(training),
                                             out.data_ptr(), T, D, K.stream_ptr()),
                "llc_adapter_forward")
        ctx.ad, ctx.yb, ctx.a, ctx.params, ctx.shape = ad, yb, a, params, x.shape
        ad.last_bottleneck = (a,) if ad.keep_bottleneck else ()
        ctx.training = training
        ctx.res_is_x = add_residual and residual is None
        ctx.res_given = add_residual and residual is not None
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        ad, D = ctx.ad, ctx.ad.n_embd
        dev = dy.device
        dx = dy.detach().float().contiguous().view(-1, D)
        T = dx.shape[0]
        dxb = _bf16(T, D, dev)
        ops.cast_bf16(dx, dxb)
        grads = [torch.zeros_like(p, dtype=torch.float32) for p in ctx.params]
        need_x = ctx.needs_input_grad[1]
        d_y = torch.empty(T, D, device=dev) if need_x else None
        da = _bf16(T, DIM, dev)
        partial = torch.empty(K.load().llc_adapter_partial_floats(D), device=dev)
        s = ad.struct(ctx.params, grads)
        K.check(K.load().llc_adapter_backward(
            C.byref(s), ctx.yb.data_ptr(), D, ctx.a.data_ptr(), dx.data_ptr(), dxb.data_ptr(), D,
            K.ptr(d_y), int(ctx.res_is_x), da.data_ptr(), partial.data_ptr(), 0,
            int(ctx.training), T, D, K.stream_ptr()), "llc_adapter_backward")
        gx = d_y.view(ctx.shape) if need_x else None
        gres = dx.view(ctx.shape) if (ctx.res_given and ctx.needs_input_grad[2]) else None
        return (None, gx, gres, None) + tuple(
            g if p.requires_grad else None for g, p in zip(grads, ctx.params))


class _AdapterBlockFn(torch.autograd.Function):
    """ResidualAttentionBlock_Adapter.forward on [L, N, D] through llc_adapter_block_forward /
    llc_adapter_block_backward."""

    @staticmethod
    def forward(ctx, blk, x, *params):
        L, N, D = x.shape
        T, M, dev = L * N, blk.mlp.c_fc.out_features, x.device
        if dev.type != "cuda":
            raise RuntimeError("lifelong_clip_b200 modules compute on CUDA (sm_100a) only; move "
                               "the module and its input to the GPU (no CPU fallback)")
        ad = blk.adaptmlp
        x2 = x.detach().float().contiguous().view(T, D)
        need_grad = any(ctx.needs_input_grad)
        training = blk.training and ad.dropout > 0
        bufs = dict(
            x_in=x2, h1=_bf16p(T, D, dev), qkv=_bf16p(T, 3 * D, dev),
            lse=torch.empty(N * blk.n_head * L, device=dev), o=_bf16p(T, D, dev),
            x_mid=torch.empty(T, D, device=dev), h2=_bf16e(T, D, dev),
            z=_bf16e(T, M, dev) if need_grad else None, g=_bf16e(T, M, dev),
            x_out=torch.empty(T, D, device=dev))
        abufs = dict(ya=_bf16e(T, D, dev), a1=_bf16e(T, DIM, dev), m=_bf16e(T, D, dev),
                     a2=_bf16e(T, DIM, dev),
                     mask1=ad.pop_mask(T, dev) if training else None,
                     mask2=ad.pop_mask(T, dev) if training else None)
        b = K.BlockBufs()
        for k, v in bufs.items():
            setattr(b, k, K.ptr(v))
        ab = K.AdapterBufs()
        for k, v in abufs.items():
            setattr(ab, k, K.ptr(v))
        lora = blk.lora_params()
        # (no LoRA refresh: the factors are zero buffers, and the pad columns of the packed
        # transposed out-projection belong to the adapter's composed columns)
        layer = blk._layer_struct(lora, [None] * 4)
        causal = blk._causal_flag(L)
        seed = ad.next_seed()
        s = ad.struct(params, seed=seed, block=blk)
        K.check(K.load().llc_adapter_block_forward(
            C.byref(blk._cfg), C.byref(layer), C.byref(s), C.byref(b), C.byref(ab), N, L, 1, N,
            causal, int(training), K.stream_ptr()), "llc_adapter_block_forward")
        out = bufs.pop("x_out")      # not needed by the backward; h2 / g are recomputed nowhere
        bufs["h2"] = bufs["g"] = None   # ...and not read by it either: free them with the call
        ctx.blk, ctx.bufs, ctx.abufs, ctx.shape, ctx.causal = blk, bufs, abufs, (L, N, D), causal
        ad.last_bottleneck = (abufs["a1"], abufs["a2"]) if ad.keep_bottleneck else ()
        ctx.x_needs_grad, ctx.params, ctx.training, ctx.seed = x.requires_grad, params, training, seed
        return out.view(L, N, D)

    @staticmethod
    def backward(ctx, dy):
        blk, bufs, abufs, (L, N, D) = ctx.blk, ctx.bufs, ctx.abufs, ctx.shape
        T, M, dev = L * N, blk.mlp.c_fc.out_features, dy.device
        if bufs["z"] is None:
            raise RuntimeError("block forward ran without grad; cannot backpropagate")
        ad = blk.adaptmlp
        # dy belongs to autograd and is only read (llc_block_bwd_bufs.dy); dx receives the
        # gradient of the block's input
        dyc = dy.detach().float().contiguous().view(T, D)
        dx = torch.empty(T, D, device=dev)
        grads = [torch.zeros_like(p, dtype=torch.float32) for p in ctx.params]
        lora = blk.lora_params()
        lgrads = [None] * 4       # frozen attention: no LoRA reductions (llc.h: llc_vit_layer)
        lib = K.load()
        scratch = dict(
            dx=dx, dxb=_bf16p(T, D, dev), dz=_bf16e(T, M, dev), dh=_bf16e(T, D, dev),
            d_o=_bf16e(T, D, dev), dqkv=_bf16p(T, 3 * D, dev),
            partial=torch.empty(ops.lora_side_max_partials() * 3 * D * 8, device=dev),
            delta=torch.empty(N * blk.n_head * L, device=dev))
        s = K.BlockBwdBufs()
        for k, v in scratch.items():
            setattr(s, k, v.data_ptr())
        s.dy = dyc.data_ptr()
        extra = dict(da=_bf16e(T, DIM, dev), d_branch=None,
                     partial=torch.empty(lib.llc_adapter_partial_floats(D), device=dev))
        ab = K.AdapterBufs()
        for k, v in {**abufs, **extra}.items():
            setattr(ab, k, K.ptr(v))
        K.check(lib.llc_cast_bf16(dyc.data_ptr(), scratch["dxb"].data_ptr(), T, D, D + PAD,
                                  K.stream_ptr()), "llc_cast_bf16")
        b = K.BlockBufs()
        for k, v in bufs.items():
            setattr(b, k, K.ptr(v))
        layer = blk._layer_struct(lora, lgrads)
        a = ad.struct(ctx.params, grads, seed=ctx.seed, block=blk)
        K.check(lib.llc_adapter_block_backward(
            C.byref(blk._cfg), C.byref(layer), C.byref(a), C.byref(b), C.byref(ab), C.byref(s), N,
            L, 1, N, ctx.causal, int(ctx.x_needs_grad), int(ctx.training), K.stream_ptr()),
            "llc_adapter_block_backward")
        gx = dx.view(L, N, D) if ctx.x_needs_grad else None
        return (None, gx) + tuple(g if p.requires_grad else None
                                  for g, p in zip(grads, ctx.params))


class ResidualAttentionBlock_Adapter(ResidualAttentionBlock):
    """model.py:418-442: x = x + adaptmlp(attention(ln_1(x))); x = x + adaptmlp(mlp(ln_2(x)))
    with ONE shared Adapter(d_model, dropout=0.1, bottleneck=ffn_num, init 'lora', scalar 0.1,
    no adapter LayerNorm). Backbone frozen; gradients reach the four adapter tensors and x."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None,
                 design_details: dict = {}):
        super().__init__(d_model, n_head, attn_mask)
        self.ffn_num = design_details.get('ffn_num', 64)
        self.adaptmlp = Adapter(d_model=d_model, dropout=0.1, bottleneck=self.ffn_num,
                                init_option='lora', adapter_scalar=0.1,
                                adapter_layernorm_option='none')

    def forward(self, x: torch.Tensor):
        if x.dim() != 3 or x.shape[-1] != self.d_model:
            raise RuntimeError(f"expected [L, N, {self.d_model}], got {tuple(x.shape)}")
        return _AdapterBlockFn.apply(self, x, *self.adaptmlp.params())
