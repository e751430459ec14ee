"""VitEngine: host-side state of one image tower on one GPU.

Owns (as torch tensors, i.e. PyTorch's allocator): the prepared bf16 weights of every block, ONE
flat fp32 buffer holding all LoRA factors (the nn.Parameters become views of it) and one flat
gradient buffer of the same layout (what the data-parallel all-reduce and the fused AdamW operate
on), and the activation arena sized by llc_vit_arena_bytes. All compute is llc_* calls.
"""
from __future__ import annotations

import os

import ctypes as C

import torch

from . import _capi as K
from . import ops

PAD = K.LORA_LD   # row-pitch pad of augmented buffers (the K extension itself is LORA_PAD)


class VitEngine:
    def __init__(self, vit):
        dev = vit.proj.device
        if dev.type != "cuda":
            raise RuntimeError("lifelong_clip_b200 computes on CUDA (sm_100a) only: move the model "
                               "to the GPU before calling it (there is no CPU fallback)")
        ops.check_device(dev.index if dev.index is not None else torch.cuda.current_device())
        self.vit, self.device = vit, dev
        blocks = list(vit.transformer.resblocks)
        b0 = blocks[0]
        self.L = (vit.input_resolution // vit.patch_size) ** 2 + 1
        self.D, self.E = vit.width, vit.output_dim
        c = K.VitCfg()
        c.image_size, c.patch, c.width, c.layers = (vit.input_resolution, vit.patch_size,
                                                    vit.width, len(blocks))
        c.heads, c.mlp_dim, c.embed_dim = vit.heads, b0.mlp.c_fc.out_features, vit.output_dim
        c.lora_r, c.lora_scale = b0.lora_r, b0.lora_alpha / b0.lora_r
        self.cfg = c

        # ---- flat LoRA parameter / gradient buffers; parameters become views -------------------
        self.lora_params = list(vit.lora_params())
        n = sum(p.numel() for p in self.lora_params)
        self.lora_flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.grad_flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.lora_grad_views, off = [], 0
        for p in self.lora_params:
            view = self.lora_flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.detach().float())
            p.data = view
            self.lora_grad_views.append(self.grad_flat[off:off + p.numel()].view(p.shape))
            off += p.numel()

        # ---- prepared weights ------------------------------------------------------------------
        self.layers = (K.VitLayer * len(blocks))()
        for i, blk in enumerate(blocks):
            blk.packed().fill(self.layers[i], blk.lora_params(),
                              self.lora_grad_views[4 * i:4 * i + 4])
        P = vit.patch_size
        kp = 3 * P * P
        self.wpatch = ops.pack_weight(
            vit.conv1.weight.detach().float().reshape(vit.width, kp).contiguous(),
            torch.zeros(vit.width, (kp + 15) // 16 * 16, dtype=torch.bfloat16, device=dev))
        f32 = lambda t: t.detach().float().contiguous()
        self.class_emb, self.pos = f32(vit.class_embedding), f32(vit.positional_embedding)
        self.ln_pre = (f32(vit.ln_pre.weight), f32(vit.ln_pre.bias))
        self.ln_post = (f32(vit.ln_post.weight), f32(vit.ln_post.bias))
        self.proj = f32(vit.proj)
        w = K.VitWeights()
        w.wpatch = self.wpatch.data_ptr()
        w.class_emb, w.pos_emb = self.class_emb.data_ptr(), self.pos.data_ptr()
        w.ln_pre_g, w.ln_pre_b = self.ln_pre[0].data_ptr(), self.ln_pre[1].data_ptr()
        w.layers = self.layers
        self.weights = w

        self.arena = None
        self.arena_key = None
        self.dx = None
        self.x_final = None
        self.N = 0
        self._lora_version = None
        self._dummy_text = torch.zeros(1, self.E, device=dev)
        self._dummy_text[0, 0] = 1.0

    # ------------------------------------------------------------------------------------------
    def mark_lora_updated(self):
        """Call after writing lora_flat outside torch's version tracking (llc_adamw)."""
        self._lora_version = None

    def _refresh_lora(self, force: bool = False):
        ver = sum(p._version for p in self.lora_params)
        if force or ver != self._lora_version:
            K.check(K.load().llc_vit_refresh_lora(C.byref(self.cfg), C.byref(self.weights),
                                                  K.stream_ptr()), "llc_vit_refresh_lora")
            self._lora_version = ver

    def _ensure_arena(self, N: int, training: bool):
        key = (N, bool(training))
        if self.arena_key == key:
            return
        if self.arena_key is not None and self.arena_key[0] == N and self.arena_key[1]:
            return  # a training arena also serves inference of the same batch size
        nbytes = K.load().llc_vit_arena_bytes(C.byref(self.cfg), N, int(training))
        if nbytes == 0:
            K.check(-1, "llc_vit_arena_bytes")
        self.arena = None
        self.arena = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.dx = torch.empty(N * self.L, self.D, device=self.device) if training else None
        self.arena_key = key

    def forward(self, images: torch.Tensor, training: bool, force_refresh: bool = False):
        if images.device != self.device or images.dtype != torch.float32:
            images = images.to(device=self.device, dtype=torch.float32)
        images = images.contiguous()
        N = images.shape[0]
        if tuple(images.shape[1:]) != (3, self.cfg.image_size, self.cfg.image_size):
            raise RuntimeError(f"expected [N, 3, {self.cfg.image_size}, {self.cfg.image_size}] "
                               f"images, got {tuple(images.shape)}")
        self._ensure_arena(N, training)
        self._refresh_lora(force_refresh)   # force: CUDA-graph capture must always contain it
        mode = int(self.arena_key[1])
        xf = C.c_void_p()
        # every consumer of this engine reads only the class-token rows (ln_post(x[:, 0]) @ proj):
        # the last block runs class-token-only unless LLC_FULL_LAST_BLOCK is set (A/B, tests)
        self._cls_only = os.environ.get("LLC_FULL_LAST_BLOCK") is None
        fwd = K.load().llc_vit_forward_cls if self._cls_only else K.load().llc_vit_forward
        K.check(fwd(C.byref(self.cfg), C.byref(self.weights), images.data_ptr(), N,
                    self.arena.data_ptr(), mode, C.byref(xf), K.stream_ptr()), "llc_vit_forward")
        off = xf.value - self.arena.data_ptr()
        T = N * self.L
        self.x_final = self.arena[off:off + T * self.D * 4].view(torch.float32).view(T, self.D)
        self.N = N
        self._trained_arena = bool(mode)
        return self.x_final

    def cls_rows(self) -> torch.Tensor:
        return self.x_final.view(self.N, self.L, self.D)[:, 0, :]

    def head(self, text, logit_scale_exp, *, cls_idx=None, add_mask=None, labels=None,
             double_softmax=True, inv_batch=None) -> ops.Head:
        return ops.Head(self.x_final, self.L, self.ln_post[0], self.ln_post[1], self.proj, text,
                        logit_scale_exp, self.N, cls_idx=cls_idx, add_mask=add_mask, labels=labels,
                        double_softmax=double_softmax, inv_batch=inv_batch).forward()

    def features_only(self) -> ops.Head:
        self._feat_head = self.head(self._dummy_text, 1.0)
        return self._feat_head

    # ------------------------------------------------------------------------------------------
    def _vit_backward(self):
        bwd = K.load().llc_vit_backward_cls if self._cls_only else K.load().llc_vit_backward
        K.check(bwd(C.byref(self.cfg), C.byref(self.weights), self.N, self.arena.data_ptr(),
                    self.dx.data_ptr(), K.stream_ptr()), "llc_vit_backward")

    def _clear_dx(self):
        # the full backward consumes dx for every token (zero except the CLS rows the head
        # writes); the class-token-only backward reads just those rows
        if not self._cls_only:
            self.dx.zero_()

    def _check_trainable(self):
        if self.dx is None or not getattr(self, "_trained_arena", False):
            raise RuntimeError("backward needs a forward(training=True) first")

    def backward_from_feat(self, d_feat: torch.Tensor):
        """d_feat [N, E] (gradient of ln_post(CLS) @ proj) -> LoRA grads in grad_flat."""
        self._check_trainable()
        h = self._feat_head
        h.keep = h.keep + (d_feat,)
        h.args.d_feat = d_feat.data_ptr()
        h.args.skip_logit_grad = 1
        self._clear_dx()
        h.backward(self.dx)
        self._vit_backward()

    def backward_from_head(self, head: ops.Head, d_probs=None, d_feat=None, loss_scale=1.0):
        """Gradient of the fused loss (d_probs=None: analytic, needs labels) or of an external
        d_probs [N, C] -> LoRA grads in grad_flat."""
        self._check_trainable()
        if d_feat is not None:
            head.keep = head.keep + (d_feat,)
            head.args.d_feat = d_feat.data_ptr()
        head.args.skip_logit_grad = 0
        self._clear_dx()
        head.backward(self.dx, d_probs, loss_scale)
        self._vit_backward()

    def bind_grads(self):
        """Point every LoRA Parameter's .grad at its slice of grad_flat (fused trainer path)."""
        for p, g in zip(self.lora_params, self.lora_grad_views):
            p.grad = g


class FlatAdamW:
    """torch.optim.AdamW semantics (utils/train_utils.py:27-28: lr, weight_decay=1e-5) on the
    engine's flat LoRA buffer: one llc_adamw launch per step; `grad_scale` folds the 1/world of
    the data-parallel mean or a GradScaler unscale into the same pass."""

    def __init__(self, engine: VitEngine, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=1e-5):
        self.engine, self.lr, self.betas, self.eps, self.wd = engine, lr, betas, eps, weight_decay
        self.m = torch.zeros_like(engine.lora_flat)
        self.v = torch.zeros_like(engine.lora_flat)
        self.t = 0

    def step(self, grad_scale: float = 1.0):
        self.t += 1
        e = self.engine
        ops.adamw(e.lora_flat, e.grad_flat, self.m, self.v, self.lr, self.betas[0], self.betas[1],
                  self.eps, self.wd, self.t, grad_scale)
        e.mark_lora_updated()

    def zero_grad(self):
        pass  # every backward overwrites grad_flat completely
