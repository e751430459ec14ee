"""Host-side state of one transformer tower on one GPU (VitEngine: image tower, TextEngine: text
tower with LoRA, peft_encoder='both').

Owns (as torch tensors, i.e. PyTorch's allocator): the prepared bf16 weights of every block, ONE
flat fp32 buffer holding all LoRA factors (the nn.Parameters become views of it) and one flat
gradient buffer of the same layout (what the data-parallel all-reduce and the fused AdamW operate
on), and the activation arenas sized by llc_vit_arena_bytes / llc_text_arena_bytes: one for
training (saved activations of every layer + backward scratch) and a separate one for evaluation
batches of another size, so that an evaluation pass never frees memory a captured training graph
still points at. All compute is llc_* calls.
"""
from __future__ import annotations

import os

import ctypes as C

import torch

from . import _capi as K
from . import ops

PAD = K.LORA_LD   # row-pitch pad of augmented buffers (the K extension itself is LORA_PAD)


class _TowerEngine:
    """Flat LoRA buffers, per-layer weight structs and arenas shared by both towers."""

    def _init_lora(self, blocks, dev):
        self.lora_params = [p for b in blocks for p in b.lora_params()]
        n = sum(p.numel() for p in self.lora_params)
        self.lora_flat = torch.empty(n, device=dev, dtype=torch.float32)
        # gradient buffer with two trailing floats (loss_sum, n_correct of the step): the
        # data-parallel exchange is then ONE all-reduce per tower instead of two
        self.grad_store = torch.zeros(n + 2, device=dev, dtype=torch.float32)
        self.grad_flat = self.grad_store[:n]
        self.scal = self.grad_store[n:]
        self.lora_grad_views, off = [], 0
        for p in self.lora_params:
            view = self.lora_flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.detach().float())
            p.data = view
            self.lora_grad_views.append(self.grad_flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
        self.layers = (K.VitLayer * len(blocks))()
        for i, blk in enumerate(blocks):
            blk.packed().fill(self.layers[i], blk.lora_params(),
                              self.lora_grad_views[4 * i:4 * i + 4])
        self.arena = None          # training arena (or the only one)
        self.arena_key = None
        self.eval_arena = None
        self.eval_key = None
        self.dx = None
        self.x_final = None
        self.N = 0
        self._lora_version = None
        self._refresh_w = K.VitWeights()
        self._refresh_w.layers = self.layers

    # ------------------------------------------------------------------------------------------
    def mark_lora_updated(self):
        """Call after writing lora_flat outside torch's version tracking (llc_adamw)."""
        self._lora_version = None

    def _refresh_lora(self, force: bool = False):
        ver = sum(p._version for p in self.lora_params)
        if force or ver != self._lora_version:
            K.check(K.load().llc_vit_refresh_lora(C.byref(self.cfg), C.byref(self._refresh_w),
                                                  K.stream_ptr()), "llc_vit_refresh_lora")
            self._lora_version = ver

    def _arena_bytes(self, N: int, training: bool) -> int:
        raise NotImplementedError

    def _pick_arena(self, N: int, training: bool):
        """Returns (arena tensor, mode flag passed to the C call). Training batches own
        self.arena; an evaluation batch reuses it when the size matches (its layout then is the
        training one) and otherwise lives in self.eval_arena."""
        if training:
            if self.arena_key != N:
                nbytes = self._arena_bytes(N, True)
                if nbytes == 0:
                    K.check(-1, "arena_bytes")
                self.arena = None
                self.dx = None
                self.arena = self._new_arena(nbytes)
                self.dx = torch.empty(N * self.L, self.D, device=self.device)
                self.arena_key = N
            return self.arena, 1
        if self.arena_key == N:
            return self.arena, 1
        if self.eval_key != N:
            nbytes = self._arena_bytes(N, False)
            if nbytes == 0:
                K.check(-1, "arena_bytes")
            self.eval_arena = None
            self.eval_arena = self._new_arena(nbytes)
            self.eval_key = N
        return self.eval_arena, 0

    def _new_arena(self, nbytes):
        a = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        a[:8192].zero_()     # flag words of the stream-K GEMM workspace (include/llc.h)
        return a

    def graph_signature(self):
        """What a captured CUDA graph of this engine's step depends on besides its inputs."""
        return (id(self), 0 if self.arena is None else self.arena.data_ptr(),
                0 if self.dx is None else self.dx.data_ptr())

    def release_eval_arena(self):
        self.eval_arena, self.eval_key = None, None

    def _set_x_final(self, arena, xf, N, mode):
        off = xf.value - arena.data_ptr()
        T = N * self.L
        self.x_final = arena[off:off + T * self.D * 4].view(torch.float32).view(T, self.D)
        self.N = N
        self._trained_arena = bool(mode)
        self._fwd_arena = arena
        return self.x_final

    def _check_trainable(self):
        if self.dx is None or not getattr(self, "_trained_arena", False) or \
                self._fwd_arena is not self.arena:
            raise RuntimeError("backward needs a forward(training=True) first")

    def bind_grads(self):
        """Point every LoRA Parameter's .grad at its slice of grad_flat (fused trainer path)."""
        for p, g in zip(self.lora_params, self.lora_grad_views):
            p.grad = g


class VitEngine(_TowerEngine):
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
        self._init_lora(blocks, dev)

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
        self._dummy_text = torch.zeros(1, self.E, device=dev)
        self._dummy_text[0, 0] = 1.0
        self._eval_head = None

    def _arena_bytes(self, N, training):
        return K.load().llc_vit_arena_bytes(C.byref(self.cfg), N, int(training))

    def forward(self, images: torch.Tensor = None, training: bool = False,
                force_refresh: bool = False, transform=None, n: int = None):
        """images fp32 [N, 3, S, S] (already transformed), or `transform` = an llc_img_transform
        describing the RAW batch (ops.make_transform): the input transform then runs fused in
        front of the patch embedding."""
        if transform is None:
            if images.device != self.device or images.dtype != torch.float32:
                images = images.to(device=self.device, dtype=torch.float32)
            images = images.contiguous()
            N = images.shape[0]
            if tuple(images.shape[1:]) != (3, self.cfg.image_size, self.cfg.image_size):
                raise RuntimeError(f"expected [N, 3, {self.cfg.image_size}, "
                                   f"{self.cfg.image_size}] images, got {tuple(images.shape)}")
        else:
            N = int(n)
        arena, mode = self._pick_arena(N, training)
        self._refresh_lora(force_refresh)   # force: CUDA-graph capture must always contain it
        xf = C.c_void_p()
        # every consumer of this engine reads only the class-token rows (ln_post(x[:, 0]) @ proj):
        # the last block runs class-token-only unless LLC_FULL_LAST_BLOCK is set (A/B, tests)
        self._cls_only = os.environ.get("LLC_FULL_LAST_BLOCK") is None
        lib = K.load()
        if transform is not None:
            K.check(lib.llc_vit_forward_tx(C.byref(self.cfg), C.byref(self.weights),
                                           C.byref(transform), N, arena.data_ptr(), mode,
                                           int(self._cls_only), C.byref(xf), K.stream_ptr()),
                    "llc_vit_forward_tx")
        else:
            fwd = lib.llc_vit_forward_cls if self._cls_only else lib.llc_vit_forward
            K.check(fwd(C.byref(self.cfg), C.byref(self.weights), images.data_ptr(), N,
                        arena.data_ptr(), mode, C.byref(xf), K.stream_ptr()), "llc_vit_forward")
        return self._set_x_final(arena, xf, N, mode)

    def cls_rows(self) -> torch.Tensor:
        return self.x_final.view(self.N, self.L, self.D)[:, 0, :]

    def head(self, text, logit_scale_exp, *, cls_idx=None, add_mask=None, labels=None,
             double_softmax=True, inv_batch=None, want_dlogits=False) -> ops.Head:
        return ops.Head(self.x_final, self.L, self.ln_post[0], self.ln_post[1], self.proj, text,
                        logit_scale_exp, self.N, cls_idx=cls_idx, add_mask=add_mask, labels=labels,
                        double_softmax=double_softmax, inv_batch=inv_batch,
                        want_dlogits=want_dlogits).forward()

    def head_compact(self, text, logit_scale_exp, *, cls_idx=None, add_mask=None, labels=None,
                     double_softmax=True, inv_batch=None, want_dlogits=False) -> ops.Head:
        """The same head on a compact copy of the class-token rows [N, D] (frozen image tower:
        its backward then writes an [N, D] gradient instead of a token-sized one)."""
        return ops.Head(self.cls_rows().contiguous(), 1, self.ln_post[0], self.ln_post[1], self.proj,
                        text, logit_scale_exp, self.N, cls_idx=cls_idx, add_mask=add_mask,
                        labels=labels, double_softmax=double_softmax, inv_batch=inv_batch,
                        want_dlogits=want_dlogits).forward()

    def features_only(self) -> ops.Head:
        self._feat_head = self.head(self._dummy_text, 1.0)
        return self._feat_head

    def eval_head(self, text, logit_scale_exp, *, cls_idx=None, add_mask=None, want_probs=True):
        """Evaluation-sized head on the tensor cores (BASELINE config 5: 4096 images x 1000
        classes): ln_post(CLS) -> bf16, feat = y @ proj and logits = s * f @ T^T as tcgen05 GEMMs,
        row softmax / arg-max kernel (model.py:782-785, 966-973; models/adapter_clip.py:99)."""
        if self._eval_head is None:
            self._eval_head = EvalHead(self)
        return self._eval_head.run(text, logit_scale_exp, cls_idx, add_mask, want_probs)

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


class EvalHead:
    """Buffers of VitEngine.eval_head, cached per (N, class list)."""

    def __init__(self, eng: VitEngine):
        self.eng = eng
        dev, D, E = eng.device, eng.D, eng.E
        self.projT = ops.pack_weight(eng.proj, torch.zeros(E, D, dtype=torch.bfloat16, device=dev),
                                     transpose=True)           # [E, D] K-major
        self.N = 0
        self.text_key = None

    def _ensure(self, N, Cn):
        dev, D, E = self.eng.device, self.eng.D, self.eng.E
        if N != self.N:
            self.y = torch.empty(N, D, dtype=torch.bfloat16, device=dev)
            self.feat = torch.empty(N, E, device=dev)
            self.fnorm = torch.empty(N, E, device=dev)
            self.fb = torch.empty(N, E, dtype=torch.bfloat16, device=dev)
            self.pred = torch.empty(N, dtype=torch.int64, device=dev)
            self.N, self.C = N, 0
        if Cn != self.C:
            self.Cp = (Cn + 7) // 8 * 8
            self.logits = torch.empty(N, self.Cp, device=dev)
            self.probs = torch.empty(N, Cn, device=dev)
            self.C = Cn

    def run(self, text, logit_scale_exp, cls_idx, add_mask, want_probs):
        eng = self.eng
        N = eng.N
        Cn = cls_idx.numel() if cls_idx is not None else text.shape[0]
        self._ensure(N, Cn)
        key = (text.data_ptr(), text._version, None if cls_idx is None else
               (cls_idx.data_ptr(), cls_idx._version, Cn), float(logit_scale_exp))
        if key != self.text_key:
            # bf16 operand of the logit GEMM: s * T_hat rows of the visible classes, zero rows up
            # to a multiple of 8 (one gather + cast per class list, not per batch)
            rows = text if cls_idx is None else text.index_select(0, cls_idx)
            tb = torch.zeros(self.Cp, eng.E, dtype=torch.bfloat16, device=eng.device)
            ops.l2norm_rows(rows.contiguous().float(), float(logit_scale_exp), y_bf16=tb[:Cn])
            self.tb, self.text_key = tb, key
        # ln_post of the class-token rows -> bf16 [N, D]
        x_cls = eng.x_final.view(N, eng.L * eng.D)[:, :eng.D]      # row stride L*D
        ops.ln_fwd(x_cls, eng.ln_post[0], eng.ln_post[1], self.y)
        ops.gemm_tn(self.y, self.projT, N, eng.E, eng.D, self.feat)
        ops.l2norm_rows(self.feat, 1.0, y=self.fnorm, y_bf16=self.fb)
        ops.gemm_tn(self.fb, self.tb, N, self.Cp, eng.E, self.logits)
        ops.softmax_argmax(self.logits, Cn, add_mask, self.probs if want_probs else None,
                           self.pred)
        return self


class TextEngine(_TowerEngine):
    """CLIP.encode_text with LoRA blocks (model.py:941-956): tokens [C, ctx] -> text features."""

    def __init__(self, clip):
        dev = clip.text_projection.device
        if dev.type != "cuda":
            raise RuntimeError("lifelong_clip_b200 computes on CUDA (sm_100a) only: move the model "
                               "to the GPU before calling it (there is no CPU fallback)")
        self.clip, self.device = clip, dev
        blocks = list(clip.transformer.resblocks)
        b0 = blocks[0]
        self.L = clip.context_length
        self.D = clip.transformer.width
        self.E = clip.text_projection.shape[1]
        c = K.VitCfg()
        c.image_size, c.patch, c.width, c.layers = 32, 16, self.D, len(blocks)   # geometry unused
        c.heads, c.mlp_dim, c.embed_dim = b0.n_head, b0.mlp.c_fc.out_features, self.E
        c.lora_r, c.lora_scale = b0.lora_r, b0.lora_alpha / b0.lora_r
        self.cfg = c
        self._init_lora(blocks, dev)
        f32 = lambda t: t.detach().float().contiguous()
        self.tok_emb = f32(clip.token_embedding.weight)
        self.pos = f32(clip.positional_embedding)
        self.ln_final = (f32(clip.ln_final.weight), f32(clip.ln_final.bias))
        self.proj = f32(clip.text_projection)
        w = K.TextWeights()
        w.tok_emb, w.pos_emb = self.tok_emb.data_ptr(), self.pos.data_ptr()
        w.layers = self.layers
        w.vocab, w.context = self.tok_emb.shape[0], self.L
        self.weights = w
        self._dummy_text = torch.zeros(1, self.E, device=dev)
        self._dummy_text[0, 0] = 1.0

    def _arena_bytes(self, N, training):
        return K.load().llc_text_arena_bytes(C.byref(self.cfg), self.L, N, int(training))

    def forward(self, tokens: torch.Tensor, eot_rows: torch.Tensor, training: bool,
                force_refresh: bool = False) -> ops.Head:
        """tokens int64 [C, ctx] on the device, eot_rows int64 [C] = c*ctx + argmax(tokens[c]).
        Returns the text-side head (feat = ln_final(x[eot]) @ text_projection, fnorm = its
        L2-normalised rows)."""
        if tokens.device != self.device or tokens.dtype != torch.int64 or tokens.dim() != 2 or \
                tokens.shape[1] != self.L:
            raise RuntimeError(f"expected int64 CUDA tokens [C, {self.L}]")
        Cn = tokens.shape[0]
        arena, mode = self._pick_arena(Cn, training)
        self._refresh_lora(force_refresh)
        xf = C.c_void_p()
        K.check(K.load().llc_text_forward(C.byref(self.cfg), C.byref(self.weights),
                                          tokens.data_ptr(), Cn, arena.data_ptr(), mode,
                                          C.byref(xf), K.stream_ptr()), "llc_text_forward")
        self._set_x_final(arena, xf, Cn, mode)
        self._tokens = tokens
        self.text_head = ops.Head(self.x_final, self.L, self.ln_final[0], self.ln_final[1],
                                  self.proj, self._dummy_text, 1.0, Cn, row_idx=eot_rows).forward()
        return self.text_head

    def backward(self, d_fnorm: torch.Tensor, loss_scale: float = 1.0):
        """d_fnorm [C, E]: gradient w.r.t. the normalised text features -> text LoRA grads."""
        self._check_trainable()
        h = self.text_head
        h.keep = h.keep + (d_fnorm,)
        h.args.d_fnorm = d_fnorm.data_ptr()
        h.args.d_feat = None
        h.args.skip_logit_grad = 1
        self.dx.zero_()
        h.backward(self.dx, None, loss_scale)
        K.check(K.load().llc_text_backward(C.byref(self.cfg), C.byref(self.weights), self.N,
                                           self.arena.data_ptr(), self.dx.data_ptr(),
                                           K.stream_ptr()), "llc_text_backward")


class FlatAdamW:
    """torch.optim.AdamW semantics (utils/train_utils.py:27-28: lr, weight_decay=1e-5) on the flat
    LoRA buffer of one or more towers: one llc_adamw launch per tower and step; `grad_scale` folds
    the 1/world of the data-parallel mean or a GradScaler unscale into the same pass.

    `towers` are objects with an .engine() method (VisualTransformer, the CLIP text side) or
    engines themselves. The engine is resolved at STEP time: model.to() / load_state_dict()
    rebuild it, and an optimizer holding the old instance would update buffers no Parameter views
    any more."""

    def __init__(self, towers, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5):
        if not isinstance(towers, (list, tuple)):
            towers = [towers]
        self.towers = list(towers)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.state = [None] * len(self.towers)
        self.t = 0

    @staticmethod
    def _resolve(t):
        return t.engine() if hasattr(t, "engine") else t

    def engines(self):
        return [self._resolve(t) for t in self.towers]

    def step(self, grad_scale: float = 1.0):
        self.t += 1
        for i, t in enumerate(self.towers):
            e = self._resolve(t)
            st = self.state[i]
            if st is None or st[0].numel() != e.lora_flat.numel() or \
                    st[0].device != e.lora_flat.device:
                if st is not None and st[0].numel() == e.lora_flat.numel():
                    st = (st[0].to(e.lora_flat.device), st[1].to(e.lora_flat.device))
                else:
                    st = (torch.zeros_like(e.lora_flat), torch.zeros_like(e.lora_flat))
                self.state[i] = st
            ops.adamw(e.lora_flat, e.grad_flat, st[0], st[1], self.lr, self.betas[0],
                      self.betas[1], self.eps, self.wd, self.t, grad_scale)
            e.mark_lora_updated()

    # back-compat accessors used by tests (first tower)
    @property
    def m(self):
        return self.state[0][0]

    @property
    def v(self):
        return self.state[0][1]

    def zero_grad(self):
        pass  # every backward overwrites grad_flat completely


class ParamAdamW:
    """torch.optim.AdamW semantics (utils/train_utils.py:27-28) for trainable tensors that live in
    modules (the adapters of --method adapter-clip): the parameters are re-homed as views of ONE
    flat fp32 buffer, so a step is one gradient gather + one llc_adamw launch, and data-parallel
    training all-reduces one buffer. `on_step` is called after every update (in-place kernel
    writes do not bump tensor versions: modules caching derived operands refresh there)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5,
                 on_step=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise RuntimeError("ParamAdamW: no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, device=dev)
        self.grad_flat = torch.zeros(n, device=dev)
        self.views, off = [], 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise RuntimeError("ParamAdamW: parameters must be fp32")
            v = self.flat[off:off + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v
            self.views.append(self.grad_flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.t, self.on_step = 0, on_step

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def gather_grads(self):
        """p.grad of every parameter -> grad_flat (zeros where autograd produced none)."""
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None]
        if len(have) != len(self.params):
            self.grad_flat.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        return self.grad_flat

    def step(self, grad_scale: float = 1.0):
        self.t += 1
        ops.adamw(self.flat, self.grad_flat, self.m, self.v, self.lr, self.betas[0],
                  self.betas[1], self.eps, self.wd, self.t, grad_scale)
        if self.on_step is not None:
            self.on_step()
