"""Model wrapper mirroring the reference's models/adapter_clip.AdapterCLIP (peft_method='lora') and
the parts of models/clip/model.CLIP it drives.

forward(image, text_tokens=None) -> (probs [N, C], image_features [N, E], text_features [C, E])
exactly as models/adapter_clip.py:94-100.

Text side, by `peft_encoder`:
  'image'  the text tower is frozen and dropout-free, so its output is a pure function of the class
           list: class text features are cached — either supplied by the
           caller (set_text_features) or computed ONCE per class by this package's text tower
           (vanilla blocks on the same kernels) — and the class restriction of
           methods/adapter_clip.py:53-61,84 is a gather index into the cache;
  'both'   (what scripts/lora_clip.sh sets) the text tower carries LoRA and is recomputed every
           step for the visible classes (CLIP.encode_text, models/clip/model.py:941-956), with
           gradients to its 48 LoRA tensors through the logit product;
  'text'   only the text tower trains: the image tower is the frozen vanilla one, runs forward
           without saved activations, and the head differentiates the class-token rows only to
           reach the text features;
  'none'   nothing trains (zero-shot evaluation).
The BPE vocabulary file of OpenAI CLIP is not part of this repository: labels_tokenize() uses the
tokenizer given to set_tokenizer() (any callable list[str] -> int64 [C, 77]); SyntheticTokenizer
is the deterministic stand-in used by the synthetic benchmarks.
"""
from __future__ import annotations

import math
import zlib

import torch
import torch.nn as nn

from . import ops
from .clip_modules import LayerNorm, Transformer, VisualTransformer

# model_name -> (image_resolution, patch, width, layers, embed_dim); heads = width // 64
# (reference models/clip/model.py:1008-1017,1036,1040 derive the same numbers from a checkpoint)
VISION_CONFIGS = {
    "ViT-B/16": (224, 16, 768, 12, 512),
    "ViT-B/32": (224, 32, 768, 12, 512),
    "ViT-L/14": (224, 14, 1024, 24, 768),
}
# model_name -> (context_length, vocab_size, transformer_width, transformer_heads, layers)
TEXT_CONFIGS = {
    "ViT-B/16": (77, 49408, 512, 8, 12),
    "ViT-B/32": (77, 49408, 512, 8, 12),
    "ViT-L/14": (77, 49408, 768, 12, 12),
}


class SyntheticTokenizer:
    """Deterministic stand-in for the BPE tokenizer: [SOT, k ids in [1, 40000), EOT, 0...], the
    EOT id (vocab-1) being the arg-max of every row as model.py:953-954 requires."""

    def __init__(self, context_length=77, vocab_size=49408, words=5):
        self.ctx, self.vocab, self.words = context_length, vocab_size, words

    def __call__(self, texts):
        out = torch.zeros(len(texts), self.ctx, dtype=torch.int64)
        hi = min(40000, self.vocab - 2)
        for i, t in enumerate(texts):
            h = zlib.crc32(t.encode("utf-8"))
            ids = []
            for k in range(self.words):
                h = (h * 1103515245 + 12345) & 0x7fffffff
                ids.append(1 + h % (hi - 1))
            row = [self.vocab - 2] + ids + [self.vocab - 1]
            out[i, :len(row)] = torch.tensor(row)
        return out


class _TextSide:
    """engine() provider of the text tower for FlatAdamW / the trainer (mirrors
    VisualTransformer.engine())."""

    def __init__(self, clip):
        self.clip = clip

    def engine(self):
        return self.clip.text_engine()


class CLIP(nn.Module):
    """models/clip/model.CLIP (:790-975): `.visual`, the text transformer with its embeddings,
    `.logit_scale`, `.dtype`, encode_image, encode_text and the cosine-logit forward. The text
    tower is optional (text_config=None): with cached text features it is never run."""

    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size,
                 context_length=None, vocab_size=None, transformer_width=None,
                 transformer_heads=None, transformer_layers=None, design_details=None):
        super().__init__()
        design_details = design_details or {}
        self.design_details = design_details
        self.visual = VisualTransformer(input_resolution=image_resolution,
                                        patch_size=vision_patch_size, width=vision_width,
                                        layers=vision_layers, heads=vision_width // 64,
                                        output_dim=embed_dim, modal='image',
                                        design_details=design_details)
        self.context_length = context_length
        self.has_text = transformer_layers is not None and transformer_layers > 0
        if self.has_text:
            self.transformer = Transformer(width=transformer_width, layers=transformer_layers,
                                           heads=transformer_heads,
                                           attn_mask=self.build_attention_mask(), modal='text',
                                           design_details=design_details)
            self.vocab_size = vocab_size
            self.token_embedding = nn.Embedding(vocab_size, transformer_width)
            self.positional_embedding = nn.Parameter(torch.empty(context_length,
                                                                 transformer_width))
            self.ln_final = LayerNorm(transformer_width)
            self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
            self._init_text(transformer_width, transformer_layers)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))  # model.py:845
        self._text_engine = None

    def _init_text(self, width, layers):
        """model.py:852-885 initialize_parameters (text side)."""
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
        attn_std = width ** -0.5
        fc_std = (2 * width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=width ** -0.5)

    def build_attention_mask(self):
        """model.py:926-932."""
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    def _apply(self, fn, *a, **k):
        self._text_engine = None
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        self._text_engine = None
        return super()._load_from_state_dict(*a, **k)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def text_engine(self):
        from .engine import TextEngine
        if not self.has_text:
            raise RuntimeError("this CLIP was built without a text tower (text_config=None)")
        if self.text_block_by_block:
            raise RuntimeError("adapter blocks run block by block (_encode_text_blocks); the "
                               "fused text engine only knows the LoRA / vanilla block")
        if self._text_engine is None:
            self._text_engine = TextEngine(self)
        return self._text_engine

    def text_side(self):
        return _TextSide(self)

    def text_lora_params(self):
        if not self.has_text:
            return ()
        return tuple(p for b in self.transformer.resblocks for p in b.lora_params())

    def logit_scale_exp(self) -> float:
        """exp(logit_scale) as a host float, read back from the device only when the (frozen)
        parameter changes - a per-step .item() would be a host sync on the hot path."""
        ver = (self.logit_scale._version, self.logit_scale.data_ptr())
        if getattr(self, "_ls_cache", (None, None))[0] != ver:
            self._ls_cache = (ver, float(self.logit_scale.detach().float().exp()))
        return self._ls_cache[1]

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))

    def encode_text(self, text):
        """model.py:941-956: text int64 [C, ctx] -> [C, E] (before normalisation)."""
        if self.text_block_by_block:
            return self._encode_text_blocks(text, normalise=False)
        lora = self.text_lora_params()
        return _TextFn.apply(self, text, *lora)

    @property
    def text_block_by_block(self) -> bool:
        from .adapter_modules import ResidualAttentionBlock_Adapter
        return self.has_text and isinstance(self.transformer.resblocks[0],
                                            ResidualAttentionBlock_Adapter)

    def _encode_text_blocks(self, text, normalise):
        """model.py:941-956 with adapter blocks: embedding gather (frozen) -> blocks under
        autograd -> ln_final(x[eot]) @ text_projection through the head kernels."""
        from .clip_modules import _RowFeatFn
        dev = self.positional_embedding.device
        text = text.to(dev)
        with torch.no_grad():
            x = self.token_embedding(text).float() + self.positional_embedding.float()
        x = x.permute(1, 0, 2).contiguous()                         # NLD -> LND
        x = self.transformer(x)
        Cn = text.shape[0]
        rows = (text.argmax(dim=-1) * Cn + torch.arange(Cn, device=dev)).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        return _RowFeatFn.apply(x, rows, f32(self.ln_final.weight), f32(self.ln_final.bias),
                                f32(self.text_projection), normalise)

    def forward(self, image, text):
        """model.py:958-975."""
        if image is None:
            return self.encode_text(text)
        if text is None:
            return self.encode_image(image)
        image_features = self.encode_image(image)
        text_features = self.encode_text(text)
        image_features = image_features / image_features.norm(dim=-1, keepdim=True)
        text_features = text_features / text_features.norm(dim=-1, keepdim=True)
        logits_per_image = self.logit_scale.exp() * image_features @ text_features.t()
        return logits_per_image, logits_per_image.t(), image_features, text_features


VisionCLIP = CLIP   # round-1 name


def build_model(state_dict: dict, design_details: dict = None) -> CLIP:
    """models/clip/model.py:1005-1066: a CLIP whose dimensions are read off an (OpenAI) checkpoint's
    state_dict, with the PEFT blocks `design_details` asks for, the checkpoint loaded and every
    parameter in fp32. ViT towers only (the ResNet towers of :16-191 are outside this package).
    The PEFT tensors (lora_A / lora_B / adaptmlp.*) are not part of a checkpoint and keep their
    reference initialisation; any OTHER missing or unexpected key raises."""
    if "visual.proj" not in state_dict:
        raise NotImplementedError("ResNet CLIP towers are not built by lifelong_clip_b200")
    sd = {k: v for k, v in state_dict.items()
          if k not in ("input_resolution", "context_length", "vocab_size")}
    vision_width = sd["visual.conv1.weight"].shape[0]
    vision_layers = len([k for k in sd if k.startswith("visual.")
                         and k.endswith(".attn.in_proj_weight")])
    patch = sd["visual.conv1.weight"].shape[-1]
    grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    embed_dim = sd["text_projection"].shape[1]
    width = sd["ln_final.weight"].shape[0]
    layers = len(set(k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")))
    model = CLIP(embed_dim, patch * grid, vision_layers, vision_width, patch,
                 sd["positional_embedding"].shape[0], sd["token_embedding.weight"].shape[0],
                 width, width // 64, layers, design_details or {})
    missing, unexpected = model.load_state_dict(sd, strict=False)
    bad = [k for k in missing if "lora" not in k and "adaptmlp" not in k]
    if bad or unexpected:
        raise RuntimeError(f"checkpoint does not match the model: missing {bad[:5]}, "
                           f"unexpected {list(unexpected)[:5]}")
    for p in model.parameters():
        p.data = p.data.float()
    return model.eval()


def eot_rows(tokens: torch.Tensor) -> torch.Tensor:
    """Row of every prompt's EOT token in the flattened [C*ctx] token axis (model.py:953-954:
    x[arange, text.argmax(-1)])."""
    Cn, ctx = tokens.shape
    return (torch.arange(Cn, device=tokens.device) * ctx + tokens.argmax(dim=-1)).contiguous()


def _text_requires_grad(model: CLIP, lora) -> bool:
    return any(p.requires_grad for p in lora)


class _TextFn(torch.autograd.Function):
    """tokens -> text features [C, E] = ln_final(x)[eot] @ text_projection (llc_text_forward)."""

    @staticmethod
    def forward(ctx, model, tokens, *lora):
        need_grad = any(ctx.needs_input_grad)
        eng = model.text_engine()
        tokens = tokens.to(eng.device).contiguous()
        head = eng.forward(tokens, eot_rows(tokens), training=need_grad)
        ctx.eng, ctx.head, ctx.lora, ctx.need_grad = eng, head, lora, need_grad
        return head.feat.clone()

    @staticmethod
    def backward(ctx, d_feat):
        if not ctx.need_grad:
            raise RuntimeError("text forward ran without grad")
        eng, h = ctx.eng, ctx.head
        d_feat = d_feat.detach().float().contiguous()
        eng._check_trainable()
        h.keep = h.keep + (d_feat,)
        h.args.d_feat = d_feat.data_ptr()
        h.args.d_fnorm = None
        h.args.skip_logit_grad = 1
        eng.dx.zero_()
        h.backward(eng.dx)
        import ctypes as C
        from . import _capi as K
        K.check(K.load().llc_text_backward(C.byref(eng.cfg), C.byref(eng.weights), eng.N,
                                           eng.arena.data_ptr(), eng.dx.data_ptr(),
                                           K.stream_ptr()), "llc_text_backward")
        return (None, None) + tuple(g.clone() if p.requires_grad else None
                                    for g, p in zip(eng.lora_grad_views, ctx.lora))


class _ProbsFn(torch.autograd.Function):
    """images (+ text tokens when the text tower trains) -> (probs, normalised image features,
    pred, normalised text features) with the tower + head kernels."""

    @staticmethod
    def forward(ctx, model, images, text, cls_idx, add_mask, tokens, n_vis, *lora):
        need_grad = any(ctx.needs_input_grad)
        lora_v, lora_t = lora[:n_vis], lora[n_vis:]
        eng = model.visual.engine()
        teng = thead = None
        if tokens is not None:       # trainable text tower: recomputed for the visible classes
            teng = model.text_engine()
            thead = teng.forward(tokens, eot_rows(tokens),
                                 training=need_grad and _text_requires_grad(model, lora_t))
            text, cls_idx = thead.fnorm, None
        img_train = any(p.requires_grad for p in lora_v)
        eng.forward(images, training=need_grad and img_train)
        if img_train or not need_grad:
            head = eng.head(text, model.logit_scale_exp(), cls_idx=cls_idx, add_mask=add_mask,
                            want_dlogits=teng is not None)
        else:   # frozen image tower: the head differentiates the compact class-token rows only
            head = eng.head_compact(text, model.logit_scale_exp(), cls_idx=cls_idx,
                                    add_mask=add_mask, want_dlogits=teng is not None)
        ctx.img_train = img_train
        ctx.eng, ctx.teng, ctx.head, ctx.thead = eng, teng, head, thead
        ctx.lora_v, ctx.lora_t, ctx.need_grad = lora_v, lora_t, need_grad
        ctx.scale = model.logit_scale_exp()
        tf = text if cls_idx is None else text.index_select(0, cls_idx)
        # clones: outputs also reachable from ctx.head would form a cycle output -> grad_fn -> ctx
        # -> output that is never collected
        probs, fnorm, pred = head.probs.clone(), head.fnorm.clone(), head.pred.clone()
        ctx.mark_non_differentiable(pred)
        return probs, fnorm, pred, tf.detach().clone()

    @staticmethod
    def backward(ctx, d_probs, d_fnorm, _, d_text):
        if not ctx.need_grad:
            raise RuntimeError("forward ran without grad")
        head = ctx.head
        d_feat = None
        if d_fnorm is not None and bool((d_fnorm != 0).any()):
            # features are returned for analysis only on the reference path; chain through
            # f = z/|z| on the [N, E] tensor: dz = (g - f (f.g)) / |z|
            f = head.fnorm
            nz = head.feat.norm(dim=-1, keepdim=True)
            d_feat = ((d_fnorm - f * (f * d_fnorm).sum(-1, keepdim=True)) / nz).contiguous()
        dp = d_probs.detach().float().contiguous() if d_probs is not None else \
            torch.zeros_like(head.probs)
        if ctx.img_train:
            ctx.eng.backward_from_head(head, d_probs=dp, d_feat=d_feat)
        else:
            head.args.skip_logit_grad = 0
            head.backward(torch.zeros(head.N, head.D, device=dp.device), dp)   # -> head.dlogits
        g_v = tuple(g.clone() if p.requires_grad else None
                    for g, p in zip(ctx.eng.lora_grad_views, ctx.lora_v))
        g_t = (None,) * len(ctx.lora_t)
        if ctx.teng is not None and ctx.teng._trained_arena and \
                any(p.requires_grad for p in ctx.lora_t):
            d_t = ops.head_dtext(head.dlogits, head.fnorm, ctx.scale)
            if d_text is not None and bool((d_text != 0).any()):
                d_t = d_t + d_text.detach().float()
            ctx.teng.backward(d_t.contiguous())
            g_t = tuple(g.clone() if p.requires_grad else None
                        for g, p in zip(ctx.teng.lora_grad_views, ctx.lora_t))
        return (None,) * 7 + g_v + g_t


class AdapterCLIP(nn.Module):
    """models/adapter_clip.py:14-104."""

    def __init__(self, model_name="ViT-B/16", peft_method='lora', peft_encoder='image',
                 device=None, vision_config=None, text_config=None):
        super().__init__()
        self._init_fields(peft_method, peft_encoder, device)
        design_details = {'method': peft_method, 'peft_encoder': peft_encoder, 'ffn_num': 64,
                          'lora_alpha': 1, 'lora_r': 4}  # models/adapter_clip.py:24-30
        res, patch, width, layers, embed = vision_config or VISION_CONFIGS[model_name]
        if text_config is None and self.text_trainable:
            text_config = TEXT_CONFIGS[model_name]
        tc = text_config or (None,) * 5
        self.model = CLIP(embed, res, layers, width, patch, tc[0], tc[1], tc[2], tc[3], tc[4],
                          design_details)
        if device is not None:
            self.model.to(device)
        self.dtype = self.model.dtype

    def _init_fields(self, peft_method, peft_encoder, device):
        if peft_method not in ('lora', 'adapter'):
            raise NotImplementedError("lifelong_clip_b200 implements the lora-clip and "
                                      "adapter-clip methods (scripts/lora_clip.sh, "
                                      "scripts/adapter_clip.sh)")
        if peft_encoder not in ('image', 'both', 'text', 'none'):   # configuration/config.py:15
            raise ValueError("peft_encoder must be one of 'none', 'both', 'text', 'image'")
        self.device = device
        self.peft_method = peft_method
        self.peft_encoder = peft_encoder
        self.text_trainable = peft_encoder in ('both', 'text')
        self.image_trainable = peft_encoder in ('both', 'image')
        self.text_tokens = None
        self.current_class_names = []
        self.prompt_template = "a bad photo of a {}."
        self._tokenizer = None
        self._text_names: list[str] = []
        self._name_to_row = {}
        self._text_all = None      # [C_all, E] normalised, device (cached path)
        self._cls_idx = None       # int64 [C] visible rows of the cache
        self._cls_key = None
        self._tok_cache = {}       # class name -> int64 [ctx] (host)
        self._tokens = None        # int64 [C, ctx] on the device ('both')
        self._add_mask = None

    @classmethod
    def from_state_dict(cls, state_dict, peft_method='lora', peft_encoder='image', device=None):
        """What clip_loader.load(model_name, design_details=...) does once the checkpoint is in
        memory (models/clip/clip_loader.py:83-139 -> build_model): the wrapper around a CLIP built
        from an OpenAI state_dict."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        cls._init_fields(self, peft_method, peft_encoder, device)
        self.model = build_model(state_dict, {'method': peft_method, 'peft_encoder': peft_encoder,
                                              'ffn_num': 64, 'lora_alpha': 1, 'lora_r': 4})
        if device is not None:
            self.model.to(device)
        self.dtype = self.model.dtype
        return self

    # ---- text side ---------------------------------------------------------------------------
    def set_tokenizer(self, fn):
        """fn(list[str]) -> int64 [C, context_length] (e.g. OpenAI CLIP's SimpleTokenizer driven
        as models/adapter_clip.py:55-70 does; SyntheticTokenizer for synthetic runs)."""
        self._tokenizer = fn
        self._tok_cache = {}

    def labels_tokenize(self, labels, context_length: int = None):
        """models/adapter_clip.py:39-74: prompt template + tokenizer -> int64 [C, ctx] on the
        model's device."""
        if self._tokenizer is None:
            raise RuntimeError("no tokenizer: call set_tokenizer() (the BPE vocabulary of OpenAI "
                               "CLIP is not shipped), or provide cached class text features with "
                               "set_text_features()")
        if isinstance(labels, str):
            labels = [labels]
        if context_length is None:
            context_length = self.model.context_length or 77
        missing = [c for c in labels if c not in self._tok_cache]
        if missing:
            toks = self._tokenizer([self.prompt_template.format(c) for c in missing])
            toks = torch.as_tensor(toks, dtype=torch.int64)
            if toks.dim() != 2 or toks.shape[1] != context_length:
                raise RuntimeError(f"tokenizer must return [C, {context_length}] ids")
            for c, row in zip(missing, toks):
                self._tok_cache[c] = row.clone()
        out = torch.stack([self._tok_cache[c] for c in labels])
        return out.to(self.model.visual.proj.device)

    def set_text_features(self, class_names, features: torch.Tensor):
        """Cache one feature row per class name (model.py:941-956 output); rows are L2-normalised
        here as model.py:968-969 does every step. peft_encoder='image' only."""
        if self.text_trainable:
            raise RuntimeError(f"peft_encoder={self.peft_encoder!r} recomputes the text features "
                               "every step; there is nothing to cache")
        feats = features.detach().float()
        feats = feats / feats.norm(dim=-1, keepdim=True)
        dev = self.model.visual.proj.device
        self._text_names = list(class_names)
        self._text_all = feats.to(dev).contiguous()
        self._name_to_row = {n: i for i, n in enumerate(self._text_names)}
        self._cls_idx = None

    def _cache_missing(self, classnames):
        """peft_encoder='image' without caller-supplied features: run the frozen text tower once
        for the classes not seen before and append their normalised features to the cache."""
        missing = [c for c in classnames if c not in self._name_to_row]
        if not missing:
            return
        if not self.model.has_text:
            raise RuntimeError(f"no text features cached for {missing[:3]}...: call "
                               "set_text_features(), or build AdapterCLIP with text_config and "
                               "set_tokenizer()")
        with torch.no_grad():
            z = self.model.encode_text(self.labels_tokenize(missing))
            z = z / z.norm(dim=-1, keepdim=True)
        base = len(self._text_names)
        self._text_all = z if self._text_all is None else torch.cat([self._text_all, z])
        for i, c in enumerate(missing):
            self._name_to_row[c] = base + i
        self._text_names += missing
        self._cls_idx = None

    def set_token(self, classnames):
        """models/adapter_clip.py:102-104: select the classes visible to forward(). 'image': the
        gather index into the cached text features (no per-step tokenisation); 'both': the token
        matrix of the visible classes (tokenised once per class name)."""
        key = tuple(classnames)
        if self.text_trainable:
            if self._tokens is None or key != self._cls_key:
                self._tokens = self.labels_tokenize(list(classnames)).contiguous()
                self._cls_key = key
            self.text_tokens = self._tokens
            return
        if self._text_all is None and not self.model.has_text:
            raise RuntimeError("call set_text_features() before set_token()")
        self._cache_missing(classnames)
        if self._cls_idx is None or key != self._cls_key:   # unchanged list: no H2D, no alloc
            rows = [self._name_to_row[c] for c in classnames]
            self._cls_idx = torch.tensor(rows, dtype=torch.int64, device=self._text_all.device)
            self._cls_key = key
        self.text_tokens = self._cls_idx

    def set_additive_mask(self, mask):
        """methods/mvp_clip.py:113-118 variant: logits + mask (0 for seen, -inf for unseen)."""
        dev = self.model.visual.proj.device
        self._add_mask = None if mask is None else mask.detach().float().to(dev).contiguous()

    def update_class_names(self, new_class_names):
        """models/adapter_clip.py:81-92 (bookkeeping only; returns None like the reference).
        Membership through a set: the reference's `c not in list` is O(classes^2) per step."""
        known = getattr(self, "_known_names", None)
        if known is None or len(known) != len(self.current_class_names):
            known = self._known_names = set(self.current_class_names)
        if len(new_class_names) == len(known) and all(c in known for c in new_class_names):
            return None
        for c in new_class_names:
            if c not in known:
                known.add(c)
                self.current_class_names.append(c)
        return None

    # ---- forward -----------------------------------------------------------------------------
    def encode_image(self, image):
        """models/adapter_clip.py:76-79: L2-normalised image features."""
        z = self.model.encode_image(image)
        return z / z.norm(dim=-1, keepdim=True)

    def forward(self, image, text_tokens=None):
        if text_tokens is None:
            text_tokens = self.text_tokens
        if text_tokens is None:
            raise RuntimeError("no visible classes: call set_token() (and set_text_features() or "
                               "set_tokenizer()) first")
        vis = self.model.visual
        if self.peft_method == 'adapter':
            probs, fnorm, _, _, tf = self._forward_blocks(image, text_tokens)
            return probs, fnorm, tf
        lv = vis.lora_params()
        if self.text_trainable:
            lt = self.model.text_lora_params()
            probs, fnorm, _, tf = _ProbsFn.apply(self.model, image, None, None, self._add_mask,
                                                 text_tokens, len(lv), *lv, *lt)
        else:
            probs, fnorm, _, tf = _ProbsFn.apply(self.model, image, self._text_all, text_tokens,
                                                 self._add_mask, None, len(lv), *lv)
        return probs, fnorm, tf

    def adapters(self):
        from .adapter_modules import Adapter
        return [m for m in self.modules() if isinstance(m, Adapter)]

    def invalidate_adapters(self):
        """After an in-place update of the adapter parameters (optimizer step): the bf16 operands
        are re-derived on the next call."""
        for a in self.adapters():
            a._ops_key = None

    def text_features_blocks(self, text_tokens):
        """Normalised text features of the visible classes on the adapter path."""
        if self.text_trainable:
            return self.model._encode_text_blocks(text_tokens, normalise=True)
        return self._text_all.index_select(0, text_tokens)

    def _forward_blocks(self, image, text_tokens, labels=None, inv_batch=None,
                        double_softmax=True, t_hat=None):
        """adapter-clip: both towers block by block under autograd (the adapters are the only
        trainable tensors), then the head kernels. Returns (probs, normalised image features,
        pred, loss_sum, normalised text features); loss_sum is 0 without labels."""
        from .clip_modules import _HeadProbsFn
        m, vis = self.model, self.model.visual
        if t_hat is None:
            t_hat = self.text_features_blocks(text_tokens)
        if vis.block_by_block:
            x = vis.forward_tokens(image.type(self.dtype))
        else:   # frozen vanilla image tower (peft_encoder='text'): fused forward, no saved
            with torch.no_grad():   # activations; the head sees the class-token rows [1, N, D]
                eng = vis.engine()
                eng.forward(image.type(self.dtype), training=False)
                x = eng.cls_rows().contiguous().unsqueeze(0)
        f32 = lambda t: t.detach().float().contiguous()
        probs, fnorm, pred, loss = _HeadProbsFn.apply(
            x, t_hat, f32(vis.ln_post.weight), f32(vis.ln_post.bias), f32(vis.proj),
            m.logit_scale_exp(), self._add_mask, labels, inv_batch, double_softmax)
        return probs, fnorm, pred, loss, t_hat.detach()
