"""Model wrapper mirroring the reference's models/adapter_clip.AdapterCLIP (peft_method='lora').

forward(image, text_tokens=None) -> (probs [N, C], image_features [N, E], text_features [C, E])
exactly as models/adapter_clip.py:94-100, with the class restriction of
methods/adapter_clip.py:53-61,84 realised as a gather of cached, L2-normalised class text features
(with peft_encoder='image' the text tower is frozen and dropout-free, so its output is a pure
function of the class list: SURVEY.md §8a row a14). The text tower itself and the BPE tokenizer are
the "next" row N1 and are not part of this package: text features are supplied by the caller
through set_text_features().
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .clip_modules import VisualTransformer

# model_name -> (image_resolution, patch, width, layers, embed_dim); heads = width // 64
# (reference models/clip/model.py:1008-1017,1036,1040 derive the same numbers from a checkpoint)
VISION_CONFIGS = {
    "ViT-B/16": (224, 16, 768, 12, 512),
    "ViT-B/32": (224, 32, 768, 12, 512),
    "ViT-L/14": (224, 14, 1024, 24, 768),
}


class VisionCLIP(nn.Module):
    """The part of models/clip/model.CLIP this path needs: `.visual`, `.logit_scale`, `.dtype`,
    encode_image (model.py:934-939) and the cosine-logit forward (model.py:958-975) against
    cached text features."""

    def __init__(self, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size,
                 design_details):
        super().__init__()
        self.design_details = design_details
        self.visual = VisualTransformer(input_resolution=image_resolution,
                                        patch_size=vision_patch_size, width=vision_width,
                                        layers=vision_layers, heads=vision_width // 64,
                                        output_dim=embed_dim, modal='image',
                                        design_details=design_details)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))  # model.py:845

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def logit_scale_exp(self) -> float:
        """exp(logit_scale) as a host float, read back from the device only when the (frozen)
        parameter changes - a per-step .item() would be a host sync on the hot path."""
        ver = (self.logit_scale._version, self.logit_scale.data_ptr())
        if getattr(self, "_ls_cache", (None, None))[0] != ver:
            self._ls_cache = (ver, float(self.logit_scale.detach().float().exp()))
        return self._ls_cache[1]

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))


class _ProbsFn(torch.autograd.Function):
    """images -> (probs, normalised image features) with the tower + head kernels."""

    @staticmethod
    def forward(ctx, model, images, text, cls_idx, add_mask, *lora):
        need_grad = any(ctx.needs_input_grad)
        eng = model.visual.engine()
        eng.forward(images, training=need_grad)
        head = eng.head(text, model.logit_scale_exp(), cls_idx=cls_idx,
                        add_mask=add_mask)
        ctx.eng, ctx.head, ctx.lora, ctx.need_grad = eng, head, lora, need_grad
        ctx.mark_non_differentiable(head.pred)
        return head.probs, head.fnorm, head.pred

    @staticmethod
    def backward(ctx, d_probs, d_fnorm, _):
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
        ctx.eng.backward_from_head(head, d_probs=dp, d_feat=d_feat)
        return (None,) * 5 + tuple(g.clone() if p.requires_grad else None
                                   for g, p in zip(ctx.eng.lora_grad_views, ctx.lora))


class AdapterCLIP(nn.Module):
    """models/adapter_clip.py:14-104."""

    def __init__(self, model_name="ViT-B/16", peft_method='lora', peft_encoder='image',
                 device=None, vision_config=None):
        super().__init__()
        if peft_method != 'lora':
            raise NotImplementedError("lifelong_clip_b200 implements the lora-clip method only")
        if peft_encoder != 'image':
            raise NotImplementedError(
                "peft_encoder='both'/'text' needs the LoRA text tower (SURVEY.md §8f N1, not "
                "built yet); this path runs peft_encoder='image' with cached text features")
        self.device = device
        design_details = {'method': peft_method, 'peft_encoder': peft_encoder, 'ffn_num': 64,
                          'lora_alpha': 1, 'lora_r': 4}  # models/adapter_clip.py:24-30
        res, patch, width, layers, embed = vision_config or VISION_CONFIGS[model_name]
        self.model = VisionCLIP(embed, res, layers, width, patch, design_details)
        if device is not None:
            self.model.to(device)
        self.text_tokens = None
        self.current_class_names = []
        self.dtype = self.model.dtype
        self.prompt_template = "a bad photo of a {}."
        self._text_names: list[str] = []
        self._text_all = None      # [C_all, E] normalised, device
        self._cls_idx = None       # int64 [C] visible rows
        self._cls_key = None
        self._add_mask = None

    # ---- cached text features ----------------------------------------------------------------
    def set_text_features(self, class_names, features: torch.Tensor):
        """Cache one feature row per class name (model.py:941-956 output); rows are L2-normalised
        here as model.py:968-969 does every step."""
        feats = features.detach().float()
        feats = feats / feats.norm(dim=-1, keepdim=True)
        dev = self.model.visual.proj.device
        self._text_names = list(class_names)
        self._text_all = feats.to(dev).contiguous()
        self._name_to_row = {n: i for i, n in enumerate(self._text_names)}
        self._cls_idx = None

    def labels_tokenize(self, labels, context_length: int = 77):
        raise NotImplementedError("BPE tokenisation belongs to the text side (SURVEY.md §8f N1); "
                                  "provide class text features with set_text_features()")

    def set_token(self, classnames):
        """models/adapter_clip.py:102-104: select the classes visible to forward(). Here it
        builds the gather index into the cached text features instead of re-tokenising."""
        if self._text_all is None:
            raise RuntimeError("call set_text_features() before set_token()")
        key = tuple(classnames)
        if self._cls_idx is None or key != self._cls_key:   # unchanged list: no H2D, no alloc
            rows = [self._name_to_row[c] for c in classnames]
            self._cls_idx = torch.tensor(rows, dtype=torch.int64, device=self._text_all.device)
            self._cls_key = key
        self.text_tokens = self._cls_idx

    def set_additive_mask(self, mask):
        """methods/mvp_clip.py:113-118 variant: logits + mask (0 for seen, -inf for unseen)."""
        self._add_mask = None if mask is None else mask.detach().float().to(
            self._text_all.device).contiguous()

    def update_class_names(self, new_class_names):
        """models/adapter_clip.py:81-92 (bookkeeping only; returns None like the reference)."""
        for c in new_class_names:
            if c not in self.current_class_names:
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
        if text_tokens is None or self._text_all is None:
            raise RuntimeError("no visible classes: call set_text_features() and set_token()")
        vis = self.model.visual
        probs, fnorm, _ = _ProbsFn.apply(self.model, image, self._text_all, text_tokens,
                                         self._add_mask, *vis.lora_params())
        return probs, fnorm, self._text_all[text_tokens]
