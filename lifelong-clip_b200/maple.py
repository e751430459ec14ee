"""MaPLe: multi-modal deep prompt tuning on frozen CLIP towers (BASELINE config 4; reference
models/maple.py:74-253 over models/maple_clip/model.py:316-401,522-589), on libllc.

    prompt learner   ctx [n_ctx, 512] (+ proj -> 768 for the image side), depth-1 compound text
                     prompts with their own projections: the ONLY trainable tensors
    text encoder     [SOS, ctx, class tokens...] + pos -> frozen blocks under the causal mask; rows
                     1..n_ctx are REPLACED by the compound prompts at layers 1..depth-1 -> ln_final
                     -> EOT row @ text_projection
    image encoder    patch tokens + n_ctx projected ctx tokens (L = 197 + 3) -> ln_pre -> frozen
                     blocks; the last n_ctx rows are replaced at layers 1..depth-1 -> ln_post(CLS)
                     @ proj
    logits           exp(logit_scale) * f_hat @ t_hat^T         (CE on the logits, methods/maple.py:96)

Everything at batch scale runs in libllc kernels: the patch embedding, every block forward and
backward (llc_block_forward / llc_block_backward through ResidualAttentionBlock on the reference's
[L, N, D] layout: attention over 200 / 77 tokens on tcgen05, the activation gradient flows through
the frozen blocks down to the prompt rows, no weight gradient is ever formed), and the feature /
logit heads. Torch autograd only carries the prompt-learner's own [n_ctx, *] tensors (its two
3-row linear layers and the row splicing), as the reference does with ~3 rows of work.
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from .adapter_clip import CLIP, TEXT_CONFIGS, VISION_CONFIGS
from .clip_modules import _CosineLogitFn, _RowFeatFn, embed_images


def _h(x):
    """The reference rounds every prompt it splices in to fp16 (`.half()`,
    models/maple_clip/model.py:306,380,395,562)."""
    return x.half().to(x.dtype)


class MultiModalPromptLearner(nn.Module):
    """models/maple.py:74-175."""

    def __init__(self, ctx_dim=512, vis_dim=768, n_ctx=3, depth=3, ctx_init=None):
        super().__init__()
        self.n_ctx, self.compound_prompts_depth = n_ctx, depth
        ctx = torch.empty(n_ctx, ctx_dim)
        nn.init.normal_(ctx, std=0.02)
        if ctx_init is not None:       # embedding of "a bad photo of a" (maple.py:96-104)
            ctx = ctx_init.detach().clone().float()
        self.ctx = nn.Parameter(ctx)
        self.proj = nn.Linear(ctx_dim, vis_dim)
        self.compound_prompts_text = nn.ParameterList(
            [nn.Parameter(torch.empty(n_ctx, ctx_dim)) for _ in range(depth - 1)])
        for p in self.compound_prompts_text:
            nn.init.normal_(p, std=0.02)
        single = nn.Linear(ctx_dim, vis_dim)
        self.compound_prompt_projections = nn.ModuleList(
            [copy.deepcopy(single) for _ in range(depth - 1)])

    def forward(self, prefix, suffix):
        ctx = self.ctx.unsqueeze(0).expand(prefix.shape[0], -1, -1)
        prompts = torch.cat([prefix, ctx, suffix], dim=1)          # construct_prompts :138-160
        visual_deep = [layer(self.compound_prompts_text[i])
                       for i, layer in enumerate(self.compound_prompt_projections)]
        return prompts, self.proj(self.ctx), list(self.compound_prompts_text), visual_deep


class MaPLe(nn.Module):
    """models/maple.py:178-253. `base_clip_model` is this package's CLIP with vanilla (frozen)
    blocks in both towers; state_dict keys of the towers equal the reference's."""

    def __init__(self, model_name="ViT-B/16", n_ctx=3, device=None, vision_config=None,
                 text_config=None, depth=3):
        super().__init__()
        self.device = device
        res, patch, width, layers, embed = vision_config or VISION_CONFIGS[model_name]
        tc = text_config or TEXT_CONFIGS[model_name]
        self.base_clip_model = CLIP(embed, res, layers, width, patch, tc[0], tc[1], tc[2], tc[3],
                                    tc[4], {"method": "vanilla", "peft_encoder": "none"})
        self.prompt_learner = MultiModalPromptLearner(tc[2], width, n_ctx, depth)
        self.image_encoder = self.base_clip_model.visual
        self.logit_scale = self.base_clip_model.logit_scale
        self.n_ctx, self.depth = n_ctx, depth
        self.register_buffer("token_prefix", torch.zeros(0), persistent=False)   # SOS
        self.register_buffer("token_suffix", torch.zeros(0), persistent=False)   # class, EOS
        self.tokenized_prompts = None
        self.current_class_names = []
        self.prompt_prefix = "a bad photo of a"
        self._tokenizer = None
        if device is not None:
            self.to(device)

    @property
    def dtype(self):
        return self.base_clip_model.dtype

    def set_tokenizer(self, fn):
        self._tokenizer = fn

    def update_class_names(self, new_class_names):
        """models/maple.py:193-203."""
        n = 0
        for c in new_class_names:
            if c not in self.current_class_names:
                self.current_class_names.append(c)
                n += 1
        if n > 0:
            self.tokenized_prompts, self.token_prefix, self.token_suffix = \
                self.get_tokenized_prompts(self.current_class_names)
        return self.tokenized_prompts

    def get_tokenized_prompts(self, classnames):
        """models/maple.py:205-224: tokens of "<prefix> <name>." and the frozen embeddings of the
        SOS token and of everything after the n_ctx context slots."""
        if self._tokenizer is None:
            raise RuntimeError("no tokenizer: call set_tokenizer() (list[str] -> int64 [C, ctx])")
        prompts = [self.prompt_prefix + " " + name.replace("_", " ") + "." for name in classnames]
        dev = self.logit_scale.device
        tok = torch.as_tensor(self._tokenizer(prompts), dtype=torch.int64).to(dev)
        with torch.no_grad():
            emb = self.base_clip_model.token_embedding(tok).type(self.dtype)
        return tok, emb[:, :1, :], emb[:, 1 + self.n_ctx:, :]

    # ------------------------------------------------------------------------------------------
    def _encode_text(self, prompts, tokenized_prompts, deep_text):
        """TextEncoder.forward models/maple.py:45-71 -> L2-normalised text features [C, E]."""
        m = self.base_clip_model
        x = prompts + m.positional_embedding.type(self.dtype)
        x = x.permute(1, 0, 2).contiguous()                        # NLD -> LND
        Cn, n = x.shape[1], self.n_ctx
        for i, blk in enumerate(m.transformer.resblocks):
            if 1 <= i <= len(deep_text):                           # maple_clip/model.py:384-399
                c = _h(deep_text[i - 1]).unsqueeze(1).expand(-1, Cn, -1)
                x = torch.cat([x[:1], c, x[1 + n:]], dim=0)
            x = blk(x)
        eot = tokenized_prompts.argmax(dim=-1)
        rows = (eot * Cn + torch.arange(Cn, device=x.device)).contiguous()   # token (l, c) = row l*C + c
        return _RowFeatFn.apply(x, rows, m.ln_final.weight.detach().float().contiguous(),
                                m.ln_final.bias.detach().float().contiguous(),
                                m.text_projection.detach().float().contiguous(), True)

    def _embed_images(self, image):
        """model.py:548-559 for the image tokens: stride-P conv + class token + positional
        embedding + ln_pre (per token, so it commutes with appending the prompt tokens)."""
        return embed_images(self.image_encoder, image)

    def forward(self, image, tokenized_prompts=None, prefix=None, suffix=None):
        if image.device.type != "cuda":
            raise RuntimeError("lifelong_clip_b200 computes on CUDA (sm_100a) only (no CPU "
                               "fallback)")
        if tokenized_prompts is None:
            tokenized_prompts, prefix, suffix = (self.tokenized_prompts, self.token_prefix,
                                                 self.token_suffix)
        v, pl = self.image_encoder, self.prompt_learner
        prompts, shared_ctx, deep_text, deep_vis = pl(prefix, suffix)
        t_hat = self._encode_text(prompts, tokenized_prompts, deep_text)
        # image side: VisionTransformer_MaPLe.forward model.py:548-589
        x0 = self._embed_images(image)
        N, n = x0.shape[0], self.n_ctx
        ctx_v = F.layer_norm(_h(shared_ctx), (v.width,), v.ln_pre.weight.float(),
                             v.ln_pre.bias.float(), v.ln_pre.eps)           # 3 rows: ln_pre(ctx)
        x = torch.cat([x0, ctx_v.unsqueeze(0).expand(N, -1, -1)], dim=1)
        x = x.permute(1, 0, 2).contiguous()                                   # NLD -> LND
        L = x.shape[0]
        for i, blk in enumerate(v.transformer.resblocks):
            if 1 <= i <= len(deep_vis):                                       # model.py:366-381
                c = _h(deep_vis[i - 1]).unsqueeze(1).expand(-1, N, -1)
                x = torch.cat([x[:L - n], c], dim=0)
            x = blk(x)
        f32 = lambda t: t.detach().float().contiguous()
        return _CosineLogitFn.apply(x, t_hat, f32(v.ln_post.weight), f32(v.ln_post.bias),
                                    f32(v.proj), self.base_clip_model.logit_scale_exp())


class GraphedStep:
    """forward + CE + backward of a MaPLe model captured once in a CUDA graph and replayed (fixed
    batch shape and class list): the 24 + 24 block calls, the prompt splicing and the prompt
    learner's small torch ops become one launch, which removes the per-block host dispatch the
    kernels do not hide (3.7 of 33.5 ms at the bench shape). Gradients land in the parameters'
    .grad (memory of the graph's pool, stable across replays); the optimizer step and, under data
    parallelism, the gradient all-reduce stay outside.

        step = GraphedStep(model, x0, y0, global_batch)
        loss = step(x, y); optimizer.step()
    """

    def __init__(self, model: MaPLe, x: torch.Tensor, y: torch.Tensor, global_batch: int = None):
        self.m = model
        self.x, self.y = x.detach().clone(), y.detach().clone()
        self.gb = float(global_batch or x.shape[0])
        self.params = [p for p in model.parameters() if p.requires_grad]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):      # warm-up: allocator pool, lazily packed operands
            for _ in range(2):
                self._zero()
                self._fwd_bwd()
        cur.wait_stream(side)
        self._zero()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()
        self.grads = [p.grad for p in self.params]      # tensors of the graph's memory pool

    def _zero(self):
        for p in self.params:
            p.grad = None

    def _fwd_bwd(self):
        logits = self.m(self.x)
        loss = F.cross_entropy(logits, self.y, reduction="sum") / self.gb
        loss.backward()
        return loss.detach()

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self.params, self.grads):       # (survives zero_grad(set_to_none=True))
            p.grad = g
        return self.loss
