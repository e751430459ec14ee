"""ORACLE — test infrastructure only. Never imported by the product path (lifelong_clip_b200/*).

CPU restatement of the reference's online-step hot path (qcNPU/LifeLong-CLIP, paths relative to
the reference root): CLIP ViT image tower with LoRA on the attention projections + cosine-logit
head + the reference's loss. Written as plain tensor math (no nn.Module from the reference), so it
is dtype-generic: run it in float64 for a truth value or float32 to mirror the reference.
Gradients come from torch autograd over this restatement.

Pinning: tests/test_oracle_golden.py checks this file against tests/golden/*.npz, which were
produced by tests/golden/make_golden.py from the reference's OWN classes (models/clip/model.py,
models/clip/lora.py imported from /root/reference). The reference ships no tests or golden vectors
of its own, so executing its classes is the only pin there is.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class VitCfg:
    image_size: int = 224
    patch: int = 16
    width: int = 768
    layers: int = 12
    heads: int = 12
    embed_dim: int = 512
    lora_r: int = 4          # models/adapter_clip.py:24-30 (hard-coded r=4, alpha=1)
    lora_alpha: float = 1.0

    @property
    def mlp_dim(self) -> int:  # model.py:219-222: c_fc is d_model -> 4*d_model
        return 4 * self.width

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def tokens(self) -> int:  # model.py:717-718: grid**2 + 1
        return self.grid ** 2 + 1

    @property
    def lora_scale(self) -> float:  # lora.py:401 scaling = alpha / r
        return self.lora_alpha / self.lora_r


VIT_B16 = VitCfg()
VIT_L14 = VitCfg(image_size=224, patch=14, width=1024, layers=24, heads=16, embed_dim=768)
VIT_TINY = VitCfg(image_size=32, patch=8, width=128, layers=2, heads=2, embed_dim=64)


def param_shapes(cfg: VitCfg) -> dict[str, tuple[int, ...]]:
    """Vision-tower parameter names/shapes exactly as the reference's state_dict has them
    (model.py:709-729 VisualTransformer, :209-222 block, lora.py:419-435 LoRA tensors)."""
    D, r, P = cfg.width, cfg.lora_r, cfg.patch
    s: dict[str, tuple[int, ...]] = {
        "visual.conv1.weight": (D, 3, P, P),
        "visual.class_embedding": (D,),
        "visual.positional_embedding": (cfg.tokens, D),
        "visual.ln_pre.weight": (D,), "visual.ln_pre.bias": (D,),
        "visual.ln_post.weight": (D,), "visual.ln_post.bias": (D,),
        "visual.proj": (D, cfg.embed_dim),
    }
    for i in range(cfg.layers):
        p = f"visual.transformer.resblocks.{i}."
        s[p + "attn.in_proj_weight"] = (3 * D, D)
        s[p + "attn.in_proj_bias"] = (3 * D,)
        s[p + "attn.in_proj_weight_lora_A"] = (r, D)
        s[p + "attn.in_proj_weight_lora_B"] = (3 * D, r)
        s[p + "attn.out_proj.weight"] = (D, D)
        s[p + "attn.out_proj.bias"] = (D,)
        s[p + "attn.out_proj.lora_A"] = (r, D)
        s[p + "attn.out_proj.lora_B"] = (D, r)
        s[p + "ln_1.weight"] = (D,); s[p + "ln_1.bias"] = (D,)
        s[p + "ln_2.weight"] = (D,); s[p + "ln_2.bias"] = (D,)
        s[p + "mlp.c_fc.weight"] = (cfg.mlp_dim, D); s[p + "mlp.c_fc.bias"] = (cfg.mlp_dim,)
        s[p + "mlp.c_proj.weight"] = (D, cfg.mlp_dim); s[p + "mlp.c_proj.bias"] = (D,)
    return s


def synth_weights(cfg: VitCfg, seed: int = 0) -> dict[str, np.ndarray]:
    """Deterministic random-init weights (numpy PCG64, stable across machines) with the scale of
    the reference's initialisers (model.py:852-885, lora.py:123-139,451-452) but with every affine
    term, bias and LoRA factor non-trivial so all gradient paths are exercised
    (out_proj.lora_B is zero at the reference's init; SURVEY.md §8c asks to randomise it)."""
    rng = np.random.default_rng(seed)
    D = cfg.width
    out: dict[str, np.ndarray] = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith(("ln_pre.weight", "ln_post.weight", "ln_1.weight", "ln_2.weight")):
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith(("ln_pre.bias", "ln_post.bias", "ln_1.bias", "ln_2.bias")):
            v = 0.1 * rng.standard_normal(shape)
        elif name.endswith("bias"):
            v = 0.02 * rng.standard_normal(shape)
        elif name.endswith("conv1.weight"):
            v = rng.standard_normal(shape) / math.sqrt(3 * cfg.patch ** 2)
        elif name.endswith(("class_embedding", "positional_embedding", "visual.proj")):
            v = D ** -0.5 * rng.standard_normal(shape)
        elif name.endswith("in_proj_weight"):
            v = D ** -0.5 * rng.standard_normal(shape)
        elif name.endswith(("out_proj.weight", "c_proj.weight")):
            v = D ** -0.5 * (2 * cfg.layers) ** -0.5 * rng.standard_normal(shape)
        elif name.endswith("c_fc.weight"):
            v = (2 * D) ** -0.5 * rng.standard_normal(shape)
        elif name.endswith(("lora_A", "lora_B")):
            fan = shape[0] + shape[1]
            bound = math.sqrt(6.0 / fan)  # xavier-uniform as lora.py:451-452
            v = rng.uniform(-bound, bound, shape)
        else:
            raise KeyError(name)
        out[name] = v.astype(np.float32)
    return out


def synth_text_features(num_classes: int, embed_dim: int, seed: int = 1) -> np.ndarray:
    """Cached class text features: L2-normalised rows (what model.py:968-969 hands the head)."""
    rng = np.random.default_rng(seed)
    t = rng.standard_normal((num_classes, embed_dim))
    t /= np.linalg.norm(t, axis=-1, keepdims=True)
    return t.astype(np.float32)


def to_torch(w: dict[str, np.ndarray], dtype=torch.float32, lora_grad: bool = True):
    """numpy weights -> torch; only tensors whose name contains 'lora' get requires_grad, the
    freeze policy of methods/adapter_clip.py:115-119."""
    out = {}
    for k, v in w.items():
        t = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
        if lora_grad and "lora" in k:
            t.requires_grad_(True)
        out[k] = t
    return out


# ------------------------------------------------------------------------------------- operators
def layer_norm(x, g, b, eps: float = 1e-5):
    """model.py:194-200 (nn.LayerNorm over the last dim, eps 1e-5, affine)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def quick_gelu(x):
    """model.py:203-206."""
    return x * torch.sigmoid(1.702 * x)


def lora_linear(x, W, b, A, B, s):
    """lora.py:837-839 (in-proj) and :1072-1074 / lora.py:163-169 (out-proj):
    F.linear(x, W, b) + F.linear(F.linear(x, A), B) * s."""
    return x @ W.T + b + ((x @ A.T) @ B.T) * s


def attention_core(q, k, v, heads: int, causal: bool = False):
    """lora.py:950 (q scaled by hd^-0.5 after the bias), :1002-1006 (head of sample n / head h is
    batch index n*H+h), :1043 bmm, :1063 softmax, :1068 bmm, :1070-1071 merge.
    q, k, v: [N, L, D] sample-major (the reference holds [L, N, D]; same math)."""
    N, L, D = q.shape
    hd = D // heads
    qh = (q * hd ** -0.5).reshape(N, L, heads, hd).permute(0, 2, 1, 3)
    kh = k.reshape(N, L, heads, hd).permute(0, 2, 1, 3)
    vh = v.reshape(N, L, heads, hd).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2)
    if causal:  # model.py:926-932 additive -inf upper triangle (text tower)
        mask = torch.full((L, L), float("-inf"), dtype=s.dtype, device=s.device).triu(1)
        s = s + mask
    p = torch.softmax(s, dim=-1)
    o = p @ vh
    return o.permute(0, 2, 1, 3).reshape(N, L, D)


def block_forward(x, w, prefix: str, cfg: VitCfg, causal: bool = False):
    """ResidualAttentionBlock.forward model.py:233-236 with the LoRA attention of :400-415."""
    D, s = cfg.width, cfg.lora_scale
    h = layer_norm(x, w[prefix + "ln_1.weight"], w[prefix + "ln_1.bias"])
    qkv = lora_linear(h, w[prefix + "attn.in_proj_weight"], w[prefix + "attn.in_proj_bias"],
                      w[prefix + "attn.in_proj_weight_lora_A"],
                      w[prefix + "attn.in_proj_weight_lora_B"], s)
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]  # lora.py:840 chunk(3)
    o = attention_core(q, k, v, cfg.heads, causal)
    x = x + lora_linear(o, w[prefix + "attn.out_proj.weight"], w[prefix + "attn.out_proj.bias"],
                        w[prefix + "attn.out_proj.lora_A"], w[prefix + "attn.out_proj.lora_B"], s)
    h2 = layer_norm(x, w[prefix + "ln_2.weight"], w[prefix + "ln_2.bias"])
    z = h2 @ w[prefix + "mlp.c_fc.weight"].T + w[prefix + "mlp.c_fc.bias"]
    x = x + quick_gelu(z) @ w[prefix + "mlp.c_proj.weight"].T + w[prefix + "mlp.c_proj.bias"]
    return x


def patch_embed(images, w, cfg: VitCfg):
    """model.py:756-766: stride-P conv (no bias) -> [N, G*G, D] -> prepend class token ->
    + positional embedding -> ln_pre.  Returns x0 [N, L, D]."""
    N = images.shape[0]
    P, G, D = cfg.patch, cfg.grid, cfg.width
    # im2col: rows (n, py, px), columns (c, i, j) == conv weight viewed [D, 3*P*P]
    pt = images.reshape(N, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(N, G * G, 3 * P * P)
    x = pt @ w["visual.conv1.weight"].reshape(D, -1).T
    cls = w["visual.class_embedding"].expand(N, 1, D)
    x = torch.cat([cls, x], dim=1) + w["visual.positional_embedding"]
    return layer_norm(x, w["visual.ln_pre.weight"], w["visual.ln_pre.bias"])


def vit_forward(images, w, cfg: VitCfg, return_tokens: bool = False):
    """VisualTransformer.forward model.py:755-787 (prompt_module=None path): image features
    [N, E] before normalisation."""
    x = patch_embed(images, w, cfg)
    for i in range(cfg.layers):
        x = block_forward(x, w, f"visual.transformer.resblocks.{i}.", cfg)
    y = layer_norm(x[:, 0, :], w["visual.ln_post.weight"], w["visual.ln_post.bias"])  # :782
    feat = y @ w["visual.proj"]                                                        # :784-785
    return (feat, x) if return_tokens else feat


def head_forward(feat, text, logit_scale_exp: float, cls_idx=None, add_mask=None):
    """model.py:966-973 + models/adapter_clip.py:99. `text` rows are already L2-normalised
    (model.py:968-969). cls_idx = visible-class gather (methods/adapter_clip.py:53-61,84);
    add_mask = additive seen-class mask (methods/mvp_clip.py:113-118). Returns
    (probs, logits, f_normalised)."""
    f = feat / feat.norm(dim=-1, keepdim=True)
    t = text if cls_idx is None else text[cls_idx]
    logits = logit_scale_exp * f @ t.T
    if add_mask is not None:
        logits = logits + add_mask
    return torch.softmax(logits, dim=-1), logits, f


def reference_loss(probs, labels, logits=None, double_softmax: bool = True):
    """methods/adapter_clip.py:89 with criterion nn.CrossEntropyLoss (methods/_trainer.py:164):
    the reference feeds the PROBABILITIES to cross-entropy (a second softmax). With
    double_softmax=False the conventional CE on the logits is returned instead."""
    if double_softmax:
        return F.cross_entropy(probs, labels)
    return F.cross_entropy(logits, labels)


def predict(probs):
    """methods/adapter_clip.py:90 topk(1) / :149 argmax -> int64 [N]."""
    return probs.argmax(dim=-1)


def label_remap(labels: np.ndarray, class_list: list[int]) -> np.ndarray:
    """methods/adapter_clip.py:75-76: y[j] = train_class_list.index(y[j]) (bit-exact int64)."""
    return np.asarray([class_list.index(int(v)) for v in labels], dtype=np.int64)


def class_lut(class_list: list[int], size: int) -> np.ndarray:
    """Dense inverse of class_list for the device kernel: lut[class id] = position, -1 if unseen."""
    lut = np.full((size,), -1, dtype=np.int64)
    for pos, c in enumerate(class_list):
        lut[int(c)] = pos
    return lut


def online_step_oracle(images: np.ndarray, labels_local: np.ndarray, w_np: dict, text: np.ndarray,
                       cfg: VitCfg, logit_scale_exp: float = 1.0 / 0.07, dtype=torch.float64,
                       double_softmax: bool = True, cls_idx=None, inv_batch: float | None = None):
    """One forward+backward of the hot path. Returns dict of numpy arrays: feat, probs, logits,
    loss, pred and the LoRA gradients keyed by parameter name."""
    w = to_torch(w_np, dtype)
    x = torch.from_numpy(images).to(dtype)
    t = torch.from_numpy(text).to(dtype)
    y = torch.from_numpy(labels_local)
    feat = vit_forward(x, w, cfg)
    idx = None if cls_idx is None else torch.from_numpy(np.asarray(cls_idx))
    probs, logits, f = head_forward(feat, t, logit_scale_exp, idx)
    loss = reference_loss(probs, y, logits, double_softmax)
    if inv_batch is not None:  # mean over a global batch larger than this shard
        loss = loss * (inv_batch * len(y))
    loss.backward()
    out = {"feat": feat, "fnorm": f, "probs": probs, "logits": logits, "loss": loss,
           "pred": predict(probs)}
    res = {k: v.detach().cpu().numpy() for k, v in out.items()}
    res["grads"] = {k: v.grad.detach().cpu().numpy() for k, v in w.items() if v.requires_grad}
    return res


# ------------------------------------------------------------------------------ text tower (N1)
@dataclass(frozen=True)
class TextCfg:
    """CLIP text transformer (model.py:833-846): context 77, vocab 49408, width 512, 8 heads,
    12 layers for ViT-B/16."""
    context: int = 77
    vocab: int = 49408
    width: int = 512
    heads: int = 8
    layers: int = 12
    embed_dim: int = 512
    lora_r: int = 4
    lora_alpha: float = 1.0

    @property
    def mlp_dim(self) -> int:
        return 4 * self.width

    @property
    def lora_scale(self) -> float:
        return self.lora_alpha / self.lora_r


TEXT_B16 = TextCfg()
TEXT_TINY = TextCfg(context=16, vocab=300, width=128, heads=2, layers=2, embed_dim=64)


def text_param_shapes(cfg: TextCfg) -> dict[str, tuple[int, ...]]:
    """Text-side parameter names/shapes as in the reference's CLIP state_dict (model.py:833-846)
    with LoRA blocks (peft_encoder='both')."""
    D, r = cfg.width, cfg.lora_r
    s: dict[str, tuple[int, ...]] = {
        "token_embedding.weight": (cfg.vocab, D),
        "positional_embedding": (cfg.context, D),
        "ln_final.weight": (D,), "ln_final.bias": (D,),
        "text_projection": (D, cfg.embed_dim),
    }
    for i in range(cfg.layers):
        p = f"transformer.resblocks.{i}."
        s[p + "attn.in_proj_weight"] = (3 * D, D)
        s[p + "attn.in_proj_bias"] = (3 * D,)
        s[p + "attn.in_proj_weight_lora_A"] = (r, D)
        s[p + "attn.in_proj_weight_lora_B"] = (3 * D, r)
        s[p + "attn.out_proj.weight"] = (D, D)
        s[p + "attn.out_proj.bias"] = (D,)
        s[p + "attn.out_proj.lora_A"] = (r, D)
        s[p + "attn.out_proj.lora_B"] = (D, r)
        s[p + "ln_1.weight"] = (D,); s[p + "ln_1.bias"] = (D,)
        s[p + "ln_2.weight"] = (D,); s[p + "ln_2.bias"] = (D,)
        s[p + "mlp.c_fc.weight"] = (cfg.mlp_dim, D); s[p + "mlp.c_fc.bias"] = (cfg.mlp_dim,)
        s[p + "mlp.c_proj.weight"] = (D, cfg.mlp_dim); s[p + "mlp.c_proj.bias"] = (D,)
    return s


def synth_text_weights(cfg: TextCfg, seed: int = 0) -> dict[str, np.ndarray]:
    """Deterministic random-init text-tower weights, scales of model.py:852-885."""
    rng = np.random.default_rng(seed)
    D = cfg.width
    out: dict[str, np.ndarray] = {}
    for name, shape in text_param_shapes(cfg).items():
        if name.endswith(("ln_final.weight", "ln_1.weight", "ln_2.weight")):
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith(("ln_final.bias", "ln_1.bias", "ln_2.bias")):
            v = 0.1 * rng.standard_normal(shape)
        elif name.endswith("bias"):
            v = 0.02 * rng.standard_normal(shape)
        elif name == "token_embedding.weight":
            v = 0.02 * rng.standard_normal(shape)
        elif name == "positional_embedding":
            v = 0.01 * rng.standard_normal(shape)
        elif name == "text_projection":
            v = D ** -0.5 * rng.standard_normal(shape)
        elif name.endswith("in_proj_weight"):
            v = D ** -0.5 * rng.standard_normal(shape)
        elif name.endswith(("out_proj.weight", "c_proj.weight")):
            v = D ** -0.5 * (2 * cfg.layers) ** -0.5 * rng.standard_normal(shape)
        elif name.endswith("c_fc.weight"):
            v = (2 * D) ** -0.5 * rng.standard_normal(shape)
        elif name.endswith(("lora_A", "lora_B")):
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            v = rng.uniform(-bound, bound, shape)
        else:
            raise KeyError(name)
        out[name] = v.astype(np.float32)
    return out


def synth_tokens(num_classes: int, cfg: TextCfg, seed: int = 0) -> np.ndarray:
    """[C, context] int64 prompts: SOT, 3..8 random ids, EOT (= vocab-1, the row arg-max that
    model.py:953-954 gathers), zero padding."""
    rng = np.random.default_rng(seed)
    out = np.zeros((num_classes, cfg.context), dtype=np.int64)
    for c in range(num_classes):
        k = int(rng.integers(3, min(9, cfg.context - 2)))
        out[c, 0] = cfg.vocab - 2
        out[c, 1:1 + k] = rng.integers(1, cfg.vocab - 2, size=k)
        out[c, 1 + k] = cfg.vocab - 1
    return out


def text_forward(tokens, w, cfg: TextCfg):
    """CLIP.encode_text model.py:941-956: token + positional embedding -> LoRA blocks under the
    causal mask (:926-932) -> ln_final -> EOT row -> @ text_projection. tokens: int64 [C, ctx]."""
    x = w["token_embedding.weight"][tokens] + w["positional_embedding"]
    bcfg = VitCfg(width=cfg.width, heads=cfg.heads, layers=cfg.layers, lora_r=cfg.lora_r,
                  lora_alpha=cfg.lora_alpha)
    for i in range(cfg.layers):
        x = block_forward(x, w, f"transformer.resblocks.{i}.", bcfg, causal=True)
    x = layer_norm(x, w["ln_final.weight"], w["ln_final.bias"])
    eot = tokens.argmax(dim=-1)
    return x[torch.arange(x.shape[0]), eot] @ w["text_projection"]


def clip_step_oracle(images: np.ndarray, labels_local: np.ndarray, tokens: np.ndarray, wv_np: dict,
                     wt_np: dict, cfg: VitCfg, tcfg: TextCfg, logit_scale_exp: float = 1.0 / 0.07,
                     dtype=torch.float64, double_softmax: bool = True):
    """One forward+backward with BOTH towers trainable (peft_encoder='both', the value
    scripts/lora_clip.sh sets): model.py:958-975 + models/adapter_clip.py:99 +
    methods/adapter_clip.py:89. Returns probs, loss, pred, text features and the LoRA gradients of
    both towers keyed by parameter name."""
    wv, wt = to_torch(wv_np, dtype), to_torch(wt_np, dtype)
    x = torch.from_numpy(images).to(dtype)
    tok = torch.from_numpy(tokens)
    y = torch.from_numpy(labels_local)
    feat = vit_forward(x, wv, cfg)
    tfeat = text_forward(tok, wt, tcfg)
    tn = tfeat / tfeat.norm(dim=-1, keepdim=True)
    probs, logits, f = head_forward(feat, tn, logit_scale_exp)
    loss = reference_loss(probs, y, logits, double_softmax)
    loss.backward()
    out = {"feat": feat, "tfeat": tfeat, "tnorm": tn, "probs": probs, "logits": logits,
           "loss": loss, "pred": predict(probs)}
    res = {k: v.detach().cpu().numpy() for k, v in out.items()}
    res["grads"] = {k: v.grad.detach().cpu().numpy() for k, v in wv.items() if v.requires_grad}
    res["grads"].update({k: v.grad.detach().cpu().numpy() for k, v in wt.items()
                         if v.requires_grad})
    return res


# --------------------------------------------------------------------------- evaluation tail (N4)
def interpret_pred(y: np.ndarray, pred: np.ndarray, n_tasks: int):
    """methods/_trainer.py:519-534: ten float bins indexed by y // n_tasks: number of samples and
    number of correct predictions per bin (IndexError past ten bins, as the reference)."""
    num = np.zeros(10, np.float32)
    ok = np.zeros(10, np.float32)
    cls = y // n_tasks
    for c in np.unique(cls):
        num[c] = float((cls == c).sum())          # IndexError when c >= 10
    good = y[y == pred] // n_tasks
    for c in np.unique(good):
        ok[c] = float((good == c).sum())
    return num, ok


def confusion(y: np.ndarray, pred: np.ndarray) -> np.ndarray:
    """sklearn.metrics.confusion_matrix(y, pred) with labels=None (methods/adapter_clip.py:166):
    rows/columns = sorted union of the values present."""
    labels = np.unique(np.concatenate([y, pred]))
    idx = {int(v): i for i, v in enumerate(labels)}
    cm = np.zeros((len(labels), len(labels)), np.int64)
    for a, b in zip(y, pred):
        cm[idx[int(a)], idx[int(b)]] += 1
    return cm


# ------------------------------------------------------------------------ MaPLe deep prompts (N3)
def strip_lora(w: dict) -> dict:
    """Weights of the vanilla (frozen) towers MaPLe runs on: the LoRA tensors dropped."""
    return {k: v for k, v in w.items() if "lora" not in k}


def synth_maple_weights(tcfg: TextCfg, vis_width: int, n_ctx: int = 3, depth: int = 3,
                        seed: int = 0) -> dict[str, np.ndarray]:
    """prompt_learner.* of models/maple.py:74-136: ctx [n_ctx, ctx_dim], proj Linear(ctx_dim ->
    vis_width), depth-1 compound text prompts and their projection layers (all trainable)."""
    rng = np.random.default_rng(seed)
    D = tcfg.width
    out = {"prompt_learner.ctx": 0.02 * rng.standard_normal((n_ctx, D)),
           "prompt_learner.proj.weight": D ** -0.5 * rng.standard_normal((vis_width, D)),
           "prompt_learner.proj.bias": 0.02 * rng.standard_normal((vis_width,))}
    for i in range(depth - 1):
        out[f"prompt_learner.compound_prompts_text.{i}"] = 0.02 * rng.standard_normal((n_ctx, D))
        out[f"prompt_learner.compound_prompt_projections.{i}.weight"] = \
            D ** -0.5 * rng.standard_normal((vis_width, D))
        out[f"prompt_learner.compound_prompt_projections.{i}.bias"] = \
            0.02 * rng.standard_normal((vis_width,))
    return {k: v.astype(np.float32) for k, v in out.items()}


def _h(x):
    """The reference rounds every prompt it splices in to fp16 (`.half()`,
    models/maple_clip/model.py:306,380,395,562) before concatenating with the activations."""
    return x.half().to(x.dtype)


def maple_forward(images, tokens, wv, wt, wp, cfg: VitCfg, tcfg: TextCfg, n_ctx: int = 3,
                  depth: int = 3, logit_scale_exp: float = 1.0 / 0.07):
    """MaPLe.forward models/maple.py:226-253 on vanilla towers:
    prompt learner (:138-175) -> text encoder (:45-71 over ResidualAttentionBlock_MaPLe,
    maple_clip/model.py:353-401: rows 1..n_ctx replaced at layers 1..depth-1) -> image encoder
    (VisionTransformer_MaPLe.forward :548-589: n_ctx shared tokens appended before ln_pre, the last
    n_ctx rows replaced at layers 1..depth-1) -> cosine logits. Returns logits [N, C]."""
    N, C = images.shape[0], tokens.shape[0]
    ctx = wp["prompt_learner.ctx"]
    # ---- text side
    emb = wt["token_embedding.weight"][tokens]                      # [C, ctx_len, D] frozen
    prompts = torch.cat([emb[:, :1], ctx.unsqueeze(0).expand(C, -1, -1), emb[:, 1 + n_ctx:]], 1)
    x = prompts + wt["positional_embedding"]
    bt = VitCfg(width=tcfg.width, heads=tcfg.heads, layers=tcfg.layers)
    zero_t = {k: torch.zeros(s, dtype=x.dtype) for k, s in text_param_shapes(tcfg).items()
              if "lora" in k}
    wtt = {**wt, **zero_t}
    for i in range(tcfg.layers):
        if 1 <= i <= depth - 1:
            c = _h(wp[f"prompt_learner.compound_prompts_text.{i - 1}"])
            x = torch.cat([x[:, :1], c.unsqueeze(0).expand(C, -1, -1), x[:, 1 + n_ctx:]], 1)
        x = block_forward(x, wtt, f"transformer.resblocks.{i}.", bt, causal=True)
    x = layer_norm(x, wt["ln_final.weight"], wt["ln_final.bias"])
    tfeat = x[torch.arange(C), tokens.argmax(-1)] @ wt["text_projection"]
    # ---- image side
    P, G, D = cfg.patch, cfg.grid, cfg.width
    pt = images.reshape(N, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(N, G * G, 3 * P * P)
    y = pt @ wv["visual.conv1.weight"].reshape(D, -1).T
    y = torch.cat([wv["visual.class_embedding"].expand(N, 1, D), y], 1) + \
        wv["visual.positional_embedding"]
    shared = ctx @ wp["prompt_learner.proj.weight"].T + wp["prompt_learner.proj.bias"]
    y = torch.cat([y, _h(shared).unsqueeze(0).expand(N, -1, -1)], 1)
    y = layer_norm(y, wv["visual.ln_pre.weight"], wv["visual.ln_pre.bias"])
    zero_v = {k: torch.zeros(s, dtype=y.dtype) for k, s in param_shapes(cfg).items() if "lora" in k}
    wvv = {**wv, **zero_v}
    for i in range(cfg.layers):
        if 1 <= i <= depth - 1:
            c = wp[f"prompt_learner.compound_prompts_text.{i - 1}"] @ \
                wp[f"prompt_learner.compound_prompt_projections.{i - 1}.weight"].T + \
                wp[f"prompt_learner.compound_prompt_projections.{i - 1}.bias"]
            y = torch.cat([y[:, :y.shape[1] - n_ctx], _h(c).unsqueeze(0).expand(N, -1, -1)], 1)
        y = block_forward(y, wvv, f"visual.transformer.resblocks.{i}.", cfg)
    feat = layer_norm(y[:, 0], wv["visual.ln_post.weight"], wv["visual.ln_post.bias"]) @ \
        wv["visual.proj"]
    f = feat / feat.norm(dim=-1, keepdim=True)
    t = tfeat / tfeat.norm(dim=-1, keepdim=True)
    return logit_scale_exp * f @ t.T


def maple_step_oracle(images, labels, tokens, wv_np, wt_np, wp_np, cfg, tcfg, n_ctx=3, depth=3,
                      logit_scale_exp=1.0 / 0.07, dtype=torch.float64):
    """logits, CE(logits, y) (methods/maple.py:94-96) and the gradients of prompt_learner.*."""
    wv = to_torch(strip_lora(wv_np), dtype, lora_grad=False)
    wt = to_torch(strip_lora(wt_np), dtype, lora_grad=False)
    wp = {k: torch.from_numpy(v).to(dtype).requires_grad_(True) for k, v in wp_np.items()}
    logits = maple_forward(torch.from_numpy(images).to(dtype), torch.from_numpy(tokens), wv, wt,
                           wp, cfg, tcfg, n_ctx, depth, logit_scale_exp)
    loss = F.cross_entropy(logits, torch.from_numpy(labels))
    loss.backward()
    return {"logits": logits.detach().numpy(), "loss": loss.detach().numpy(),
            "pred": logits.argmax(-1).numpy(),
            "grads": {k: (v.grad if v.grad is not None else torch.zeros_like(v)).detach().numpy()
                      for k, v in wp.items()}}


# ------------------------------------------------------------------- adapter-clip blocks (N4)
ADAPTER_DIM = 64        # models/clip/adapter.py:39 (down_proj hard-coded to 64 outputs)
ADAPTER_SCALE = 0.1     # models/clip/model.py:437 adapter_scalar
ADAPTER_DROPOUT = 0.1   # models/clip/model.py:434


def synth_adapter_weights(width: int, layers: int, prefix: str = "visual.transformer.resblocks.",
                          seed: int = 0) -> dict[str, np.ndarray]:
    """adaptmlp.* of every block (models/clip/adapter.py:39-52). The reference's init zeroes
    up_proj (the adapter starts as the identity); every tensor is randomised here so that all four
    gradients are exercised."""
    rng = np.random.default_rng(seed)
    out = {}
    for i in range(layers):
        p = f"{prefix}{i}.adaptmlp."
        out[p + "down_proj.weight"] = rng.uniform(-1, 1, (ADAPTER_DIM, width)) * width ** -0.5
        out[p + "down_proj.bias"] = 0.05 * rng.standard_normal(ADAPTER_DIM)
        out[p + "up_proj.weight"] = rng.uniform(-1, 1, (width, ADAPTER_DIM)) * ADAPTER_DIM ** -0.5
        out[p + "up_proj.bias"] = 0.05 * rng.standard_normal(width)
    return {k: v.astype(np.float32) for k, v in out.items()}


def adapter_forward(y, w, prefix: str, mask=None, p: float = 0.0, gate=None):
    """Adapter.forward models/clip/adapter.py:53-73 with adapter_layernorm_option='none' and
    add_residual=True: y + scale * up(dropout(relu(down(y)))). mask: keep flags (same shape as
    the bottleneck) standing in for torch's dropout draw; the kept values are scaled by
    1 / (1 - p) as nn.functional.dropout does.
    gate (kernel tests only): 0/1 flags replacing BOTH ReLU's own decision and the mask - the
    pre-activation is multiplied by the gate. ReLU is discontinuous in its gradient at 0, so an
    implementation that rounds the pre-activation (bf16) places a few gates differently from this
    fp64 restatement; with the implementation's gates handed in, what is compared is arithmetic."""
    down = y @ w[prefix + "down_proj.weight"].T + w[prefix + "down_proj.bias"]
    if gate is not None:
        down = down * gate.to(down.dtype) / (1.0 - p)
        mask = None
    else:
        down = torch.relu(down)
    if mask is not None:
        down = down * mask.to(down.dtype) / (1.0 - p)
    up = down @ w[prefix + "up_proj.weight"].T + w[prefix + "up_proj.bias"]
    return up * ADAPTER_SCALE + y


def adapter_block_forward(x, w, prefix: str, cfg: VitCfg, causal: bool = False, masks=None,
                          p: float = 0.0, gates=None):
    """ResidualAttentionBlock_Adapter.forward models/clip/model.py:440-442 (vanilla attention and
    MLP, one shared adaptmlp on both branches). masks: (mask1, mask2) or None; gates: see
    adapter_forward."""
    D = cfg.width
    m1, m2 = masks if masks is not None else (None, None)
    g1, g2 = gates if gates is not None else (None, None)
    h = layer_norm(x, w[prefix + "ln_1.weight"], w[prefix + "ln_1.bias"])
    qkv = h @ w[prefix + "attn.in_proj_weight"].T + w[prefix + "attn.in_proj_bias"]
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    o = attention_core(q, k, v, cfg.heads, causal)
    ya = o @ w[prefix + "attn.out_proj.weight"].T + w[prefix + "attn.out_proj.bias"]
    x = x + adapter_forward(ya, w, prefix + "adaptmlp.", m1, p, g1)
    h2 = layer_norm(x, w[prefix + "ln_2.weight"], w[prefix + "ln_2.bias"])
    z = h2 @ w[prefix + "mlp.c_fc.weight"].T + w[prefix + "mlp.c_fc.bias"]
    m = quick_gelu(z) @ w[prefix + "mlp.c_proj.weight"].T + w[prefix + "mlp.c_proj.bias"]
    return x + adapter_forward(m, w, prefix + "adaptmlp.", m2, p, g2)


def adapter_step_oracle(images, labels_local, w_np, wa_np, text, cfg: VitCfg,
                        logit_scale_exp: float = 1.0 / 0.07, dtype=torch.float64, masks=None,
                        p: float = 0.0, tokens=None, wt_np=None, wta_np=None, tcfg=None,
                        tmasks=None):
    """One forward + reference loss + backward of the adapter-clip method (scripts/
    adapter_clip.sh): adapter blocks in the image tower and, when tokens / wt_np / wta_np are
    given (peft_encoder='both', the reference's default, models/adapter_clip.py:19), in the text
    tower; else `text` are cached normalised features. masks / tmasks: per layer (mask1, mask2)
    keep flags [N, L, 64] (training-mode dropout), None = eval-mode. Returns probs, loss, pred and
    the gradients of every adaptmlp tensor."""
    wv = to_torch(strip_lora(w_np), dtype, lora_grad=False)
    wa = {k: torch.from_numpy(v).to(dtype).requires_grad_(True) for k, v in (wa_np or {}).items()}
    w = {**wv, **wa}
    x = patch_embed(torch.from_numpy(images).to(dtype), w, cfg)
    if not wa:   # peft_encoder='text': the image tower is the vanilla (frozen) one
        w = {**w, **{k: torch.zeros(s, dtype=dtype) for k, s in param_shapes(cfg).items()
                     if "lora" in k}}
    for i in range(cfg.layers):
        if not wa:
            x = block_forward(x, w, f"visual.transformer.resblocks.{i}.", cfg)
            continue
        x = adapter_block_forward(x, w, f"visual.transformer.resblocks.{i}.", cfg,
                                  masks=None if masks is None else masks[i], p=p)
    feat = layer_norm(x[:, 0, :], w["visual.ln_post.weight"], w["visual.ln_post.bias"]) @ \
        w["visual.proj"]
    wta = {}
    if tokens is not None:
        wt = to_torch(strip_lora(wt_np), dtype, lora_grad=False)
        wta = {k: torch.from_numpy(v).to(dtype).requires_grad_(True) for k, v in wta_np.items()}
        wtt = {**wt, **wta}
        tok = torch.from_numpy(tokens)
        t = wtt["token_embedding.weight"][tok] + wtt["positional_embedding"]
        bcfg = VitCfg(width=tcfg.width, heads=tcfg.heads, layers=tcfg.layers)
        for i in range(tcfg.layers):
            t = adapter_block_forward(t, wtt, f"transformer.resblocks.{i}.", bcfg, causal=True,
                                      masks=None if tmasks is None else tmasks[i], p=p)
        t = layer_norm(t, wtt["ln_final.weight"], wtt["ln_final.bias"])
        tfeat = t[torch.arange(t.shape[0]), tok.argmax(dim=-1)] @ wtt["text_projection"]
        tn = tfeat / tfeat.norm(dim=-1, keepdim=True)
    else:
        tn = torch.from_numpy(text).to(dtype)
    y = torch.from_numpy(labels_local)
    probs, logits, f = head_forward(feat, tn, logit_scale_exp)
    loss = reference_loss(probs, y, logits, True)
    loss.backward()
    out = {"feat": feat, "fnorm": f, "probs": probs, "logits": logits, "loss": loss,
           "pred": predict(probs), "tnorm": tn}
    np_ = lambda v: (v.float() if v.dtype == torch.bfloat16 else v).detach().cpu().numpy()
    res = {k: np_(v) for k, v in out.items()}
    res["grads"] = {k: np_(v.grad) for k, v in {**wa, **wta}.items()}
    return res


def adapter_masks(seed: int, layers: int, L: int, N: int, p: float = ADAPTER_DROPOUT):
    """Deterministic dropout keep masks standing in for torch's draws: per layer (mask1, mask2),
    uint8 [L, N, 64] in the reference's sequence-first layout (the order nn.functional.dropout is
    called in: attention branch, then MLP branch, layer by layer)."""
    rng = np.random.default_rng(seed)
    return [tuple((rng.random((L, N, ADAPTER_DIM)) >= p).astype(np.uint8) for _ in range(2))
            for _ in range(layers)]


def masks_sample_major(masks):
    """[L, N, 64] -> torch [N, L, 64] (this module's activations are sample-major)."""
    return [tuple(torch.from_numpy(np.ascontiguousarray(m.transpose(1, 0, 2))) for m in pair)
            for pair in masks]
