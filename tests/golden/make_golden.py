"""Generate tests/golden/*.npz by executing the REFERENCE's own classes (not the oracle).

Run in the authoring container (where /root/reference is mounted):
    python tests/golden/make_golden.py

Recipe: put <reference>/models on sys.path and import `clip.model` / `clip.lora` as
files (the package __init__ of `models` pulls timm/clip, which are not installed); build the
reference `CLIP` with design_details={'method':'lora','peft_encoder':'image',...}; load the
deterministic synthetic weights of oracle.vit_oracle.synth_weights into the vision tower; freeze as
methods/adapter_clip.py:115-119; run the sound lines of VisualTransformer.forward
(model.py:756-767, Transformer.forward :685-686, :782-785), the head (model.py:966-973,
models/adapter_clip.py:99) and the reference loss (methods/adapter_clip.py:89) in fp32; backward.
Only OUTPUTS are stored (weights/inputs are re-synthesised from seeds by the tests).

Production-shape cases (vitb16_b16 = BASELINE config C1, _b32 = C2 per GPU, _b256 = the bench
shape) store probs / loss / pred / the 48 LoRA gradients only. Batches above 64 images are run
through the reference in micro-batches of 64 with the per-sample CE summed and divided by the
full batch (the samples are independent, so this IS CrossEntropyLoss(mean) over the whole batch up
to fp32 summation order; a 256-image autograd graph of the reference needs > 50 GB of host memory).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("LLC_REFERENCE", "/root/reference")

from oracle import vit_oracle as vo  # noqa: E402


def load_reference():
    sys.path.insert(0, os.path.join(REF, "models"))
    from clip import model as ref_model  # type: ignore
    return ref_model


def synth_inputs(cfg: vo.VitCfg, n: int, num_classes: int, seed: int):
    rng = np.random.default_rng(seed)
    images = rng.standard_normal((n, 3, cfg.image_size, cfg.image_size)).astype(np.float32)
    labels = rng.integers(0, num_classes, size=(n,)).astype(np.int64)
    return images, labels


def build_reference(cfg: vo.VitCfg, seed: int):
    ref_model = load_reference()
    torch.manual_seed(0)
    # text tower kept minimal (unused: text features are cached inputs on this path)
    clip = ref_model.CLIP(cfg.embed_dim, cfg.image_size, cfg.layers, cfg.width, cfg.patch,
                          77, 64, 64, 1, 1,
                          {"method": "lora", "peft_encoder": "image",
                           "lora_alpha": cfg.lora_alpha, "lora_r": cfg.lora_r})
    clip = clip.float()
    w = vo.synth_weights(cfg, seed)
    sd = clip.state_dict()
    for k, v in w.items():
        assert k in sd and tuple(sd[k].shape) == v.shape, (k, v.shape)
        sd[k] = torch.from_numpy(v)
    clip.load_state_dict(sd)
    for k, p in clip.named_parameters():  # methods/adapter_clip.py:117-119
        if "adaptmlp" not in k and "lora" not in k:
            p.requires_grad = False
    return clip


def reference_forward(clip, images: np.ndarray, text: np.ndarray):
    """images -> (feat, logits, probs) through the reference's own modules."""
    vis = clip.visual
    x = torch.from_numpy(images)
    # model.py:756-767
    x = vis.conv1(x)
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    x = torch.cat([vis.class_embedding.to(x.dtype) + torch.zeros(
        x.shape[0], 1, x.shape[-1], dtype=x.dtype), x], dim=1)
    x = x + vis.positional_embedding.to(x.dtype)
    x = vis.ln_pre(x)
    x = x.permute(1, 0, 2)
    x = vis.transformer(x)          # model.py:685-686 -> ResidualAttentionBlock_LoRA blocks
    x = x.permute(1, 0, 2)
    x = vis.ln_post(x[:, 0, :])     # model.py:782
    feat = x @ vis.proj             # model.py:784-785
    # model.py:966-973 (text features arrive normalised), models/adapter_clip.py:99
    f = feat / feat.norm(dim=-1, keepdim=True)
    t = torch.from_numpy(text)
    logit_scale = clip.logit_scale.exp()
    logits = logit_scale * f @ t.t()
    probs = logits.softmax(dim=-1)
    return feat, logits, probs, logit_scale


def run_reference(cfg: vo.VitCfg, n: int, num_classes: int, seed: int, slim: bool = False,
                  chunk: int = 64):
    clip = build_reference(cfg, seed)
    images, labels = synth_inputs(cfg, n, num_classes, seed + 100)
    text = vo.synth_text_features(num_classes, cfg.embed_dim, seed + 200)
    feats, logit_l, prob_l, loss = [], [], [], 0.0
    for i in range(0, n, chunk):
        feat, logits, probs, logit_scale = reference_forward(clip, images[i:i + chunk], text)
        y = torch.from_numpy(labels[i:i + chunk])
        if n <= chunk:
            part = torch.nn.CrossEntropyLoss()(probs, y)           # adapter_clip.py:89
        else:
            part = torch.nn.CrossEntropyLoss(reduction="sum")(probs, y) / n
        part.backward()                                              # grads accumulate
        loss += float(part.detach())
        feats.append(feat.detach().numpy()); logit_l.append(logits.detach().numpy())
        prob_l.append(probs.detach().numpy())
    probs = np.concatenate(prob_l)
    out = {"probs": probs, "loss": np.float32(loss), "pred": probs.argmax(-1),
           "logit_scale_exp": logit_scale.detach().numpy()}
    if not slim:
        out["feat"] = np.concatenate(feats)
        out["logits"] = np.concatenate(logit_l)
    ng = 0
    total = sum(p.grad.numel() for p in clip.parameters() if p.grad is not None)
    for k, p in clip.named_parameters():
        if p.grad is not None:
            assert "lora" in k
            g = p.grad.numpy()
            if total > 400_000:     # ViT-L/14: per-tensor scaled fp16 (6e-4 relative rounding)
                sc = float(np.abs(g).max()) or 1.0
                out["grad16:" + k] = (g / sc).astype(np.float16)
                out["gscale:" + k] = np.float32(sc)
            else:
                out["grad:" + k] = g
            ng += 1
    assert ng == 4 * cfg.layers, ng
    return out


def load_grads(gold) -> dict:
    """name -> fp32 gradient from a golden file (undoes the per-tensor fp16 packing)."""
    out = {k[5:]: gold[k] for k in gold.files if k.startswith("grad:")}
    for k in gold.files:
        if k.startswith("grad16:"):
            out[k[7:]] = gold[k].astype(np.float32) * float(gold["gscale:" + k[7:]])
    return out


CASES = {
    # name: (cfg, batch, classes, seed)
    "tiny": (vo.VIT_TINY, 3, 10, 11),
    "vitb16": (vo.VIT_B16, 8, 100, 7),
}
# production shapes (outputs only): C1 batch, C2 per-GPU batch, the bench batch
BIG_CASES = {
    "vitb16_b16": (vo.VIT_B16, 16, 100, 7),
    "vitb16_b32": (vo.VIT_B16, 32, 100, 7),
    "vitb16_b256": (vo.VIT_B16, 256, 100, 7),
    # BASELINE config 3 geometry: all 24 layers of ViT-L/14 (257 tokens), 200 classes
    "vitl14": (vo.VIT_L14, 3, 200, 9),
}


def run_reference_both(cfg: vo.VitCfg, tcfg: vo.TextCfg, n: int, num_classes: int, seed: int,
                       peft: str = "both"):
    """peft_encoder='both' (scripts/lora_clip.sh:10): the reference's CLIP with LoRA blocks in BOTH
    towers; text features from its own encode_text (model.py:941-956), head model.py:966-973.
    peft='text': LoRA blocks in the text tower only (the image tower is the vanilla one)."""
    ref_model = load_reference()
    torch.manual_seed(0)
    clip = ref_model.CLIP(cfg.embed_dim, cfg.image_size, cfg.layers, cfg.width, cfg.patch,
                          tcfg.context, tcfg.vocab, tcfg.width, tcfg.heads, tcfg.layers,
                          {"method": "lora", "peft_encoder": peft,
                           "lora_alpha": cfg.lora_alpha, "lora_r": cfg.lora_r}).float()
    wv, wt = vo.synth_weights(cfg, seed), vo.synth_text_weights(tcfg, seed + 1)
    if peft == "text":
        wv = vo.strip_lora(wv)
    sd = clip.state_dict()
    for k, v in {**wv, **wt}.items():
        assert k in sd and tuple(sd[k].shape) == v.shape, (k, v.shape)
        sd[k] = torch.from_numpy(v)
    clip.load_state_dict(sd)
    for k, p in clip.named_parameters():  # methods/adapter_clip.py:117-119
        if "adaptmlp" not in k and "lora" not in k:
            p.requires_grad = False
    images, labels = synth_inputs(cfg, n, num_classes, seed + 100)
    tokens = vo.synth_tokens(num_classes, tcfg, seed + 300)
    vis = clip.visual
    x = vis.conv1(torch.from_numpy(images))
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    x = torch.cat([vis.class_embedding.to(x.dtype) + torch.zeros(
        x.shape[0], 1, x.shape[-1], dtype=x.dtype), x], dim=1)
    x = x + vis.positional_embedding.to(x.dtype)
    x = vis.ln_pre(x).permute(1, 0, 2)
    x = vis.transformer(x).permute(1, 0, 2)
    feat = vis.ln_post(x[:, 0, :]) @ vis.proj
    tfeat = clip.encode_text(torch.from_numpy(tokens))          # model.py:941-956
    f = feat / feat.norm(dim=-1, keepdim=True)                  # model.py:966-973
    t = tfeat / tfeat.norm(dim=-1, keepdim=True)
    logit_scale = clip.logit_scale.exp()
    logits = logit_scale * f @ t.t()
    probs = logits.softmax(dim=-1)                              # models/adapter_clip.py:99
    loss = torch.nn.CrossEntropyLoss()(probs, torch.from_numpy(labels))
    loss.backward()
    out = {"feat": feat.detach().numpy(), "tfeat": tfeat.detach().numpy(),
           "logits": logits.detach().numpy(), "probs": probs.detach().numpy(),
           "loss": loss.detach().numpy(), "pred": probs.argmax(-1).numpy(),
           "logit_scale_exp": logit_scale.detach().numpy()}
    ng = 0
    for k, p in clip.named_parameters():
        if p.grad is not None:
            assert "lora" in k
            out["grad:" + k] = p.grad.numpy()
            ng += 1
    assert ng == 4 * ((cfg.layers if peft == "both" else 0) + tcfg.layers), ng
    return out


BOTH_CASES = {
    # name: (vision cfg, text cfg, batch, classes, seed)
    "both_tiny": (vo.VIT_TINY, vo.TEXT_TINY, 4, 6, 21),
    "both_vitb16": (vo.VIT_B16, vo.TEXT_B16, 8, 20, 23),
}


# peft_encoder='text' (scripts/*.sh list it next to 'both' and 'image'): name -> (..., method)
TEXT_ONLY_CASES = {
    "textonly_lora_tiny": (vo.VIT_TINY, vo.TEXT_TINY, 5, 6, 51, "lora"),
    "textonly_adapter_tiny": (vo.VIT_TINY, vo.TEXT_TINY, 5, 6, 53, "adapter"),
}


MAPLE_CASES = {
    # name: (vision cfg, text cfg, batch, classes, seed)   (n_ctx = 3, prompt depth = 3)
    # (the reference hard-codes the 512 / 768 widths of ViT-B/16, models/maple.py:118,127,134)
    "maple_small": (vo.VitCfg(image_size=64, patch=16, width=768, layers=4, heads=12,
                              embed_dim=512),
                    vo.TextCfg(context=16, vocab=300, width=512, heads=8, layers=4, embed_dim=512),
                    5, 6, 31),
    "maple_vitb16": (vo.VIT_B16, vo.TEXT_B16, 4, 10, 33),
}


def run_reference_maple(cfg: vo.VitCfg, tcfg: vo.TextCfg, n: int, num_classes: int, seed: int):
    """BASELINE config 4: the reference's MaPLe (models/maple.py:74-253) on its own
    models/maple_clip/model.py CLIP (VisionTransformer_MaPLe, ResidualAttentionBlock_MaPLe).
    maple_clip's package __init__ pulls the BPE tokenizer (ftfy, not installed), so model.py is
    loaded as a file and models/maple.py is imported with a stub `clip` module whose tokenize()
    returns fixed ids (only used to initialise ctx, which is overwritten with synthetic values)."""
    import importlib.util
    import types
    spec = importlib.util.spec_from_file_location(
        "maple_clip_model", os.path.join(REF, "models", "maple_clip", "model.py"))
    mm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mm)
    stub_pkg = types.ModuleType("models")
    stub_pkg.__path__ = [os.path.join(REF, "models")]
    stub_mc = types.ModuleType("models.maple_clip")
    stub_mc.__path__ = []
    stub_clip = types.ModuleType("models.maple_clip.clip")
    stub_clip.tokenize = lambda text: torch.tensor([[tcfg.vocab - 2, 5, 6, 7, 8, 9,
                                                     tcfg.vocab - 1] + [0] * (tcfg.context - 7)])
    stub_tok = types.ModuleType("models.maple_clip.simple_tokenizer")
    stub_tok.SimpleTokenizer = lambda: None
    stub_mc.clip = stub_clip
    saved = {k: sys.modules.get(k) for k in ("models", "models.maple_clip",
                                             "models.maple_clip.clip",
                                             "models.maple_clip.simple_tokenizer")}
    sys.modules.update({"models": stub_pkg, "models.maple_clip": stub_mc,
                        "models.maple_clip.clip": stub_clip,
                        "models.maple_clip.simple_tokenizer": stub_tok})
    try:
        spec2 = importlib.util.spec_from_file_location("ref_maple",
                                                       os.path.join(REF, "models", "maple.py"))
        maple = importlib.util.module_from_spec(spec2)
        spec2.loader.exec_module(maple)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    torch.manual_seed(0)
    clip_model = mm.CLIP(cfg.embed_dim, cfg.image_size, cfg.layers, cfg.width, cfg.patch,
                         tcfg.context, tcfg.vocab, tcfg.width, tcfg.heads, tcfg.layers,
                         {"trainer": "MaPLe", "vision_depth": 0, "language_depth": 0,
                          "vision_ctx": 0, "language_ctx": 0, "maple_length": 3}).float()
    wv = vo.strip_lora(vo.synth_weights(cfg, seed))
    wt = vo.strip_lora(vo.synth_text_weights(tcfg, seed + 1))
    wp = vo.synth_maple_weights(tcfg, cfg.width, seed=seed + 2)
    sd = clip_model.state_dict()
    for k, v in {**wv, **wt}.items():
        assert k in sd and tuple(sd[k].shape) == v.shape, (k, v.shape)
        sd[k] = torch.from_numpy(v)
    clip_model.load_state_dict(sd)
    # MaPLe.__init__ downloads a checkpoint (models/maple.py:181): assemble the module by hand
    m = maple.MaPLe.__new__(maple.MaPLe)
    torch.nn.Module.__init__(m)
    maple.MultiModalPromptLearner.__init__.__globals__["clip"] = stub_clip
    m.prompt_learner = maple.MultiModalPromptLearner(clip_model, n_ctx=3)
    psd = m.prompt_learner.state_dict()
    for k, v in wp.items():
        kk = k[len("prompt_learner."):]
        assert kk in psd and tuple(psd[kk].shape) == v.shape, (kk, v.shape, psd[kk].shape)
        psd[kk] = torch.from_numpy(v)
    m.prompt_learner.load_state_dict(psd)
    m.image_encoder, m.text_encoder = clip_model.visual, maple.TextEncoder(clip_model)
    m.logit_scale, m.dtype, m.n_ctx = clip_model.logit_scale, clip_model.dtype, 3
    for k, p in m.named_parameters():       # methods/maple.py: only the prompt learner trains
        p.requires_grad = "prompt_learner" in k
    images, labels = synth_inputs(cfg, n, num_classes, seed + 100)
    tokens = torch.from_numpy(vo.synth_tokens(num_classes, tcfg, seed + 300))
    with torch.no_grad():                   # MaPLe.get_tokenized_prompts (:205-224)
        emb = clip_model.token_embedding(tokens).type(m.dtype)
    prefix, suffix = emb[:, :1, :], emb[:, 1 + 3:, :]
    logits = m(torch.from_numpy(images), tokens, prefix, suffix)      # MaPLe.forward :226-253
    loss = torch.nn.CrossEntropyLoss()(logits, torch.from_numpy(labels))   # methods/maple.py:96
    loss.backward()
    out = {"logits": logits.detach().numpy(), "loss": loss.detach().numpy(),
           "pred": logits.argmax(-1).numpy(),
           "logit_scale_exp": clip_model.logit_scale.exp().detach().numpy()}
    ng = 0
    for k, p in m.named_parameters():
        if p.grad is not None:
            g = p.grad.numpy()
            ng += 1
            if g.size > 100_000:
                # the three [768, 512] projection-weight gradients: kept as scaled fp16 (6e-4
                # relative rounding, far below the 1e-2 parity tolerance) to keep fixtures small
                sc = float(np.abs(g).max()) or 1.0
                out["grad16:" + k] = (g / sc).astype(np.float16)
                out["gscale:" + k] = np.float32(sc)
            else:
                out["grad:" + k] = g
    assert ng == 9
    return out


def load_maple_grads(gold) -> dict:
    """name -> fp32 gradient from a ref_maple_*.npz (undoes the fp16 packing above)."""
    out = {k[5:]: gold[k] for k in gold.files if k.startswith("grad:")}
    for k in gold.files:
        if k.startswith("grad16:"):
            out[k[7:]] = gold[k].astype(np.float32) * float(gold["gscale:" + k[7:]])
    return out


def run_interpret_pred():
    """Execute the reference's OWN _interpret_pred (methods/_trainer.py:519-534; the module itself
    cannot be imported here - it pulls randaugment / timm - so the function's source is compiled
    out of the file) and sklearn's confusion_matrix as methods/adapter_clip.py:166 calls it, on
    seeded labels/predictions. Stores inputs and outputs."""
    import ast
    import types
    from sklearn.metrics import confusion_matrix
    path = os.path.join(REF, "methods", "_trainer.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)
              and n.name == "_interpret_pred")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch}
    exec(compile(mod, path, "exec"), ns)
    out = {}
    rng = np.random.default_rng(5)
    for i, (n, ncls, n_tasks) in enumerate([(257, 100, 10), (64, 37, 10), (1000, 100, 10),
                                            (33, 50, 5)]):
        y = rng.integers(0, ncls, size=n).astype(np.int64)
        pred = np.where(rng.random(n) < 0.6, y, rng.integers(0, ncls, size=n)).astype(np.int64)
        self_ = types.SimpleNamespace(n_tasks=n_tasks)
        num, ok = ns["_interpret_pred"](self_, torch.from_numpy(y), torch.from_numpy(pred))
        out[f"y{i}"], out[f"pred{i}"] = y, pred
        out[f"meta{i}"] = np.asarray([ncls, n_tasks], np.int64)
        out[f"num{i}"], out[f"ok{i}"] = num.numpy(), ok.numpy()
        out[f"cm{i}"] = confusion_matrix(y.tolist(), pred.tolist())
    return out


ADAPTER_CASES = {
    # name: (vision cfg, text cfg, batch, classes, seed): adapter-clip, peft_encoder='both',
    # training mode with injected dropout masks + the eval-mode forward of the same inputs
    "adapter_tiny": (vo.VIT_TINY, vo.TEXT_TINY, 4, 6, 41),
    "adapter_vitb16": (vo.VIT_B16, vo.TEXT_B16, 4, 10, 43),
}


def run_reference_adapter(cfg: vo.VitCfg, tcfg: vo.TextCfg, n: int, num_classes: int, seed: int,
                          peft: str = "both"):
    """--method adapter-clip (scripts/adapter_clip.sh): the reference's CLIP with
    ResidualAttentionBlock_Adapter in both towers (model.py:418-442, adapter.py:11-73). Dropout
    draws are replaced by oracle.adapter_masks (nn.functional.dropout is patched for the run: the
    reference's own draws come from torch's generator and cannot be reproduced by another
    implementation); everything else is the reference's code."""
    ref_model = load_reference()
    torch.manual_seed(0)
    clip = ref_model.CLIP(cfg.embed_dim, cfg.image_size, cfg.layers, cfg.width, cfg.patch,
                          tcfg.context, tcfg.vocab, tcfg.width, tcfg.heads, tcfg.layers,
                          {"method": "adapter", "peft_encoder": peft, "ffn_num": 64}).float()
    wv = vo.strip_lora(vo.synth_weights(cfg, seed))
    wt = vo.strip_lora(vo.synth_text_weights(tcfg, seed + 1))
    wa = vo.synth_adapter_weights(cfg.width, cfg.layers, "visual.transformer.resblocks.", seed + 2)
    wta = vo.synth_adapter_weights(tcfg.width, tcfg.layers, "transformer.resblocks.", seed + 3)
    if peft == "text":
        wa = {}
    sd = clip.state_dict()
    for k, v in {**wv, **wt, **wa, **wta}.items():
        assert k in sd and tuple(sd[k].shape) == v.shape, (k, v.shape)
        sd[k] = torch.from_numpy(v)
    clip.load_state_dict(sd)
    for k, p in clip.named_parameters():  # methods/adapter_clip.py:117-119
        if "adaptmlp" not in k and "lora" not in k:
            p.requires_grad = False
    images, labels = synth_inputs(cfg, n, num_classes, seed + 100)
    tokens = vo.synth_tokens(num_classes, tcfg, seed + 300)
    p_drop = vo.ADAPTER_DROPOUT
    queue = []

    def fake_dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        assert abs(p - p_drop) < 1e-12
        m = queue.pop(0)
        assert m.shape == x.shape, (m.shape, x.shape)
        return x * m.to(x.dtype) / (1.0 - p)

    def forward():
        vis = clip.visual
        x = vis.conv1(torch.from_numpy(images))
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        x = torch.cat([vis.class_embedding.to(x.dtype) + torch.zeros(
            x.shape[0], 1, x.shape[-1], dtype=x.dtype), x], dim=1)
        x = x + vis.positional_embedding.to(x.dtype)
        x = vis.ln_pre(x).permute(1, 0, 2)
        x = vis.transformer(x).permute(1, 0, 2)
        feat = vis.ln_post(x[:, 0, :]) @ vis.proj
        tfeat = clip.encode_text(torch.from_numpy(tokens))          # model.py:941-956
        f = feat / feat.norm(dim=-1, keepdim=True)                  # model.py:966-973
        t = tfeat / tfeat.norm(dim=-1, keepdim=True)
        logits = clip.logit_scale.exp() * f @ t.t()
        return feat, tfeat, logits, logits.softmax(dim=-1)          # models/adapter_clip.py:99

    real = torch.nn.functional.dropout
    torch.nn.functional.dropout = fake_dropout
    try:
        clip.train()
        for pair in (vo.adapter_masks(seed + 400, cfg.layers, cfg.tokens, n) if wa else []):
            queue += [torch.from_numpy(m) for m in pair]
        for pair in vo.adapter_masks(seed + 500, tcfg.layers, tcfg.context, num_classes):
            queue += [torch.from_numpy(m) for m in pair]
        feat, tfeat, logits, probs = forward()
        assert not queue
        loss = torch.nn.CrossEntropyLoss()(probs, torch.from_numpy(labels))
        loss.backward()
        clip.eval()
        with torch.no_grad():
            _, _, _, probs_eval = forward()
    finally:
        torch.nn.functional.dropout = real
    out = {"feat": feat.detach().numpy(), "tfeat": tfeat.detach().numpy(),
           "logits": logits.detach().numpy(), "probs": probs.detach().numpy(),
           "probs_eval": probs_eval.numpy(),
           "loss": loss.detach().numpy(), "pred": probs.argmax(-1).numpy(),
           "logit_scale_exp": clip.logit_scale.exp().detach().numpy()}
    ng = 0
    for k, p in clip.named_parameters():
        if p.grad is not None:
            assert "adaptmlp" in k
            g = p.grad.numpy()
            ng += 1
            layer = int(k.split("resblocks.")[1].split(".")[0])
            if g.size > 4096 and cfg.layers > 4 and layer not in (0, cfg.layers // 2,
                                                                  cfg.layers - 1):
                continue            # full-size case: weight gradients of three layers per tower
            if g.size > 4096:       # scaled fp16 keeps the file small (as load_grads undoes)
                sc = float(np.abs(g).max()) / 32768.0 or 1.0
                out["grad16:" + k] = (g / sc).astype(np.float16)
                out["gscale:" + k] = np.float32(sc)
            else:
                out["grad:" + k] = g
    assert ng == 4 * ((cfg.layers if wa else 0) + tcfg.layers), ng
    return out


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:]
    for name, (cfg, n, c, seed) in {**CASES, **BIG_CASES}.items():
        if only and name not in only:
            continue
        out = run_reference(cfg, n, c, seed, slim=name in BIG_CASES)
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "loss", float(out["loss"]), "->", path, os.path.getsize(path), "bytes")
    for name, (cfg, tcfg, n, c, seed) in BOTH_CASES.items():
        if only and name not in only:
            continue
        out = run_reference_both(cfg, tcfg, n, c, seed)
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "loss", float(out["loss"]), "->", path, os.path.getsize(path), "bytes")
    for name, (cfg, tcfg, n, c, seed) in MAPLE_CASES.items():
        if only and name not in only:
            continue
        out = run_reference_maple(cfg, tcfg, n, c, seed)
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "loss", float(out["loss"]), "->", path, os.path.getsize(path), "bytes")
    for name, (cfg, tcfg, n, c, seed) in ADAPTER_CASES.items():
        if only and name not in only:
            continue
        out = run_reference_adapter(cfg, tcfg, n, c, seed)
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "loss", float(out["loss"]), "->", path, os.path.getsize(path), "bytes")
    for name, (cfg, tcfg, n, c, seed, method) in TEXT_ONLY_CASES.items():
        if only and name not in only:
            continue
        run = run_reference_both if method == "lora" else run_reference_adapter
        out = run(cfg, tcfg, n, c, seed, peft="text")
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "loss", float(out["loss"]), "->", path, os.path.getsize(path), "bytes")
    if not only or "interpret_pred" in only:
        path = os.path.join(HERE, "ref_interpret_pred.npz")
        np.savez_compressed(path, **run_interpret_pred())
        print("interpret_pred ->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
