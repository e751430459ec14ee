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


def run_reference(cfg: vo.VitCfg, n: int, num_classes: int, seed: int):
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
    vis = clip.visual
    images, labels = synth_inputs(cfg, n, num_classes, seed + 100)
    text = vo.synth_text_features(num_classes, cfg.embed_dim, seed + 200)
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
    loss = torch.nn.CrossEntropyLoss()(probs, torch.from_numpy(labels))  # adapter_clip.py:89
    loss.backward()
    out = {
        "feat": feat.detach().numpy(), "logits": logits.detach().numpy(),
        "probs": probs.detach().numpy(), "loss": loss.detach().numpy(),
        "pred": probs.argmax(-1).numpy(), "logit_scale_exp": logit_scale.detach().numpy(),
    }
    ng = 0
    for k, p in clip.named_parameters():
        if p.grad is not None:
            assert "lora" in k
            out["grad:" + k] = p.grad.numpy()
            ng += 1
    assert ng == 4 * cfg.layers, ng
    return out


CASES = {
    # name: (cfg, batch, classes, seed)
    "tiny": (vo.VIT_TINY, 3, 10, 11),
    "vitb16": (vo.VIT_B16, 8, 100, 7),
}


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for name, (cfg, n, c, seed) in CASES.items():
        out = run_reference(cfg, n, c, seed)
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "loss", float(out["loss"]), "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
