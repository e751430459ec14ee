"""ORACLE side — test / baseline infrastructure only (never imported by lifelong_clip_b200/*).

Runs the REFERENCE's own modules (models/clip/model.py + lora.py of qcNPU/LifeLong-CLIP) when a
checkout is reachable ($LLC_REFERENCE, /root/reference, baseline/_ref): used by
tests/golden/make_golden.py to produce the golden vectors and by `bench.py --impl reference` /
`cpu_baseline` as the "reference" kind. On a box without the checkout (the GPU box) callers fall
back to the oracle port (oracle/vit_oracle.py, kind "port").

Recipe: put <reference>/models on sys.path and import `clip.model` as a file-level
package (the package __init__ of `models` pulls timm / clip, which this image lacks).
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    for cand in (os.environ.get("LLC_REFERENCE"), "/root/reference",
                 os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "models", "clip", "model.py")):
            return cand
    return None


def load_reference_model(ref_root: str):
    path = os.path.join(ref_root, "models")
    if path not in sys.path:
        sys.path.insert(0, path)
    from clip import model as ref_model  # type: ignore
    return ref_model


class ReferenceStep:
    """One online step of the reference's PyTorch path on the CPU: its own CLIP (vision tower with
    LoRA blocks), freeze policy of methods/adapter_clip.py:115-119, forward lines
    model.py:756-767 / :685-686 / :782-785, head :966-973 + models/adapter_clip.py:99, loss
    methods/adapter_clip.py:89, AdamW(utils/train_utils.py:27-28)."""

    def __init__(self, ref_root, cfg, weights_np, text_np, lr=1e-3):
        import torch
        self.torch = torch
        m = load_reference_model(ref_root)
        torch.manual_seed(0)
        clip = m.CLIP(cfg.embed_dim, cfg.image_size, cfg.layers, cfg.width, cfg.patch, 77, 64, 64,
                      1, 1, {"method": "lora", "peft_encoder": "image",
                             "lora_alpha": cfg.lora_alpha, "lora_r": cfg.lora_r}).float()
        sd = clip.state_dict()
        for k, v in weights_np.items():
            sd[k] = torch.from_numpy(v)
        clip.load_state_dict(sd)
        for k, p in clip.named_parameters():
            if "adaptmlp" not in k and "lora" not in k:
                p.requires_grad = False
        self.clip = clip
        self.text = torch.from_numpy(text_np)
        self.opt = torch.optim.AdamW([p for p in clip.parameters() if p.requires_grad], lr=lr,
                                     weight_decay=1e-5)
        self.crit = torch.nn.CrossEntropyLoss()

    def step(self, x, y) -> float:
        torch, vis = self.torch, self.clip.visual
        self.opt.zero_grad(set_to_none=True)
        h = vis.conv1(x)
        h = h.reshape(h.shape[0], h.shape[1], -1).permute(0, 2, 1)
        h = torch.cat([vis.class_embedding.to(h.dtype) + torch.zeros(
            h.shape[0], 1, h.shape[-1], dtype=h.dtype), h], dim=1)
        h = h + vis.positional_embedding.to(h.dtype)
        h = vis.ln_pre(h).permute(1, 0, 2)
        h = vis.transformer(h).permute(1, 0, 2)
        feat = vis.ln_post(h[:, 0, :]) @ vis.proj
        f = feat / feat.norm(dim=-1, keepdim=True)
        logits = self.clip.logit_scale.exp() * f @ self.text.t()
        probs = logits.softmax(dim=-1)
        loss = self.crit(probs, y)
        loss.backward()
        self.opt.step()
        return float(loss.detach())
