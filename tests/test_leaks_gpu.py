"""Steady-state device memory of every training path: after warm-up, more steps must not grow
torch's allocated bytes. (An autograd.Function that returns a tensor it also keeps on `ctx` closes
a cycle output -> grad_fn -> ctx -> output that Python's garbage collector cannot see through: the
whole graph, i.e. every block's saved activations, then outlives the step. The adapter path did
exactly that in its first version - 21 GB per step at the bench shape.)"""
import gc

import numpy as np
import pytest
import torch

from oracle import vit_oracle as vo

pytestmark = pytest.mark.gpu

CFG = vo.VitCfg(image_size=64, patch=16, width=256, layers=3, heads=4, embed_dim=128)
TCFG = (16, 300, 128, 2, 3)           # context, vocab, width, heads, layers
N, C = 12, 6


def _steady(step, warm=4, more=24, slack=1 << 20):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    gc.collect()
    base = torch.cuda.memory_allocated()
    for _ in range(more):
        step()
    torch.cuda.synchronize()
    gc.collect()
    grown = torch.cuda.memory_allocated() - base
    assert grown < slack, f"{grown / 2 ** 20:.1f} MiB more after {more} steps"


def _model(method, peft):
    from lifelong_clip_b200.adapter_clip import AdapterCLIP, SyntheticTokenizer
    torch.manual_seed(0)
    m = AdapterCLIP(peft_method=method, peft_encoder=peft,
                    vision_config=(CFG.image_size, CFG.patch, CFG.width, CFG.layers,
                                   CFG.embed_dim),
                    text_config=TCFG if peft != "image" else None).cuda()
    names = [f"class{i}" for i in range(C)]
    if peft == "image":
        m.set_text_features(names, torch.randn(C, CFG.embed_dim))
    else:
        m.set_tokenizer(SyntheticTokenizer(TCFG[0], TCFG[1]))
    return m, names


def _batch():
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((N, 3, CFG.image_size, CFG.image_size))
                         .astype(np.float32))
    return x, torch.from_numpy(rng.integers(0, C, N))


@pytest.mark.parametrize("method,peft", [("lora", "image"), ("lora", "both"), ("lora", "text"),
                                         ("adapter", "image"), ("adapter", "both"),
                                         ("adapter", "text")])
def test_trainer_step_memory_is_steady(method, peft):
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    m, names = _model(method, peft)
    tr = LoRAClipTrainer(m, names, n_classes=C, n_tasks=2, lr=1e-3, visible_classes="all")
    tr.online_before_task(0)
    x, y = _batch()
    _steady(lambda: tr.online_step(x, y, torch.arange(N)))
    _steady(lambda: tr.online_evaluate([(x, y)]), warm=2, more=8)


@pytest.mark.parametrize("method,peft", [("lora", "image"), ("lora", "both"), ("adapter", "both")])
def test_module_forward_backward_memory_is_steady(method, peft):
    """The reference's own sequence on the module (methods/adapter_clip.py:87-96)."""
    m, names = _model(method, peft)
    for k, p in m.named_parameters():
        p.requires_grad = "lora" in k or "adaptmlp" in k
    m.set_token(names)
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3)
    x, y = _batch()
    x, y = x.cuda(), y.cuda()

    def step():
        opt.zero_grad()
        probs, _, _ = m(x)
        torch.nn.functional.cross_entropy(probs, y).backward()
        opt.step()
        if method == "lora":
            m.model.visual.engine().mark_lora_updated()
        else:
            m.invalidate_adapters()

    _steady(step)


def test_maple_memory_is_steady():
    from lifelong_clip_b200.adapter_clip import SyntheticTokenizer
    from lifelong_clip_b200.maple import MaPLe
    m = MaPLe(vision_config=(CFG.image_size, CFG.patch, CFG.width, CFG.layers, CFG.embed_dim),
              text_config=TCFG).cuda()
    m.set_tokenizer(SyntheticTokenizer(TCFG[0], TCFG[1]))
    for k, p in m.named_parameters():
        p.requires_grad = "prompt_learner" in k
    m.update_class_names([f"class{i}" for i in range(C)])
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3)
    x, y = _batch()
    x, y = x.cuda(), y.cuda()

    def step():
        opt.zero_grad()
        torch.nn.functional.cross_entropy(m(x), y).backward()
        opt.step()

    _steady(step)
