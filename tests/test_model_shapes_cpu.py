"""Host-side structure of the drop-in modules (no GPU, no compute: parameters on the meta device):
parameter names and counts against what the reference itself reports."""
import numpy as np
import torch

from oracle import vit_oracle as vo


def _build(**kw):
    from lifelong_clip_b200.adapter_clip import AdapterCLIP
    with torch.device("meta"):
        m = AdapterCLIP(**kw)
    for k, p in m.named_parameters():           # methods/adapter_clip.py:117-119
        if "adaptmlp" not in k and "lora" not in k:
            p.requires_grad = False
    return m


def test_adapter_clip_vitl14_parameter_counts_match_the_reference_log():
    """The reference's own run log of --method adapter-clip on ViT-L/14 (nohup.out:28-29):
    'Total parameters: 431977985', 'Trainable parameters: 4361472'."""
    m = _build(model_name="ViT-L/14", peft_method="adapter", peft_encoder="both")
    total = sum(p.numel() for p in m.parameters())
    trainable = sum(p.numel() for p in m.parameters() if p.requires_grad)
    assert total == 431977985
    assert trainable == 4361472
    names = [k for k, p in m.named_parameters() if p.requires_grad]
    assert len(names) == 4 * (24 + 12) and all(".adaptmlp." in k for k in names)


def test_lora_clip_vitb16_parameter_counts():
    """lora-clip ViT-B/16 with both towers: 149,989,377 parameters, 368,640 of them LoRA
    (221,184 vision + 147,456 text) - counted on the reference's classes."""
    m = _build(model_name="ViT-B/16", peft_method="lora", peft_encoder="both")
    assert sum(p.numel() for p in m.parameters()) == 149989377
    lora = {k: p.numel() for k, p in m.named_parameters() if p.requires_grad}
    assert sum(lora.values()) == 368640
    assert sum(v for k, v in lora.items() if ".visual." in k) == 221184


def test_adapter_state_dict_keys_are_the_oracles():
    """adaptmlp.* names and shapes = the reference's (oracle.synth_adapter_weights mirrors
    models/clip/adapter.py:39-41; tests/golden/make_golden.py loads them into the reference's
    own state_dict with an exact key / shape assertion)."""
    cfg, tcfg = vo.VIT_TINY, vo.TEXT_TINY
    m = _build(peft_method="adapter", peft_encoder="both",
               vision_config=(cfg.image_size, cfg.patch, cfg.width, cfg.layers, cfg.embed_dim),
               text_config=(tcfg.context, tcfg.vocab, tcfg.width, tcfg.heads, tcfg.layers))
    sd = m.model.state_dict()
    want = {**vo.strip_lora(vo.synth_weights(cfg, 0)), **vo.strip_lora(vo.synth_text_weights(tcfg, 0)),
            **vo.synth_adapter_weights(cfg.width, cfg.layers, "visual.transformer.resblocks.", 0),
            **vo.synth_adapter_weights(tcfg.width, tcfg.layers, "transformer.resblocks.", 0)}
    assert set(want) | {"logit_scale"} == set(sd)
    for k, v in want.items():
        assert tuple(sd[k].shape) == np.shape(v), k


def test_adapter_init_is_the_references():
    """init_option='lora' (adapter.py:44-51): up_proj and both biases zero - a fresh adapter is
    the identity - and down_proj kaiming-uniform(a = sqrt 5)."""
    from lifelong_clip_b200.adapter_modules import Adapter
    torch.manual_seed(0)
    a = Adapter(d_model=256, dropout=0.1, bottleneck=64, init_option="lora", adapter_scalar=0.1,
                adapter_layernorm_option="none")
    assert float(a.up_proj.weight.abs().max()) == 0 and float(a.up_proj.bias.abs().max()) == 0
    assert float(a.down_proj.bias.abs().max()) == 0
    bound = 1.0 / np.sqrt(256)          # kaiming_uniform(a=sqrt(5)): U(-1/sqrt(fan_in), +)
    w = a.down_proj.weight
    assert float(w.abs().max()) <= bound and float(w.abs().max()) > 0.95 * bound
    assert a.scale == 0.1 and a.dropout == 0.1 and a.down_proj.out_features == 64
