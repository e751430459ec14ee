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


def test_build_model_reads_dimensions_off_a_checkpoint():
    """models/clip/model.py:1005-1066 build_model: dimensions from an OpenAI-style state_dict,
    PEFT blocks from design_details, checkpoint loaded (the PEFT tensors are not in it)."""
    from lifelong_clip_b200.adapter_clip import AdapterCLIP, build_model
    cfg = vo.VitCfg(image_size=48, patch=16, width=128, layers=3, heads=2, embed_dim=64)
    tcfg = vo.TextCfg(context=12, vocab=50, width=128, heads=2, layers=2, embed_dim=64)
    sd = {k: torch.from_numpy(v) for k, v in {**vo.strip_lora(vo.synth_weights(cfg, 1)),
                                              **vo.strip_lora(vo.synth_text_weights(tcfg, 2))}.items()}
    sd["logit_scale"] = torch.tensor(2.0)
    for extra in ("input_resolution", "context_length", "vocab_size"):
        sd[extra] = torch.tensor(0)
    for method, key in (("lora", "lora"), ("adapter", "adaptmlp")):
        clip = build_model(dict(sd), {"method": method, "peft_encoder": "both", "ffn_num": 64,
                                      "lora_alpha": 1, "lora_r": 4})
        v = clip.visual
        assert (v.input_resolution, v.patch_size, v.width, v.layers) == (48, 16, 128, 3)
        assert clip.context_length == 12 and clip.vocab_size == 50
        assert len(clip.transformer.resblocks) == 2 and not clip.training
        assert torch.equal(clip.visual.conv1.weight, sd["visual.conv1.weight"])
        assert torch.equal(clip.token_embedding.weight, sd["token_embedding.weight"])
        assert float(clip.logit_scale) == 2.0
        peft = [k for k, _ in clip.named_parameters() if key in k]
        assert len(peft) == 4 * (3 + 2)
    m = AdapterCLIP.from_state_dict(dict(sd), peft_method="adapter", peft_encoder="text")
    assert m.text_trainable and not m.image_trainable and m.model.context_length == 12
    assert not any("adaptmlp" in k for k, _ in m.model.visual.named_parameters())
    bad = dict(sd)
    del bad["visual.ln_post.weight"]
    import pytest
    with pytest.raises(RuntimeError, match="does not match"):
        build_model(bad, {"method": "lora", "peft_encoder": "image"})
