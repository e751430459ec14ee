"""Pins oracle/vit_oracle.py against golden vectors produced by the REFERENCE's own classes
(tests/golden/make_golden.py). CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import vit_oracle as vo
from tests.golden.make_golden import CASES, synth_inputs


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("name", ["tiny", "vitb16"])
def test_oracle_matches_reference_golden(name, golden_dir):
    cfg, n, c, seed = CASES[name]
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    torch.set_num_threads(os.cpu_count() or 1)
    w = vo.synth_weights(cfg, seed)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 200)
    # fp32 like the reference for the big case (speed); fp64 truth for the tiny one
    dtype = torch.float64 if name == "tiny" else torch.float32
    out = vo.online_step_oracle(images, labels, w, text, cfg,
                                logit_scale_exp=float(gold["logit_scale_exp"]), dtype=dtype)
    # the reference ran in fp32: agreement is limited by ITS rounding
    assert _rel(out["feat"], gold["feat"]) < 2e-5
    assert _rel(out["logits"], gold["logits"]) < 2e-5
    assert _rel(out["probs"], gold["probs"]) < 2e-5
    assert abs(float(out["loss"]) - float(gold["loss"])) < 1e-5
    np.testing.assert_array_equal(out["pred"], gold["pred"])  # integer: bit-exact
    grads = {k[5:]: gold[k] for k in gold.files if k.startswith("grad:")}
    assert set(grads) == set(out["grads"]) and len(grads) == 4 * cfg.layers
    worst = max(_rel(out["grads"][k], v) for k, v in grads.items())
    assert worst < 5e-4, worst


def test_label_remap_matches_python_loop():
    # methods/adapter_clip.py:75-76 restated; class list in order of first exposure
    rng = np.random.default_rng(0)
    class_list = [17, 3, 99, 42, 0, 8]
    y = rng.choice(class_list, size=64).astype(np.int64)
    local = vo.label_remap(y, class_list)
    lut = vo.class_lut(class_list, 100)
    np.testing.assert_array_equal(local, lut[y])
    assert local.dtype == np.int64 and lut[5] == -1


def test_double_softmax_loss_is_ce_on_probs():
    # methods/adapter_clip.py:89: CE applied to probabilities
    rng = np.random.default_rng(1)
    logits = torch.from_numpy(rng.standard_normal((5, 7)))
    y = torch.from_numpy(rng.integers(0, 7, size=(5,)))
    p = logits.softmax(-1)
    want = (-p[torch.arange(5), y] + torch.logsumexp(p, -1)).mean()
    got = vo.reference_loss(p, y)
    assert abs(float(want) - float(got)) < 1e-12


@pytest.mark.parametrize("name", ["both_tiny", "both_vitb16"])
def test_text_tower_oracle_matches_reference_golden(name, golden_dir):
    """peft_encoder='both': the oracle's text tower (model.py:941-956 restated) and the two-tower
    step against the reference's own CLIP with LoRA blocks in both transformers."""
    from tests.golden.make_golden import BOTH_CASES
    cfg, tcfg, n, c, seed = BOTH_CASES[name]
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    torch.set_num_threads(os.cpu_count() or 1)
    wv, wt = vo.synth_weights(cfg, seed), vo.synth_text_weights(tcfg, seed + 1)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    tokens = vo.synth_tokens(c, tcfg, seed + 300)
    dtype = torch.float64 if name == "both_tiny" else torch.float32
    out = vo.clip_step_oracle(images, labels, tokens, wv, wt, cfg, tcfg,
                              logit_scale_exp=float(gold["logit_scale_exp"]), dtype=dtype)
    assert _rel(out["tfeat"], gold["tfeat"]) < 2e-5
    assert _rel(out["probs"], gold["probs"]) < 2e-5
    assert abs(float(out["loss"]) - float(gold["loss"])) < 1e-5
    np.testing.assert_array_equal(out["pred"], gold["pred"])
    grads = {k[5:]: gold[k] for k in gold.files if k.startswith("grad:")}
    assert set(grads) == set(out["grads"]) and len(grads) == 4 * (cfg.layers + tcfg.layers)
    worst = max(_rel(out["grads"][k], v) for k, v in grads.items())
    assert worst < 1e-3, worst


def test_interpret_pred_and_confusion_match_reference(golden_dir):
    """oracle.interpret_pred / confusion against the outputs of the reference's own
    _interpret_pred (methods/_trainer.py:519-534) and sklearn's confusion_matrix: bit-exact."""
    gold = np.load(os.path.join(golden_dir, "ref_interpret_pred.npz"))
    i = 0
    while f"y{i}" in gold.files:
        y, pred = gold[f"y{i}"], gold[f"pred{i}"]
        n_tasks = int(gold[f"meta{i}"][1])
        num, ok = vo.interpret_pred(y, pred, n_tasks)
        np.testing.assert_array_equal(num, gold[f"num{i}"])
        np.testing.assert_array_equal(ok, gold[f"ok{i}"])
        np.testing.assert_array_equal(vo.confusion(y, pred), gold[f"cm{i}"])
        i += 1
    assert i == 4


@pytest.mark.parametrize("name", ["maple_small", "maple_vitb16"])
def test_maple_oracle_matches_reference_golden(name, golden_dir):
    """BASELINE config 4: the oracle's MaPLe restatement (deep multi-modal prompts over frozen
    towers) against the reference's own models/maple.py + models/maple_clip/model.py."""
    from tests.golden.make_golden import MAPLE_CASES, load_maple_grads
    cfg, tcfg, n, c, seed = MAPLE_CASES[name]
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    torch.set_num_threads(os.cpu_count() or 1)
    wv, wt = vo.synth_weights(cfg, seed), vo.synth_text_weights(tcfg, seed + 1)
    wp = vo.synth_maple_weights(tcfg, cfg.width, seed=seed + 2)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    tokens = vo.synth_tokens(c, tcfg, seed + 300)
    out = vo.maple_step_oracle(images, labels, tokens, wv, wt, wp, cfg, tcfg,
                               logit_scale_exp=float(gold["logit_scale_exp"]),
                               dtype=torch.float32)
    assert _rel(out["logits"], gold["logits"]) < 2e-5
    assert abs(float(out["loss"]) - float(gold["loss"])) < 1e-5
    np.testing.assert_array_equal(out["pred"], gold["pred"])
    want = load_maple_grads(gold)
    assert {"prompt_learner." + k if not k.startswith("prompt_learner.") else k
            for k in want} == set(out["grads"])
    for k, v in want.items():
        kk = k if k.startswith("prompt_learner.") else "prompt_learner." + k
        assert _rel(out["grads"][kk], v) < 2e-3, (k, _rel(out["grads"][kk], v))


def test_vitl14_oracle_matches_reference_golden(golden_dir):
    """BASELINE config 3 geometry: all 24 layers of ViT-L/14 (257 tokens, width 1024, 16 heads,
    embed 768) through the reference's own classes vs the oracle in fp32."""
    from tests.golden.make_golden import BIG_CASES, load_grads
    cfg, n, c, seed = BIG_CASES["vitl14"]
    gold = np.load(os.path.join(golden_dir, "ref_vitl14.npz"))
    torch.set_num_threads(os.cpu_count() or 1)
    w = vo.synth_weights(cfg, seed)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 200)
    out = vo.online_step_oracle(images, labels, w, text, cfg,
                                logit_scale_exp=float(gold["logit_scale_exp"]),
                                dtype=torch.float32)
    assert _rel(out["probs"], gold["probs"]) < 2e-5
    assert abs(float(out["loss"]) - float(gold["loss"])) < 1e-5
    np.testing.assert_array_equal(out["pred"], gold["pred"])
    want = load_grads(gold)
    assert set(want) == set(out["grads"]) and len(want) == 96
    worst = max(_rel(out["grads"][k], v) for k, v in want.items())
    assert worst < 2e-3, worst      # fp16-packed fixture: 6e-4 rounding


def adapter_case_inputs(name):
    """Inputs of an ADAPTER_CASES golden (shared with the GPU parity test)."""
    from tests.golden.make_golden import ADAPTER_CASES
    cfg, tcfg, n, c, seed = ADAPTER_CASES[name]
    wv, wt = vo.synth_weights(cfg, seed), vo.synth_text_weights(tcfg, seed + 1)
    wa = vo.synth_adapter_weights(cfg.width, cfg.layers, "visual.transformer.resblocks.", seed + 2)
    wta = vo.synth_adapter_weights(tcfg.width, tcfg.layers, "transformer.resblocks.", seed + 3)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    tokens = vo.synth_tokens(c, tcfg, seed + 300)
    masks = vo.adapter_masks(seed + 400, cfg.layers, cfg.tokens, n)
    tmasks = vo.adapter_masks(seed + 500, tcfg.layers, tcfg.context, c)
    return cfg, tcfg, wv, wt, wa, wta, images, labels, tokens, masks, tmasks


@pytest.mark.parametrize("name", ["adapter_tiny", "adapter_vitb16"])
def test_adapter_oracle_matches_reference_golden(name, golden_dir):
    """--method adapter-clip (N4): the oracle's adapter blocks (model.py:418-442, adapter.py:11-73
    restated) in both towers against the reference's own classes, training mode with the same
    injected dropout masks, and the eval-mode forward."""
    from tests.golden.make_golden import load_grads
    cfg, tcfg, wv, wt, wa, wta, images, labels, tokens, masks, tmasks = adapter_case_inputs(name)
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    torch.set_num_threads(os.cpu_count() or 1)
    dtype = torch.float64 if name == "adapter_tiny" else torch.float32
    out = vo.adapter_step_oracle(
        images, labels, wv, wa, None, cfg, logit_scale_exp=float(gold["logit_scale_exp"]),
        dtype=dtype, masks=vo.masks_sample_major(masks), p=vo.ADAPTER_DROPOUT, tokens=tokens,
        wt_np=wt, wta_np=wta, tcfg=tcfg, tmasks=vo.masks_sample_major(tmasks))
    assert _rel(out["probs"], gold["probs"]) < 2e-5
    assert abs(float(out["loss"]) - float(gold["loss"])) < 1e-5
    np.testing.assert_array_equal(out["pred"], gold["pred"])
    want = load_grads(gold)
    assert set(want) <= set(out["grads"])
    assert len(out["grads"]) == 4 * (cfg.layers + tcfg.layers)
    worst = max(_rel(out["grads"][k], v) for k, v in want.items())
    assert worst < 2e-3, worst      # fp16-packed fixture: 6e-4 rounding
    ev = vo.adapter_step_oracle(images, labels, wv, wa, None, cfg,
                                logit_scale_exp=float(gold["logit_scale_exp"]), dtype=dtype,
                                tokens=tokens, wt_np=wt, wta_np=wta, tcfg=tcfg)
    assert _rel(ev["probs"], gold["probs_eval"]) < 2e-5


def text_only_case_inputs(name):
    """Inputs of a TEXT_ONLY_CASES golden (peft_encoder='text'; shared with the GPU test)."""
    from tests.golden.make_golden import TEXT_ONLY_CASES
    cfg, tcfg, n, c, seed, method = TEXT_ONLY_CASES[name]
    wv, wt = vo.synth_weights(cfg, seed), vo.synth_text_weights(tcfg, seed + 1)
    wta = vo.synth_adapter_weights(tcfg.width, tcfg.layers, "transformer.resblocks.", seed + 3)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    tokens = vo.synth_tokens(c, tcfg, seed + 300)
    tmasks = vo.adapter_masks(seed + 500, tcfg.layers, tcfg.context, c)
    return cfg, tcfg, method, wv, wt, wta, images, labels, tokens, tmasks


@pytest.mark.parametrize("name", ["textonly_lora_tiny", "textonly_adapter_tiny"])
def test_text_only_oracle_matches_reference_golden(name, golden_dir):
    """peft_encoder='text' (scripts/lora_clip.sh:10 / adapter_clip.sh:10 list it): PEFT blocks in
    the text tower only, the image tower vanilla - against the reference's own CLIP."""
    from tests.golden.make_golden import load_grads
    cfg, tcfg, method, wv, wt, wta, images, labels, tokens, tmasks = text_only_case_inputs(name)
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    ls = float(gold["logit_scale_exp"])
    if method == "lora":
        wv0 = {k: (np.zeros_like(v) if "lora" in k else v) for k, v in wv.items()}
        out = vo.clip_step_oracle(images, labels, tokens, wv0, wt, cfg, tcfg, logit_scale_exp=ls)
        out["grads"] = {k: v for k, v in out["grads"].items() if not k.startswith("visual.")}
    else:
        out = vo.adapter_step_oracle(images, labels, wv, None, None, cfg, logit_scale_exp=ls,
                                     p=vo.ADAPTER_DROPOUT, tokens=tokens, wt_np=wt, wta_np=wta,
                                     tcfg=tcfg, tmasks=vo.masks_sample_major(tmasks))
    assert _rel(out["probs"], gold["probs"]) < 2e-5
    assert abs(float(out["loss"]) - float(gold["loss"])) < 1e-5
    want = load_grads(gold)
    assert set(want) == set(out["grads"]) and len(want) == 4 * tcfg.layers
    assert max(_rel(out["grads"][k], v) for k, v in want.items()) < 2e-3
