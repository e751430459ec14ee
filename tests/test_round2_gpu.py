"""Round-2 GPU parity tests (all through the C-ABI):
  * the production shapes (B = 16 / 32 / 256, ViT-B/16) against goldens produced by the REFERENCE's
    own classes, on the autograd path and on the fused CUDA-graph trainer path;
  * the reference's AMP sequence (autocast + GradScaler, methods/adapter_clip.py:87-96);
  * boundary classes: LoRA Linear.forward, MultiheadAttention.forward, the vanilla
    ResidualAttentionBlock and attention();
  * evaluation tail: _interpret_pred / confusion matrix bit-exact against the reference's own
    function, prediction index space, tensor-core evaluation head at the C5 shape;
  * the text tower with LoRA (peft_encoder='both') against reference goldens;
  * the GPU input transform against torchvision.
"""
import os

import numpy as np
import pytest
import torch

from oracle import vit_oracle as vo
from tests.golden.make_golden import BIG_CASES, BOTH_CASES, synth_inputs
from tests.test_e2e_gpu import (COS, TOL, TOL_WORST_TENSOR, build_model, check_step, cos,
                                grads_by_name, rel)

pytestmark = pytest.mark.gpu


def _gold(golden_dir, name):
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    want = {"probs": gold["probs"], "loss": gold["loss"], "pred": gold["pred"],
            "grads": {k[5:]: gold[k] for k in gold.files if k.startswith("grad:")}}
    return gold, want


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["vitb16_b16", "vitb16_b32", "vitb16_b256"])
@pytest.mark.parametrize("path", ["autograd", "fused_graph"])
def test_production_shapes_match_reference_golden(name, path, golden_dir):
    """B = 16 (BASELINE C1), 32 (C2 per GPU: T = 6304 selects the CTA-pair GEMM, the fused LoRA
    passes and multi-pair attention) and 256 (the bench shape) against the reference's own
    modules. 'fused_graph' is the trainer's production path: llc_vit_forward_cls + analytic loss
    gradient, captured as a CUDA graph and REPLAYED (lr = 0 keeps the weights at the golden's)."""
    cfg, n, c, seed = BIG_CASES[name]
    gold, want = _gold(golden_dir, name)
    w = vo.synth_weights(cfg, seed)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 200)
    m = build_model(cfg, w)
    with torch.no_grad():
        m.model.logit_scale.fill_(float(np.log(gold["logit_scale_exp"])))
    names = [f"c{i}" for i in range(c)]
    m.set_text_features(names, torch.from_numpy(text))
    m.set_token(names)
    x, y = torch.from_numpy(images).cuda(), torch.from_numpy(labels).cuda()
    if path == "autograd":
        probs, _, _ = m(x)
        loss = torch.nn.CrossEntropyLoss()(probs, y)
        loss.backward()
        torch.cuda.synchronize()
        check_step(probs.detach().cpu().numpy(), loss.item(), probs.argmax(-1).cpu().numpy(),
                   grads_by_name(m), want, cfg)
        return
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    tr = LoRAClipTrainer(m, names, n_classes=c, lr=0.0, visible_classes="all",
                         use_cuda_graph=True)
    tr.online_before_task(0)
    for _ in range(3):           # capture, then two replays
        loss_sum, n_correct = tr.fused_step(x, y, n)
    assert tr._graph is not None and tr._graph_hits >= 3
    eng = m.model.visual.engine()
    assert eng._cls_only
    head = tr.last_head
    names_l = [k for k in w if "lora" in k]
    grads = {k: g.cpu().numpy() for k, g in zip(names_l, eng.lora_grad_views)}
    check_step(head.probs.cpu().numpy(), loss_sum, head.pred.cpu().numpy(), grads, want, cfg)
    assert n_correct == float((head.pred == y).sum())


def test_amp_autocast_gradscaler_sequence():
    """methods/adapter_clip.py:85-96 verbatim: zero_grad; with autocast: forward + criterion;
    topk; zero_grad; scaler.scale(loss).backward(); scaler.step(optimizer); scaler.update() -
    against the fp64 oracle followed by a plain AdamW step."""
    cfg = vo.VitCfg(image_size=64, patch=16, width=256, layers=2, heads=4, embed_dim=128)
    n, c, seed = 6, 10, 71
    w = vo.synth_weights(cfg, seed)
    images, labels = synth_inputs(cfg, n, c, seed + 1)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 2)
    want = vo.online_step_oracle(images, labels, w, text, cfg)
    m = build_model(cfg, w)
    names = [f"c{i}" for i in range(c)]
    m.set_text_features(names, torch.from_numpy(text))
    m.set_token(names)
    params = [p for p in m.parameters() if p.requires_grad]
    before = {k: p.detach().clone() for k, p in m.named_parameters() if p.requires_grad}
    optimizer = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-5)
    scaler = torch.cuda.amp.GradScaler()
    criterion = torch.nn.CrossEntropyLoss()
    x, y = torch.from_numpy(images).cuda(), torch.from_numpy(labels).cuda()
    optimizer.zero_grad()
    with torch.cuda.amp.autocast(enabled=True):
        logit, image_features, text_features = m(x)
        loss = criterion(logit, y)
    _, preds = logit.topk(1, 1, True, True)
    optimizer.zero_grad()
    scaler.scale(loss).backward()
    scale = scaler.get_scale()
    grads = {k[len("model."):]: (p.grad.detach().float() / scale).cpu().numpy()
             for k, p in m.named_parameters() if p.grad is not None}
    scaler.step(optimizer)
    scaler.update()
    torch.cuda.synchronize()
    assert scale == 65536.0 and scaler.get_scale() == 65536.0     # no inf/nan: step not skipped
    check_step(logit.detach().float().cpu().numpy(), loss.item(), preds[:, 0].cpu().numpy(),
               grads, want, cfg)
    # the step: AdamW on the oracle's gradients from the same start
    ref_p = {k: torch.from_numpy(w[k[len("model."):]]).double().requires_grad_(True)
             for k in before}
    ref_opt = torch.optim.AdamW(list(ref_p.values()), lr=1e-3, weight_decay=1e-5)
    for k, p in ref_p.items():
        p.grad = torch.from_numpy(want["grads"][k[len("model."):]]).double()
    ref_opt.step()
    for k, p in m.named_parameters():
        if p.requires_grad:
            moved = (p.detach().cpu().double() - before[k].cpu().double())
            moved_ref = ref_p[k].detach() - before[k].cpu().double()
            assert cos(moved, moved_ref) > 0.98, k     # Adam's first step is lr * sign(grad)


# ------------------------------------------------------------------------------------------------
def test_lora_linear_forward_backward():
    """lora.Linear.forward (lora.py:162-173): F.linear + (x A^T B^T) * scaling, with gradients to
    x, lora_A, lora_B and none to the frozen weight."""
    from lifelong_clip_b200.clip_modules import Linear
    torch.manual_seed(0)
    lin = Linear(256, 384, r=4, lora_alpha=1, merge_weights=False).cuda()
    with torch.no_grad():
        lin.lora_B.normal_(0, 0.05)       # zero at init: randomise so every path is exercised
        lin.lora_A.normal_(0, 0.05)
    x = torch.randn(7, 33, 256, device="cuda", requires_grad=True)
    dy = torch.randn(7, 33, 384, device="cuda")
    y = lin(x)
    y.backward(dy)
    xd = x.detach().double().requires_grad_(True)
    A, B = (lin.lora_A.detach().double().requires_grad_(True),
            lin.lora_B.detach().double().requires_grad_(True))
    yd = xd @ lin.weight.detach().double().T + lin.bias.detach().double() + \
        (xd @ A.T @ B.T) * lin.scaling
    yd.backward(dy.double())
    assert rel(y, yd) < 5e-3
    assert rel(x.grad, xd.grad) < TOL
    assert rel(lin.lora_A.grad, A.grad) < TOL and rel(lin.lora_B.grad, B.grad) < TOL
    assert lin.weight.grad is None and lin.bias.grad is None   # frozen: never computed


@pytest.mark.parametrize("causal", [False, True])
def test_multihead_attention_forward_is_the_reference_call(causal):
    """lora.MultiheadAttention.forward(x, x, x, need_weights=False, attn_mask) on [L, N, D]
    (model.py:226-231 -> lora.py:454-702, :732-1082) with autograd through x and the four LoRA
    tensors; and ResidualAttentionBlock_LoRA.attention(x) is that same call."""
    from lifelong_clip_b200.clip_modules import ResidualAttentionBlock_LoRA
    cfg = vo.VitCfg(image_size=32, patch=8, width=256, layers=1, heads=4, embed_dim=64)
    w = vo.synth_weights(cfg, 15)
    pre = "visual.transformer.resblocks.0."
    L, N = 19, 3
    mask = torch.full((L, L), float("-inf")).triu(1) if causal else None
    blk = ResidualAttentionBlock_LoRA(cfg.width, cfg.heads, mask, {"lora_alpha": 1, "lora_r": 4})
    blk.load_state_dict({k[len(pre):]: torch.from_numpy(v) for k, v in w.items()
                         if k.startswith(pre)})
    blk.cuda()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(L, N, cfg.width, generator=g)
    dy = torch.randn(L, N, cfg.width, generator=g)
    xc = x.cuda().requires_grad_(True)
    out, weights = blk.attn(xc, xc, xc, need_weights=False, attn_mask=mask)
    assert weights is None
    out.backward(dy.cuda())
    with torch.no_grad():
        out2 = blk.attention(x.cuda())
    assert torch.equal(out2, out.detach())
    wd = vo.to_torch(w, torch.float64)
    xd = x.double().transpose(0, 1).contiguous().requires_grad_(True)     # oracle is [N, L, D]
    D, s = cfg.width, cfg.lora_scale
    qkv = vo.lora_linear(xd, wd[pre + "attn.in_proj_weight"], wd[pre + "attn.in_proj_bias"],
                         wd[pre + "attn.in_proj_weight_lora_A"],
                         wd[pre + "attn.in_proj_weight_lora_B"], s)
    o = vo.attention_core(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], cfg.heads, causal)
    yd = vo.lora_linear(o, wd[pre + "attn.out_proj.weight"], wd[pre + "attn.out_proj.bias"],
                        wd[pre + "attn.out_proj.lora_A"], wd[pre + "attn.out_proj.lora_B"], s)
    yd.backward(dy.double().transpose(0, 1))
    assert rel(out.detach().cpu().transpose(0, 1), yd.detach()) < TOL
    assert rel(xc.grad.cpu().transpose(0, 1), xd.grad) < TOL
    for k, p in blk.attn.named_parameters():
        if "lora" in k:
            assert rel(p.grad.cpu(), wd[pre + "attn." + k].grad) < TOL, k
        else:
            assert p.grad is None, k
    with pytest.raises(NotImplementedError):
        blk.attn(xc, xc, xc, need_weights=True)


def test_vanilla_residual_attention_block():
    """ResidualAttentionBlock(d_model, n_head, attn_mask) (model.py:209-236): nn.MultiheadAttention
    state_dict keys, forward and attention() on [L, N, D], input gradient, no weight gradients."""
    from lifelong_clip_b200.clip_modules import ResidualAttentionBlock
    cfg = vo.VitCfg(image_size=32, patch=8, width=128, layers=1, heads=2, embed_dim=64)
    w = vo.synth_weights(cfg, 17)
    pre = "visual.transformer.resblocks.0."
    blk = ResidualAttentionBlock(cfg.width, cfg.heads, None)
    sd = {k[len(pre):]: torch.from_numpy(v) for k, v in w.items()
          if k.startswith(pre) and "lora" not in k}
    assert set(sd) == set(blk.state_dict().keys())
    blk.load_state_dict(sd)
    blk.cuda()
    L, N = 17, 4
    g = torch.Generator().manual_seed(3)
    x = torch.randn(L, N, cfg.width, generator=g)
    dy = torch.randn(L, N, cfg.width, generator=g)
    xc = x.cuda().requires_grad_(True)
    y = blk(xc)
    y.backward(dy.cuda())
    w0 = {k: (np.zeros_like(v) if "lora" in k else v) for k, v in w.items()}
    wd = vo.to_torch(w0, torch.float64, lora_grad=False)
    xd = x.double().transpose(0, 1).contiguous().requires_grad_(True)
    yd = vo.block_forward(xd, wd, pre, cfg)
    yd.backward(dy.double().transpose(0, 1))
    assert rel(y.detach().cpu().transpose(0, 1), yd.detach()) < 5e-3
    assert rel(xc.grad.cpu().transpose(0, 1), xd.grad) < TOL
    assert all(p.grad is None for p in blk.parameters())
    with torch.no_grad():
        a = blk.attention(blk.ln_1(x.cuda()))
    h = vo.layer_norm(x.double().transpose(0, 1), wd[pre + "ln_1.weight"], wd[pre + "ln_1.bias"])
    qkv = h @ wd[pre + "attn.in_proj_weight"].T + wd[pre + "attn.in_proj_bias"]
    D = cfg.width
    o = vo.attention_core(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], cfg.heads)
    want = o @ wd[pre + "attn.out_proj.weight"].T + wd[pre + "attn.out_proj.bias"]
    assert rel(a.cpu().transpose(0, 1), want) < TOL


# ------------------------------------------------------------------------------------------------
def test_interpret_pred_and_confusion_bit_exact(golden_dir):
    """llc_eval_accum against the outputs of the reference's own _interpret_pred
    (methods/_trainer.py:519-534) and sklearn's confusion_matrix (methods/adapter_clip.py:166)."""
    from lifelong_clip_b200 import ops
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    gold = np.load(os.path.join(golden_dir, "ref_interpret_pred.npz"))
    i = 0
    while f"y{i}" in gold.files:
        y = torch.from_numpy(gold[f"y{i}"]).cuda()
        pred = torch.from_numpy(gold[f"pred{i}"]).cuda()
        ncls, n_tasks = (int(v) for v in gold[f"meta{i}"])
        tr = object.__new__(LoRAClipTrainer)
        tr.n_tasks, tr.n_classes = n_tasks, ncls
        num, ok = tr._interpret_pred(y, pred)
        np.testing.assert_array_equal(num.numpy(), gold[f"num{i}"])
        np.testing.assert_array_equal(ok.numpy(), gold[f"ok{i}"])
        assert num.dtype == torch.float32 and tuple(num.shape) == (10,)
        cm = torch.zeros(ncls, ncls, dtype=torch.int64, device="cuda")
        counts = torch.zeros(22, dtype=torch.int64, device="cuda")
        half = y.numel() // 2          # accumulated over two batches, like online_evaluate
        ops.eval_accum(y[:half].contiguous(), pred[:half].contiguous(), n_tasks, ncls, cm, counts)
        ops.eval_accum(y[half:].contiguous(), pred[half:].contiguous(), n_tasks, ncls, cm, counts)
        cmh = cm.cpu()
        present = ((cmh.sum(0) + cmh.sum(1)) > 0).nonzero().flatten()
        np.testing.assert_array_equal(cmh[present][:, present].numpy(), gold[f"cm{i}"])
        i += 1
    assert i == 4
    tr.n_tasks = 2                      # label 40 // 2 = bin 20: the reference raises IndexError
    with pytest.raises(IndexError):
        tr._interpret_pred(torch.tensor([40], device="cuda"), torch.tensor([40], device="cuda"))


def test_online_evaluate_predicts_raw_class_ids():
    """classes exposed in non-identity order [7, 2, 9, 4]; after
    online_after_task the evaluation prediction must live in the RAW class-id space of the test
    labels (methods/adapter_clip.py:129-130 sets all_classnames[:_total_classes])."""
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    cfg, c = vo.VIT_TINY, 10
    w = vo.synth_weights(cfg, 3)
    m = build_model(cfg, w)
    names = [f"class{i}" for i in range(c)]
    text = vo.synth_text_features(c, cfg.embed_dim, 4)
    m.set_text_features(names, torch.from_numpy(text))
    tr = LoRAClipTrainer(m, names, n_classes=c, n_tasks=10, lr=0.0, visible_classes="all")
    tr.online_before_task(0)
    rng = np.random.default_rng(0)
    images = torch.from_numpy(rng.standard_normal((8, 3, 32, 32)).astype(np.float32))
    tr.online_step(images, torch.tensor([7, 2, 7, 9, 2, 4, 4, 7]), torch.arange(8))
    assert tr.exposed_classes == [7, 2, 9, 4]
    tr._total_classes = 10              # what the driver loop sets (methods/_trainer.py:322)
    tr.online_after_task(0)
    test_x = torch.from_numpy(rng.standard_normal((40, 3, 32, 32)).astype(np.float32))
    wd = vo.to_torch(w, torch.float64, lora_grad=False)
    feat = vo.vit_forward(test_x.double(), wd, cfg)
    probs, _, _ = vo.head_forward(feat, torch.from_numpy(text).double(), 1.0 / 0.07)
    srt = torch.sort(probs, dim=-1).values
    safe = (srt[:, -1] - srt[:, -2]) > 0.05 * srt[:, -1]
    test_x, pred_ref = test_x[safe], probs.argmax(-1)[safe]
    assert test_x.shape[0] >= 20
    # labels: the oracle's prediction for even samples (correct), a different class for odd ones
    y = pred_ref.clone()
    y[1::2] = (y[1::2] + 3) % c
    loader = [(test_x[:16], y[:16]), (test_x[16:], y[16:])]       # ragged second batch
    res = tr.online_evaluate(loader, 1000)
    n = test_x.shape[0]
    assert abs(float(res["avg_acc"]) - ((n + 1) // 2) / n) < 1e-6
    np.testing.assert_array_equal(np.asarray(res["confusion_matrix"]),
                                  vo.confusion(y.numpy(), pred_ref.numpy()))
    num, ok = vo.interpret_pred(y.numpy(), np.where(np.arange(n) % 2 == 0, y.numpy(), -1), 10)
    np.testing.assert_allclose(res["task_acc"], (ok / (num + 1e-5)).tolist(), rtol=1e-6)
    # an evaluation batch of another size must not invalidate the captured
    # training graph (separate evaluation arena), and the next training step is still correct
    l1, _ = tr.online_step(images, torch.tensor([7, 2, 7, 9, 2, 4, 4, 7]), torch.arange(8))
    l2, _ = tr.online_step(images, torch.tensor([7, 2, 7, 9, 2, 4, 4, 7]), torch.arange(8))
    assert np.isfinite(l1) and abs(l1 - l2) < 1e-6      # lr = 0: identical steps


def test_optimizer_survives_engine_rebuild():
    """model.to()/load_state_dict rebuild the engine; the optimizer must
    follow it instead of updating buffers no Parameter views any more."""
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    cfg, c = vo.VIT_TINY, 6
    m = build_model(cfg, vo.synth_weights(cfg, 3))
    names = [f"class{i}" for i in range(c)]
    m.set_text_features(names, torch.from_numpy(vo.synth_text_features(c, cfg.embed_dim, 4)))
    tr = LoRAClipTrainer(m, names, n_classes=c, lr=5e-3, visible_classes="all",
                         use_cuda_graph=False)
    tr.online_before_task(0)
    rng = np.random.default_rng(0)
    images = torch.from_numpy(rng.standard_normal((6, 3, 32, 32)).astype(np.float32))
    labels = torch.arange(6)
    tr.online_step(images, labels, labels)
    old = m.model.visual.engine()
    m.cuda()                                   # _apply drops the engine
    assert m.model.visual.engine() is not old
    p = next(p for k, p in m.named_parameters() if "lora_B" in k)
    before = p.detach().clone()
    losses = [tr.online_step(images, labels, labels)[0] for _ in range(8)]
    assert not torch.equal(p.detach(), before)         # the live parameters moved
    assert losses[-1] < losses[0]


def test_eval_head_tensor_core_c5_shape():
    """BASELINE config 5 shape of the head: 1000 cached class text embeddings, logits on the
    tcgen05 GEMM (VitEngine.eval_head), gather and additive-mask restrictions. Probabilities
    within tolerance, predictions bit-exact where the oracle's top-2 gap exceeds the bf16 floor."""
    cfg = vo.VitCfg(image_size=64, patch=16, width=768, layers=2, heads=12, embed_dim=512)
    n, c, seed = 300, 1000, 51
    w = vo.synth_weights(cfg, seed)
    images, _ = synth_inputs(cfg, n, c, seed + 1)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 2)
    m = build_model(cfg, w)
    eng = m.model.visual.engine()
    wd = vo.to_torch(w, torch.float64, lora_grad=False)
    feat = vo.vit_forward(torch.from_numpy(images).double(), wd, cfg)
    seen = np.sort(np.random.default_rng(seed).choice(c, size=300, replace=False))
    mask = np.full((c,), -np.inf); mask[seen] = 0.0
    tt = torch.from_numpy(text).cuda()
    with torch.no_grad():
        eng.forward(torch.from_numpy(images).cuda(), training=False)
        for idx, msk in ((None, None), (torch.from_numpy(seen), None),
                         (None, torch.from_numpy(mask))):
            h = eng.eval_head(tt, 1.0 / 0.07, cls_idx=None if idx is None else idx.cuda(),
                              add_mask=None if msk is None else msk.float().cuda())
            torch.cuda.synchronize()
            probs, _, _ = vo.head_forward(feat, torch.from_numpy(text).double(), 1.0 / 0.07,
                                          idx, msk)
            assert rel(h.probs, probs) < TOL
            srt = torch.sort(probs, dim=-1).values
            safe = ((srt[:, -1] - srt[:, -2]) > 0.05 * srt[:, -1]).numpy()
            assert safe.sum() > n // 2
            np.testing.assert_array_equal(h.pred.cpu().numpy()[safe],
                                          probs.argmax(-1).numpy()[safe])
            # same decisions as the fp32 CUDA-core head of the training path
            h32 = eng.head(tt, 1.0 / 0.07, cls_idx=None if idx is None else idx.cuda(),
                           add_mask=None if msk is None else msk.float().cuda())
            assert rel(h.probs, h32.probs) < TOL
            np.testing.assert_array_equal(h.pred.cpu().numpy()[safe], h32.pred.cpu().numpy()[safe])


# ------------------------------------------------------------------------------------------------
def build_both(cfg, tcfg, wv, wt):
    from lifelong_clip_b200.adapter_clip import AdapterCLIP
    m = AdapterCLIP(peft_encoder="both",
                    vision_config=(cfg.image_size, cfg.patch, cfg.width, cfg.layers,
                                   cfg.embed_dim),
                    text_config=(tcfg.context, tcfg.vocab, tcfg.width, tcfg.heads, tcfg.layers))
    sd = {k: torch.from_numpy(v) for k, v in {**wv, **wt}.items()}
    missing, unexpected = m.model.load_state_dict(sd, strict=False)
    assert not unexpected and set(missing) <= {"logit_scale"}, (missing, unexpected)
    m.cuda()
    for k, p in m.named_parameters():
        if "adaptmlp" not in k and "lora" not in k:
            p.requires_grad = False
    return m


# With BOTH towers in bf16 the logit gradient only sees the DIFFERENCES between the class text
# features (softmax gradients sum to zero), and random-init prompts give features with pairwise
# cosine ~0.77: relative errors are amplified ~2x compared with the cached-text case. Calibration
# (tools/parity_both_calib.py -> profiles/r02_parity_both_vs_autocast.txt): PyTorch's OWN bf16
# autocast of the same two-tower step measures, against the reference's fp32 output,
#   both_vitb16: flat 1.4e-2 / 1.5e-2 (image / text tower), median tensor 1.5e-2 / 1.8e-2, worst
#                tensor 3.4e-2 / 4.3e-2 (the same out_proj.lora_A tensors that are worst here);
#   both_tiny:   flat 2.7e-2 / 3.0e-2, worst 4.0e-2.
# This path measures 1.2e-2 flat / 3.5e-2 worst on both_vitb16. The bounds below are autocast's
# level, not a looser one.
TOL_BOTH = {"both_vitb16": (2e-2, 5e-2), "both_tiny": (3.5e-2, 6e-2)}


def check_both(probs, loss, pred, grads, gold, n_layers, name):
    tol_flat, tol_worst = TOL_BOTH[name]
    assert rel(probs, gold["probs"]) < TOL
    assert abs(float(loss) - float(gold["loss"])) < TOL * abs(float(gold["loss"]))
    p = np.asarray(gold["probs"], np.float64)
    srt = np.sort(p, axis=-1)
    safe = (srt[:, -1] - srt[:, -2]) > 0.05 * srt[:, -1]
    np.testing.assert_array_equal(np.asarray(pred)[safe], np.asarray(gold["pred"])[safe])
    want = {k[5:]: gold[k] for k in gold.files if k.startswith("grad:")}
    assert set(grads) == set(want) and len(want) == 4 * n_layers
    for tower in ("visual.", "transformer."):
        keys = sorted(k for k in want if k.startswith(tower))
        fg = np.concatenate([np.asarray(grads[k], np.float64).ravel() for k in keys])
        fw = np.concatenate([np.asarray(want[k], np.float64).ravel() for k in keys])
        assert rel(fg, fw) < tol_flat, (tower, rel(fg, fw))
        assert cos(fg, fw) > 1.0 - tol_flat ** 2
        rels = [rel(grads[k], want[k]) for k in keys]
        assert float(np.median(rels)) < tol_flat
        assert max(rels) < tol_worst, (max(rels), keys[int(np.argmax(rels))])


@pytest.mark.parametrize("name", ["both_tiny", "both_vitb16"])
@pytest.mark.parametrize("path", ["autograd", "fused"])
def test_text_tower_lora_matches_reference_golden(name, path, golden_dir):
    """peft_encoder='both' (scripts/lora_clip.sh:10): text features recomputed by the LoRA text
    tower (model.py:941-956, causal mask :926-932), gradients to the 48 + 48 LoRA tensors of both
    towers, against the reference's own CLIP."""
    cfg, tcfg, n, c, seed = BOTH_CASES[name]
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    wv, wt = vo.synth_weights(cfg, seed), vo.synth_text_weights(tcfg, seed + 1)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    tokens = vo.synth_tokens(c, tcfg, seed + 300)
    m = build_both(cfg, tcfg, wv, wt)
    with torch.no_grad():
        m.model.logit_scale.fill_(float(np.log(gold["logit_scale_exp"])))
    names = [f"c{i}" for i in range(c)]
    table = {m.prompt_template.format(nm): torch.from_numpy(tokens[i]) for i, nm in enumerate(names)}
    m.set_tokenizer(lambda texts: torch.stack([table[t] for t in texts]))
    assert torch.equal(m.labels_tokenize(names).cpu(), torch.from_numpy(tokens))
    m.set_token(names)
    x, y = torch.from_numpy(images).cuda(), torch.from_numpy(labels).cuda()
    with torch.no_grad():
        tf = m.model.encode_text(torch.from_numpy(tokens).cuda())
    assert rel(tf.cpu(), gold["tfeat"]) < TOL
    if path == "autograd":
        probs, fi, ft = m(x)
        loss = torch.nn.CrossEntropyLoss()(probs, y)
        loss.backward()
        torch.cuda.synchronize()
        assert tuple(ft.shape) == (c, cfg.embed_dim)
        check_both(probs.detach().cpu().numpy(), loss.item(), probs.argmax(-1).cpu().numpy(),
                   grads_by_name(m), gold, cfg.layers + tcfg.layers, name)
        return
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    tr = LoRAClipTrainer(m, names, n_classes=c, lr=0.0, visible_classes="all",
                         use_cuda_graph=True)
    tr.online_before_task(0)
    for _ in range(3):
        loss_sum, _ = tr.fused_step(x, y, n)
    assert tr._graph is not None
    veng, teng = m.model.visual.engine(), m.model.text_engine()
    grads = {k: g.cpu().numpy() for k, g in zip([k for k in wv if "lora" in k],
                                                veng.lora_grad_views)}
    grads.update({k: g.cpu().numpy() for k, g in zip([k for k in wt if "lora" in k],
                                                     teng.lora_grad_views)})
    head = tr.last_head
    check_both(head.probs.cpu().numpy(), loss_sum, head.pred.cpu().numpy(), grads, gold,
               cfg.layers + tcfg.layers, name)


def test_text_tower_trains_and_evaluates():
    """online_step / online_evaluate with both towers trainable: the loss goes down, text LoRA
    parameters move, evaluation runs on the text features of the evaluated class list."""
    from lifelong_clip_b200.adapter_clip import SyntheticTokenizer
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    cfg, tcfg, c = vo.VIT_TINY, vo.TEXT_TINY, 6
    m = build_both(cfg, tcfg, vo.synth_weights(cfg, 3), vo.synth_text_weights(tcfg, 4))
    m.set_tokenizer(SyntheticTokenizer(tcfg.context, tcfg.vocab))
    names = [f"class{i}" for i in range(c)]
    tr = LoRAClipTrainer(m, names, n_classes=c, n_tasks=10, lr=5e-3, visible_classes="all")
    tr.online_before_task(0)
    rng = np.random.default_rng(0)
    images = torch.from_numpy(rng.standard_normal((8, 3, 32, 32)).astype(np.float32))
    labels = torch.tensor([0, 1, 2, 3, 4, 5, 0, 1])
    pt = next(p for k, p in m.named_parameters() if k.startswith("model.transformer") and "lora_B" in k)
    before = pt.detach().clone()
    losses = [tr.online_step(images, labels, torch.arange(8))[0] for _ in range(12)]
    assert losses[-1] < losses[0] - 1e-4 and not torch.equal(pt.detach(), before)
    tr._total_classes = c
    tr.online_after_task(0)
    res = tr.online_evaluate([(images, labels)], 1000)
    assert 0.0 <= float(res["avg_acc"]) <= 1.0


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("src_dtype", ["uint8", "float32"])
@pytest.mark.parametrize("mode", ["train", "test"])
def test_gpu_transform_matches_torchvision(src_dtype, mode):
    """methods/_trainer.py:236-247: Resize((224,224)) -> RandomCrop(224, padding=4) ->
    RandomHorizontalFlip -> Normalize on a CIFAR-shaped batch; seeded, so torchvision on the CPU
    and GpuTransform consume the same draws. fp32 output within 2e-6 (bilinear weights are
    computed in fp32 on both sides; summation order of the four taps may differ by an ulp); the
    bf16 patch rows are the rounded fp32 values."""
    from torchvision import transforms
    from lifelong_clip_b200 import ops
    from lifelong_clip_b200.transform import GpuTransform
    mean, std = (0.5071, 0.4867, 0.4408), (0.2675, 0.2565, 0.2761)
    g = torch.Generator().manual_seed(0)
    raw_u8 = torch.randint(0, 256, (5, 3, 32, 32), generator=g, dtype=torch.uint8)
    host = raw_u8.float().div(255)                     # what ToTensor hands the reference
    if mode == "train":
        ref_tf = transforms.Compose([transforms.Resize((224, 224)),
                                     transforms.RandomCrop(224, padding=4),
                                     transforms.RandomHorizontalFlip(),
                                     transforms.Normalize(mean, std)])
        ours = GpuTransform.train(224, mean, std)
    else:
        ref_tf = transforms.Compose([transforms.Resize((224, 224)),
                                     transforms.Normalize(mean, std)])
        ours = GpuTransform.test(224, mean, std)
    src = raw_u8.cuda() if src_dtype == "uint8" else host.cuda()
    draws = set()
    for trial in range(6):
        torch.manual_seed(100 + trial)
        want = ref_tf(host)
        torch.manual_seed(100 + trial)
        got = ours(src)
        draws.add(ours.last_draw)
        assert got.dtype == torch.float32 and tuple(got.shape) == (5, 3, 224, 224)
        assert float((got.cpu() - want).abs().max()) < 2e-6 * float(want.abs().max())
        # fused producer: the same transform straight into the patch rows
        torch.manual_seed(100 + trial)
        tx = ours.struct(src)
        patches = torch.empty(5 * 14 * 14, 768, dtype=torch.bfloat16, device="cuda")
        ops.transform_patchify(tx, 5, 16, patches)
        direct = torch.empty(5 * 14 * 14, 768, dtype=torch.bfloat16, device="cuda")
        ops.patchify(got, 16, direct)
        assert torch.equal(patches, direct)
    if mode == "train":
        assert len(draws) > 2            # crops / flips actually varied


def test_trainer_with_gpu_transform_and_graph():
    """online_step fed RAW uint8 batches with GpuTransform as train_transform: the transform runs
    inside the captured step (llc_vit_forward_tx), fresh crop/flip draws reach every replay, and
    the result equals transforming first and feeding the fp32 images."""
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    from lifelong_clip_b200.transform import GpuTransform
    cfg = vo.VitCfg(image_size=64, patch=16, width=128, layers=2, heads=2, embed_dim=64)
    c = 6
    mean, std = (0.5, 0.5, 0.5), (0.25, 0.25, 0.25)
    names = [f"class{i}" for i in range(c)]
    text = torch.from_numpy(vo.synth_text_features(c, cfg.embed_dim, 4))
    g = torch.Generator().manual_seed(0)
    raw = torch.randint(0, 256, (6, 3, 16, 16), generator=g, dtype=torch.uint8)
    labels = torch.arange(6)

    def run(fused):
        m = build_model(cfg, vo.synth_weights(cfg, 3))
        m.set_text_features(names, text)
        tf = GpuTransform.train(64, mean, std)
        tr = LoRAClipTrainer(m, names, n_classes=c, lr=1e-3, visible_classes="all",
                             train_transform=tf if fused else (lambda x: tf(x)),
                             use_cuda_graph=True)
        tr.online_before_task(0)
        torch.manual_seed(7)
        out = [tr.online_step(raw.cuda(), labels, labels)[0] for _ in range(5)]
        return out, tr

    a, tr_a = run(True)
    b, _ = run(False)
    assert tr_a._graph is not None and tr_a._graph_hits >= 4
    np.testing.assert_allclose(a, b, rtol=1e-5)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["maple_small", "maple_vitb16"])
def test_maple_matches_reference_golden(name, golden_dir):
    """BASELINE config 4: MaPLe deep multi-modal prompts (models/maple.py:74-253 over
    models/maple_clip/model.py) on frozen towers: logits, CE-on-logits loss and the gradients of
    the nine prompt_learner tensors against the reference's own classes. Attention runs over
    197 + 3 image tokens and 16 / 77 text tokens; the activation gradient flows through every
    frozen block down to the prompt rows."""
    from lifelong_clip_b200.maple import MaPLe
    from tests.golden.make_golden import MAPLE_CASES, load_maple_grads
    cfg, tcfg, n, c, seed = MAPLE_CASES[name]
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    wv = vo.strip_lora(vo.synth_weights(cfg, seed))
    wt = vo.strip_lora(vo.synth_text_weights(tcfg, seed + 1))
    wp = vo.synth_maple_weights(tcfg, cfg.width, seed=seed + 2)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    tokens = torch.from_numpy(vo.synth_tokens(c, tcfg, seed + 300))
    m = MaPLe(vision_config=(cfg.image_size, cfg.patch, cfg.width, cfg.layers, cfg.embed_dim),
              text_config=(tcfg.context, tcfg.vocab, tcfg.width, tcfg.heads, tcfg.layers))
    missing, unexpected = m.base_clip_model.load_state_dict(
        {k: torch.from_numpy(v) for k, v in {**wv, **wt}.items()}, strict=False)
    assert not unexpected and set(missing) <= {"logit_scale"}, (missing, unexpected)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in wp.items()}, strict=False)
    m.cuda()
    with torch.no_grad():
        m.logit_scale.fill_(float(np.log(gold["logit_scale_exp"])))
    for k, p in m.named_parameters():          # methods/maple.py: only the prompt learner trains
        p.requires_grad = "prompt_learner" in k
    assert sum(p.requires_grad for p in m.parameters()) == 9
    tok = tokens.cuda()
    with torch.no_grad():
        emb = m.base_clip_model.token_embedding(tok)
    prefix, suffix = emb[:, :1, :], emb[:, 1 + 3:, :]
    logits = m(torch.from_numpy(images).cuda(), tok, prefix, suffix)
    loss = torch.nn.CrossEntropyLoss()(logits, torch.from_numpy(labels).cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert rel(logits.detach().cpu(), gold["logits"]) < TOL
    assert abs(loss.item() - float(gold["loss"])) < TOL * abs(float(gold["loss"]))
    lg = np.sort(gold["logits"].astype(np.float64), axis=-1)
    safe = (lg[:, -1] - lg[:, -2]) > 0.05
    np.testing.assert_array_equal(logits.argmax(-1).cpu().numpy()[safe], gold["pred"][safe])
    want = load_maple_grads(gold)
    got = {k: p.grad.detach().cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(want)
    # prompt gradients pass through EVERY frozen block in bf16 (12 layers back to the prompt
    # rows): same two-sided error amplification as the two-tower LoRA case above
    rels = {k: rel(got[k], want[k]) for k in want}
    assert float(np.median(list(rels.values()))) < 2e-2, rels
    assert max(rels.values()) < 5e-2, rels
    # frozen towers: no gradient was formed for any backbone tensor
    assert all(p.grad is None for k, p in m.named_parameters() if "prompt_learner" not in k)


def test_maple_prompt_training_step():
    """A few AdamW steps on the prompt learner only (methods/maple.py:85-105 flow): update_class_names
    -> tokenized prompts, forward(image) -> logits, CE, backward, step; the loss goes down and the
    frozen towers stay bit-identical."""
    from lifelong_clip_b200.adapter_clip import SyntheticTokenizer
    from lifelong_clip_b200.maple import MaPLe
    cfg = vo.VitCfg(image_size=64, patch=16, width=256, layers=3, heads=4, embed_dim=128)
    m = MaPLe(vision_config=(cfg.image_size, cfg.patch, cfg.width, cfg.layers, cfg.embed_dim),
              text_config=(16, 300, 128, 2, 3)).cuda()
    m.set_tokenizer(SyntheticTokenizer(16, 300))
    for k, p in m.named_parameters():
        p.requires_grad = "prompt_learner" in k
    frozen = {k: p.detach().clone() for k, p in m.named_parameters() if "prompt_learner" not in k}
    names = [f"class{i}" for i in range(5)]
    assert m.update_class_names(names).shape == (5, 16)
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=2e-2)
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((6, 3, 64, 64)).astype(np.float32)).cuda()
    y = torch.tensor([0, 1, 2, 3, 4, 0]).cuda()
    losses = []
    for _ in range(10):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(m(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0] - 1e-3, losses
    for k, p in m.named_parameters():
        if "prompt_learner" not in k:
            assert torch.equal(p.detach(), frozen[k]), k


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,L,H,seq_first", [
    (20, 257, 16, False),      # ViT-L/14: more pairs than SMs
    (3, 257, 16, True),        # the block module's [L, N, D] layout
    (2, 258, 4, False), (2, 264, 2, False),     # 2 and 8 side tokens
    (2, 265, 2, False)])       # past the side-kernel range: legacy mma.sync kernels
def test_attention_longer_than_256_tokens(N, L, H, seq_first):
    """256 < L <= 264 (ViT-L/14's 257 tokens, BASELINE config 3): the TMEM kernels on the first
    256 tokens + the log-sum-exp merge / side-term kernels of attention_long.cu, against the
    oracle's attention core in fp64."""
    from lifelong_clip_b200 import ops
    from tests.test_kernels_gpu import _attn_case
    _attn_case(ops, N, L, H, False, seq_first, seed=300 + L)


def test_vitl14_full_depth_matches_reference_golden(golden_dir):
    """BASELINE config 3: all 24 layers of ViT-L/14 on the production path (257-token attention on
    the TMEM kernels + side kernels, class-token-only last block, analytic loss gradient) against
    the reference's own modules."""
    from tests.golden.make_golden import load_grads
    cfg, n, c, seed = BIG_CASES["vitl14"]
    gold = np.load(os.path.join(golden_dir, "ref_vitl14.npz"))
    want = {"probs": gold["probs"], "loss": gold["loss"], "pred": gold["pred"],
            "grads": load_grads(gold)}
    w = vo.synth_weights(cfg, seed)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 200)
    m = build_model(cfg, w)
    eng = m.model.visual.engine()
    eng.forward(torch.from_numpy(images).cuda(), training=True)
    head = eng.head(torch.from_numpy(text).cuda(), float(gold["logit_scale_exp"]),
                    labels=torch.from_numpy(labels).cuda())
    eng.backward_from_head(head)
    torch.cuda.synchronize()
    names = [k for k in w if "lora" in k]
    grads = {k: g.cpu().numpy() for k, g in zip(names, eng.lora_grad_views)}
    # 3 images through 24 layers: little averaging (cf. test_single_image_batches)
    check_step(head.probs.cpu().numpy(), float(head.loss_rows.sum()), head.pred.cpu().numpy(),
               grads, want, cfg, tol=2e-2)


def test_online_evaluate_with_gpu_test_transform():
    """online_evaluate fed RAW uint8 batches with GpuTransform.test as test_transform (Resize +
    Normalize fused into the tower's first kernel) equals evaluating the pre-transformed images."""
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    from lifelong_clip_b200.transform import GpuTransform
    cfg = vo.VitCfg(image_size=64, patch=16, width=128, layers=2, heads=2, embed_dim=64)
    c = 10
    mean, std = (0.5, 0.45, 0.4), (0.25, 0.2, 0.3)
    m = build_model(cfg, vo.synth_weights(cfg, 3))
    names = [f"class{i}" for i in range(c)]
    m.set_text_features(names, torch.from_numpy(vo.synth_text_features(c, cfg.embed_dim, 4)))
    g = torch.Generator().manual_seed(0)
    raw = torch.randint(0, 256, (37, 3, 16, 16), generator=g, dtype=torch.uint8)
    y = torch.randint(0, c, (37,), generator=g)
    tf = GpuTransform.test(64, mean, std)
    tr = LoRAClipTrainer(m, names, n_classes=c, n_tasks=10, visible_classes="all",
                         test_transform=tf)
    tr._total_classes = c
    tr.online_after_task(0)
    fused = tr.online_evaluate([(raw[:20], y[:20]), (raw[20:], y[20:])], 0)
    tr.test_transform = lambda x: x
    pre = tf(raw.cuda())
    plain = tr.online_evaluate([(pre[:20], y[:20]), (pre[20:], y[20:])], 0)
    assert float(fused["avg_acc"]) == float(plain["avg_acc"])
    assert fused["confusion_matrix"] == plain["confusion_matrix"]
    assert fused["task_acc"] == plain["task_acc"]


def test_maple_graphed_step_equals_eager():
    """maple.GraphedStep: forward + CE + backward replayed from one CUDA graph gives the eager
    gradients bit for bit, on new inputs too."""
    from lifelong_clip_b200.adapter_clip import SyntheticTokenizer
    from lifelong_clip_b200.maple import GraphedStep, MaPLe
    cfg = vo.VitCfg(image_size=64, patch=16, width=256, layers=3, heads=4, embed_dim=128)
    torch.manual_seed(1)
    m = MaPLe(vision_config=(cfg.image_size, cfg.patch, cfg.width, cfg.layers, cfg.embed_dim),
              text_config=(16, 300, 128, 2, 3)).cuda()
    m.set_tokenizer(SyntheticTokenizer(16, 300))
    for k, p in m.named_parameters():
        p.requires_grad = "prompt_learner" in k
    m.update_class_names([f"class{i}" for i in range(5)])
    xs = [torch.randn(6, 3, 64, 64, device="cuda") for _ in range(2)]
    ys = [torch.randint(0, 5, (6,), device="cuda") for _ in range(2)]
    step = GraphedStep(m, xs[0], ys[0])
    params = [p for p in m.parameters() if p.requires_grad]
    for x, y in zip(xs[::-1], ys[::-1]):
        loss_g = float(step(x, y))
        got = [p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        loss = torch.nn.functional.cross_entropy(m(x), y)
        loss.backward()
        assert abs(loss_g - float(loss)) < 1e-6 * abs(float(loss))
        for a, p in zip(got, params):
            assert torch.equal(a, p.grad)
