"""End-to-end parity of the CUDA path (through the C-ABI) against
  (a) the golden vectors produced by the REFERENCE's own classes (tests/golden/*.npz), and
  (b) the oracle restatement evaluated in fp64 on the same seeded inputs.
Tolerances (north star): logits / loss / LoRA gradients rel-L2 <= 1e-2 and gradient cosine
>= 0.9999 for bf16 tensor-core operands with fp32 accumulation; integer work bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import vit_oracle as vo
from tests.golden.make_golden import CASES, synth_inputs

pytestmark = pytest.mark.gpu

TOL = 1e-2          # tensor rel-L2, stated by BASELINE.json north_star ("about 1e-2 relative")
COS = 0.9999
# Calibration measured on B200 (tools/parity_e2e_diag.py, gpurun_out/diag6.log): on the ViT-B/16
# golden case PyTorch's OWN bf16 autocast of the same math sits at probs 3.8e-3 / per-tensor LoRA
# grads 6e-3..1.5e-2 against the fp64 oracle; this path measures 1.2e-3 / 3.4e-3..9e-3, with ONE
# small-magnitude tensor (a near-cancelling out_proj.lora_B gradient) at 1.8e-2 when the batch is
# only 2 images (autocast: 1.5e-2 on the same tensor). Hence: the flat LoRA-gradient buffer (what
# the optimizer and the all-reduce see) and the median tensor must meet 1e-2; a single tensor
# may reach 2e-2.
TOL_WORST_TENSOR = 2e-2


def _t(a):
    return (a.detach().cpu() if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))).double()


def rel(a, b):
    a = _t(a).flatten()
    b = _t(b).flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cos(a, b):
    a = _t(a).flatten()
    b = _t(b).flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def build_model(cfg, w_np):
    """AdapterCLIP mirror with the oracle's synthetic weights loaded by reference key name."""
    from lifelong_clip_b200.adapter_clip import AdapterCLIP
    m = AdapterCLIP(vision_config=(cfg.image_size, cfg.patch, cfg.width, cfg.layers,
                                   cfg.embed_dim))
    sd = {k: torch.from_numpy(v) for k, v in w_np.items()}
    missing, unexpected = m.model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert set(missing) <= {"logit_scale"}, missing
    m.cuda()
    for k, p in m.named_parameters():  # methods/adapter_clip.py:117-119
        if "adaptmlp" not in k and "lora" not in k:
            p.requires_grad = False
    return m


def grads_by_name(m):
    return {k[len("model."):]: p.grad.detach().float().cpu().numpy()
            for k, p in m.named_parameters() if p.grad is not None}


def check_step(out_probs, out_loss, out_pred, grads, want, cfg, tol=TOL):
    assert rel(out_probs, want["probs"]) < tol
    assert abs(float(out_loss) - float(want["loss"])) < tol * abs(float(want["loss"]))
    # argmax is bit-exact wherever the reference's own top-2 gap exceeds the bf16 noise floor
    p = np.asarray(want["probs"], np.float64)
    srt = np.sort(p, axis=-1)
    safe = (srt[:, -1] - srt[:, -2]) > 0.05 * srt[:, -1]
    np.testing.assert_array_equal(np.asarray(out_pred)[safe], np.asarray(want["pred"])[safe])
    assert len(grads) == 4 * cfg.layers
    keys = sorted(want["grads"])
    rels = [rel(grads[k], want["grads"][k]) for k in keys]
    flat_got = np.concatenate([np.asarray(grads[k], np.float64).ravel() for k in keys])
    flat_want = np.concatenate([np.asarray(want["grads"][k], np.float64).ravel() for k in keys])
    assert rel(flat_got, flat_want) < tol, rel(flat_got, flat_want)
    assert cos(flat_got, flat_want) > COS
    assert float(np.median(rels)) < tol, float(np.median(rels))
    assert max(rels) < TOL_WORST_TENSOR, (max(rels), keys[int(np.argmax(rels))])


@pytest.mark.parametrize("name", ["tiny", "vitb16"])
def test_online_step_matches_reference_golden(name, golden_dir):
    """AdapterCLIP.forward -> CE on probs -> backward (torch autograd over the fused kernels)
    against the outputs of the reference's own classes."""
    cfg, n, c, seed = CASES[name]
    gold = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    w = vo.synth_weights(cfg, seed)
    images, labels = synth_inputs(cfg, n, c, seed + 100)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 200)
    m = build_model(cfg, w)
    with torch.no_grad():
        m.model.logit_scale.fill_(float(np.log(gold["logit_scale_exp"])))
    names = [f"c{i}" for i in range(c)]
    m.set_text_features(names, torch.from_numpy(text))
    m.set_token(names)
    probs, feats, tfeats = m(torch.from_numpy(images).cuda())
    loss = torch.nn.CrossEntropyLoss()(probs, torch.from_numpy(labels).cuda())  # reference loss
    loss.backward()
    torch.cuda.synchronize()
    want = {"probs": gold["probs"], "loss": gold["loss"], "pred": gold["pred"],
            "grads": {k[5:]: gold[k] for k in gold.files if k.startswith("grad:")}}
    check_step(probs.detach().cpu().numpy(), loss.item(), probs.argmax(-1).cpu().numpy(),
               grads_by_name(m), want, cfg)
    assert tuple(feats.shape) == (n, cfg.embed_dim) and tuple(tfeats.shape) == (c, cfg.embed_dim)
    assert rel(feats.norm(dim=-1).cpu(), torch.ones(n)) < 1e-5


@pytest.mark.parametrize("cfg,n,c,gather", [
    (vo.VIT_TINY, 5, 12, True),
    (vo.VitCfg(image_size=28, patch=14, width=256, layers=3, heads=4, embed_dim=96), 4, 9, False),
    (vo.VitCfg(image_size=64, patch=16, width=768, layers=2, heads=12, embed_dim=512), 7, 33,
     True),
])
def test_fused_trainer_step_matches_fp64_oracle(cfg, n, c, gather):
    """The fully fused path (engine forward -> head with analytic loss gradient -> backward),
    class restriction by gather, against the fp64 oracle."""
    seed = 31
    w = vo.synth_weights(cfg, seed)
    images, _ = synth_inputs(cfg, n, c, seed + 1)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 2)
    rng = np.random.default_rng(seed + 3)
    cls_idx = np.sort(rng.choice(c, size=max(2, c // 2), replace=False)) if gather else None
    cv = len(cls_idx) if gather else c
    labels = rng.integers(0, cv, size=(n,)).astype(np.int64)
    want = vo.online_step_oracle(images, labels, w, text, cfg, cls_idx=cls_idx)
    m = build_model(cfg, w)
    eng = m.model.visual.engine()
    eng.forward(torch.from_numpy(images).cuda(), training=True)
    idx = torch.from_numpy(cls_idx).cuda() if gather else None
    head = eng.head(torch.from_numpy(text).cuda(), 1.0 / 0.07, cls_idx=idx,
                    labels=torch.from_numpy(labels).cuda())
    eng.backward_from_head(head)
    torch.cuda.synchronize()
    names = [k for k in w if "lora" in k]
    grads = {k: g.cpu().numpy() for k, g in zip(names, eng.lora_grad_views)}
    # order of lora_grad_views == order of named lora parameters (block-major, A_in B_in A_o B_o)
    check_step(head.probs.cpu().numpy(), float(head.loss_rows.sum()), head.pred.cpu().numpy(),
               grads, want, cfg)
    assert rel(head.feat.cpu(), want["feat"]) < TOL


@pytest.mark.parametrize("cfg,n", [
    (vo.VitCfg(image_size=64, patch=16, width=768, layers=2, heads=12, embed_dim=512), 9),
    (vo.VitCfg(image_size=28, patch=14, width=128, layers=1, heads=2, embed_dim=64), 3)])
def test_class_token_only_last_block_equals_full(cfg, n, monkeypatch):
    """llc_vit_forward_cls / llc_vit_backward_cls (the last block computed for the class-token rows
    only) against the full-size last block: same features, probabilities and LoRA gradients (the
    two paths differ only in the rounding of the last block's attention probabilities, which
    stay fp32 on the one-query path), and both within the parity policy of the fp64 oracle."""
    c, seed = 10, 77
    w = vo.synth_weights(cfg, seed)
    images, _ = synth_inputs(cfg, n, c, seed + 1)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 2)
    labels = np.random.default_rng(seed + 3).integers(0, c, size=(n,)).astype(np.int64)
    want = vo.online_step_oracle(images, labels, w, text, cfg)
    names = [k for k in w if "lora" in k]

    def run(full):
        if full:
            monkeypatch.setenv("LLC_FULL_LAST_BLOCK", "1")
        else:
            monkeypatch.delenv("LLC_FULL_LAST_BLOCK", raising=False)
        m = build_model(cfg, w)
        eng = m.model.visual.engine()
        eng.forward(torch.from_numpy(images).cuda(), training=True)
        head = eng.head(torch.from_numpy(text).cuda(), 1.0 / 0.07,
                        labels=torch.from_numpy(labels).cuda())
        eng.backward_from_head(head)
        torch.cuda.synchronize()
        assert eng._cls_only == (not full)
        grads = {k: g.cpu().numpy().copy() for k, g in zip(names, eng.lora_grad_views)}
        return head.feat.cpu().numpy().copy(), head.probs.cpu().numpy().copy(), grads, head

    f_cls, p_cls, g_cls, h_cls = run(False)
    f_full, p_full, g_full, _ = run(True)
    assert rel(f_cls, f_full) < 3e-3
    assert rel(p_cls, p_full) < 3e-3
    flat = lambda g: np.concatenate([np.asarray(g[k], np.float64).ravel() for k in names])
    assert rel(flat(g_cls), flat(g_full)) < 5e-3
    check_step(p_cls, float(h_cls.loss_rows.sum()), h_cls.pred.cpu().numpy(), g_cls, want, cfg)


@pytest.mark.parametrize("n", [1, 2])
def test_single_image_batches(n):
    """Smallest batches (one sample: the 3-D token maps degenerate to one slice)."""
    cfg = vo.VitCfg(image_size=64, patch=16, width=256, layers=2, heads=4, embed_dim=128)
    c, seed = 5, 61
    w = vo.synth_weights(cfg, seed)
    images, labels = synth_inputs(cfg, n, c, seed + 1)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 2)
    want = vo.online_step_oracle(images, labels, w, text, cfg)
    m = build_model(cfg, w)
    eng = m.model.visual.engine()
    eng.forward(torch.from_numpy(images).cuda(), training=True)
    head = eng.head(torch.from_numpy(text).cuda(), 1.0 / 0.07,
                    labels=torch.from_numpy(labels).cuda())
    eng.backward_from_head(head)
    torch.cuda.synchronize()
    names = [k for k in w if "lora" in k]
    grads = {k: g.cpu().numpy() for k, g in zip(names, eng.lora_grad_views)}
    check_step(head.probs.cpu().numpy(), float(head.loss_rows.sum()), head.pred.cpu().numpy(),
               grads, want, cfg, tol=2e-2)   # 17 tokens x 1-2 images: almost no averaging


def test_vit_l14_shapes_two_layers():
    """BASELINE config 3 geometry (ViT-L/14: width 1024, 16 heads, 257 tokens, embed 768) at two
    layers: the 257-token attention runs on the mma.sync kernels, everything else as ViT-B/16."""
    cfg = vo.VitCfg(image_size=224, patch=14, width=1024, layers=2, heads=16, embed_dim=768)
    n, c, seed = 3, 200, 41
    w = vo.synth_weights(cfg, seed)
    images, labels = synth_inputs(cfg, n, c, seed + 1)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 2)
    want = vo.online_step_oracle(images, labels, w, text, cfg)
    m = build_model(cfg, w)
    eng = m.model.visual.engine()
    eng.forward(torch.from_numpy(images).cuda(), training=True)
    head = eng.head(torch.from_numpy(text).cuda(), 1.0 / 0.07,
                    labels=torch.from_numpy(labels).cuda())
    eng.backward_from_head(head)
    torch.cuda.synchronize()
    names = [k for k in w if "lora" in k]
    grads = {k: g.cpu().numpy() for k, g in zip(names, eng.lora_grad_views)}
    check_step(head.probs.cpu().numpy(), float(head.loss_rows.sum()), head.pred.cpu().numpy(),
               grads, want, cfg)


def test_inference_1000_classes_masked_logits():
    """BASELINE config 5 shape of the head: eval (no grad), 1000 cached class text embeddings,
    both class-restriction variants: gather of the seen classes (methods/adapter_clip.py:129-130)
    and the additive -inf mask (methods/mvp_clip.py:113-118). Predictions bit-exact where the
    oracle's top-2 gap exceeds the bf16 noise floor."""
    cfg = vo.VitCfg(image_size=64, patch=16, width=768, layers=2, heads=12, embed_dim=512)
    n, c, seed = 48, 1000, 51
    w = vo.synth_weights(cfg, seed)
    images, _ = synth_inputs(cfg, n, c, seed + 1)
    text = vo.synth_text_features(c, cfg.embed_dim, seed + 2)
    m = build_model(cfg, w)
    eng = m.model.visual.engine()
    wd = vo.to_torch(w, torch.float64, lora_grad=False)
    feat = vo.vit_forward(torch.from_numpy(images).double(), wd, cfg)
    seen = np.sort(np.random.default_rng(seed).choice(c, size=300, replace=False))
    mask = np.full((c,), -np.inf); mask[seen] = 0.0
    with torch.no_grad():
        eng.forward(torch.from_numpy(images).cuda(), training=False)
        h_all = eng.head(torch.from_numpy(text).cuda(), 1.0 / 0.07)
        h_gat = eng.head(torch.from_numpy(text).cuda(), 1.0 / 0.07,
                         cls_idx=torch.from_numpy(seen).cuda())
        h_msk = eng.head(torch.from_numpy(text).cuda(), 1.0 / 0.07,
                         add_mask=torch.from_numpy(mask).float().cuda())
    torch.cuda.synchronize()
    for h, idx, msk in ((h_all, None, None), (h_gat, torch.from_numpy(seen), None),
                        (h_msk, None, torch.from_numpy(mask))):
        probs, _, _ = vo.head_forward(feat, torch.from_numpy(text).double(), 1.0 / 0.07, idx, msk)
        assert rel(h.probs, probs) < TOL
        srt = torch.sort(probs, dim=-1).values
        safe = ((srt[:, -1] - srt[:, -2]) > 0.05 * srt[:, -1]).numpy()
        np.testing.assert_array_equal(h.pred.cpu().numpy()[safe], probs.argmax(-1).numpy()[safe])
    assert float(h_msk.probs[:, np.setdiff1d(np.arange(c), seen)].abs().max()) == 0.0


def test_block_module_is_dropin_on_seq_first_layout():
    """ResidualAttentionBlock_LoRA on the reference's [L, N, D] layout, autograd through x and the
    four LoRA tensors, against oracle.block_forward (model.py:233-236, :400-415)."""
    from lifelong_clip_b200.clip_modules import ResidualAttentionBlock_LoRA
    cfg = vo.VitCfg(image_size=32, patch=8, width=256, layers=1, heads=4, embed_dim=64)
    w = vo.synth_weights(cfg, 5)
    pre = "visual.transformer.resblocks.0."
    blk = ResidualAttentionBlock_LoRA(cfg.width, cfg.heads, None,
                                      {"lora_alpha": 1, "lora_r": 4})
    blk.load_state_dict({k[len(pre):]: torch.from_numpy(v) for k, v in w.items()
                         if k.startswith(pre)})
    blk.cuda()
    assert [k for k, _ in blk.named_parameters() if "lora" in k] == [
        "attn.in_proj_weight_lora_A", "attn.in_proj_weight_lora_B", "attn.out_proj.lora_A",
        "attn.out_proj.lora_B"]
    L, N = 17, 3
    g = torch.Generator().manual_seed(0)
    x = torch.randn(L, N, cfg.width, generator=g)
    dy = torch.randn(L, N, cfg.width, generator=g)
    xc = x.cuda().requires_grad_(True)
    y = blk(xc)
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    wd = vo.to_torch(w, torch.float64)
    xd = x.double().transpose(0, 1).contiguous().requires_grad_(True)  # oracle is [N, L, D]
    yd = vo.block_forward(xd, wd, pre, cfg)
    yd.backward(dy.double().transpose(0, 1))
    assert rel(y.detach().cpu().transpose(0, 1), yd.detach()) < 5e-3
    assert rel(xc.grad.cpu().transpose(0, 1), xd.grad) < TOL
    for k, p in blk.named_parameters():
        if "lora" in k:
            assert rel(p.grad.cpu(), wd[pre + k].grad) < TOL, k
        else:
            assert p.grad is None, k   # frozen backbone: weight gradients are never computed


def test_causal_block_text_tower_mask():
    """attn_mask = the text tower's causal mask (model.py:926-932) maps onto the causal kernel."""
    from lifelong_clip_b200.clip_modules import ResidualAttentionBlock_LoRA
    cfg = vo.VitCfg(image_size=32, patch=8, width=128, layers=1, heads=2, embed_dim=64)
    w = vo.synth_weights(cfg, 6)
    pre = "visual.transformer.resblocks.0."
    L, N = 21, 2
    mask = torch.full((L, L), float("-inf")).triu(1)
    blk = ResidualAttentionBlock_LoRA(cfg.width, cfg.heads, mask, {"lora_alpha": 1, "lora_r": 4})
    blk.load_state_dict({k[len(pre):]: torch.from_numpy(v) for k, v in w.items()
                         if k.startswith(pre)})
    blk.cuda()
    x = torch.randn(L, N, cfg.width, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y = blk(x.cuda())
    wd = vo.to_torch(w, torch.float64, lora_grad=False)
    yd = vo.block_forward(x.double().transpose(0, 1), wd, pre, cfg, causal=True)
    assert rel(y.cpu().transpose(0, 1), yd) < 5e-3


def test_visual_transformer_encode_image_and_state_dict_keys():
    """VisualTransformer mirror: reference state_dict keys, encode_image -> [N, E] features."""
    cfg = vo.VIT_TINY
    w = vo.synth_weights(cfg, 9)
    m = build_model(cfg, w)
    keys = set(m.model.state_dict().keys())
    assert set(w.keys()) <= keys and keys - set(w.keys()) == {"logit_scale"}
    images, _ = synth_inputs(cfg, 4, 3, 1)
    with torch.no_grad():
        f = m.model.encode_image(torch.from_numpy(images).cuda())
    wd = vo.to_torch(w, torch.float64, lora_grad=False)
    want = vo.vit_forward(torch.from_numpy(images).double(), wd, cfg)
    assert rel(f.cpu(), want) < TOL


def test_trainer_online_step_interface_and_learning():
    """online_step(images, labels, idx) -> (loss, acc) floats; visible-class bookkeeping in order
    of first exposure; label remap bit-exact; the loss goes down on a repeated batch."""
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    cfg = vo.VIT_TINY
    c = 10
    m = build_model(cfg, vo.synth_weights(cfg, 3))
    names = [f"class{i}" for i in range(c)]
    m.set_text_features(names, torch.from_numpy(vo.synth_text_features(c, cfg.embed_dim, 4)))
    tr = LoRAClipTrainer(m, names, n_classes=c, n_tasks=2, lr=5e-3, online_iter=1,
                         visible_classes="all")
    tr.online_before_task(0)
    assert all(("lora" in k) == p.requires_grad for k, p in m.named_parameters())
    rng = np.random.default_rng(0)
    images = torch.from_numpy(rng.standard_normal((8, 3, 32, 32)).astype(np.float32))
    labels = torch.tensor([7, 2, 7, 9, 2, 4, 4, 7])
    losses = []
    for _ in range(12):
        loss, acc = tr.online_step(images, labels, torch.arange(8))
        assert isinstance(loss, float) and isinstance(acc, float) and 0.0 <= acc <= 1.0
        losses.append(loss)
    assert tr.exposed_classes == [7, 2, 9, 4]           # order of first exposure
    assert tr.last_head.args.C == 4
    want_local = vo.label_remap(labels.numpy(), tr.exposed_classes)
    got_local = torch.zeros(8, dtype=torch.int64)
    from lifelong_clip_b200 import ops
    got_local = ops.label_remap(labels.cuda(), tr._class_lut(tr.exposed_classes)).cpu().numpy()
    np.testing.assert_array_equal(got_local, want_local)
    assert losses[-1] < losses[0] - 1e-4, losses
    # the prefetching loader wrapper hands device images to the same interface
    from lifelong_clip_b200.trainer import DevicePrefetcher
    pinned = [(images.pin_memory(), labels, torch.arange(8)) for _ in range(3)]
    seen = 0
    for im, lb, ids in DevicePrefetcher(pinned, "cuda:0"):
        assert im.is_cuda and not lb.is_cuda
        loss2, _ = tr.online_step(im, lb, ids)
        seen += 1
    assert seen == 3 and loss2 < losses[0]
    tr.online_after_task(0)
    res = tr.online_evaluate([(images, labels)])
    assert set(res) == {"avg_loss", "avg_acc", "cls_acc", "task_acc", "confusion_matrix"}


def test_trainer_replay_concat_and_batch_visible_classes():
    """Replay path of online_train (methods/adapter_clip.py:65-73: the memory batch is concatenated
    to the stream batch, unseen replay classes join the visible list) with device-resident stream
    images and host-resident replay images; visible_classes='batch' changes the class list every
    step, which must quietly drop the CUDA-graph path instead of re-capturing forever."""
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    cfg = vo.VIT_TINY
    c = 10
    m = build_model(cfg, vo.synth_weights(cfg, 3))
    names = [f"class{i}" for i in range(c)]
    m.set_text_features(names, torch.from_numpy(vo.synth_text_features(c, cfg.embed_dim, 4)))
    rng = np.random.default_rng(1)

    class FakeMemory:            # the two members online_train touches (utils/memory.py)
        filled = 0
        def __len__(self): return self.filled
        def add_new_class(self, cls_list): self.cls = list(cls_list)

    def provider():
        while True:
            yield (torch.from_numpy(rng.standard_normal((4, 3, 32, 32)).astype(np.float32)),
                   torch.tensor([1, 1, 3, 5]))

    tr = LoRAClipTrainer(m, names, n_classes=c, n_tasks=2, lr=1e-3, visible_classes="batch",
                         memory=FakeMemory(), memory_provider=provider(), memory_batchsize=4)
    tr.online_before_task(0)
    # replay memory only ever holds classes the stream has shown: expose 1, 3, 5 first
    first = torch.tensor([1, 3, 5, 1, 3, 5])
    tr.online_step(torch.from_numpy(rng.standard_normal((6, 3, 32, 32)).astype(np.float32)).cuda(),
                   first, torch.arange(6))
    tr.memory.filled = 4
    for step in range(8):
        labels = torch.from_numpy(rng.integers(0, c, size=(6,)))
        images = torch.from_numpy(rng.standard_normal((6, 3, 32, 32)).astype(np.float32)).cuda()
        loss, acc = tr.online_step(images, labels, torch.arange(6))
        assert np.isfinite(loss) and 0.0 <= acc <= 1.0
        assert tr.last_head.args.N == 10                      # 6 stream + 4 replay samples
        assert {1, 3, 5} <= set(tr.batch_exposed_classes)     # replay classes became visible
    assert tr.use_cuda_graph is False                         # churned signatures -> eager


def test_no_cpu_fallback():
    """A CPU tensor / CPU module must fail loudly: the product path has no eager fallback."""
    from lifelong_clip_b200 import ops
    from lifelong_clip_b200.adapter_clip import AdapterCLIP
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.ln_fwd(torch.zeros(4, 128), torch.ones(128), torch.zeros(128),
                   torch.zeros(4, 128, dtype=torch.bfloat16))
    cfg = vo.VIT_TINY
    m = AdapterCLIP(vision_config=(cfg.image_size, cfg.patch, cfg.width, cfg.layers,
                                   cfg.embed_dim))
    with pytest.raises(RuntimeError, match="CUDA"):
        m.model.encode_image(torch.zeros(1, 3, 32, 32))
