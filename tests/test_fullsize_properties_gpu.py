"""The hot path at BASELINE.json's full size (ViT-B/16, 256 images per GPU, 100 classes), where the
fp64 oracle no longer finishes in seconds: size-independent properties of the online step
instead of an element-wise comparison.

  determinism        the same step twice -> bit-identical probabilities, loss and LoRA gradients
                     (every reduction has a fixed order: partials + ordered finish, no atomics)
  loss-scale         gradients with loss_scale = 2 are bit-exactly twice the gradients (a power
                     of two commutes with every bf16 / fp32 rounding on the way back)
  batch additivity   grad(256 images) = grad(first 128) + grad(last 128) with the same 1/256 loss
                     normalisation; per-image outputs do not depend on the rest of the batch
  permutation        permuting the images permutes the probability rows and leaves the gradient
                     unchanged, both up to fp32 summation order
  class restriction  probabilities are a distribution over exactly the visible classes
  last block         class-token-only last block == full last block (the engine's two code paths)
"""
import numpy as np
import pytest
import torch

from oracle import vit_oracle as vo
from tests.test_e2e_gpu import build_model, rel

pytestmark = pytest.mark.gpu

N, C = 256, 100


@pytest.fixture(scope="module")
def setup():
    cfg = vo.VIT_B16
    w = vo.synth_weights(cfg, 5)
    rng = np.random.default_rng(6)
    images = torch.from_numpy(
        rng.standard_normal((N, 3, cfg.image_size, cfg.image_size)).astype(np.float32)).cuda()
    labels = torch.from_numpy(rng.integers(0, C, size=(N,)).astype(np.int64)).cuda()
    text = torch.from_numpy(vo.synth_text_features(C, cfg.embed_dim, 7)).cuda()
    m = build_model(cfg, w)
    eng = m.model.visual.engine()
    return cfg, eng, images, labels, text


def step(eng, images, labels, text, inv_batch=None, loss_scale=1.0, cls_idx=None):
    eng.forward(images, training=True)
    head = eng.head(text, 1.0 / 0.07, labels=labels, inv_batch=inv_batch, cls_idx=cls_idx)
    eng.backward_from_head(head, loss_scale=loss_scale)
    torch.cuda.synchronize()
    return (head.probs.clone(), head.loss_rows.sum().clone(), head.pred.clone(),
            eng.grad_flat.clone())


def test_determinism_and_loss_scale(setup):
    cfg, eng, images, labels, text = setup
    p1, l1, a1, g1 = step(eng, images, labels, text)
    p2, l2, a2, g2 = step(eng, images, labels, text)
    assert torch.equal(p1, p2) and torch.equal(l1, l2) and torch.equal(a1, a2)
    assert torch.equal(g1, g2)
    assert float(g1.abs().max()) > 0 and bool(torch.isfinite(g1).all())
    _, _, _, g3 = step(eng, images, labels, text, loss_scale=2.0)
    assert torch.equal(g3, 2.0 * g1)


def test_batch_additivity_and_independence(setup):
    cfg, eng, images, labels, text = setup
    p, l, _, g = step(eng, images, labels, text)
    h = N // 2
    pa, la, _, ga = step(eng, images[:h], labels[:h], text, inv_batch=1.0 / N)
    pb, lb, _, gb = step(eng, images[h:], labels[h:], text, inv_batch=1.0 / N)
    # an image's probabilities do not depend on what else is in the batch: the per-sample
    # arithmetic has no cross-sample term. Bit-exact under the whole-tile GEMM schedule; the
    # stream-K schedule cuts a tile's K range at shape-dependent points, so the fp32 summation
    # order (and, rarely, a bf16 rounding) changes with the batch size
    assert rel(p[:h], pa) < 2e-3 and rel(p[h:], pb) < 2e-3
    from lifelong_clip_b200 import _capi
    old = _capi.load().llc_gemm_set_stream_k(0)
    try:
        p0, _, _, g0 = step(eng, images, labels, text)
        pa0, _, _, ga0 = step(eng, images[:h], labels[:h], text, inv_batch=1.0 / N)
        pb0, _, _, gb0 = step(eng, images[h:], labels[h:], text, inv_batch=1.0 / N)
        assert torch.equal(p0[:h], pa0) and torch.equal(p0[h:], pb0)
        # whole-tile schedule: only the fp32 order of the token sums differs
        assert rel(ga0 + gb0, g0) < 2e-4
    finally:
        _capi.load().llc_gemm_set_stream_k(old)
    assert abs(float(l) - float(la) - float(lb)) < 1e-5 * abs(float(l))
    # the gradient is a sum over images (stream-K on: a few bf16 roundings flip with the batch
    # size, as in test_permutation - far below the bf16 noise floor of the parity policy)
    assert rel(ga + gb, g) < 5e-3


def test_permutation(setup):
    cfg, eng, images, labels, text = setup
    p, l, a, g = step(eng, images, labels, text)
    perm = torch.from_numpy(np.random.default_rng(8).permutation(N)).cuda()
    pp, lp, ap, gp = step(eng, images[perm].contiguous(), labels[perm].contiguous(), text)
    # rows follow the permutation; not bit-exactly: the head walks the projection rows from a
    # start row that depends on the sample's position in the batch (fp32 summation order)
    assert rel(pp, p[perm]) < 1e-5
    assert float((ap == a[perm]).float().mean()) > 0.995
    assert abs(float(lp) - float(l)) < 1e-5 * abs(float(l))
    # the LoRA gradients are sums of 50 k token terms that largely cancel, so a different fp32
    # summation order (every tile now holds other images) shows up at ~1e-3 of the result -
    # well below the bf16 noise floor of the parity policy (1e-2), far above bit-exactness
    assert rel(gp, g) < 5e-3


def test_class_restriction_is_a_distribution(setup):
    cfg, eng, images, labels, text = setup
    vis = torch.arange(0, C, 3, device="cuda")                   # every third class visible
    lab = torch.randint(0, vis.numel(), (N,), device="cuda")
    p, _, a, g = step(eng, images, lab, text, cls_idx=vis)
    assert p.shape == (N, vis.numel())
    assert float((p.sum(-1) - 1).abs().max()) < 1e-5 and float(p.min()) >= 0
    assert int(a.min()) >= 0 and int(a.max()) < vis.numel()
    assert bool(torch.isfinite(g).all())


def test_class_token_only_last_block_at_full_size(setup, monkeypatch):
    cfg, eng, images, labels, text = setup
    monkeypatch.delenv("LLC_FULL_LAST_BLOCK", raising=False)
    p_cls, l_cls, _, g_cls = step(eng, images, labels, text)
    monkeypatch.setenv("LLC_FULL_LAST_BLOCK", "1")
    p_full, l_full, _, g_full = step(eng, images, labels, text)
    assert rel(p_cls, p_full) < 3e-3
    assert abs(float(l_cls) - float(l_full)) < 1e-4 * abs(float(l_full))
    assert rel(g_cls, g_full) < 5e-3
