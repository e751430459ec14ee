"""Drives the host-side bookkeeping of LoRAClipTrainer side by side with the REFERENCE's own
methods/adapter_clip.AdapterCLIP.online_step on the same Si-Blurry stream, built from the
reference's real utils/online_sampler.OnlineSampler, utils/indexed_dataset.IndexedDataset and
utils/memory.Memory ("run unchanged" of the north star). CPU only; the numerics are stubbed out on
both sides (a recording model), what is compared is integer work: class exposure order, the visible
class list handed to set_token, the remapped labels - bit-exact.

Needs /root/reference (present in the authoring container, where the CPU suite runs); skipped
elsewhere. The reference modules are imported as files with stubs for the packages it imports
but this image lacks (randaugment, torch_optimizer, timm, clip)."""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = os.environ.get("LLC_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "methods")),
                                reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    """(AdapterCLIP trainer class, Memory, OnlineSampler, IndexedDataset) of the reference."""
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    sys.path.insert(0, REF)
    for name in ("randaugment", "torch_optimizer", "datasets", "models"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["randaugment"].RandAugment = object
    sys.modules["datasets"].get_dataset = lambda *a, **k: None
    sys.modules["models"].get_model = lambda *a, **k: None
    pkg = types.ModuleType("methods")          # bypass methods/__init__.py (imports every method)
    pkg.__path__ = [os.path.join(REF, "methods")]
    sys.modules["methods"] = pkg
    try:
        trainer_mod = importlib.import_module("methods.adapter_clip")
        memory_mod = importlib.import_module("utils.memory")
        sampler_mod = importlib.import_module("utils.online_sampler")
        indexed_mod = importlib.import_module("utils.indexed_dataset")
        yield (trainer_mod.AdapterCLIP, memory_mod.Memory, sampler_mod.OnlineSampler,
               indexed_mod.IndexedDataset)
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]
        sys.modules.update(saved_mods)


class SynthCifar(torch.utils.data.Dataset):
    """CIFAR-100-shaped synthetic set: 100 classes x 20 items, 3x8x8 images."""

    def __init__(self, n_classes=100, per_class=20):
        g = torch.Generator().manual_seed(0)
        self.targets = [c for c in range(n_classes) for _ in range(per_class)]
        self.data = torch.rand(len(self.targets), 3, 8, 8, generator=g)
        self.classes = list(range(n_classes))
        self.classes_names = [f"class_{i}" for i in range(n_classes)]

    def __len__(self):
        return len(self.targets)

    def __getitem__(self, i):
        return self.data[i], self.targets[i]


class Recorder:
    """Stands in for custom_clip(.module) on the reference side: records set_token calls."""

    def __init__(self):
        self.tokens, self.current_class_names = [], []
        self.module = self

    def update_class_names(self, names):
        for c in names:
            if c not in self.current_class_names:
                self.current_class_names.append(c)

    def set_token(self, names):
        self.tokens.append(list(names))
        self._c = len(names)

    def train(self):
        pass

    def __call__(self, x):
        n = x.shape[0]
        logit = torch.full((n, self._c), 1.0 / self._c, requires_grad=True)
        return logit, torch.zeros(n, 4), torch.zeros(self._c, 4)


@pytest.mark.parametrize("visible", ["all", "batch"])
def test_bookkeeping_matches_reference_trainer(ref, visible, monkeypatch):
    RefTrainer, Memory, OnlineSampler, IndexedDataset = ref
    ds = SynthCifar()
    train = IndexedDataset(ds)
    sampler = OnlineSampler(data_source=train, num_tasks=5, m=10, n=50, rnd_seed=1,
                            varing_NM=False)
    loader = torch.utils.data.DataLoader(train, batch_size=16, sampler=sampler, num_workers=0)

    # ---- the reference trainer, constructed without its CLI plumbing -------------------------
    r = object.__new__(RefTrainer)
    rec = Recorder()
    r.custom_clip, r.memory, r.train_dataset = rec, Memory(), ds
    r.exposed_classes, r.exposed_classes_names = [], []
    r.batch_exposed_classes, r.batch_exposed_classes_names = [], []
    r.visible_classes, r.memory_size, r.memory_batchsize = visible, 0, 0
    r.sched_name, r.online_iter, r.n_tasks, r.topk = "default", 1, 5, 1
    r.device, r.train_transform, r.use_amp = torch.device("cpu"), (lambda x: x), False
    r.criterion = torch.nn.CrossEntropyLoss()
    r.args = {}
    r.update_schedule = lambda *a, **k: None

    class _Opt:
        def zero_grad(self): pass
    class _Scaler:
        def scale(self, l): return l
        def step(self, o): pass
        def update(self): pass
    r.optimizer, r.scaler = _Opt(), _Scaler()
    ref_labels = []
    real_ce = r.criterion
    def rec_criterion(logit, y):
        ref_labels.append(y.clone())
        return real_ce(logit, y)
    r.criterion = rec_criterion

    # ---- ours, numerics stubbed out the same way -----------------------------------------------
    from lifelong_clip_b200 import ops
    from lifelong_clip_b200.adapter_clip import AdapterCLIP
    from lifelong_clip_b200.trainer import LoRAClipTrainer
    m = AdapterCLIP(vision_config=(32, 8, 128, 1, 64))
    m.set_text_features(ds.classes_names, torch.randn(100, 64))
    ours_tokens, ours_labels = [], []
    real_set_token = m.set_token
    def rec_set_token(names):
        ours_tokens.append(list(names))
        real_set_token(names)
    m.set_token = rec_set_token
    t = LoRAClipTrainer(m, ds.classes_names, n_classes=100, n_tasks=5, visible_classes=visible,
                        memory=Memory(), use_cuda_graph=False)
    t.optimizer = object()
    monkeypatch.setattr(ops, "label_remap", lambda y, lut: lut[y])   # the LUT gather, on the CPU
    def rec_step(x, y_local, B, sync=True):
        ours_labels.append(y_local.clone())
        return 0.0, 0
    t.fused_step = rec_step

    steps = 0
    for task in range(2):
        sampler.set_task(task)
        for images, labels, idx in loader:
            r.online_step(images.clone(), labels.clone(), idx)
            t.online_step(images, labels, idx)
            steps += 1
            assert t.exposed_classes == r.exposed_classes
            assert t.exposed_classes_names == r.exposed_classes_names
            assert t.batch_exposed_classes == r.batch_exposed_classes
            assert m.current_class_names == rec.current_class_names
    assert steps > 20
    assert ours_tokens == rec.tokens                      # visible class list of every step
    assert len(ours_labels) == len(ref_labels) == steps
    for a, b in zip(ours_labels, ref_labels):             # label remap, bit-exact
        assert a.dtype == b.dtype == torch.int64 and torch.equal(a, b)
    # after the task the reference evaluates over all_classnames[:_total_classes]
    total = 0
    for task in range(2):
        total += int(sampler.disjoint_class_num[task])
    r._total_classes = t._total_classes = total
    r.all_classnames = ds.classes_names
    r.online_after_task(1)
    t.online_after_task(1)
    assert ours_tokens[-1] == rec.tokens[-1] == ds.classes_names[:total]
