"""Experiment (dev tool): does running the step as TWO half-batch replicas on two CUDA streams beat
one full-batch step? The HBM-bound kernels of one half could fill the tails and the power headroom
of the tensor-bound kernels of the other. Each replica is a complete trainer (own model copy, arena
and CUDA graph); the result is the time for 2 x (B/2) images against 1 x B.
usage: python tools/dual_stream.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import _capi as K
from lifelong_clip_b200 import ops
from lifelong_clip_b200.adapter_clip import AdapterCLIP
from lifelong_clip_b200.trainer import LoRAClipTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C, E, S = 100, 512, 224
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = K.load()
lib.llc_gemm_set_stream_k(0)      # two concurrent stream-K grids could wait on each other


def make(b):
    torch.manual_seed(0)
    model = AdapterCLIP(model_name="ViT-B/16", peft_encoder="image",
                        vision_config=(224, 16, 768, 12, 512)).to(dev)
    names = [f"class {i}" for i in range(C)]
    model.set_text_features(names, torch.randn(C, E, generator=torch.Generator().manual_seed(1)))
    tr = LoRAClipTrainer(model, names, n_classes=C, n_tasks=5, lr=1e-3, online_iter=1,
                         visible_classes="all", sharded_input=True, use_cuda_graph=True)
    tr.online_before_task(0)
    tr.add_new_class(torch.arange(C))
    model.set_token(tr.exposed_classes_names)
    lut = tr._class_lut(tr.exposed_classes)
    x = torch.randn(b, 3, S, S, device=dev)
    y = ops.label_remap(torch.randint(0, C, (b,)).to(dev), lut)
    return tr, x, y


def timed(fn, n=20, w=5):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


full = make(B)
t_full = timed(lambda: full[0].fused_step(full[1], full[2], B, sync=False))
print(f"one step of {B}: {t_full:.3f} ms")
del full
torch.cuda.empty_cache()

halves = [make(B // 2) for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def serial():
    for tr, x, y in halves:
        tr.fused_step(x, y, B, sync=False)


def dual():
    cur = torch.cuda.current_stream()
    for (tr, x, y), s in zip(halves, streams):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            tr.fused_step(x, y, B, sync=False)
    for s in streams:
        cur.wait_stream(s)


t_serial = timed(serial)
print(f"two steps of {B // 2}, one stream: {t_serial:.3f} ms")
# graphs were captured on the default stream; replaying them on side streams is allowed
t_dual = timed(dual)
print(f"two steps of {B // 2}, two streams: {t_dual:.3f} ms")
t_serial2 = timed(serial)
print(f"two steps of {B // 2}, one stream (again): {t_serial2:.3f} ms")
