"""Host-side cost of one online_step (dev tool): cProfile over 40 e2e steps (raw uint8 host
batches, GpuTransform, CUDA graph), plus the wall time between the end of one step's result read
and the next graph replay - the window in which the GPU idles."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200.adapter_clip import AdapterCLIP
from lifelong_clip_b200.trainer import DevicePrefetcher, LoRAClipTrainer
from lifelong_clip_b200.transform import GpuTransform

B, C = int(os.environ.get("B", 256)), 100
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = AdapterCLIP(vision_config=(224, 16, 768, 12, 512)).to(dev)
names = [f"c{i}" for i in range(C)]
model.set_text_features(names, torch.randn(C, 512))
tr = LoRAClipTrainer(model, names, n_classes=C, visible_classes="all", use_cuda_graph=True,
                     sharded_input=True)
tr.online_before_task(0)
tr.add_new_class(torch.arange(C))
tr.train_transform = GpuTransform.train(224, (0.5071, 0.4867, 0.4408), (0.2675, 0.2565, 0.2761))
hx = [torch.randint(0, 256, (B, 3, 32, 32), dtype=torch.uint8).pin_memory() for _ in range(3)]
hy = [torch.randint(0, C, (B,)).pin_memory() for _ in range(3)]
idx = torch.arange(B)


def loader(n):
    for i in range(n):
        yield hx[i % 3], hy[i % 3], idx


def run(n):
    for im, lb, ids in DevicePrefetcher(loader(n), dev):
        tr.online_step(im, lb, ids)


run(6)
torch.cuda.synchronize()
t0 = time.perf_counter()
run(40)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 40 * 1e3
# device-only time of the same step
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
x = hx[0].to(dev); y = hy[0]
print(f"e2e wall per step {wall:.3f} ms")
pr = cProfile.Profile()
pr.enable()
run(40)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
st.sort_stats("tottime").print_stats(18)
