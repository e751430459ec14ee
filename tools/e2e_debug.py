import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import ops
from lifelong_clip_b200.adapter_clip import AdapterCLIP
from lifelong_clip_b200.trainer import LoRAClipTrainer, DevicePrefetcher
B, C = int(os.environ.get("B", 256)), 100
graph = os.environ.get("GRAPH", "1") == "1"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = AdapterCLIP(vision_config=(224, 16, 768, 12, 512)).to(dev)
names = [f"c{i}" for i in range(C)]
model.set_text_features(names, torch.randn(C, 512))
tr = LoRAClipTrainer(model, names, n_classes=C, visible_classes="all", use_cuda_graph=graph)
tr.online_before_task(0)
tr.add_new_class(torch.arange(C))
hx = [torch.randn(B, 3, 224, 224).pin_memory() for _ in range(3)]
hy = [torch.randint(0, C, (B,)).pin_memory() for _ in range(3)]
idx = torch.arange(B)
def loader(n):
    for i in range(n): yield hx[i % 3], hy[i % 3], idx
for im, lb, ids in DevicePrefetcher(loader(4), dev): tr.online_step(im, lb, ids)
torch.cuda.synchronize()
ts = []
t0 = time.perf_counter()
caps = []
for im, lb, ids in DevicePrefetcher(loader(8), dev):
    k0 = tr._graph_key
    tr.online_step(im, lb, ids)
    t1 = time.perf_counter(); ts.append((t1 - t0) * 1e3); t0 = t1
    caps.append(tr._graph_key is not k0)
print("graph", graph, "per-step wall ms:", [round(t, 1) for t in ts], "recaptured:", caps)
# same loop but images already on device
dx = [h.to(dev) for h in hx]
torch.cuda.synchronize(); ts = []; t0 = time.perf_counter()
for i in range(8):
    tr.online_step(dx[i % 3], hy[i % 3], idx)
    t1 = time.perf_counter(); ts.append((t1 - t0) * 1e3); t0 = t1
print("device-resident images:", [round(t, 1) for t in ts])
