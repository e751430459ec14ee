import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import ops
N, L, D, E, C = 256, 197, 768, 512, 100
x = torch.randn(N * L, D, device="cuda"); g = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
proj = torch.randn(D, E, device="cuda") * D ** -0.5
text = torch.nn.functional.normalize(torch.randn(C, E, device="cuda"), dim=-1)
labels = torch.randint(0, C, (N,), device="cuda")
h = ops.Head(x, L, g, b, proj, text, 14.3, N, labels=labels)
dx = torch.zeros(N * L, D, device="cuda")
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(f"head fwd {timeit(h.forward):7.1f} us   bwd {timeit(lambda: h.backward(dx)):7.1f} us")
