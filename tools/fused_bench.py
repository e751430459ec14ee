"""Fused LoRA column-sum + row-product pass at the step's shapes (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import ops
T = 50432
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for C in (768, 2304):
    X = torch.randn(T, C + 64, device="cuda").to(torch.bfloat16)
    w = torch.randn(T, 832, device="cuda").to(torch.bfloat16)
    F = torch.zeros(16, C, device="cuda", dtype=torch.bfloat16); F[:8] = torch.randn(8, C, device="cuda") * 0.05
    partial = torch.empty(ops.lora_side_max_partials() * C * 8, device="cuda")
    U = X[:, C:]
    n = [0]
    def f(): n[0] = ops.lora_side_fused(X, T, C, 8, w[:, 768:], 832, F, U, partial)
    us = timeit(f)
    print(f"fused C={C}: {us:7.1f} us  {T*C*2/us/1e3:7.1f} GB/s  partials {n[0]}")
    out = torch.empty(C, 8, device="cuda")
    us = timeit(lambda: ops.lora_colsum_finish(partial, n[0], C, 8, 1.0, out, 8, 1))
    print(f"finish C={C} ({n[0]} partials): {us:7.1f} us")
