"""Adapter forward / backward at the bench shape (T = 197 x images tokens, D = 768) through the
module (llc_adapter_forward / llc_adapter_backward), CUDA-event timed (dev tool).
usage: python tools/adapter_bench.py [images] > profiles/r02_adapter_bench.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import ops
from lifelong_clip_b200.adapter_modules import Adapter

imgs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
D, T = 768, 197 * imgs
torch.manual_seed(0)
ad = Adapter(d_model=D, dropout=0.1, bottleneck=64, init_option="lora", adapter_scalar=0.1,
             adapter_layernorm_option="none").cuda().train()
with torch.no_grad():
    ad.up_proj.weight.normal_(0, 0.02)
x = torch.randn(T, D, device="cuda", requires_grad=True)
dy = torch.randn(T, D, device="cuda")


def timeit(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def fwd_bwd():
    ad.zero_grad()
    x.grad = None
    ad(x).backward(dy)


print(f"images {imgs}  T {T}  D {D}")
print(f"forward + backward (module, incl. casts and allocations) {timeit(fwd_bwd):8.1f} us")
ops.prof_enable(True)
fwd_bwd()
torch.cuda.synchronize()
recs = ops.prof_read()
ops.prof_enable(False)
for kind, m, n, k, ms, fl, by in recs:
    print(f"  {kind:10s} m {m:6d} n {n:5d} k {k:5d}  {ms * 1e3:7.1f} us  {by / (ms * 1e-3) / 1e9:7.0f} GB/s"
          f"  {fl / (ms * 1e-3) / 1e12:7.1f} TFLOP/s")
