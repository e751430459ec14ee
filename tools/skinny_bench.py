"""Skinny (N = 16) GEMM timing at the step's shapes against the HBM stream time (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import ops

T = int(os.environ.get("T", 50432))
torch.manual_seed(0)
for (K, ld) in [(768, 768 + 64), (2304, 2304 + 64), (768, 768)]:
    A = torch.randn(T, ld, device="cuda").to(torch.bfloat16)
    B = (torch.randn(16, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    out = torch.empty(T, 64, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops.gemm_tn(A, B, T, 16, K, out)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    ref = A[:2048, :K].float() @ B.float().T
    err = float((out[:2048, :16].float() - ref).norm() / ref.norm())
    print(f"K={K} ld={ld}: {us:7.1f} us  {T*K*2/us/1e3:7.1f} GB/s  rel {err:.1e}")
