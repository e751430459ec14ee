"""Attention fwd/bwd timing at ViT-L/14's shape (64 images x 16 heads x 257 tokens) (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import ops
N, L, H = int(os.environ.get("N", 64)), int(os.environ.get("L", 257)), int(os.environ.get("H", 16))
D = H * 64; T = N * L
torch.manual_seed(0)
qkv = torch.randn(T, 3 * D + 64, device="cuda").to(torch.bfloat16)
o = torch.zeros(T, D + 64, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(N * H * L, device="cuda")
d_o = torch.randn(T, D, device="cuda").to(torch.bfloat16)
dqkv = torch.zeros(T, 3 * D + 64, device="cuda", dtype=torch.bfloat16)
delta = torch.empty(N * H * L, device="cuda")
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(f"N {N} H {H} L {L}")
print(f"fwd {timeit(lambda: ops.attn_fwd(qkv, o, lse, N, L, H, L, 1, False)):8.1f} us")
print(f"bwd {timeit(lambda: ops.attn_bwd(qkv, o, d_o, lse, dqkv, N, L, H, L, 1, False, delta)):8.1f} us")
