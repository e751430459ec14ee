import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import ops
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
x = torch.empty(512 * 1024 * 1024, dtype=torch.bfloat16, device="cuda")   # 1 GiB
y = torch.empty_like(x)
us = timeit(lambda: x.zero_()); print(f"pure write  1 GiB: {us:7.1f} us  {x.numel()*2/us/1e3:7.1f} GB/s")
us = timeit(lambda: y.copy_(x)); print(f"copy        1 GiB: {us:7.1f} us  {2*x.numel()*2/us/1e3:7.1f} GB/s (r+w)")
us = timeit(lambda: x.sum()); print(f"pure read   1 GiB: {us:7.1f} us  {x.numel()*2/us/1e3:7.1f} GB/s")
M, N, K = 50432, 3072, 768
A = torch.randn(M, K, device="cuda").to(torch.bfloat16); B = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.zeros(N, device="cuda")
z = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); g = torch.empty_like(z)
us = timeit(lambda: ops.gemm_tn(A, B, M, N, K, z, bias=bias, act=1, out2=g)); print(f"gelu z+g : {us:7.1f} us")
us = timeit(lambda: ops.gemm_tn(A, B, M, N, K, None, bias=bias, act=1, out2=g)); print(f"gelu g only: {us:7.1f} us")
us = timeit(lambda: ops.gemm_tn(A, B, M, N, K, z, bias=bias)); print(f"bf16 plain (one output): {us:7.1f} us")
