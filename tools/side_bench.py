"""LoRA side kernels + LN timing at the step's shapes (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import ops
T = 50432
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for C in (768, 2304):
    X = torch.randn(T, C + 64, device="cuda").to(torch.bfloat16)
    Mrd = torch.randn(C, 4, device="cuda") * 0.05
    w = torch.randn(T, 832, device="cuda").to(torch.bfloat16)
    partial = torch.empty(ops.lora_side_max_partials() * C * 8, device="cuda")
    out = torch.empty(C, 4, device="cuda")
    us = timeit(lambda: ops.lora_side(X, T, C, 4, Mrd=Mrd, rd_sc=4, rd_sj=1, rd_scale=0.25))
    print(f"rowdot C={C}: {us:7.1f} us  {T*C*2/us/1e3:7.1f} GB/s")
    n = [0]
    def cs(): n[0] = ops.lora_side(X, T, C, 4, w=w[:, 768:], ld_w=832, partial=partial)
    us = timeit(cs)
    print(f"colsum C={C}: {us:7.1f} us  {T*C*2/us/1e3:7.1f} GB/s  partials {n[0]}")
    us = timeit(lambda: ops.lora_colsum_finish(partial, n[0], C, 4, 1.0, out, 4, 1))
    print(f"finish C={C}: {us:7.1f} us")
# rowdot as a skinny GEMM on the tensor cores: u[T,16] = X[T,C] . F[16,C]^T written into X's pad
for C in (768, 2304):
    X = torch.randn(T, C + 64, device="cuda").to(torch.bfloat16)
    F = torch.zeros(16, C, device="cuda", dtype=torch.bfloat16); F[:4] = torch.randn(4, C, device="cuda") * 0.05
    out = X[:, C:C + 16]
    us = timeit(lambda: ops.gemm_tn(X, F, T, 16, C, out))
    ref = X[:4096, :C].float() @ F.float().T
    err = float((out[:4096].float() - ref).norm() / ref.norm())
    print(f"rowdot-as-GEMM C={C}: {us:7.1f} us  {T*C*2/us/1e3:7.1f} GB/s  rel {err:.2e}")
# LayerNorm forward / backward at the step's shape
D = 768
x = torch.randn(T, D, device="cuda"); g = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
A = torch.randn(4, D, device="cuda") * 0.05; Bm = torch.randn(D, 4, device="cuda") * 0.05
y = torch.empty(T, D + 64, device="cuda", dtype=torch.bfloat16)
us = timeit(lambda: ops.ln_fwd(x, g, b, y, lora_A=A, r=4)); print(f"ln_fwd: {us:7.1f} us  {T*D*6/us/1e3:7.1f} GB/s")
dy = torch.randn(T, D, device="cuda").to(torch.bfloat16); dxin = torch.randn(T, D, device="cuda"); dxo = torch.empty(T, D, device="cuda")
us = timeit(lambda: ops.ln_bwd(x, g, dy, dxin, dxo, dxb=y, lora_B=Bm, r=4, scale=0.25)); print(f"ln_bwd: {us:7.1f} us  {T*D*16/us/1e3:7.1f} GB/s")
y2 = torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
us = timeit(lambda: ops.ln_fwd(x, g, b, y2)); print(f"ln_fwd (no lora): {us:7.1f} us  {T*D*6/us/1e3:7.1f} GB/s")
us = timeit(lambda: ops.ln_bwd(x, g, dy, dxin, dxo, dxb=y)); print(f"ln_bwd (no lora): {us:7.1f} us  {T*D*16/us/1e3:7.1f} GB/s")
