"""Per-shape timing + correctness of llc_gemm_bf16_tn on the step's GEMM shapes (dev tool).
usage: python tools/gemm_bench.py [--check]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import ops

T = int(os.environ.get("T", 50432))
SHAPES = [  # (M, N, K, mode)
    (T, 2304, 784, "bf16"), (T, 768, 784, "f32"), (T, 3072, 768, "gelu"), (T, 768, 3072, "f32"),
    (T, 3072, 768, "dgelu"), (T, 768, 3072, "bf16"), (T, 768, 784, "bf16"), (T, 768, 2320, "bf16"),
]
check = "--check" in sys.argv
if os.environ.get("ONLY"):
    SHAPES = [SHAPES[int(i)] for i in os.environ["ONLY"].split(",")]
NIT = int(os.environ.get("NIT", 10))
WARM = int(os.environ.get("WARM", 3))
CUBLAS = "--cublas" in sys.argv
torch.manual_seed(0)
for (M, N, K, mode) in SHAPES:
    A = (torch.randn(M, K, device="cuda")).to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") * 0.1
    kw = {}
    if mode == "bf16":
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); kw = dict(bias=bias)
    elif mode == "f32":
        out = torch.empty(M, N, device="cuda"); kw = dict(bias=bias, resid=torch.randn(M, N, device="cuda"))
    elif mode == "gelu":
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        kw = dict(bias=bias, act=1, out2=torch.empty(M, N, device="cuda", dtype=torch.bfloat16))
    else:
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        kw = dict(act=2, aux=torch.randn(M, N, device="cuda").to(torch.bfloat16))
    for _ in range(WARM):
        ops.gemm_tn(A, B, M, N, K, out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = NIT
    e0.record()
    for _ in range(n):
        ops.gemm_tn(A, B, M, N, K, out, **kw)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    msg = f"M={M} N={N} K={K} {mode:6s} {us:8.1f} us  {2.0*M*N*K/us/1e6:7.1f} TF/s"
    if check:
        ref = A[:4096].float() @ B.float().T
        if mode == "bf16": want = ref + bias
        elif mode == "f32": want = ref + bias + kw["resid"][:4096]
        elif mode == "gelu": want = ref + bias
        else:
            z = kw["aux"][:4096].float(); s = torch.sigmoid(1.702 * z)
            want = ref * (s * (1 + 1.702 * z * (1 - s)))
        got = out[:4096].float()
        err = float((got - want).norm() / want.norm())
        # last rows too (tile tails)
        msg += f"  rel {err:.2e}"
    if CUBLAS:  # reference point only: cuBLAS on the same shape (plain, no epilogue)
        for _ in range(3):
            torch.matmul(A, B.T)
        e0.record()
        for _ in range(n):
            torch.matmul(A, B.T)
        e1.record(); torch.cuda.synchronize()
        cu = e0.elapsed_time(e1) / n * 1e3
        msg += f"   | cuBLAS plain {cu:7.1f} us {2.0*M*N*K/cu/1e6:7.1f} TF/s"
    print(msg, flush=True)
