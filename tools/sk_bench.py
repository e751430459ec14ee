"""Stream-K vs whole-tile schedule of the CTA-pair GEMM, per shape, same process (dev tool):
alternates the two schedules over the GEMM shapes of the step at several token counts and checks
that both give the same result (to fp32 summation order) against an fp32 matmul.
usage: python tools/sk_bench.py > profiles/r02_streamk_ab.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lifelong_clip_b200 import _capi as K
from lifelong_clip_b200 import ops

lib = K.load()
ws_bytes = lib.llc_gemm_ws_bytes()
ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
NIT = 30


def run(A, B, M, N, Kd, out, mode, resid):
    e = K.GemmEpi()
    e.out = out.data_ptr(); e.ld_out = out.stride(0); e.out_fp32 = int(out.dtype == torch.float32)
    if resid is not None:
        e.resid = resid.data_ptr(); e.ld_resid = resid.stride(0)
    e.ws = ws.data_ptr(); e.ws_bytes = ws_bytes
    import ctypes as C
    K.check(lib.llc_gemm_bf16_tn(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), M, N, Kd,
                                 C.byref(e), K.stream_ptr()), "gemm")


def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(NIT):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / NIT * 1e3


torch.manual_seed(0)
print(f"# workspace {ws_bytes} bytes; us per launch (CUDA events, {NIT} launches back to back)")
for imgs in (16, 32, 64, 128, 192, 256):
    T = 197 * imgs
    for (N, Kd, mode) in ((768, 3072, "f32"), (768, 3072, "bf16"), (768, 2320, "bf16"),
                          (3072, 768, "bf16"), (2304, 784, "bf16"), (768, 784, "bf16")):
        A = torch.randn(T, Kd, device="cuda").to(torch.bfloat16)
        B = (torch.randn(N, Kd, device="cuda") * Kd ** -0.5).to(torch.bfloat16)
        out = torch.empty(T, N, device="cuda", dtype=torch.float32 if mode == "f32" else torch.bfloat16)
        resid = torch.randn(T, N, device="cuda") if mode == "f32" else None
        res = {}
        for sk in (0, 1):
            lib.llc_gemm_set_stream_k(sk)
            us = timed(lambda: run(A, B, T, N, Kd, out, mode, resid))
            ref = A[:2048].float() @ B.float().T + (resid[:2048] if resid is not None else 0)
            tail = A[-300:].float() @ B.float().T + (resid[-300:] if resid is not None else 0)
            err = max(float((out[:2048].float() - ref).norm() / ref.norm()),
                      float((out[-300:].float() - tail).norm() / tail.norm()))
            res[sk] = (us, err)
        lib.llc_gemm_set_stream_k(1)
        tiles = ((T + 255) // 256) * (N // 256)
        print(f"images {imgs:3d} T {T:6d} N {N:4d} K {Kd:4d} {mode:4s} tiles {tiles:4d} "
              f"({tiles / 74:5.2f} waves)  whole-tile {res[0][0]:7.1f} us  stream-K {res[1][0]:7.1f} us"
              f"  ({res[0][0] / res[1][0]:4.2f}x)  rel err {res[0][1]:.1e} / {res[1][1]:.1e}",
              flush=True)
