import ctypes, os, sys, torch
here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, "_probe.so"))
lib.probe_run.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 2
torch.manual_seed(0)
for K in (16, 64, 208):
    A = torch.randn(128, K, device="cuda"); B = torch.randn(K, 64, device="cuda")
    ref = A.bfloat16().float() @ B.bfloat16().float()
    for variant in (0, 1):
        D = torch.zeros(128, 64, device="cuda")
        rc = lib.probe_run(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, variant)
        err = float((D - ref).norm() / ref.norm())
        print(f"K={K} variant={variant} rc={rc} rel={err:.3e}", flush=True)
