"""lifelong_clip_b200 — B200 (sm_100a) implementation of LifeLong-CLIP's online-step hot path.

Host code is Python + PyTorch (device memory, streams, torch.distributed); all compute goes through
the C-ABI library libllc.so (hand-written CUDA: TMA/tcgen05 GEMMs, fused attention, LayerNorm,
LoRA side reductions, head). There is no CPU or eager fallback.
"""
__version__ = "0.1.0"
