"""GPU input transforms mirroring methods/_trainer.py:236-247.

    train_transform = Compose([Resize((S, S)), RandomCrop(S, padding=4), RandomHorizontalFlip(),
                               Normalize(mean, std)])        # applied to the whole batch tensor
    test_transform  = Compose([Resize((S, S)), Normalize(mean, std)])

The reference applies these to a [B, 3, h, w] float batch already on the device through ~6 ATen
ops; the DataLoader before it ships float32 pixels. Here the RAW batch (uint8 as the dataset stores
it, or float 0..1) goes over PCIe and ONE kernel produces the normalised 224x224 tensor
(GpuTransform.__call__, drop-in for `trainer.train_transform(x)`), or — on the fused trainer path —
directly the bf16 patch rows the patch-embedding GEMM reads (llc_vit_forward_tx).

Random draws stay on the host, from torch's global CPU generator and in torchvision's order
(RandomCrop.get_params: i, j; RandomHorizontalFlip: one rand(1) per batch), so a seeded run makes
the decisions torchvision would make. The AutoAugment / RandAugment / Cutout policies of
methods/_trainer.py:214-233 are PIL-style RNG-heavy ops and stay outside this package.
"""
from __future__ import annotations

import torch

from . import ops


class GpuTransform:
    def __init__(self, inp_size: int, mean, std, padding: int = 0, flip_p: float = 0.0,
                 device=None):
        self.S, self.mean, self.std = int(inp_size), tuple(mean), tuple(std)
        self.padding, self.flip_p = int(padding), float(flip_p)
        self.device = device
        self._dyn = None       # device int32 [3]: crop_i, crop_j, flip (CUDA-graph replays)
        self.last_draw = (0, 0, False)

    @classmethod
    def train(cls, inp_size, mean, std, device=None):
        """Resize -> RandomCrop(inp_size, padding=4) -> RandomHorizontalFlip -> Normalize."""
        return cls(inp_size, mean, std, padding=4, flip_p=0.5, device=device)

    @classmethod
    def test(cls, inp_size, mean, std, device=None):
        """Resize -> Normalize."""
        return cls(inp_size, mean, std, padding=0, flip_p=0.0, device=device)

    # ------------------------------------------------------------------------------------------
    def draw(self):
        """The batch-level random decisions, consumed from the global CPU generator exactly as
        torchvision's RandomCrop.get_params and RandomHorizontalFlip.forward do."""
        i = j = 0
        if self.padding > 0:
            i = int(torch.randint(0, 2 * self.padding + 1, size=(1,)).item())
            j = int(torch.randint(0, 2 * self.padding + 1, size=(1,)).item())
        flip = bool(torch.rand(1) < self.flip_p) if self.flip_p > 0 else False
        self.last_draw = (i, j, flip)
        return self.last_draw

    def _check(self, x):
        if not x.is_cuda:
            raise RuntimeError("GpuTransform runs on CUDA tensors only (no CPU fallback)")
        if x.dtype not in (torch.uint8, torch.float32):
            x = x.float()
        return x.contiguous()

    def push(self, draw, device):
        """Write a draw into the device parameter buffer (stable pointer: a captured CUDA graph
        reads it on every replay). The source is a fresh pageable tensor, so the copy is staged
        before this call returns and the next draw cannot overwrite it."""
        if self._dyn is None or self._dyn.device != torch.device(device):
            self._dyn = torch.zeros(3, dtype=torch.int32, device=device)
        i, j, flip = draw
        self._dyn.copy_(torch.tensor([i, j, int(flip)], dtype=torch.int32))
        return self._dyn

    def struct(self, x, dynamic: bool = False, draw=None):
        """Describe the transform of raw batch `x` for the C call; draws this batch's parameters
        unless `draw` is given. dynamic=True routes them through the device buffer."""
        if draw is None:
            draw = self.draw()
        i, j, flip = draw
        dyn = self.push(draw, x.device) if dynamic else None
        return ops.make_transform(x, self.S, self.mean, self.std, pad=self.padding, crop=(i, j),
                                  flip=flip, dyn=dyn)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """Raw batch [N, 3, h, w] (uint8 0..255 or float 0..1) -> fp32 [N, 3, S, S]."""
        x = self._check(x)
        out = torch.empty(x.shape[0], 3, self.S, self.S, device=x.device)
        return ops.transform_images(self.struct(x), x.shape[0], out)
