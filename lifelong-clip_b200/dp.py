"""Data-parallel plumbing of the online step: one process per GPU, persistent
replicas, the combined stream+replay minibatch split rank::world, ONE all-reduce(sum) of the flat
LoRA gradient buffer (221,184 fp32 for ViT-B/16) plus a 2-float all-reduce for (loss, #correct).
Replaces nn.DataParallel (methods/_trainer.py:167-168), which re-broadcasts ~600 MB of parameters
every step. Pure torch.distributed calls: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_batch(x, y, rank: int, world: int):
    """rank::world slice of the global batch (every rank iterates the same sampler)."""
    if world == 1:
        return x, y
    return x[rank::world], y[rank::world]


def gather_labels(labels: torch.Tensor, world: int, device=None) -> torch.Tensor:
    """All ranks' label shards (equal sizes), for the class bookkeeping of
    methods/_trainer.py:404-416 which must see the GLOBAL batch. Returns a CPU tensor."""
    if world == 1:
        return labels.cpu()
    mine = labels.to(device) if device is not None else labels
    allb = torch.empty(world * mine.numel(), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(allb, mine.contiguous())
    return allb.cpu()


def allreduce_step(grad_flat: torch.Tensor, scalars: torch.Tensor, world: int) -> None:
    """Sum the per-shard LoRA gradients and the (loss_sum, n_correct) pair. Every shard's loss was
    already scaled by 1/global_batch, so the sum IS the global-mean gradient: no division."""
    if world == 1:
        return
    dist.all_reduce(grad_flat)
    dist.all_reduce(scalars)
