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
    """All ranks' label shards (ragged allowed: the last batch of a stream), for the class
    bookkeeping of methods/_trainer.py:404-416 which must see the GLOBAL batch. Returns a CPU
    tensor."""
    if world == 1:
        return labels.cpu()
    mine = (labels.to(device) if device is not None else labels).contiguous()
    n = torch.tensor([mine.numel()], dtype=torch.int64, device=mine.device)
    counts = torch.empty(world, dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(counts, n)
    counts = counts.cpu().tolist()
    cap = max(counts)
    if cap == 0:
        return mine.cpu()
    padded = torch.full((cap,), -1, dtype=mine.dtype, device=mine.device)
    padded[:mine.numel()] = mine
    allb = torch.empty(world * cap, dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(allb, padded)
    allb = allb.cpu().view(world, cap)
    return torch.cat([allb[r, :c] for r, c in enumerate(counts)])


def global_count(n_local: int, world: int, device=None) -> int:
    """Sum of the per-rank shard sizes (global batch when the loader yields ragged shards)."""
    if world == 1:
        return n_local
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    return int(t.item())


def allreduce_step(grad_flats, scalars: torch.Tensor, world: int) -> None:
    """Sum the per-shard LoRA gradients (one flat buffer per trainable tower) and the
    (loss_sum, n_correct) pair. Every shard's loss was already scaled by 1/global_batch, so the
    sum IS the global-mean gradient: no division."""
    if world == 1:
        return
    if torch.is_tensor(grad_flats):
        grad_flats = [grad_flats]
    for g in grad_flats:
        dist.all_reduce(g)
    if scalars is not None:
        dist.all_reduce(scalars)
