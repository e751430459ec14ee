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


_CAP = {}   # (world, device) -> agreed shard capacity of the padded label gather


def gather_labels(labels: torch.Tensor, world: int, device=None) -> torch.Tensor:
    """All ranks' label shards (ragged allowed: the last batch of a stream), for the class
    bookkeeping of methods/_trainer.py:404-416 which must see the GLOBAL batch. Returns a CPU
    tensor, rank-major. ONE collective and one host read per step: shards are padded with -1 to a
    capacity the ranks agree on once (re-agreed only if a shard ever exceeds it)."""
    if world == 1:
        return labels.cpu()
    mine = (labels.to(device) if device is not None else labels).contiguous()
    key = (world, str(mine.device))
    cap = _CAP.get(key, 0)
    # every rank must take the same branch: a shard larger than the agreed capacity is announced
    # through the gather itself (its first element is then -2 - count)
    if cap == 0:
        t = torch.tensor([mine.numel()], dtype=torch.int64, device=mine.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cap = _CAP[key] = max(int(t.item()), 1)
    over = mine.numel() > cap
    padded = torch.full((cap,), -1, dtype=torch.int64, device=mine.device)
    if over:
        padded[0] = -2 - mine.numel()
    else:
        padded[:mine.numel()] = mine
    allb = torch.empty(world * cap, dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(allb, padded)
    allb = allb.cpu().view(world, cap)
    if bool((allb[:, 0] < -1).any()):          # some shard outgrew the capacity: re-agree, redo
        _CAP[key] = int((-2 - allb[:, 0]).max().item())
        return gather_labels(labels, world, device)
    return torch.cat([row[row >= 0] for row in allb])


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
