"""world_size-2 gloo test of the data-parallel protocol (lifelong_clip_b200/dp.py): rank::world
shards + per-shard loss scaled by 1/global_batch + ONE all-reduce(sum) of the flat LoRA gradient
reproduce the single-process gradient, and the class bookkeeping sees the global labels.
The per-shard compute is the fp64 oracle here (no GPU in this container); on the GPU box the same
protocol runs over NCCL with the CUDA path (tests/test_e2e_gpu.py covers the 1-GPU numerics)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vit_oracle as vo

CFG = vo.VIT_TINY
N, C, SEED = 6, 7, 17


def _inputs():
    rng = np.random.default_rng(SEED)
    images = rng.standard_normal((N, 3, CFG.image_size, CFG.image_size)).astype(np.float32)
    labels = rng.integers(0, C, size=(N,)).astype(np.int64)
    return images, labels, vo.synth_weights(CFG, SEED), vo.synth_text_features(C, CFG.embed_dim, 3)


def _flat(grads: dict) -> torch.Tensor:
    return torch.cat([torch.from_numpy(grads[k]).flatten() for k in sorted(grads)])


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lifelong_clip_b200 import dp
    assert dp.world_info() == (rank, world)
    images, labels, w, text = _inputs()
    x, y = dp.shard_batch(torch.from_numpy(images), torch.from_numpy(labels), rank, world)
    seen = dp.gather_labels(y, world)
    res = vo.online_step_oracle(x.numpy(), y.numpy(), w, text, CFG, inv_batch=1.0 / N)
    grad_flat = _flat(res["grads"])
    scal = torch.tensor([float(res["loss"]), float((res["pred"] == y.numpy()).sum())],
                        dtype=torch.float64)
    dp.allreduce_step(grad_flat, scal, world)
    torch.save({"grad": grad_flat, "scal": scal, "seen": seen},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    images, labels, w, text = _inputs()
    full = vo.online_step_oracle(images, labels, w, text, CFG)
    want = _flat(full["grads"])
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    assert torch.equal(r0["grad"], r1["grad"])            # replicas stay identical
    rel = float((r0["grad"] - want).norm() / want.norm())
    assert rel < 1e-10, rel                               # fp64: reduction order only
    assert abs(float(r0["scal"][0]) - float(full["loss"])) < 1e-12
    assert int(r0["scal"][1]) == int((full["pred"] == labels).sum())
    # gathered labels = rank-major concatenation of the rank::2 shards = a permutation of the batch
    assert sorted(r0["seen"].tolist()) == sorted(labels.tolist())
    assert r0["seen"].tolist() == labels[0::2].tolist() + labels[1::2].tolist()


def _ragged_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lifelong_clip_b200 import dp
    # last batch of a stream: rank 0 holds 3 labels, rank 1 none (ragged shards)
    mine = torch.tensor([5, 1, 5]) if rank == 0 else torch.empty(0, dtype=torch.int64)
    seen = dp.gather_labels(mine, world)
    total = dp.global_count(mine.numel(), world)
    # an empty shard contributes zeros and still joins the collectives
    g1 = torch.full((8,), float(rank + 1)) if mine.numel() else torch.zeros(8)
    g2 = torch.full((4,), 2.0) if mine.numel() else torch.zeros(4)
    scal = torch.tensor([0.5, 2.0]) if mine.numel() else torch.zeros(2)
    dp.allreduce_step([g1, g2], scal, world)
    torch.save({"seen": seen, "total": total, "g1": g1, "g2": g2, "scal": scal},
               os.path.join(out_dir, f"q{rank}.pt"))
    dist.destroy_process_group()


def test_ragged_shards_and_multi_tower_allreduce(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_ragged_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        q = torch.load(os.path.join(tmp_path, f"q{r}.pt"))
        assert q["seen"].tolist() == [5, 1, 5] and q["total"] == 3
        assert torch.equal(q["g1"], torch.ones(8)) and torch.equal(q["g2"], torch.full((4,), 2.0))
        assert q["scal"].tolist() == [0.5, 2.0]
