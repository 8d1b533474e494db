"""Data-parallel sharding of independent scans across GPUs.

Scans never interact (predict.py:93-121 handles one target at a time), so the path shards
with no data-path collective: rank r owns the contiguous range ``shard_range(B, r, G)``; the
model is replicated at load.  The only exchange is ONE all-gather of the int32 labels per
batch (NCCL over NVLink on GPUs; gloo in the CPU tests).  One process per GPU, launched with
``python -m torch.distributed.run``.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous [lo, hi) of rank; the first ``total % world`` ranks get one extra scan."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int):
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun sets them)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def allgather_labels(local: torch.Tensor, total: int | None = None, out: torch.Tensor | None = None,
                     group=None) -> torch.Tensor:
    """Gather per-rank label shards (rank order = scan order) into one [total] tensor.

    Equal shards use a single ``all_gather_into_tensor`` straight into ``out``; ragged shards
    (total % world != 0) are padded to the largest shard and compacted afterwards.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        if out is not None:
            out.copy_(local)
            return out
        return local
    n = local.numel()
    if total is None:
        total = n * world
    sizes = shard_sizes(total, world)
    if len(set(sizes)) == 1:
        if out is None:
            out = torch.empty((total,), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    big = max(sizes)
    padded = torch.zeros((big,), dtype=local.dtype, device=local.device)
    padded[:n] = local
    buf = torch.empty((world * big,), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    parts = [buf[r * big:r * big + sizes[r]] for r in range(world)]
    res = torch.cat(parts)
    if out is not None:
        out.copy_(res)
        return out
    return res


class ShardedClassifier:
    """Each rank scores its shard on its own GPU, then all ranks get every label."""

    def __init__(self, engine, rank: int, world: int):
        self.engine, self.rank, self.world = engine, rank, world

    def predict_shard(self, local_cubes, total: int | None = None, **kw):
        proba, label, known = self.engine.predict(local_cubes, **kw)
        # non-integral voxels / bad slice indices must surface on the rank that saw them BEFORE
        # its labels are published to every other rank
        self.engine.check_status()
        labels_all = allgather_labels(label, total)
        return proba, label, known, labels_all
