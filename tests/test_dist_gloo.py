"""CPU test of the N>1 path: two processes (gloo) shard a batch, score their shard with the
ORACLE standing in for the GPU engine, and all-gather the labels; the result must equal the
single-process labels element for element (SURVEY.md §8e determinism check)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, ragged, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from radar_ml_b200.dist import allgather_labels, init_from_env, shard_range
    r, _, w = init_from_env("gloo")
    assert (r, w) == (rank, world)
    lo, hi = shard_range(total, rank, world)
    g = torch.Generator().manual_seed(99)
    all_labels = torch.randint(0, 3, (total,), generator=g, dtype=torch.int32)  # "scored" labels
    local = all_labels[lo:hi].clone()
    out = allgather_labels(local, total)
    ok = bool(torch.equal(out, all_labels))
    # preallocated output path (what bench.py uses for equal shards)
    if not ragged:
        buf = torch.empty((total,), dtype=torch.int32)
        allgather_labels(local, total, out=buf)
        ok = ok and bool(torch.equal(buf, all_labels))
    q.put((rank, ok, int(out.numel())))
    dist.destroy_process_group()


@pytest.mark.parametrize("total,ragged", [(64, False), (37, True)])
def test_two_rank_label_allgather(total, ragged):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, ragged, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res) and all(r[2] == total for r in res)
