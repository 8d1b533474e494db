"""Worker for tests/test_gpu_dist.py (launched with torch.distributed.run, one rank per GPU):
every rank scores its contiguous shard, labels are all-gathered over NCCL, and the result must
equal scoring the whole batch on one GPU, element for element."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from radar_ml_b200.dist import ShardedClassifier, init_from_env, shard_range  # noqa: E402
from radar_ml_b200.engine import Engine  # noqa: E402
from radar_ml_b200.model import ModelParams  # noqa: E402


def main():
    rank, local, world = init_from_env("nccl")
    z = np.load(os.path.join(ROOT, "tests", "golden", "svc_max.npz"))
    sv = (z["m_sv_u8"].astype(np.float32) / np.float32(255)).astype(np.float64)
    params = ModelParams(kind="svc_rbf", n_classes=3, n_features=sv.shape[1], classes=np.arange(3),
                         platt_a=z["m_platt_a"], platt_b=z["m_platt_b"], gamma=float(z["m_gamma"]),
                         sv=sv, dual_coef=z["m_dual_coef"], rho=z["m_rho"], n_support=z["m_n_support"])
    eng = Engine(local)
    eng.load_model(params)
    total = 1003                                   # ragged on purpose
    rng = np.random.default_rng(5)                 # same batch on every rank
    base = z["cubes_u8"].astype(np.float32)
    idx = rng.integers(0, base.shape[0], total)
    cubes = torch.from_numpy(base[idx] * (rng.random((total, 1, 1, 1)) > 0.1)).float().cuda(local)
    _, full_label, _ = eng.predict(cubes)
    lo, hi = shard_range(total, rank, world)
    sc = ShardedClassifier(eng, rank, world)
    _, _, _, gathered = sc.predict_shard(cubes[lo:hi].contiguous(), total)
    eng.check_status()
    torch.cuda.synchronize()
    ok = bool(torch.equal(gathered.cpu(), full_label.cpu())) and gathered.numel() == total
    # equal shards take the single all_gather_into_tensor path
    even = total - total % world
    lo2, hi2 = shard_range(even, rank, world)
    _, _, _, g2 = sc.predict_shard(cubes[lo2:hi2].contiguous(), even)
    ok = ok and bool(torch.equal(g2.cpu(), full_label.cpu()[:even]))
    if "--c-abi" in sys.argv:
        # the same exchange through the C ABI: labels written in place into the gather buffer by
        # rml_predict, then rml_allgather_labels (ncclAllGather via dlopen) — SURVEY.md §8b/§8e
        box = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(rank, world, box[0])
        n = even // world
        buf = torch.full((even,), -1, device="cuda", dtype=torch.int32)
        mine = buf[rank * n:(rank + 1) * n]
        out = (torch.empty((n, 3), device="cuda"), mine, torch.empty((n,), device="cuda", dtype=torch.uint8))
        eng.predict(cubes[lo2:hi2].contiguous(), out=out)
        eng.check_status()
        eng.allgather_labels(mine, buf)
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(buf.cpu(), full_label.cpu()[:even]))
        eng.lib.rml_comm_destroy(eng.ctx)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_OK" if int(flag.item()) == 1 else "DIST_MISMATCH", "world", world, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
