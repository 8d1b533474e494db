#!/usr/bin/env python
"""k6_tower alone (rml_net_forward on precomputed scaled projections): time vs batch, CUDA events."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import device_cubes  # noqa: E402
from oracle import nets  # noqa: E402
from radar_ml_b200.engine import Engine  # noqa: E402
from radar_ml_b200.nets import GpuNetClassifier  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "dnn"
eng = Engine(0)
spec = nets.random_dnn(0) if kind == "dnn" else nets.random_sgan(0)
cubes = device_cubes(8192, 7, eng.device)
eng.set_affine(127.5, 127.5, True)
feats_all = eng.project(cubes, mode="max")
eng.set_affine(0.0, 255.0, True)
del cubes
for n in (512, 1024, 2048, 4096, 8192):
    net = GpuNetClassifier(spec, engine=eng, chunk=n)
    ws = net._workspace()
    feats = feats_all[:n].contiguous()
    proba = torch.empty((n, 3), device="cuda"); label = torch.empty((n,), device="cuda", dtype=torch.int32)
    def run():
        rc = eng.lib.rml_net_forward(eng.ctx, C.c_void_p(feats.data_ptr()), n, C.c_void_p(ws.data_ptr()), ws.numel(),
                                     C.c_void_p(proba.data_ptr()), None, C.c_void_p(label.data_ptr()), eng._stream())
        assert rc == 0, eng.lib.rml_last_error(eng.ctx)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%s towers+dense  n=%5d  %.3f ms  %.2f M scans/s  (%.1f us per 1024)" % (kind, n, ms, n / ms / 1e3, ms * 1e3 * 1024 / n), flush=True)
