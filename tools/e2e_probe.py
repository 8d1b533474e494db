#!/usr/bin/env python
"""rml_predict_host on pinned float32 cubes: with and without host-side narrowing, per-chunk trace."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import synth
from radar_ml_b200.engine import Engine
from radar_ml_b200.model import from_sklearn
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    cal = synth.standard_model()
eng = Engine(0)
eng.load_model(from_sklearn(cal))
n = 4096
dev = torch.device("cuda:0")
host = torch.empty((n, 22, 31, 176), dtype=torch.float32, pin_memory=True)
host.copy_(bench.device_cubes(n, 3, dev))
hn = host.numpy()
out = (np.empty((n, 3), np.float32), np.empty((n,), np.int32), np.empty((n,), np.uint8))
for mode in (True, False, True):
    eng.set_host_narrowing(mode)
    eng.predict_host(hn, out=out)
    os.environ.pop("RML_HOST_NARROW_TRACE", None)
    t0 = time.perf_counter()
    for _ in range(5):
        eng.predict_host(hn, out=out)
    dt = (time.perf_counter() - t0) / 5
    print("narrowing %s: %.2f ms per %d scans = %.1f k scans/s  %s" % (mode, dt * 1e3, n, n / dt / 1e3, eng.last_host_transfer()), flush=True)
