#!/usr/bin/env python
"""Per-kernel device times of the dnn / sgan forward in a REAL back-to-back run (CUPTI through
torch.profiler; ncu serialises and cold-starts every launch) + SM clock under load."""
import argparse, os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import device_cubes, ClockSampler  # noqa: E402
from oracle import nets  # noqa: E402
from radar_ml_b200.engine import Engine  # noqa: E402
from radar_ml_b200.nets import GpuNetClassifier  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scans", type=int, default=16384)
ap.add_argument("--chunk", type=int, default=4096)
ap.add_argument("--kind", default="dnn")
args = ap.parse_args()
eng = Engine(0)
spec = nets.random_dnn(0) if args.kind == "dnn" else nets.random_sgan(0)
net = GpuNetClassifier(spec, engine=eng, chunk=args.chunk)
cubes = device_cubes(args.scans, 7, eng.device)
for _ in range(3):
    net.predict_cubes(cubes)
torch.cuda.synchronize()
cs = ClockSampler(0)
t0 = time.perf_counter()
for _ in range(10):
    net.predict_cubes(cubes)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
clk = cs.stop()
print("%s scans=%d chunk=%d: %.3f ms/pass = %.3f M scans/s; clocks %s" % (args.kind, args.scans, args.chunk, dt * 1e3, args.scans / dt / 1e6, clk))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        net.predict_cubes(cubes)
    torch.cuda.synchronize()
agg = {}
for ev in prof.events():
    if ev.device_type.name == "CUDA":
        a = agg.setdefault(ev.name[:60], [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-60s x%-3d total %9.1f us  avg %8.1f us  %4.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
print("  sum of kernels per pass: %.3f ms" % (tot / 3 / 1e3))
