#!/usr/bin/env python
"""Throughput of the dnn.py / sgan.py forward pass (configs[2], configs[4]) on one GPU:
cubes resident in HBM -> labels.  Not the headline bench; prints one JSON line per network."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import device_cubes  # noqa: E402
from oracle import nets  # noqa: E402
from radar_ml_b200.engine import Engine  # noqa: E402
from radar_ml_b200.nets import GpuNetClassifier  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=4096)
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    eng = Engine(0)
    for kind, n in (("dnn", args.scans), ("sgan_c", args.scans // 4)):
        spec = nets.random_dnn(0) if kind == "dnn" else nets.random_sgan(0)
        net = GpuNetClassifier(spec, engine=eng, chunk=args.chunk)
        cubes = device_cubes(n, 7, eng.device)
        net.predict_cubes(cubes)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            net.predict_cubes(cubes)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        flop = 54690176 if kind == "dnn" else 512762240
        print(json.dumps({"net": kind, "scans": n, "chunk": args.chunk, "igemm": net.uses_igemm,
                          "scans_per_s": n / dt, "ms": dt * 1e3, "tflops": n * flop / dt / 1e12}), flush=True)


if __name__ == "__main__":
    main()
