#!/usr/bin/env python
"""uint8 cubes resident in HBM: scans/s of rml_predict_u8 for several K1/K2 SM splits of the fused
pipeline, the serial order, and K1 alone (GB/s against the 120 032 B/scan it must read)."""
import argparse
import ctypes as C
import json
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import device_cubes  # noqa: E402
from oracle import synth  # noqa: E402
from radar_ml_b200.engine import Engine  # noqa: E402
from radar_ml_b200.model import from_sklearn  # noqa: E402


def timed(fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--splits", default="24,48")
    ap.add_argument("--only-default", action="store_true", help="profiling aid: one configuration")
    args = ap.parse_args()
    B = args.scans
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cal = synth.standard_model(n_train=int(os.environ.get("RML_BENCH_TRAIN", "909")), mode="max")
    eng = Engine(0)
    eng.load_model(from_sklearn(cal))
    cubes8 = torch.empty((B, 22, 31, 176), device=eng.device, dtype=torch.uint8)
    for lo in range(0, B, 4096):
        n = min(4096, B - lo)
        cubes8[lo:lo + n] = device_cubes(n, 1234 + lo, eng.device).to(torch.uint8)
    out = eng.predict(cubes8)
    eng.check_status()
    ref = tuple(t.clone() for t in out)
    rows = []
    if args.only_default:
        ms = timed(lambda: eng.predict(cubes8, out=out), args.steps)
        print(json.dumps({"scans": B, "ms": ms, "scans_per_s": B / ms * 1e3}), flush=True)
        return
    for sms in [int(x) for x in args.splits.split(",")]:
        assert eng.lib.rml_set_fused_u8(eng.ctx, sms) == 0
        eng.predict(cubes8, out=out)
        ms = timed(lambda: eng.predict(cubes8, out=out), args.steps)
        eng.check_status()
        ok = all(torch.equal(a, b) for a, b in zip(out, ref))
        rows.append({"k2_sms": sms, "ms": ms, "scans_per_s": B / ms * 1e3, "identical": ok})
    assert eng.lib.rml_set_fused_u8(eng.ctx, 0) == 0
    eng.predict(cubes8, out=out)
    ms = timed(lambda: eng.predict(cubes8, out=out), args.steps)
    rows.append({"k2_sms": "serial", "ms": ms, "scans_per_s": B / ms * 1e3,
                 "identical": all(torch.equal(a, b) for a, b in zip(out, ref))})
    # K1 alone, u8 rows and f32 rows
    q, norms = eng.project(cubes8, dtype=1)
    ms = timed(lambda: eng.project(cubes8, dtype=1, out=q, norms=norms), args.steps)
    rows.append({"k1_u8in_u8out_ms": ms, "GBps": B * 120032 / ms / 1e6, "scans_per_s": B / ms * 1e3})
    del q, norms
    Bf = min(B, 16384)
    f = eng.project(cubes8[:Bf])
    ms = timed(lambda: eng.project(cubes8[:Bf], out=f), args.steps)
    rows.append({"k1_u8in_f32out_ms": ms, "scans": Bf, "GBps": Bf * 120032 / ms / 1e6, "scans_per_s": Bf / ms * 1e3})
    for r in rows:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
