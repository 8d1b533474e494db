#!/usr/bin/env python
"""Reference-exact SLICE mode (predict.py:102-107) at configs[1] size: scans/s with cubes resident."""
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


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cal = synth.standard_model(mode="slice")
    eng = Engine(0)
    eng.load_model(from_sklearn(cal))
    cubes = device_cubes(B, 5, eng.device)
    ijk = eng.derive_targets(cubes, 1)[:, 0, :].contiguous()        # targets from the cube itself
    out = eng.predict(cubes, mode="slice", ijk=ijk)
    eng.check_status()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        eng.predict(cubes, mode="slice", ijk=ijk, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        eng.derive_targets(cubes, 1)
    t1.record()
    torch.cuda.synchronize()
    print(json.dumps({"mode": "slice", "scans": B, "ms_per_step": ms, "scans_per_s": B / ms * 1e3,
                      "derive_targets_ms": t0.elapsed_time(t1) / 5,
                      "derive_targets_GBps": B * 480128 / (t0.elapsed_time(t1) / 5 * 1e-3) / 1e9}), flush=True)


if __name__ == "__main__":
    main()
