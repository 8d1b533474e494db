#!/usr/bin/env python
"""Throughput of the float32-feature scorers: exact multi-digit tensor-core path vs the float64
CUDA-core path (the first version's only general-precision scorer)."""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402
from radar_ml_b200.engine import Engine  # noqa: E402
from radar_ml_b200.model import from_sklearn  # noqa: E402


def main():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cal = synth.standard_model()
    eng = Engine(0)
    eng.load_model(from_sklearn(cal))
    for n, exact in ((32768, False), (1024, True)):
        x = torch.rand((n, 10010), device="cuda") * (torch.rand((n, 10010), device="cuda") < 0.1)
        eng.score(x, None, 0.7, exact=exact)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            eng.score(x, None, 0.7, exact=exact)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(json.dumps({"scorer": "float64 CUDA cores" if exact else "multi-digit tcgen05", "scans": n,
                          "n_sv": int(eng.params.n_sv), "ms": dt * 1e3, "scans_per_s": n / dt}), flush=True)


if __name__ == "__main__":
    main()
