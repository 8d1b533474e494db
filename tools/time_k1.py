#!/usr/bin/env python
"""K1 alone (rml_project, 148 SMs): float32 cubes -> u8 operand rows / float32 feature rows."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radar_ml_b200.engine import Engine
from radar_ml_b200 import _lib
import bench

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    eng = Engine(0)
    cubes = bench.device_cubes(n, 7, torch.device("cuda:0"))
    for name, dt in (("u8 rows", _lib.U8), ("f32 rows", _lib.F32)):
        out = eng.project(cubes, dtype=dt)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            eng.project(cubes, dtype=dt, out=out[0] if dt == _lib.U8 else out, norms=out[1] if dt == _lib.U8 else None)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        print("K1 %-9s n=%d  %.3f ms  %.2f TB/s  %.2f M scans/s" % (name, n, ms, n * 480128 / ms / 1e9, n / ms / 1e3))

main()
