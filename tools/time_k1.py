#!/usr/bin/env python
"""K1 alone (rml_project, 148 SMs): float32 cubes -> u8 operand rows / float32 feature rows; and the
DerivedTarget axis sums as a separate pass against the fused one (rml_project_derive)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from radar_ml_b200 import _lib  # noqa: E402
from radar_ml_b200.engine import Engine  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    eng = Engine(0)
    cubes = bench.device_cubes(n, 7, torch.device("cuda:0"))
    for name, dt in (("u8 rows", _lib.U8), ("f32 rows", _lib.F32)):
        out = eng.project(cubes, dtype=dt)
        ms = timed(lambda: eng.project(cubes, dtype=dt, out=out[0] if dt == _lib.U8 else out,
                                       norms=out[1] if dt == _lib.U8 else None))
        print("K1 %-9s n=%d  %.3f ms  %.2f TB/s  %.2f M scans/s" % (name, n, ms, n * 480128 / ms / 1e9, n / ms / 1e3))
    t_p = timed(lambda: eng.project(cubes, dtype=_lib.U8))
    t_d = timed(lambda: eng.derive_targets(cubes, num_targets=3))
    t_f = timed(lambda: eng.project_derive(cubes, num_targets=3, dtype=_lib.U8))
    print("project (u8 rows) %.3f ms + derive_targets %.3f ms = %.3f ms;  project_derive (one pass) %.3f ms = %.2f TB/s"
          % (t_p, t_d, t_p + t_d, t_f, n * 480128 / t_f / 1e9))
    # where do the two differ (debug aid)
    ijk_r, s_r = eng.derive_targets(cubes[:512], num_targets=3, want_sums=True)
    _, ijk_f, s_f = eng.project_derive(cubes[:512], num_targets=3, dtype=_lib.F32, want_sums=True)
    torch.cuda.synchronize()
    bad = (s_r != s_f).nonzero()
    print("sum mismatches:", bad.shape[0], "first:", bad[:8].tolist(),
          [(float(s_r[i, j]), float(s_f[i, j])) for i, j in bad[:8].tolist()])
    print("ijk mismatches:", int((ijk_r != ijk_f).sum()))


main()
