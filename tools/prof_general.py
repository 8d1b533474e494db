#!/usr/bin/env python
"""Per-kernel device times of the general-precision path (real-valued cubes, non-integral support
vectors): K1 float32 rows -> digit planes -> exact multi-digit tcgen05 scorer."""
import copy, os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import device_cubes  # noqa: E402
from oracle import synth  # noqa: E402
from radar_ml_b200.engine import Engine  # noqa: E402
from radar_ml_b200.model import from_sklearn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    cal = synth.standard_model()
p2 = copy.deepcopy(from_sklearn(cal))
p2.sv = np.clip(p2.sv + 3e-5 * np.sin(np.arange(p2.sv.size, dtype=np.float64)).reshape(p2.sv.shape), 0.0, 1.0)
eng = Engine(0)
eng.load_model(p2)
cg = device_cubes(n, 777, eng.device, integer=False)
out = eng.predict(cg)
eng.check_status()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        eng.predict(cg, out=out)
    torch.cuda.synchronize()
agg = {}
for ev in prof.events():
    if ev.device_type.name == "CUDA":
        a = agg.setdefault(ev.name[:70], [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-70s x%-3d avg %8.1f us  %4.1f%%" % (k, v[0], v[1] / v[0], 100 * v[1] / tot))
print("  %d scans: sum of kernels per pass %.3f ms = %.2f M scans/s" % (n, tot / 3 / 1e3, n / (tot / 3 / 1e3) / 1e3))
