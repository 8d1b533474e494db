#!/usr/bin/env python
"""A small pass over every hand-written kernel family, sized for compute-sanitizer (memcheck /
racecheck / synccheck): projection kernels (f32 and u8 cubes, MAX and SLICE), the integer and
multi-digit tensor-core scorers, the co-resident K1 || K2 pipeline (forced at a small batch), the
one-scan host entry points, and the dnn / sgan towers + dense stack.  Results are checked against
the CPU oracle so that a sanitizer run is also a parity run."""
import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nets, restate, synth  # noqa: E402
from radar_ml_b200.engine import Engine  # noqa: E402
from radar_ml_b200.model import from_sklearn  # noqa: E402
from radar_ml_b200.nets import GpuNetClassifier  # noqa: E402

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    cubes, y, ijk = synth.make_cubes(160 + 300, seed=5)
    X = synth.features(*synth.project_max(cubes))
    cal = synth.build_svc(X[:128], y[:128], X[128:160], y[128:160])
p = restate.export_params(cal)
eng = Engine(0)
eng.load_model(from_sklearn(cal))
test = cubes[160:]
d = torch.from_numpy(test).cuda()
_, lab_o, _, known_o, P_o = restate.scan_path(test, p, mode="max")
# serial K1 -> K2
proba, label, known = eng.predict(d)
eng.check_status()
assert np.array_equal(label.cpu().numpy(), lab_o)
# co-resident pipeline forced at this batch size
eng.lib.rml_set_fused(eng.ctx, 1, 32, 128)
proba, label, known = eng.predict(d)
eng.check_status()
torch.cuda.synchronize()
assert np.array_equal(label.cpu().numpy(), lab_o) and np.abs(proba.cpu().numpy() - P_o).max() < 1e-5
eng.lib.rml_set_fused(eng.ctx, 1, 32, 8192)
# uint8 cubes, SLICE mode, host entries
p8, l8, _ = eng.predict(torch.from_numpy(test.astype(np.uint8)).cuda())
eng.check_status()
assert torch.equal(l8, label)
_, lab_s, _, _, _ = restate.scan_path(test[:40], p, mode="slice", ijk=ijk[160:200])
ps, ls, _ = eng.predict(d[:40], mode="slice", ijk=torch.from_numpy(ijk[160:200]).cuda())
eng.check_status()
assert np.array_equal(ls.cpu().numpy(), lab_s)
Ph, lh, _ = eng.predict_host(test[:64])
assert np.array_equal(lh, lab_o[:64])
Pt, lt, _ = eng.predict_targets_host(test[0], ijk[160:164])
Pg, lg, _ = eng.score_features_host(X[160:200])
# DerivedTarget axis sums inside the projection pass (both output types) against the stand-alone kernel
from radar_ml_b200._lib import F32 as _F32, U8 as _U8  # noqa: E402
ijk_ref = eng.derive_targets(d, num_targets=3)
f_a, ijk_a = eng.project_derive(d, num_targets=3, dtype=_F32)
u_b, n_b, ijk_b = eng.project_derive(d, num_targets=3, dtype=_U8)
eng.check_status()
assert torch.equal(ijk_a, ijk_ref) and torch.equal(ijk_b, ijk_ref) and torch.equal(f_a, eng.project(d, dtype=_F32))
# real-valued cubes -> float32 features -> multi-digit scorer
eng.set_precision(True)
real = (test[:64] * 0.73).astype(np.float32)
_, lab_r, _, _, P_r = restate.scan_path(real, p, mode="max")
pr, lr, _ = eng.predict(torch.from_numpy(real).cuda())
eng.check_status()
assert np.array_equal(lr.cpu().numpy(), lab_r) and np.abs(pr.cpu().numpy() - P_r).max() < 1e-5
eng.set_precision(False)
# networks
for kind, n in (("dnn", 24), ("sgan_c", 8)):
    spec = nets.random_dnn(1) if kind == "dnn" else nets.random_sgan(1)
    net = GpuNetClassifier(spec, engine=eng, chunk=16)
    pn, ln = net.predict_cubes(d[:n])
    torch.cuda.synchronize()
    xz, yz, xy = synth.project_max(test[:n])
    Xn = nets.preprocess([(xz[i], yz[i], xy[i]) for i in range(n)], spec.R)
    P_bf, _ = nets.forward_bf16_towers(spec, Xn)
    assert np.abs(pn.cpu().numpy() - P_bf).max() < 5e-4, kind
print("san_small ok")
