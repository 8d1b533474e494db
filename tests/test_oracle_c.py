"""CPU tests: the plain-C oracle (oracle/c/radar_oracle.c) against the golden vectors minted
from the UNMODIFIED reference (tests/golden/) and against the numpy oracle (oracle/restate.py) —
two independent restatements of the same path must agree bit for bit on features / labels and to
1e-12 on probabilities."""
import os
import warnings

import numpy as np
import pytest

from oracle import c_oracle, restate, synth

G = os.path.join(os.path.dirname(__file__), "golden")


def load_model(z, prefix="m_"):
    sv = (z[prefix + "sv_u8"].astype(np.float32) / np.float32(255)).astype(np.float64)
    return restate.SvcParams(
        n_classes=int(z[prefix + "n_classes"]), gamma=float(z[prefix + "gamma"]), sv=sv,
        dual_coef=z[prefix + "dual_coef"], rho=z[prefix + "rho"], n_support=z[prefix + "n_support"],
        platt_a=z[prefix + "platt_a"], platt_b=z[prefix + "platt_b"])


@pytest.mark.parametrize("mode", ["max", "slice"])
def test_c_oracle_golden_svc(mode):
    z = np.load(os.path.join(G, "svc_%s.npz" % mode))
    cubes = z["cubes_u8"].astype(np.float32)
    p = load_model(z)
    X, lab, pr, known, P = c_oracle.scan_path(cubes, p, mode=mode, ijk=z["ijk"])
    assert X.dtype == np.float32 and np.array_equal(X, z["ref_features"])      # bit-exact
    assert np.abs(P - z["sk_predict_proba"]).max() < 1e-12
    _, dec = c_oracle.predict_proba(X, p, want_decision=True)
    assert np.abs(dec - z["sk_decision"]).max() < 1e-11
    names = np.where(known, z["classes"][lab], "Unknown")
    assert list(names) == list(z["ref_names"])
    assert np.abs(pr - z["ref_proba"]).max() < 1e-12


def test_c_oracle_golden_real_xy_and_generated():
    z = np.load(os.path.join(G, "real_xy.npz"))
    p = load_model(z)
    xy = z["xy_u8"].astype(np.float32)
    te = z["test_idx"]
    X = c_oracle.process_samples([(None, None, xy[i]) for i in te], mask=(False, False, True), scale=True)
    assert X.shape == (len(te), 682) and np.array_equal(X, z["ref_features_test"])
    P = c_oracle.predict_proba(X, p)
    assert np.abs(P - z["sk_predict_proba"]).max() < 1e-12
    g = np.load(os.path.join(G, "generated.npz"))
    samples = [(g["xz"][i], g["yz"][i], g["xy"][i]) for i in range(g["xz"].shape[0])]
    for tag, mask in (("all", (True, True, True)), ("xz_xy", (True, False, True)), ("yz", (False, True, False))):
        for sc in (0, 1):
            got = c_oracle.process_samples(samples, mask=mask, scale=bool(sc))
            want = g["feat_%s_%d" % (tag, sc)]
            assert got.dtype == want.dtype == np.float32 and np.array_equal(got, want)


def test_c_oracle_golden_indices():
    z = np.load(os.path.join(G, "indices.npz"))
    got = np.array([c_oracle.matrix_indices(*row, 22, 31, 176) for row in z["xyz"]])
    assert np.array_equal(got, z["ijk"])


@pytest.mark.parametrize("n_classes", [3, 2])
def test_c_oracle_equals_numpy_oracle_and_sklearn(n_classes):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cubes, y, ijk = synth.make_cubes(150, seed=5 + n_classes, n_classes=n_classes)
        X = synth.features(*synth.project_max(cubes))
        cal = synth.build_svc(X[:90], y[:90], X[90:120], y[90:120])
    p = restate.export_params(cal)
    for mode, ij in (("max", None), ("slice", ijk[120:])):
        Xn, labn, prn, knownn, Pn = restate.scan_path(cubes[120:], p, mode=mode, ijk=ij)
        Xc, labc, prc, knownc, Pc = c_oracle.scan_path(cubes[120:], p, mode=mode, ijk=ij)
        assert np.array_equal(Xn, Xc) and np.array_equal(labn, labc) and np.array_equal(knownn, knownc)
        assert np.abs(Pn - Pc).max() < 1e-12
    assert np.abs(c_oracle.predict_proba(X[120:], p) - cal.predict_proba(X[120:])).max() < 1e-12
    Pd, dec = c_oracle.predict_proba(X[120:], p, want_decision=True)
    assert np.abs(dec - restate.decision_function(X[120:], p)).max() < 1e-11


def test_c_oracle_linear_masks_and_index_errors():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cubes, y, ijk = synth.make_cubes(150, seed=9)
        X = synth.features(*synth.project_max(cubes))
        cal = synth.build_linear(X[:90], y[:90], X[90:120], y[90:120])
    p = restate.export_params(cal)
    assert p.kind == "linear"
    assert np.abs(c_oracle.predict_proba(X[120:], p) - restate.predict_proba(X[120:], p)).max() < 1e-12
    # projections: every mask, numpy negative-index wrap, IndexError beyond it
    t = (5, -2, -176)
    for mode, ij in (("max", None), ("slice", t)):
        a = restate.project(cubes[0], mode, ij)
        b = c_oracle.project(cubes[0], mode, ij)
        for u, v in zip(a, b):
            assert np.array_equal(np.asarray(u), v)
        for mask in ((True, True, True), (True, False, True), (False, True, False), (False, False, True)):
            want = restate.process_samples([a], proj_mask=restate.ProjMask(*mask), scale=True)
            got = c_oracle.process_samples([b], mask=mask, scale=True)
            assert want.dtype == got.dtype == np.float32 and np.array_equal(want, got)
    with pytest.raises(IndexError):
        c_oracle.project(cubes[0], "slice", (22, 0, 0))
    with pytest.raises(IndexError):
        c_oracle.project(cubes[0], "slice", (0, 0, -177))
    with pytest.raises(IndexError):
        cubes[0][:, :, -177]            # what numpy (the reference) does with the same index


@pytest.mark.parametrize("seed", range(6))
def test_c_oracle_random_arenas_models_and_values(seed):
    """Seeded random arenas (odd sizes), real-valued / negative voxels, random calibrated models with
    2..6 classes: the two restatements must agree everywhere the reference's arithmetic is defined."""
    rng = np.random.default_rng(100 + seed)
    sx, sy, sz = (int(v) for v in rng.integers(2, 13, size=3))
    n = 17
    cubes = rng.normal(40.0, 70.0, size=(n, sx, sy, sz)).astype(np.float32)
    if seed % 2:
        cubes = np.rint(np.clip(cubes, 0, 255)).astype(np.float32)
    ijk = np.stack([rng.integers(-d, d, size=n) for d in (sx, sy, sz)], axis=1).astype(np.int32)
    mask = [(True, True, True), (True, False, True), (False, True, False), (False, False, True),
            (True, True, False), (False, True, True)][seed]
    F = (sx * sz if mask[0] else 0) + (sy * sz if mask[1] else 0) + (sx * sy if mask[2] else 0)
    C = 2 + seed % 5
    n_support = rng.integers(1, 5, size=C).astype(np.int32)
    n_sv = int(n_support.sum())
    p = restate.SvcParams(
        n_classes=C, gamma=float(rng.uniform(1e-4, 5e-2)),
        sv=rng.uniform(0, 1, size=(n_sv, F)), dual_coef=rng.normal(0, 2, size=(C - 1, n_sv)),
        rho=rng.normal(0, 1, size=C * (C - 1) // 2), n_support=n_support,
        platt_a=rng.normal(-1.5, 0.5, size=(1 if C == 2 else C)),
        platt_b=rng.normal(0, 0.5, size=(1 if C == 2 else C)), classes=np.arange(C))
    for mode, ij in (("max", None), ("slice", ijk)):
        Xn, labn, prn, knownn, Pn = restate.scan_path(cubes, p, mode=mode, ijk=ij, mask=restate.ProjMask(*mask))
        Xc, labc, prc, knownc, Pc = c_oracle.scan_path(cubes, p, mode=mode, ijk=ij, mask=mask)
        assert Xn.dtype == Xc.dtype == np.float32 and np.array_equal(Xn, Xc)
        assert np.abs(Pn - Pc).max() < 1e-12
        assert np.array_equal(labn, labc) and np.array_equal(knownn, knownc)
        assert np.abs(Pc.sum(axis=1) - 1.0).max() < 1e-12
