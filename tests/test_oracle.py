"""CPU tests: the oracle (oracle/restate.py) against the golden vectors minted from the
UNMODIFIED reference (tests/golden/make_golden.py) and, when /root/reference is present,
against the reference itself."""
import json
import os

import numpy as np
import pytest

from oracle import refimport, restate, synth

G = os.path.join(os.path.dirname(__file__), "golden")


def load_model(z, prefix="m_"):
    sv = (z[prefix + "sv_u8"].astype(np.float32) / np.float32(255)).astype(np.float64)
    return restate.SvcParams(
        n_classes=int(z[prefix + "n_classes"]), gamma=float(z[prefix + "gamma"]), sv=sv,
        dual_coef=z[prefix + "dual_coef"], rho=z[prefix + "rho"], n_support=z[prefix + "n_support"],
        platt_a=z[prefix + "platt_a"], platt_b=z[prefix + "platt_b"])


@pytest.mark.parametrize("mode", ["max", "slice"])
def test_golden_svc(mode):
    z = np.load(os.path.join(G, "svc_%s.npz" % mode))
    cubes = z["cubes_u8"].astype(np.float32)
    p = load_model(z)
    X, lab, pr, known, P = restate.scan_path(cubes, p, mode=mode, ijk=z["ijk"])
    assert X.dtype == np.float32 and np.array_equal(X, z["ref_features"])      # bit-exact
    assert np.abs(P - z["sk_predict_proba"]).max() < 1e-12
    assert np.abs(restate.decision_function(X, p) - z["sk_decision"]).max() < 1e-11
    names = np.where(known, z["classes"][lab], "Unknown")
    assert list(names) == list(z["ref_names"])
    assert np.abs(pr - z["ref_proba"]).max() < 1e-12
    assert (names == "Unknown").any() and (names != "Unknown").any()   # both branches pinned


def test_golden_real_xy():
    z = np.load(os.path.join(G, "real_xy.npz"))
    p = load_model(z)
    xy = z["xy_u8"].astype(np.float32)
    te = z["test_idx"]
    mask = restate.ProjMask(False, False, True)
    X = restate.process_samples([(None, None, xy[i]) for i in te], proj_mask=mask, scale=True)
    assert X.shape == (len(te), 682) and np.array_equal(X, z["ref_features_test"])
    lab, pr, known, P = restate.classify_batch(X, p, 0.7)
    assert np.abs(P - z["sk_predict_proba"]).max() < 1e-12
    assert list(np.where(known, z["classes"][lab], "Unknown")) == list(z["ref_names"])
    # value distribution of the real sensor: integers, thresholded at 13
    assert xy.max() <= 255 and xy[xy > 0].min() >= 13 and 0.5 < (xy == 0).mean() < 0.8


def test_golden_generated_nonintegral():
    z = np.load(os.path.join(G, "generated.npz"))
    samples = [(z["xz"][i], z["yz"][i], z["xy"][i]) for i in range(z["xz"].shape[0])]
    for tag, mask in (("all", (True, True, True)), ("xz_xy", (True, False, True)), ("yz", (False, True, False))):
        for sc in (0, 1):
            got = restate.process_samples(samples, proj_mask=restate.ProjMask(*mask), scale=bool(sc))
            want = z["feat_%s_%d" % (tag, sc)]
            assert got.dtype == want.dtype == np.float32 and np.array_equal(got, want)


def test_golden_indices_and_shapes():
    z = np.load(os.path.join(G, "indices.npz"))
    got = np.array([restate.calculate_matrix_indices(*row, 22, 31, 176) for row in z["xyz"]])
    assert np.array_equal(got, z["ijk"])
    with open(os.path.join(G, "shapes.json")) as f:
        s = json.load(f)
    assert list(restate.arena_dims()) == s["raw_image"] == s["train_size"]
    sx, sy, sz = restate.arena_dims()
    assert sx * sz + sy * sz + sx * sy == s["feature_vector_length"]
    assert list(synth.CLASSES) == s["classes"]


def test_oracle_vs_sklearn_three_and_two_classes():
    import warnings
    for n_classes in (3, 2):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cubes, y, _ = synth.make_cubes(150, seed=5 + n_classes, n_classes=n_classes)
            X = synth.features(*synth.project_max(cubes))
            cal = synth.build_svc(X[:90], y[:90], X[90:120], y[90:120])
        p = restate.export_params(cal)
        assert np.abs(restate.predict_proba(X[120:], p) - cal.predict_proba(X[120:])).max() < 1e-12


def test_oracle_linear_vs_sklearn():
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cubes, y, _ = synth.make_cubes(150, seed=9)
        X = synth.features(*synth.project_max(cubes))
        cal = synth.build_linear(X[:90], y[:90], X[90:120], y[90:120])
    p = restate.export_params(cal)
    assert p.kind == "linear"
    # float64 input = what scikit-learn 0.24 (the reference's pin) computes for float32 features
    assert np.abs(restate.predict_proba(X[120:], p)
                  - cal.predict_proba(X[120:].astype(np.float64))).max() < 1e-12
    # scikit-learn 1.9 on float32 input rounds in float32: still far inside the 1e-5 tolerance
    assert np.abs(restate.predict_proba(X[120:], p) - cal.predict_proba(X[120:])).max() < 1e-5


@pytest.mark.skipif(not refimport.available(), reason="reference tree not present on this box")
def test_oracle_vs_live_reference():
    rc, rp = refimport.load()
    cubes, y, ijk = synth.make_cubes(6, seed=77)
    for mode in ("max", "slice"):
        for s in range(6):
            t = restate.project(cubes[s], mode, tuple(int(v) for v in ijk[s]))
            for mask in ((True, True, True), (False, True, True), (True, False, False)):
                a = rc.process_samples([t], proj_mask=rc.ProjMask(*mask),
                                       proj_zoom=rp.calc_proj_zoom(22, 31, 176, 22, 31, 176), scale=True)
                b = restate.process_samples([t], proj_mask=restate.ProjMask(*mask), scale=True)
                assert a.dtype == b.dtype and np.array_equal(a, b)
    assert rp.calc_proj_zoom(22, 31, 176, 11, 31, 88) == restate.calc_proj_zoom(22, 31, 176, 11, 31, 88)
    assert rc.calculate_matrix_indices(12.5, -3.0, 140.0, 22, 31, 176) == \
        restate.calculate_matrix_indices(12.5, -3.0, 140.0, 22, 31, 176)
    # constants of the layout contract
    assert (rc.R_MIN, rc.R_MAX, rc.R_RES, rc.RADAR_MAX) == (restate.R_MIN, restate.R_MAX, restate.R_RES, restate.RADAR_MAX)
    assert rc.ProjMask._fields == restate.ProjMask._fields == ("xz", "yz", "xy")
