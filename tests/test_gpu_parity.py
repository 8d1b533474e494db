"""-m gpu parity tests: CUDA path (through the C ABI) vs the CPU oracle on the same inputs.
Bit-exact for projections / u8 features / labels; probabilities within 1e-5 (north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PROBA_TOL = 1e-5


@pytest.fixture(scope="module")
def eng():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from radar_ml_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _oracle_feats(cubes, mode, ijk, mask, scale=True):
    from oracle import restate
    out = []
    for s in range(cubes.shape[0]):
        t = restate.project(cubes[s], mode, None if ijk is None else tuple(int(v) for v in ijk[s]))
        out.append(restate.process_samples([t], proj_mask=restate.ProjMask(*mask), scale=scale)[0])
    return np.asarray(out, dtype=np.float32)


@pytest.mark.parametrize("mask", [(True, True, True), (True, False, False), (False, True, False),
                                  (False, False, True), (True, False, True), (False, True, True)])
def test_k1_max_f32_bit_exact(eng, small_problem, mask):
    import torch
    cubes = small_problem["cubes"][:150]
    d = torch.from_numpy(cubes).cuda()
    got = eng.project(d, mode="max", mask=mask).cpu().numpy()
    want = _oracle_feats(cubes, "max", None, mask)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want)


def test_k1_max_u8_bit_exact_and_norms(eng, small_problem):
    import torch
    cubes = small_problem["cubes"][:301]   # not a multiple of the grid
    d = torch.from_numpy(cubes).cuda()
    q, norms = eng.project(d, mode="max", dtype=1)
    eng.check_status()
    want = _oracle_feats(cubes, "max", None, (True, True, True), scale=False)
    q = q.cpu().numpy()
    assert q.shape == (301, 10112)
    assert np.array_equal(q[:, :10010].astype(np.float32), want)
    assert not q[:, 10010:].any()
    assert np.array_equal(norms.cpu().numpy().astype(np.int64),
                          (want.astype(np.int64) ** 2).sum(axis=1))


def test_k1_unscaled_and_negative_values(eng):
    import torch
    rng = np.random.default_rng(5)
    cubes = rng.normal(50.0, 60.0, size=(9, 22, 31, 176)).astype(np.float32)
    d = torch.from_numpy(cubes).cuda()
    eng.set_affine(0.0, 255.0, False)
    try:
        got = eng.project(d, mode="max").cpu().numpy()
    finally:
        eng.set_affine(0.0, 255.0, True)
    want = _oracle_feats(cubes, "max", None, (True, True, True), scale=False)
    assert np.array_equal(got, want)
    # the u8 path must refuse non-integral data loudly
    from radar_ml_b200._lib import NonIntegralInput
    eng.project(d, mode="max", dtype=1)
    with pytest.raises(NonIntegralInput):
        eng.check_status()


@pytest.mark.parametrize("dtype", [0, 1])
def test_k1_slice_bit_exact(eng, small_problem, dtype):
    import torch
    cubes = small_problem["cubes"][:100]
    ijk = small_problem["ijk"][:100].copy()
    ijk[0] = (0, 0, 0)
    ijk[1] = (21, 30, 175)
    ijk[2] = (-1, -2, -3)          # numpy negative-index wrap (predict.py:102-107 on ndarray)
    d = torch.from_numpy(cubes).cuda()
    res = eng.project(d, mode="slice", ijk=torch.from_numpy(ijk).cuda(), dtype=dtype)
    eng.check_status()
    want = _oracle_feats(cubes, "slice", ijk, (True, True, True), scale=(dtype == 0))
    got = res.cpu().numpy() if dtype == 0 else res[0].cpu().numpy()[:, :10010].astype(np.float32)
    assert np.array_equal(got, want)


def test_k2_rbf_i8_matches_oracle(eng, small_problem):
    import torch
    from oracle import restate
    from radar_ml_b200.model import from_sklearn
    eng.load_model(from_sklearn(small_problem["cal"]))
    assert eng.model_is_integral
    cubes = small_problem["cubes"][small_problem["test"]]
    d = torch.from_numpy(cubes).cuda()
    q, norms = eng.project(d, mode="max", dtype=1)
    proba, label, known, dec = eng.score(q, norms, min_proba=0.7, want_decision=True)
    eng.check_status()
    p = small_problem["params"]
    X = small_problem["X"][small_problem["test"]]
    lab_o, pr_o, known_o, P_o = restate.classify_batch(X, p, 0.7)
    dec_o = restate.decision_function(X, p)
    assert np.abs(dec.cpu().numpy() - dec_o).max() < 1e-5
    assert np.abs(proba.cpu().numpy().astype(np.float64) - P_o).max() < PROBA_TOL
    assert np.array_equal(label.cpu().numpy(), lab_o)
    assert np.array_equal(known.cpu().numpy(), known_o)


def test_k2_general_matches_oracle(eng, small_problem):
    import torch
    from oracle import restate
    from radar_ml_b200.model import from_sklearn
    eng.load_model(from_sklearn(small_problem["cal"]))
    X = small_problem["X"][small_problem["test"]]
    proba, label, known = eng.score(torch.from_numpy(X).cuda(), None, min_proba=0.7)
    lab_o, pr_o, known_o, P_o = restate.classify_batch(X, small_problem["params"], 0.7)
    assert np.abs(proba.cpu().numpy().astype(np.float64) - P_o).max() < PROBA_TOL
    assert np.array_equal(label.cpu().numpy(), lab_o)
    assert np.array_equal(known.cpu().numpy(), known_o)


def test_predict_pipeline_and_host_entry(eng, small_problem):
    import torch
    from oracle import restate
    from radar_ml_b200.model import from_sklearn
    eng.load_model(from_sklearn(small_problem["cal"]))
    cubes = small_problem["cubes"][small_problem["test"]]
    p = small_problem["params"]
    _, lab_o, pr_o, known_o, P_o = restate.scan_path(cubes, p, mode="max")
    proba, label, known = eng.predict(torch.from_numpy(cubes).cuda(), mode="max")
    eng.check_status()
    assert np.array_equal(label.cpu().numpy(), lab_o)
    assert np.abs(proba.cpu().numpy().astype(np.float64) - P_o).max() < PROBA_TOL
    ph, lh, kh = eng.predict_host(np.ascontiguousarray(cubes), mode="max")
    assert np.array_equal(lh, lab_o) and np.array_equal(kh.astype(bool), known_o)
    assert np.array_equal(ph, proba.cpu().numpy())


# --------------------------------------------------------------------------- golden fixtures
import os  # noqa: E402

G = os.path.join(os.path.dirname(__file__), "golden")


def _golden_model(z, prefix="m_"):
    from radar_ml_b200.model import ModelParams
    sv = (z[prefix + "sv_u8"].astype(np.float32) / np.float32(255)).astype(np.float64)
    return ModelParams(kind="svc_rbf", n_classes=int(z[prefix + "n_classes"]),
                       n_features=sv.shape[1], classes=np.arange(int(z[prefix + "n_classes"])),
                       platt_a=z[prefix + "platt_a"], platt_b=z[prefix + "platt_b"],
                       gamma=float(z[prefix + "gamma"]), sv=sv, dual_coef=z[prefix + "dual_coef"],
                       rho=z[prefix + "rho"], n_support=z[prefix + "n_support"])


@pytest.mark.parametrize("mode", ["max", "slice"])
def test_golden_reference_outputs(eng, mode):
    """CUDA path vs outputs of the UNMODIFIED reference (tests/golden/make_golden.py)."""
    import torch
    z = np.load(os.path.join(G, "svc_%s.npz" % mode))
    eng.load_model(_golden_model(z))
    cubes = torch.from_numpy(z["cubes_u8"].astype(np.float32)).cuda()
    ijk = torch.from_numpy(z["ijk"]).cuda()
    feats = eng.project(cubes, mode=mode, ijk=ijk).cpu().numpy()
    assert np.array_equal(feats, z["ref_features"])                    # process_samples, bit-exact
    proba, label, known = eng.predict(cubes, mode=mode, ijk=ijk, min_proba=0.7)
    eng.check_status()
    P = proba.cpu().numpy().astype(np.float64)
    assert np.abs(P - z["sk_predict_proba"]).max() < PROBA_TOL
    names = np.where(known.cpu().numpy().astype(bool), z["classes"][label.cpu().numpy()], "Unknown")
    assert list(names) == list(z["ref_names"])                         # predict.classifier names
    best = P[np.arange(len(P)), label.cpu().numpy()]
    assert np.abs(best - z["ref_proba"]).max() < PROBA_TOL


def test_golden_real_sensor_xy_only(eng):
    """491 real xy projections from ground_truth_samples.log, ProjMask(False, False, True)."""
    import torch
    z = np.load(os.path.join(G, "real_xy.npz"))
    eng.load_model(_golden_model(z))
    xy = torch.from_numpy(z["xy_u8"][z["test_idx"]].astype(np.float32)).cuda()
    feats = eng.process_samples(None, None, xy, mask=(False, False, True), scale=True)
    assert np.array_equal(feats.cpu().numpy(), z["ref_features_test"])
    q, norms = eng.quantize(feats)
    eng.check_status()
    assert q.shape[1] == 768
    proba, label, known = eng.score(q, norms, 0.7)
    P = proba.cpu().numpy().astype(np.float64)
    assert np.abs(P - z["sk_predict_proba"]).max() < PROBA_TOL
    names = np.where(known.cpu().numpy(), z["classes"][label.cpu().numpy()], "Unknown")
    assert list(names) == list(z["ref_names"])
    # same thing through the general-precision kernel
    proba2, label2, _ = eng.score(feats, None, 0.7)
    assert np.array_equal(label2.cpu().numpy(), label.cpu().numpy())
    assert np.abs(proba2.cpu().numpy() - proba.cpu().numpy()).max() < PROBA_TOL


def test_golden_generated_nonintegral_process_samples(eng):
    import torch
    z = np.load(os.path.join(G, "generated.npz"))
    xz, yz, xy = (torch.from_numpy(z[k]).cuda() for k in ("xz", "yz", "xy"))
    for tag, mask in (("all", (True, True, True)), ("xz_xy", (True, False, True)), ("yz", (False, True, False))):
        for sc in (0, 1):
            got = eng.process_samples(xz if mask[0] else None, yz if mask[1] else None,
                                      xy if mask[2] else None, mask=mask, scale=bool(sc))
            assert np.array_equal(got.cpu().numpy(), z["feat_%s_%d" % (tag, sc)])


def test_golden_matrix_indices(eng):
    import torch
    z = np.load(os.path.join(G, "indices.npz"))
    got = eng.matrix_indices(torch.from_numpy(z["xyz"]).cuda()).cpu().numpy()
    assert np.array_equal(got, z["ijk"])


# --------------------------------------------------------------------------- other model kinds
def test_linear_model_matches_oracle(eng, small_problem):
    import warnings
    import torch
    from oracle import restate, synth
    from radar_ml_b200.model import from_sklearn
    X, y = small_problem["X"], small_problem["y"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lin = synth.build_linear(X[:300], y[:300], X[300:360], y[300:360])
    eng.load_model(from_sklearn(lin))
    p = restate.export_params(lin)
    cubes = small_problem["cubes"][small_problem["test"]]
    _, lab_o, _, known_o, P_o = restate.scan_path(cubes, p, mode="max")
    proba, label, known = eng.predict(torch.from_numpy(cubes).cuda(), mode="max")
    eng.check_status()
    assert np.array_equal(label.cpu().numpy(), lab_o)
    assert np.array_equal(known.cpu().numpy().astype(bool), known_o)
    assert np.abs(proba.cpu().numpy().astype(np.float64) - P_o).max() < PROBA_TOL
    # float32 feature input takes the same kernel
    Xt = torch.from_numpy(X[small_problem["test"]]).cuda()
    proba2, label2, _ = eng.score(Xt, None, 0.7)
    assert np.array_equal(label2.cpu().numpy(), lab_o)
    assert np.abs(proba2.cpu().numpy().astype(np.float64) - P_o).max() < PROBA_TOL


def test_two_class_svc(eng):
    import warnings
    import torch
    from oracle import restate, synth
    from radar_ml_b200.model import from_sklearn
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cubes, y, _ = synth.make_cubes(220, seed=31, n_classes=2)
        X = synth.features(*synth.project_max(cubes))
        cal = synth.build_svc(X[:120], y[:120], X[120:160], y[120:160])
    eng.load_model(from_sklearn(cal))
    p = restate.export_params(cal)
    _, lab_o, _, known_o, P_o = restate.scan_path(cubes[160:], p, mode="max")
    proba, label, known = eng.predict(torch.from_numpy(cubes[160:]).cuda(), mode="max")
    eng.check_status()
    assert np.array_equal(label.cpu().numpy(), lab_o)
    assert np.abs(proba.cpu().numpy().astype(np.float64) - P_o).max() < PROBA_TOL


def test_many_support_vectors_multi_chunk(eng, small_problem):
    """n_sv > 256 exercises several TMEM accumulator chunks (C=1000 keeps most points as SVs)."""
    import warnings
    import torch
    from oracle import restate, synth
    from radar_ml_b200.model import from_sklearn
    X, y = small_problem["X"], small_problem["y"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cal = synth.build_svc(X[:340], y[:340], X[340:360], y[340:360], C=10.0, gamma=0.5)
    p = restate.export_params(cal)
    assert p.sv.shape[0] > 256
    eng.load_model(from_sklearn(cal))
    cubes = small_problem["cubes"][small_problem["test"]]
    _, lab_o, _, known_o, P_o = restate.scan_path(cubes, p, mode="max")
    proba, label, known = eng.predict(torch.from_numpy(cubes).cuda(), mode="max")
    eng.check_status()
    assert np.array_equal(label.cpu().numpy(), lab_o)
    assert np.abs(proba.cpu().numpy().astype(np.float64) - P_o).max() < PROBA_TOL


# --------------------------------------------------------------------------- reference-facing API
def test_drop_in_common_and_predict_api(eng, small_problem):
    """common.process_samples / predict.classifier called exactly like predict.py:112-119."""
    from oracle import restate, synth
    from radar_ml_b200 import common, predict
    common.set_engine(eng)
    cal = small_problem["cal"]
    le = synth.LabelEncoderLike()
    p = small_problem["params"]
    cubes = small_problem["cubes"][small_problem["test"]][:12]
    ijk = small_problem["ijk"][small_problem["test"]][:12]
    for s in range(12):
        raw = cubes[s]
        i, j, k = (int(v) for v in ijk[s])
        yz, xz, xy = raw[i, :, :], raw[:, j, :], raw[:, :, k]
        zoom = predict.calc_proj_zoom(22, 31, 176, 22, 31, 176)
        obs = common.process_samples([(xz, yz, xy)], proj_mask=common.ProjMask(True, True, True),
                                     proj_zoom=zoom, scale=True)
        want = restate.process_samples([(xz, yz, xy)], scale=True)
        assert obs.dtype == np.float32 and np.array_equal(obs, want)
        name, proba = predict.classifier(obs, cal, le, 0.7)       # sklearn object accepted as-is
        name_o, proba_o = restate.classifier(obs, p, le.classes_, 0.7)
        assert name == name_o and abs(proba - proba_o) < PROBA_TOL
    assert common.calculate_matrix_indices(12.5, -3.0, 140.0, 22, 31, 176) == \
        restate.calculate_matrix_indices(12.5, -3.0, 140.0, 22, 31, 176)


def test_predict_loop_with_recorded_radar(eng, small_problem):
    """predict.predict (predict.py:72-131) replayed from a fake Walabot SDK object."""
    from oracle import restate, synth
    from radar_ml_b200 import common, predict
    common.set_engine(eng)
    cubes = small_problem["cubes"][small_problem["test"]][:5]

    class Target:
        def __init__(self, x, y, z):
            self.xPosCm, self.yPosCm, self.zPosCm, self.amplitude = x, y, z, 1.0

    class FakeRadar:
        def __init__(self):
            self.n = -1
            self.stopped = False
        def Trigger(self):
            self.n += 1
        def GetSensorTargets(self):
            if self.n == 1:
                return []                       # predict.py:86-87: no targets -> continue
            return [Target(5.0 + self.n, -4.0, 120.0 + 10 * self.n), Target(-20.0, 10.0, 200.0)]
        def GetRawImage(self):
            return cubes[self.n].tolist(), 22, 31, 176, 0.0
        def Stop(self):
            self.stopped = True
        def Disconnect(self):
            pass
        def Clean(self):
            pass

    radar = FakeRadar()
    le = synth.LabelEncoderLike()
    res = predict.predict(0.7, small_problem["cal"], le, common.ProjMask(True, True, True),
                          radar=radar, max_scans=5)
    assert radar.stopped and len(res) == 8
    p = small_problem["params"]
    want = []
    for n in (0, 2, 3, 4):
        for (x, y, z) in ((5.0 + n, -4.0, 120.0 + 10 * n), (-20.0, 10.0, 200.0)):
            ijk = restate.calculate_matrix_indices(x, y, z, 22, 31, 176)
            t = restate.project(cubes[n], "slice", ijk)
            obs = restate.process_samples([t], scale=True)
            want.append(restate.classifier(obs, p, le.classes_, 0.7))
    for (n1, p1), (n2, p2) in zip(res, want):
        assert n1 == n2 and abs(p1 - p2) < PROBA_TOL


def test_error_paths(eng, small_problem):
    import torch
    from radar_ml_b200._lib import RadarMLError
    from radar_ml_b200.model import from_sklearn
    eng.load_model(from_sklearn(small_problem["cal"]))
    d = torch.zeros((2, 22, 31, 176), device="cuda")
    with pytest.raises(RadarMLError):
        eng.predict(d, mask=(True, False, False))            # F mismatch with the model
    with pytest.raises(ValueError):
        eng.project(torch.zeros((2, 22, 31, 100), device="cuda"))
    with pytest.raises(ValueError):
        eng.project(d, mode="slice")                         # slice needs ijk
    bad = torch.tensor([[0, 0, 176], [0, 0, 0]], dtype=torch.int32, device="cuda")
    eng.project(d, mode="slice", ijk=bad)
    with pytest.raises(RadarMLError):
        eng.check_status()                                   # numpy would raise IndexError
    empty = torch.zeros((0, 22, 31, 176), device="cuda")
    assert eng.project(empty).shape == (0, 10010)            # empty batch is a no-op


# --------------------------------------------------------------------------- §8f callers
def test_derived_targets_match_reference_formula(eng, small_problem):
    """common.py:45-80: axis sums + top-k (exact for integer-valued sensor data)."""
    import torch
    from radar_ml_b200 import common
    common.set_engine(eng)
    cubes = small_problem["cubes"][:40]
    ijk, sums = eng.derive_targets(torch.from_numpy(cubes).cuda(), num_targets=3, want_sums=True)
    ijk, sums = ijk.cpu().numpy(), sums.cpu().numpy()
    for s in range(40):
        d = cubes[s]
        sx = np.sum(np.sum(d, axis=1), axis=1)      # find_max_indices(1, 1)
        sy = np.sum(np.sum(d, axis=0), axis=1)      # find_max_indices(0, 1)
        sz = np.sum(np.sum(d, axis=0), axis=0)      # find_max_indices(0, 0)
        assert np.array_equal(sums[s], np.concatenate([sx, sy, sz]))
        for a, sm in enumerate((sx, sy, sz)):
            top = np.argpartition(sm, -3)[-3:]
            want = top[np.argsort(sm[top])]
            if len(set(sm[want].tolist())) == 3 and sm[want[0]] > np.sort(sm)[-4]:   # no ties
                assert np.array_equal(ijk[s, :, a], want)
    # reference-shaped API: one cube -> [DerivedTarget]
    t = common.DerivedTarget.get_derived_targets(cubes[0].tolist(), 22, 31, 176, num_targets=1)
    assert len(t) == 1 and (t[0].i, t[0].j, t[0].k) == tuple(int(v) for v in ijk[0, -1])
    th = common.THETA_MIN + t[0].i * (common.THETA_MAX - common.THETA_MIN) / 21
    ph = common.PHI_MIN + t[0].j * (common.PHI_MAX - common.PHI_MIN) / 30
    r = common.R_MIN + t[0].k * (common.R_MAX - common.R_MIN) / 175
    assert np.allclose((t[0].xPosCm, t[0].yPosCm, t[0].zPosCm), common.spherical_to_cartesian(r, th, ph))
    assert t[0].amplitude is None


def test_zoomed_process_samples_matches_scipy(eng):
    """common.py:143 ndimage.zoom for a scan arena that differs from the training arena."""
    from scipy import ndimage
    from radar_ml_b200 import common, predict
    common.set_engine(eng)
    rng = np.random.default_rng(12)
    sx, sy, sz = 11, 31, 150                       # smaller arena than the 22x31x176 training one
    samples = [(rng.integers(0, 256, (sx, sz)).astype(np.float32),
                rng.integers(0, 256, (sy, sz)).astype(np.float32),
                rng.integers(0, 256, (sx, sy)).astype(np.float32)) for _ in range(5)]
    zoom = predict.calc_proj_zoom(22, 31, 176, sx, sy, sz)
    for mask in ((True, True, True), (False, True, True)):
        got = common.process_samples(samples, proj_mask=common.ProjMask(*mask), proj_zoom=zoom, scale=True)
        want = np.array([np.concatenate([ndimage.zoom(p, zoom[i]) for i, p in enumerate(t) if mask[i]],
                                        axis=None) / 255. for t in samples])
        assert got.shape == want.shape and got.dtype == np.float32
        assert got.shape[1] == sum(n for n, m in zip((22 * 176, 31 * 176, 22 * 31), mask) if m)
        assert np.abs(got - want).max() < 1e-6


def test_predict_loop_with_arena_mismatch(eng, small_problem):
    """predict.predict when GetRawImage returns a smaller arena (zoom path, README.md:207)."""
    from scipy import ndimage
    from oracle import restate, synth
    from radar_ml_b200 import common, predict
    common.set_engine(eng)
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, (11, 31, 88)).astype(np.float32)

    class Target:
        xPosCm, yPosCm, zPosCm, amplitude = 8.0, -6.0, 150.0, 1.0

    class FakeRadar:
        def Trigger(self): pass
        def GetSensorTargets(self): return [Target()]
        def GetRawImage(self): return raw.tolist(), 11, 31, 88, 0.0
        def Stop(self): pass
        def Disconnect(self): pass
        def Clean(self): pass

    le = synth.LabelEncoderLike()
    res = predict.predict(0.7, small_problem["cal"], le, common.ProjMask(True, True, True),
                          radar=FakeRadar(), max_scans=1)
    i, j, k = restate.calculate_matrix_indices(8.0, -6.0, 150.0, 11, 31, 88)
    zoom = restate.calc_proj_zoom(22, 31, 176, 11, 31, 88)
    t = (raw[:, j, :], raw[i, :, :], raw[:, :, k])
    obs = restate.process_samples([t], proj_zoom=zoom, scale=True)       # scipy, like the reference
    name, proba = restate.classifier(obs, small_problem["params"], le.classes_, 0.7)
    assert len(res) == 1 and res[0][0] == name and abs(res[0][1] - proba) < PROBA_TOL


def test_dataset_roundtrip_and_evaluate_model(eng, small_problem, tmp_path):
    """datasets/README.md format + train.py:215-228 evaluate_model on GPU-scored labels."""
    from sklearn import metrics
    from oracle import synth
    from radar_ml_b200 import common, dataset
    common.set_engine(eng)
    cubes, y, ijk = (small_problem[k] for k in ("cubes", "y", "ijk"))
    sl = small_problem["test"]
    xz, yz, xy = synth.project_max(cubes[sl])
    names = np.array(synth.CLASSES)[y[sl]]
    samples = [(xz[i], yz[i], xy[i]) for i in range(len(names))]
    path = tmp_path / "radar_samples.pickle"
    dataset.save_dataset(path, samples, list(names) + [])
    data = dataset.load_datasets([str(path)], prj_dir="")
    smp, enc, class_names = dataset.filter_and_encode(data, ["cat", "dog", "person"])
    assert class_names == ["cat", "dog", "person"] and np.array_equal(enc, y[sl])
    X = dataset.features(smp)
    assert np.array_equal(X, small_problem["X"][sl])
    cal = small_problem["cal"]
    acc, cm, report = dataset.evaluate_model(cal, X, enc, class_names)
    y_pred = cal.predict(X)
    assert acc == metrics.accuracy_score(enc, y_pred)
    assert np.array_equal(cm, metrics.confusion_matrix(enc, y_pred, labels=[0, 1, 2]))
    ref = metrics.classification_report(enc, y_pred, target_names=class_names, output_dict=True, zero_division=0)
    for row in report["per_class"]:
        assert abs(row["precision"] - ref[row["class"]]["precision"]) < 1e-12
        assert abs(row["recall"] - ref[row["class"]]["recall"]) < 1e-12
        assert row["support"] == ref[row["class"]]["support"]


def test_tiny_model_and_four_classes(eng):
    """Edge shapes of the tensor-core scorer: a handful of support vectors (one 16-column MMA),
    and C = 4 (six OvO pairs through the per-pair weight table)."""
    import warnings
    import torch
    from oracle import restate, synth
    from radar_ml_b200.model import from_sklearn
    for n_classes, n_fit in ((3, 9), (4, 160)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cubes, y, _ = synth.make_cubes(n_fit + 60 + 33, seed=70 + n_classes, n_classes=n_classes)
            y[:n_classes] = np.arange(n_classes)               # every class present in the tiny fit
            X = synth.features(*synth.project_max(cubes))
            cal = synth.build_svc(X[:n_fit], y[:n_fit], X[n_fit:n_fit + 60], y[n_fit:n_fit + 60])
        p = restate.export_params(cal)
        eng.load_model(from_sklearn(cal))
        assert eng.model_is_integral
        test = cubes[n_fit + 60:]
        _, lab_o, _, known_o, P_o = restate.scan_path(test, p, mode="max")
        proba, label, known = eng.predict(torch.from_numpy(test).cuda(), mode="max")
        eng.check_status()
        assert proba.shape == (33, n_classes)
        assert np.abs(proba.cpu().numpy().astype(np.float64) - P_o).max() < PROBA_TOL
        srt = np.sort(P_o, axis=1)
        clear = (srt[:, -1] - srt[:, -2]) > 1e-6            # exact ties are arg-max-order dependent
        assert np.array_equal(label.cpu().numpy()[clear], lab_o[clear])
        assert np.array_equal(known.cpu().numpy().astype(bool)[clear], known_o[clear])


def test_multidigit_scorer_on_nonintegral_data(eng):
    """Augmented-style model (non-integer support vectors, train.py:84-185) and non-integer
    inputs: the exact 24-bit fixed-point tensor-core scorer vs the float64 oracle, and vs the
    float64 CUDA-core scorer; out-of-range values are reported, never scored silently."""
    import warnings
    import torch
    from oracle import restate, synth
    from radar_ml_b200._lib import OutOfRangeInput
    from radar_ml_b200.model import from_sklearn
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cubes, y, _ = synth.make_cubes(330, seed=404, integer=False)
        cubes = np.clip(cubes, 0.0, 255.0).astype(np.float32)         # real-valued, in range
        X = synth.features(*synth.project_max(cubes))
        cal = synth.build_svc(X[:200], y[:200], X[200:260], y[200:260])
    p = restate.export_params(cal)
    eng.load_model(from_sklearn(cal))
    assert not eng.model_is_integral
    Xt = X[260:]
    lab_o, _, known_o, P_o = restate.classify_batch(Xt, p, 0.7)
    xt = torch.from_numpy(Xt).cuda()
    proba, label, known = eng.score(xt, None, 0.7)                    # multi-digit tensor-core path
    eng.check_status()
    proba_x, label_x, known_x = eng.score(xt, None, 0.7, exact=True)  # float64 CUDA cores
    for pr, lb, kn in ((proba, label, known), (proba_x, label_x, known_x)):
        assert np.abs(pr.cpu().numpy().astype(np.float64) - P_o).max() < PROBA_TOL
        assert np.array_equal(lb.cpu().numpy(), lab_o)
        assert np.array_equal(kn.cpu().numpy(), known_o)
    assert np.abs(proba.cpu().numpy() - proba_x.cpu().numpy()).max() < 1e-6
    # sklearn-style entry point on host features
    from radar_ml_b200.model import GpuCalibratedClassifier
    gm = GpuCalibratedClassifier(from_sklearn(cal), engine=eng)
    assert np.abs(gm.predict_proba(Xt) - P_o).max() < PROBA_TOL
    # values the fixed-point digits cannot hold: flagged, then the exact scorer takes over
    bad = Xt.copy()
    bad[3, 17] = -0.02
    bad[5, 100] = 1.5
    eng.score(torch.from_numpy(bad).cuda(), None, 0.7)
    with pytest.raises(OutOfRangeInput):
        eng.check_status()
    P_bad = restate.predict_proba(bad, p)
    assert np.abs(gm.predict_proba(bad) - P_bad).max() < PROBA_TOL


@pytest.mark.parametrize("dims", [(10, 12, 40), (7, 9, 30), (22, 31, 176)])
def test_other_arenas_max_and_slice(dims):
    """rml_set_arena: any cube size goes through the generic / vectorised-slice kernels."""
    import torch
    from oracle import restate
    from radar_ml_b200.engine import Engine
    e = Engine(0)
    sx, sy, sz = dims
    e.set_arena(sx, sy, sz)
    rng = np.random.default_rng(sx * sz)
    n = 37
    cubes = rng.integers(0, 256, (n, sx, sy, sz)).astype(np.float32)
    ijk = np.stack([rng.integers(-sx, sx, n), rng.integers(-sy, sy, n), rng.integers(-sz, sz, n)], axis=1).astype(np.int32)
    d = torch.from_numpy(cubes).cuda()
    for mask in ((True, True, True), (False, True, True), (True, False, False)):
        for mode in ("max", "slice"):
            want = np.asarray([restate.process_samples(
                [restate.project(cubes[s], mode, tuple(int(v) for v in ijk[s]))],
                proj_mask=restate.ProjMask(*mask), scale=True)[0] for s in range(n)], dtype=np.float32)
            got = e.project(d, mode=mode, ijk=torch.from_numpy(ijk).cuda(), mask=mask).cpu().numpy()
            assert got.shape == want.shape and np.array_equal(got, want), (dims, mask, mode)
            q, norms = e.project(d, mode=mode, ijk=torch.from_numpy(ijk).cuda(), mask=mask, dtype=1)
            e.check_status()
            F = want.shape[1]
            raw = np.rint(want * 255.0)
            assert np.array_equal(q.cpu().numpy()[:, :F].astype(np.float32), raw), (dims, mask, mode)
            assert not q.cpu().numpy()[:, F:].any()
            assert np.array_equal(norms.cpu().numpy().astype(np.int64), (raw.astype(np.int64) ** 2).sum(axis=1))
    e.close()
