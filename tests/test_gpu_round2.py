"""-m gpu tests of the round-2 boundary work: per-feature affine (a fitted StandardScaler), NaN/Inf
voxels, the one-scan entry points of the reference's live loop, model interleaving, the
integrality threshold, real-valued cubes through the batched API and the C-ABI label exchange."""
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


def _score(X, p, min_proba=0.7):
    """(None, label, best proba, known, proba matrix) — the tuple layout of restate.scan_path."""
    from oracle import restate
    return (None,) + restate.classify_batch(np.asarray(X, dtype=np.float32), p, min_proba)


def _raw_feats(cubes):
    from oracle import synth
    return synth.features(*synth.project_max(cubes), scale=False)


# --------------------------------------------------------------------------- per-feature affine
def test_standard_scaler_svc_matches_sklearn_pipeline(eng, small_problem):
    """north_star's "StandardScaler-normalised" feature vector: rml_load_affine(mean_, scale_) ->
    K1 emits (u - mean_f)/scale_f -> SVC-RBF trained behind the scaler -> Platt calibration.
    Checked against scikit-learn's own StandardScaler + SVC + CalibratedClassifierCV chain."""
    import torch
    import warnings
    from sklearn import svm
    from sklearn.calibration import CalibratedClassifierCV
    from sklearn.frozen import FrozenEstimator
    from sklearn.preprocessing import StandardScaler
    from oracle import restate
    from radar_ml_b200.model import from_sklearn

    cubes, y = small_problem["cubes"], small_problem["y"]
    raw = _raw_feats(cubes)                                   # integers 0..255 as float32
    scaler = StandardScaler().fit(raw[:300].astype(np.float64))
    off = scaler.mean_.astype(np.float32)
    scl = scaler.scale_.astype(np.float32)
    Xs = ((raw - off) / scl).astype(np.float32)               # what K1 must emit, IEEE float32
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = svm.SVC(C=10.0, gamma=1e-4, kernel="rbf", probability=True, class_weight="balanced",
                      random_state=1234)
        clf.fit(Xs[:300], y[:300])
        cal = CalibratedClassifierCV(estimator=FrozenEstimator(clf)).fit(Xs[300:360], y[300:360])
    test = slice(360, 480)
    P_sk = cal.predict_proba(Xs[test])
    p = restate.export_params(cal)
    _, lab_o, _, known_o, P_o = _score(Xs[test], p)
    assert np.abs(P_o - P_sk).max() < 1e-9

    eng.load_affine(off, scl)
    try:
        params = from_sklearn(cal)
        eng.load_model(params)
        assert not eng.model_is_integral                      # standardised space has no integer form
        d = torch.from_numpy(cubes[test]).cuda()
        feats = eng.project(d, mode="max").cpu().numpy()
        assert np.array_equal(feats, Xs[test])                # bit-identical features
        proba, label, known = eng.predict(d, mode="max")
        eng.check_status()                                    # the digit scorer held every value
        assert np.array_equal(label.cpu().numpy(), lab_o)
        assert np.array_equal(known.cpu().numpy().astype(bool), known_o)
        assert np.abs(proba.cpu().numpy().astype(np.float64) - P_sk).max() < PROBA_TOL
        # uint8 cubes take the same path
        p8, l8, _ = eng.predict(torch.from_numpy(cubes[test].astype(np.uint8)).cuda(), mode="max")
        eng.check_status()
        assert torch.equal(l8, label) and torch.equal(p8, proba)
        # host entry + the float64 scorer agree too
        Ph, lh, _ = eng.predict_host(cubes[test], mode="max")
        assert np.array_equal(lh, lab_o) and np.abs(Ph - P_sk).max() < PROBA_TOL
        pe, le_, _ = eng.score(torch.from_numpy(Xs[test]).cuda(), None, exact=True)
        assert np.array_equal(le_.cpu().numpy(), lab_o)
        assert np.abs(pe.cpu().numpy() - P_sk).max() < PROBA_TOL
        # a mask whose F differs from the table is an error, not a silent misindex
        from radar_ml_b200._lib import RadarMLError
        with pytest.raises(RadarMLError):
            eng.project(d, mode="max", mask=(True, False, False))
    finally:
        eng.load_affine(None, None)
    # default restored: /255 again, bit-exact
    eng.load_model(from_sklearn(small_problem["cal"]))
    got = eng.project(torch.from_numpy(cubes[:16]).cuda(), mode="max").cpu().numpy()
    assert np.array_equal(got, (raw[:16] / np.float32(255.0)).astype(np.float32))


def test_per_feature_affine_all_projection_kernels(eng, small_problem):
    """The table is applied by every float32 feature producer: streaming MAX kernel, SLICE kernel,
    generic kernel (other mask), uint8-cube kernel, process_samples."""
    import torch
    from oracle import synth
    rng = np.random.default_rng(3)
    cubes, ijk = small_problem["cubes"][:40], small_problem["ijk"][:40]
    F = 10010
    off = rng.uniform(0, 50, F).astype(np.float32)
    scl = rng.uniform(0.5, 90, F).astype(np.float32)
    eng.load_affine(off, scl)
    try:
        d = torch.from_numpy(cubes).cuda()
        raw_max = _raw_feats(cubes)
        want = ((raw_max - off) / scl).astype(np.float32)
        assert np.array_equal(eng.project(d, mode="max").cpu().numpy(), want)
        d8 = torch.from_numpy(cubes.astype(np.uint8)).cuda()
        assert np.array_equal(eng.project(d8, mode="max").cpu().numpy(), want)
        xz, yz, xy = synth.project_slice(cubes, ijk)
        raw_sl = synth.features(xz, yz, xy, scale=False)
        want_sl = ((raw_sl - off) / scl).astype(np.float32)
        got_sl = eng.project(d, mode="slice", ijk=torch.from_numpy(ijk)).cpu().numpy()
        assert np.array_equal(got_sl, want_sl)
        got_ps = eng.process_samples(torch.from_numpy(xz).cuda(), torch.from_numpy(yz).cuda(),
                                     torch.from_numpy(xy).cuda(), scale=True).cpu().numpy()
        assert np.array_equal(got_ps, want_sl)
    finally:
        eng.load_affine(None, None)


# --------------------------------------------------------------------------- NaN / Inf voxels
def test_nan_and_inf_voxels_follow_numpy(eng, small_problem):
    """np.max propagates NaN (common.py / predict.py would hand NaN features to sklearn); the
    float32 path reproduces numpy element for element, the u8 path reports the scan."""
    import torch
    from radar_ml_b200._lib import NonIntegralInput
    cubes = small_problem["cubes"][:12].copy()
    cubes[1, 3, 4, 5] = np.nan
    cubes[2, 0, 0, 0] = np.inf
    cubes[3, 21, 30, 175] = -np.inf
    cubes[4, 10, 15, 100] = np.nan
    cubes[4, 10, 15, 101] = np.inf
    d = torch.from_numpy(cubes).cuda()
    with np.errstate(invalid="ignore"):
        want = (_raw_feats(cubes) / np.float32(255.0)).astype(np.float32)
    got = eng.project(d, mode="max").cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got, want, equal_nan=True)
    # generic kernel (other arena path is the same code): a partial mask goes through it too
    got_xy = eng.project(d, mode="max", mask=(False, False, True)).cpu().numpy()
    assert np.array_equal(got_xy, want[:, 3872 + 5456:], equal_nan=True)
    eng.project(d, mode="max", dtype=1)
    with pytest.raises(NonIntegralInput):
        eng.check_status()
    eng.project(d[5:], mode="max", dtype=1)       # the clean scans pass
    eng.check_status()


# --------------------------------------------------------------------------- one-scan entry points
def test_predict_targets_host_and_score_host(eng, small_problem):
    """rml_predict_targets_host = one predict.py:93-119 iteration (one cube, T targets);
    rml_score_host = predict.py:56-70 for host features.  Both against the oracle."""
    from oracle import restate, synth
    from radar_ml_b200.model import from_sklearn
    eng.load_model(from_sklearn(small_problem["cal"]))
    p = small_problem["params"]
    cube = small_problem["cubes"][400]
    rng = np.random.default_rng(11)
    T = 9
    ijk = np.stack([rng.integers(0, 22, T), rng.integers(0, 31, T), rng.integers(0, 176, T)], axis=1).astype(np.int32)
    ijk[0] = (-1, -2, -3)                                     # numpy negative-index wrap
    P, lab, known = eng.predict_targets_host(cube, ijk, min_proba=0.6)
    cubes = np.broadcast_to(cube, (T,) + cube.shape)
    _, lab_o, _, known_o, P_o = restate.scan_path(np.ascontiguousarray(cubes), p, mode="slice", ijk=ijk, min_proba=0.6)
    assert np.array_equal(lab, lab_o) and np.array_equal(known.astype(bool), known_o)
    assert np.abs(P - P_o).max() < PROBA_TOL
    from radar_ml_b200._lib import RadarMLError
    bad = ijk.copy()
    bad[2, 2] = 176
    with pytest.raises(RadarMLError):                         # numpy would raise IndexError
        eng.predict_targets_host(cube, bad)
    # score_host: integral rows (u8 scorer), non-integral rows (digit scorer), out-of-range rows (float64)
    X = small_problem["X"][360:420]
    _, lab_x, _, known_x, P_x = _score(X, p)
    Pg, lg, kg = eng.score_features_host(X)
    assert np.array_equal(lg, lab_x) and np.array_equal(kg.astype(bool), known_x)
    assert np.abs(Pg - P_x).max() < PROBA_TOL
    Xn = (X + np.float32(1e-3) * rng.random(X.shape, dtype=np.float32)).astype(np.float32)
    _, lab_n, _, _, P_n = _score(Xn, p)
    Pg, lg, _ = eng.score_features_host(Xn)
    assert np.array_equal(lg, lab_n) and np.abs(Pg - P_n).max() < PROBA_TOL
    Xneg = (X - np.float32(0.01)).astype(np.float32)          # negative values: float64 scorer
    _, lab_m, _, _, P_m = _score(Xneg, p)
    Pg, lg, _ = eng.score_features_host(Xneg)
    assert np.array_equal(lg, lab_m) and np.abs(Pg - P_m).max() < PROBA_TOL
    # more rows than one staging chunk
    big = np.tile(X, (40, 1))[:2100]
    Pg, lg, _ = eng.score_features_host(big)
    assert np.array_equal(lg, np.tile(lab_x, 40)[:2100])


def test_two_models_interleaved_on_one_engine(eng, small_problem):
    """Two converted models sharing the process-wide engine must never score with each other's
    weights (the native context holds one model at a time)."""
    import warnings
    from oracle import restate, synth
    from radar_ml_b200.model import GpuCalibratedClassifier
    X, y = small_problem["X"], small_problem["y"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lin = synth.build_linear(X[:300], y[:300], X[300:360], y[300:360])
    a = GpuCalibratedClassifier.from_sklearn(small_problem["cal"], engine=eng)
    b = GpuCalibratedClassifier.from_sklearn(lin, engine=eng)      # replaces a's weights in the context
    Xt = X[360:400]
    Pa = small_problem["cal"].predict_proba(Xt)
    Pb = lin.predict_proba(Xt.astype(np.float64))
    for _ in range(2):
        assert np.abs(a.predict_proba(Xt) - Pa).max() < PROBA_TOL
        assert np.abs(b.predict_proba(Xt) - Pb).max() < PROBA_TOL
    assert np.abs(Pa - Pb).max() > 1e-3                       # the two models really differ


def test_integrality_threshold(eng, small_problem):
    """float32(u/255)*255 is within 2e-5 of u; a support vector 5e-5 counts off an integer is a
    genuinely non-integral model and must take the digit scorer."""
    import copy
    import torch
    from oracle import restate
    from radar_ml_b200.model import from_sklearn
    params = from_sklearn(small_problem["cal"])
    eng.load_model(params)
    assert eng.model_is_integral
    p2 = copy.deepcopy(params)
    p2.sv = p2.sv.copy()
    p2.sv[0, 7] += 5e-5 / 255.0
    eng.load_model(p2)
    assert not eng.model_is_integral
    po = copy.deepcopy(small_problem["params"])
    po.sv = p2.sv
    cubes = small_problem["cubes"][360:400]
    _, lab_o, _, _, P_o = restate.scan_path(cubes, po, mode="max")
    proba, label, _ = eng.predict(torch.from_numpy(cubes).cuda())
    eng.check_status()
    assert np.array_equal(label.cpu().numpy(), lab_o)
    assert np.abs(proba.cpu().numpy() - P_o).max() < PROBA_TOL
    eng.load_model(params)


def test_real_valued_cubes_through_batched_api(eng, small_problem):
    """The reference accepts any float32 cube.  classify_cubes / predict_host on an integral model
    see the non-integral voxels on the device and re-run with float32 features."""
    import torch
    from oracle import restate
    from radar_ml_b200 import predict as rp
    from radar_ml_b200.model import GpuCalibratedClassifier
    gm = GpuCalibratedClassifier.from_sklearn(small_problem["cal"], engine=eng)
    rng = np.random.default_rng(8)
    cubes = small_problem["cubes"][360:420] * rng.uniform(0.3, 0.999, (60, 1, 1, 1)).astype(np.float32)
    _, lab_o, _, known_o, P_o = restate.scan_path(cubes, small_problem["params"], mode="max")
    lab, best, known, P = rp.classify_cubes(torch.from_numpy(cubes).cuda(), gm)
    assert np.array_equal(lab, lab_o) and np.array_equal(known, known_o)
    assert np.abs(P - P_o).max() < PROBA_TOL
    lab, best, known, P = rp.classify_cubes(cubes, gm)         # host entry
    assert np.array_equal(lab, lab_o) and np.abs(P - P_o).max() < PROBA_TOL
    # and the engine is back on the integer path afterwards
    q = small_problem["cubes"][360:380]
    _, lab_q, _, _, _ = restate.scan_path(q, small_problem["params"], mode="max")
    lab2, _, _, _ = rp.classify_cubes(torch.from_numpy(q).cuda(), gm)
    assert np.array_equal(lab2, lab_q)


def test_derive_targets_nan_sums_do_not_corrupt(eng, small_problem):
    import torch
    from radar_ml_b200._lib import RadarMLError
    cubes = small_problem["cubes"][:4].copy()
    cubes[1] = np.nan
    ijk = eng.derive_targets(torch.from_numpy(cubes).cuda(), num_targets=2)
    with pytest.raises(RadarMLError):
        eng.check_status()
    ijk = ijk.cpu().numpy()
    assert (ijk >= 0).all() and (ijk[..., 0] < 22).all() and (ijk[..., 1] < 31).all() and (ijk[..., 2] < 176).all()


def test_project_derive_shares_one_pass(eng, small_problem):
    """DerivedTarget's axis sums (common.py:45-80) computed inside K1's single pass
    over the cube.  Same features / norms as rml_project, same indices and sums as rml_derive_targets
    and as the reference formula (integer-valued cubes: float32 sums are exact in any order)."""
    import torch
    from radar_ml_b200._lib import F32, U8
    cubes = small_problem["cubes"][:301]                       # ragged against the persistent grid
    d = torch.from_numpy(cubes).cuda()
    ijk_ref, sums_ref = eng.derive_targets(d, num_targets=3, want_sums=True)
    f32_ref = eng.project(d, dtype=F32)
    u8_ref, norms_ref = eng.project(d, dtype=U8)
    f32, ijk_a, sums_a = eng.project_derive(d, num_targets=3, dtype=F32, want_sums=True)
    u8, norms, ijk_b = eng.project_derive(d, num_targets=3, dtype=U8)
    eng.check_status()
    assert torch.equal(f32, f32_ref) and torch.equal(u8, u8_ref) and torch.equal(norms, norms_ref)
    assert torch.equal(sums_a, sums_ref)
    assert torch.equal(ijk_a, ijk_ref) and torch.equal(ijk_b, ijk_ref)
    # the reference formula itself: sums over the other two axes, argsort, last = strongest
    c64 = cubes.astype(np.float64)
    want = np.concatenate([c64.sum(axis=(2, 3)), c64.sum(axis=(1, 3)), c64.sum(axis=(1, 2))], axis=1)
    assert np.array_equal(sums_a.cpu().numpy().astype(np.float64), want)
    assert np.array_equal(sums_ref.cpu().numpy().astype(np.float64), want)
    # NaN sums are reported, never turned into out-of-range indices
    bad = cubes[:4].copy()
    bad[2] = np.nan
    _, ijk_n = eng.project_derive(torch.from_numpy(bad).cuda(), num_targets=2, dtype=F32)
    from radar_ml_b200._lib import RadarMLError
    with pytest.raises(RadarMLError):
        eng.check_status()
    ijk_n = ijk_n.cpu().numpy()
    assert (ijk_n >= 0).all() and (ijk_n[..., 0] < 22).all() and (ijk_n[..., 1] < 31).all() and (ijk_n[..., 2] < 176).all()


def test_host_narrowing_same_results_quarter_of_the_bytes(eng, small_problem):
    """rml_predict_host: float32 cubes of the sensor's integers cross the bus as bytes (converted and
    checked on the host); results are those of the float32 transfer bit for bit, and a chunk with a
    real value in it is copied as float32 and handled exactly as before."""
    from radar_ml_b200.model import from_sklearn
    eng.load_model(from_sklearn(small_problem["cal"]))
    cubes = np.ascontiguousarray(np.concatenate([small_problem["cubes"][300:]] * 7)[:700])    # 2 chunks (512 + 188)
    assert cubes.shape[0] == 700
    eng.set_host_narrowing(True, threads=4, min_gbs=1e-3)          # never switch off in this test
    P1, l1, k1 = eng.predict_host(cubes)
    x = eng.last_host_transfer()
    assert x["active"] and x["narrowed_scans"] == 700 and x["h2d_bytes"] == 700 * 22 * 31 * 176
    eng.set_host_narrowing(False)
    P0, l0, k0 = eng.predict_host(cubes)
    x0 = eng.last_host_transfer()
    assert x0["narrowed_scans"] == 0 and x0["h2d_bytes"] == 700 * 22 * 31 * 176 * 4
    assert np.array_equal(P1, P0) and np.array_equal(l1, l0) and np.array_equal(k1, k0)
    # SLICE mode through the narrowed path
    ijk = small_problem["ijk"][300:][:64]
    eng.set_host_narrowing(True, threads=4, min_gbs=1e-3)
    Ps, ls, _ = eng.predict_host(cubes[:64], mode="slice", ijk=ijk)
    eng.set_host_narrowing(False)
    Pq, lq, _ = eng.predict_host(cubes[:64], mode="slice", ijk=ijk)
    assert np.array_equal(Ps, Pq) and np.array_equal(ls, lq)
    # one real-valued voxel in the second chunk: that chunk goes over as float32, the engine then
    # re-scores the batch on the general-precision path, exactly as without narrowing
    real = cubes.copy()
    real[600, 3, 4, 5] += 0.25
    eng.set_host_narrowing(True, threads=4, min_gbs=1e-3)
    Pr, lr, _ = eng.predict_host(real)
    eng.set_host_narrowing(False)
    Pn, ln, _ = eng.predict_host(real)
    assert np.array_equal(lr, ln) and np.array_equal(Pr, Pn)


# --------------------------------------------------------------------------- label exchange
def test_allgather_labels_single_rank_and_in_place(eng, small_problem):
    """world == 1: rml_allgather_labels degenerates to a copy; rml_predict writes its labels
    straight into a slice of a larger gather buffer."""
    import torch
    from radar_ml_b200.model import from_sklearn
    eng.load_model(from_sklearn(small_problem["cal"]))
    d = torch.from_numpy(small_problem["cubes"][360:424]).cuda()
    _, ref, _ = eng.predict(d)
    gathered = torch.full((3 * 64,), -7, device="cuda", dtype=torch.int32)
    mine = gathered[64:128]
    out = (torch.empty((64, 3), device="cuda"), mine, torch.empty((64,), device="cuda", dtype=torch.uint8))
    eng.predict(d, out=out)
    assert torch.equal(gathered[64:128], ref) and int((gathered[:64] != -7).sum()) == 0
    recv = torch.empty((64,), device="cuda", dtype=torch.int32)
    eng.allgather_labels(mine, recv)
    torch.cuda.synchronize()
    assert torch.equal(recv, ref)


def test_allgather_labels_two_ranks():
    """The C-ABI exchange on real NCCL (needs >= 2 GPUs; the round-end 1-GPU box skips it, the
    bench's N >= 2 lines carry the same check as parity.gathered_equal)."""
    import os
    import subprocess
    import sys
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(root, "tests", "_dist_worker.py"), "--c-abi"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "DIST_OK" in res.stdout


def test_predict_host_many_chunks_through_pinned_result_mirrors(eng, small_problem):
    """rml_predict_host with more chunks than staging slots (3): the results of a chunk wait in a pinned
    mirror until its slot comes round again, so every slot is drained, reused and drained at the end.
    2 100 float32 scans = 5 chunks (4 x 512 + 52), as uint8 3 chunks (2 x 1 024 + 52); each host result
    must equal the device-resident call bit for bit, with the narrowing on (the default) and off, with
    and without the `known` output, and a second call must not see leftovers of the first."""
    import torch
    from radar_ml_b200.model import from_sklearn
    eng.load_model(from_sklearn(small_problem["cal"]))
    base = small_problem["cubes"][300:]
    reps = -(-2100 // base.shape[0])
    cubes = np.ascontiguousarray(np.concatenate([base] * reps)[:2100])
    cubes[1::2] = cubes[1::2][:, ::-1].copy()            # not periodic with the chunk size
    Pd, ld, kd = (t.cpu().numpy() for t in eng.predict(torch.from_numpy(cubes).cuda()))
    eng.check_status()
    for narrowing in (True, False):
        eng.set_host_narrowing(narrowing, threads=4, min_gbs=1e-3)
        P, l, k = eng.predict_host(cubes)
        assert eng.last_host_transfer()["narrowed_scans"] == (2100 if narrowing else 0)
        assert np.array_equal(P, Pd) and np.array_equal(l, ld) and np.array_equal(k, kd.astype(np.uint8))
        # shorter second call into the same output arrays: rows beyond it keep the first call's values
        P2, l2, k2 = eng.predict_host(cubes[:600], out=(P, l, k))
        assert np.array_equal(P2, Pd) and np.array_equal(l2, ld)
    P8, l8, k8 = eng.predict_host(cubes.astype(np.uint8))
    assert np.array_equal(P8, Pd) and np.array_equal(l8, ld) and np.array_equal(k8, kd.astype(np.uint8))
    eng.set_host_narrowing(False)
