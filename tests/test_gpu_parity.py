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
