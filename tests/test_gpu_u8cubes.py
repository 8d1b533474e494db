"""-m gpu parity of the uint8-cube entry points (rml_project_u8 / rml_predict_u8 /
rml_predict_host_u8 / rml_net_predict_u8).

predict.py:90-91 widens the sensor's integer voxels to float32; a caller may keep them as
uint8.  The contract is: every result equals, bit for bit, what the float32 entry point (and
the CPU oracle) gives on ``cubes.astype(np.float32)`` — features, u8 operand rows, norms,
labels, known flags and probabilities."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MASKS = [(True, True, True), (True, False, False), (False, True, False), (False, False, True),
         (True, False, True), (False, True, True)]


@pytest.fixture(scope="module")
def eng():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from radar_ml_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _oracle_feats(cubes_u8, mode, ijk, mask, scale=True):
    from oracle import restate
    out = []
    for s in range(cubes_u8.shape[0]):
        c = cubes_u8[s].astype(np.float32)          # predict.py:91
        t = restate.project(c, mode, None if ijk is None else tuple(int(v) for v in ijk[s]))
        out.append(restate.process_samples([t], proj_mask=restate.ProjMask(*mask), scale=scale)[0])
    return np.asarray(out, dtype=np.float32)


def _dense_u8(n, seed, dims=(22, 31, 176)):
    """Adversarial, non-sparse voxels: every byte value occurs, maxima are rarely unique."""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, 256, size=(n,) + dims, dtype=np.uint8)
    c[0] = 0
    if n > 1:
        c[1] = 255
    return c


@pytest.mark.parametrize("mask", MASKS)
def test_k1_max_u8cubes_f32_features_bit_exact(eng, small_problem, mask):
    import torch
    cu = np.concatenate([small_problem["cubes"][:90].astype(np.uint8), _dense_u8(61, 3)])
    assert np.array_equal(cu[:90].astype(np.float32), small_problem["cubes"][:90])   # integral fixture
    d8 = torch.from_numpy(cu).cuda()
    got = eng.project(d8, mode="max", mask=mask).cpu().numpy()
    want = _oracle_feats(cu, "max", None, mask)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want)
    # and the float32 entry point on the widened cube
    ref = eng.project(d8.float(), mode="max", mask=mask).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("mask", [(True, True, True), (False, True, True)])
def test_k1_max_u8cubes_u8_rows_and_norms(eng, mask):
    import torch
    cu = _dense_u8(301, 7)                      # not a multiple of the grid
    d8 = torch.from_numpy(cu).cuda()
    q, norms = eng.project(d8, mode="max", mask=mask, dtype=1)
    eng.check_status()
    want = _oracle_feats(cu, "max", None, mask, scale=False)
    F = want.shape[1]
    q = q.cpu().numpy()
    assert q.shape == (301, (F + 127) // 128 * 128)
    assert np.array_equal(q[:, :F].astype(np.float32), want)
    assert not q[:, F:].any()
    assert np.array_equal(norms.cpu().numpy().astype(np.int64), (want.astype(np.int64) ** 2).sum(axis=1))
    q2, n2 = eng.project(d8.float(), mode="max", mask=mask, dtype=1)
    assert torch.equal(q2.cpu(), torch.from_numpy(q)) and torch.equal(n2, norms)


def test_k1_unscaled_and_net_affine_u8cubes(eng):
    import torch
    cu = _dense_u8(9, 11)
    d8 = torch.from_numpy(cu).cuda()
    for off, sc, en in ((0.0, 255.0, False), (127.5, 127.5, True)):   # raw, dnn.py:202-205
        eng.set_affine(off, sc, en)
        try:
            got = eng.project(d8, mode="max")
            ref = eng.project(d8.float(), mode="max")
        finally:
            eng.set_affine(0.0, 255.0, True)
        assert torch.equal(got, ref)
    raw = _oracle_feats(cu, "max", None, (True, True, True), scale=False)
    eng.set_affine(0.0, 255.0, False)
    try:
        assert np.array_equal(eng.project(d8, mode="max").cpu().numpy(), raw)
    finally:
        eng.set_affine(0.0, 255.0, True)


@pytest.mark.parametrize("dtype", [0, 1])
def test_k1_slice_u8cubes(eng, small_problem, dtype):
    import torch
    cu = small_problem["cubes"][:100].astype(np.uint8)
    ijk = small_problem["ijk"][:100].copy()
    ijk[0] = (0, 0, 0)
    ijk[1] = (21, 30, 175)
    ijk[2] = (-1, -2, -3)
    res = eng.project(torch.from_numpy(cu).cuda(), mode="slice", ijk=torch.from_numpy(ijk).cuda(), dtype=dtype)
    eng.check_status()
    want = _oracle_feats(cu, "slice", ijk, (True, True, True), scale=(dtype == 0))
    got = res.cpu().numpy() if dtype == 0 else res[0].cpu().numpy()[:, :10010].astype(np.float32)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("dims", [(10, 12, 40), (7, 9, 30)])
def test_other_arenas_u8cubes(dims):
    import torch
    from radar_ml_b200.engine import Engine
    e = Engine(0)
    try:
        e.set_arena(*dims)
        cu = _dense_u8(37, 13, dims)
        rng = np.random.default_rng(1)
        ijk = np.stack([rng.integers(0, d, size=37) for d in dims], axis=1).astype(np.int32)
        d8 = torch.from_numpy(cu).cuda()
        for mode, ij in (("max", None), ("slice", ijk)):
            got = e.project(d8, mode=mode, ijk=None if ij is None else torch.from_numpy(ij).cuda())
            ref = e.project(d8.float(), mode=mode, ijk=None if ij is None else torch.from_numpy(ij).cuda())
            assert torch.equal(got, ref)
            q, n = e.project(d8, mode=mode, ijk=None if ij is None else torch.from_numpy(ij).cuda(), dtype=1)
            q2, n2 = e.project(d8.float(), mode=mode, ijk=None if ij is None else torch.from_numpy(ij).cuda(), dtype=1)
            e.check_status()
            assert torch.equal(q, q2) and torch.equal(n, n2)
    finally:
        e.close()


def test_predict_u8cubes_device_and_host(eng, small_problem):
    import torch
    from oracle import restate
    from radar_ml_b200.model import from_sklearn
    eng.load_model(from_sklearn(small_problem["cal"]))
    cubes = small_problem["cubes"][small_problem["test"]]
    cu = cubes.astype(np.uint8)
    _, lab_o, _, known_o, P_o = restate.scan_path(cubes, small_problem["params"], mode="max")
    pf, lf, kf = (t.clone() for t in eng.predict(torch.from_numpy(cubes).cuda(), mode="max"))
    p8, l8, k8 = eng.predict(torch.from_numpy(cu).cuda(), mode="max")
    eng.check_status()
    assert torch.equal(p8, pf) and torch.equal(l8, lf) and torch.equal(k8, kf)
    assert np.array_equal(l8.cpu().numpy(), lab_o)
    assert np.array_equal(k8.cpu().numpy().astype(bool), known_o)
    assert np.abs(p8.cpu().numpy().astype(np.float64) - P_o).max() < 1e-5
    # host entry: several chunks of the pipeline plus a ragged tail
    reps = np.concatenate([cu] * 19)[:2077]
    ph, lh, kh = eng.predict_host(np.ascontiguousarray(reps), mode="max")
    idx = np.arange(2077) % cu.shape[0]
    assert np.array_equal(lh, lab_o[idx]) and np.array_equal(kh.astype(bool), known_o[idx])
    assert np.array_equal(ph, p8.cpu().numpy()[idx])
    # slice mode through the host entry
    ijk = small_problem["ijk"][small_problem["test"]]
    ps, ls, ks = eng.predict_host(cu, mode="slice", ijk=ijk)
    pr, lr, kr = eng.predict_host(np.ascontiguousarray(cubes), mode="slice", ijk=ijk)
    assert np.array_equal(ps, pr) and np.array_equal(ls, lr) and np.array_equal(ks, kr)
    # empty batch
    p0, l0, k0 = eng.predict(torch.empty((0, 22, 31, 176), dtype=torch.uint8, device="cuda"))
    assert p0.shape == (0, 3) and l0.shape == (0,)


def test_linear_model_u8cubes(eng, small_problem):
    import torch
    from oracle import synth
    from radar_ml_b200.model import from_sklearn
    X, y = small_problem["X"], small_problem["y"]
    cal = synth.build_linear(X[:300], y[:300], X[300:360], y[300:360])
    eng.load_model(from_sklearn(cal))
    cubes = small_problem["cubes"][small_problem["test"]]
    pf, lf, kf = (t.clone() for t in eng.predict(torch.from_numpy(cubes).cuda()))
    p8, l8, k8 = eng.predict(torch.from_numpy(cubes.astype(np.uint8)).cuda())
    eng.check_status()
    assert torch.equal(p8, pf) and torch.equal(l8, lf) and torch.equal(k8, kf)


def test_fused_pipeline_u8cubes_large_batch(small_problem):
    """>= 8192 scans: the serial order (default for uint8 cubes) == K1(u8) || K2 as one device-side
    pipeline, bit for bit, and a scan's result does not depend on the batch it is in."""
    import ctypes as C
    import torch
    from oracle import restate
    from radar_ml_b200.engine import Engine
    from radar_ml_b200.model import from_sklearn
    eng = Engine(0)
    try:
        eng.load_model(from_sklearn(small_problem["cal"]))
        base = torch.from_numpy(small_problem["cubes"].astype(np.uint8)).cuda()     # 480 scans
        g = torch.Generator(device="cpu").manual_seed(5)
        idx = torch.randint(0, base.shape[0], (16384 + 77,), generator=g).cuda()
        cubes = base[idx].contiguous()                                               # 1.98 GB
        eng.lib.rml_enable_timing(eng.ctx, 1)
        fused = C.c_int()
        # default for uint8 cubes: serial order, two projection CTAs per SM
        p1, l1, k1 = (t.clone() for t in eng.predict(cubes))
        eng.check_status()
        assert eng.lib.rml_last_timing(eng.ctx, None, None, C.byref(fused)) == 0 and fused.value == 0
        for sms in (32, 96):      # the co-resident pipeline at two SM splits
            assert eng.lib.rml_set_fused_u8(eng.ctx, sms) == 0
            p, l, k = eng.predict(cubes)
            eng.check_status()
            assert eng.lib.rml_last_timing(eng.ctx, None, None, C.byref(fused)) == 0 and fused.value == 1
            assert torch.equal(p, p1) and torch.equal(l, l1) and torch.equal(k, k1)
        assert eng.lib.rml_set_fused_u8(eng.ctx, 0) == 0
        # every copy of a scan scores the same, and like the oracle
        pb, lb, kb = eng.predict(base)
        eng.check_status()
        assert torch.equal(p1, pb[idx]) and torch.equal(l1, lb[idx]) and torch.equal(k1, kb[idx])
        _, lab_o, _, known_o, P_o = restate.scan_path(small_problem["cubes"][:128], small_problem["params"], mode="max")
        assert np.array_equal(lb[:128].cpu().numpy(), lab_o)
        assert np.abs(pb[:128].cpu().numpy().astype(np.float64) - P_o).max() < 1e-5
    finally:
        eng.close()


def test_net_predict_u8cubes(eng):
    import torch
    from oracle import nets, synth
    from radar_ml_b200.nets import GpuNetClassifier
    net = GpuNetClassifier(nets.random_dnn(9), engine=eng, chunk=32)
    cubes, _, ijk = synth.make_cubes(70, seed=61)
    d = torch.from_numpy(cubes).cuda()
    d8 = torch.from_numpy(cubes.astype(np.uint8)).cuda()
    for mode, ij in (("max", None), ("slice", torch.from_numpy(ijk).cuda())):
        pf, lf = (t.clone() for t in net.predict_cubes(d, mode=mode, ijk=ij))
        p8, l8 = net.predict_cubes(d8, mode=mode, ijk=ij)
        assert torch.equal(p8, pf) and torch.equal(l8, lf)


def test_u8_entry_points_reject_bad_arguments(eng, small_problem):
    import torch
    from radar_ml_b200._lib import RadarMLError
    with pytest.raises(ValueError):
        eng.project(torch.zeros((2, 22, 31, 175), dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        eng.predict_host(np.zeros((2, 22, 31, 176), dtype=np.int16))
    with pytest.raises(ValueError):
        eng.derive_targets(torch.zeros((2, 22, 31, 176), dtype=torch.uint8, device="cuda"))
    rc = eng.lib.rml_predict_host_u8(eng.ctx, None, 4, 0, None, 7, 0.7, None, None, None)
    assert rc != 0
    with pytest.raises(RadarMLError):
        from radar_ml_b200._lib import check
        check(eng.ctx, rc)
