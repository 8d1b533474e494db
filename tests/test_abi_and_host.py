"""CPU tests: the C-ABI library loads and exports every symbol include/radarml.h declares,
fails loudly without a GPU, and the host-side mirrors behave like the reference's."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from radar_ml_b200 import _lib
    _lib.build()
    return _lib.load()


def header_symbols():
    with open(os.path.join(ROOT, "include", "radarml.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(rml_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from radar_ml_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libradarml.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms      # the ctypes table covers the header exactly


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    rc = lib.rml_create(0, ctypes.byref(ctx))
    assert rc == -2 and not ctx
    assert b"no CPU path" in lib.rml_last_error(None)
    from radar_ml_b200.engine import Engine
    from radar_ml_b200._lib import RadarMLError
    with pytest.raises(RadarMLError):
        Engine(0)


def test_host_narrowing_conversion_is_exact_and_strict(lib):
    """rml_predict_host narrows float32 cubes of the sensor's integers to bytes before the H2D copy
    (host threads, no GPU involved): exact for integers in [0,255], refused for anything else."""
    from radar_ml_b200._lib import E_NONINTEGRAL
    rng = np.random.default_rng(3)
    n = 3 * 120032 + 17                                   # ragged against the 32-element blocks
    src = rng.integers(0, 256, n).astype(np.float32)
    src[5] = -0.0
    raw = np.empty(n + 64, np.uint8)
    off = (-raw.ctypes.data) % 32
    dst = raw[off:off + n]
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for threads in (1, 3, 0):
        dst[:] = 7
        assert lib.rml_host_narrow_f32_to_u8(p(src), p(dst), n, threads) == 0
        assert np.array_equal(dst, src.astype(np.uint8))
    for pos, v in ((0, 0.5), (n - 1, 256.0), (70001, -1.0), (n // 2, np.nan), (n - 40, np.inf), (33, 254.99998)):
        bad = src.copy()
        bad[pos] = v
        for threads in (1, 4):
            assert lib.rml_host_narrow_f32_to_u8(p(bad), p(dst), n, threads) == E_NONINTEGRAL, (pos, v, threads)
    assert lib.rml_host_narrow_f32_to_u8(p(src), p(raw[off + 1:]), 64, 1) != 0      # misaligned destination


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "radar_ml_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_host_mirror_constants_and_zoom():
    from radar_ml_b200 import common, predict
    assert common.ProjMask._fields == common.ProjZoom._fields == ("xz", "yz", "xy")
    assert common.RADAR_MAX == 255.0 and common.arena_size() == (22, 31, 176)
    z = predict.calc_proj_zoom(22, 31, 176, 22, 31, 176)
    assert z == common.ProjZoom(xz=[1.0, 1.0], yz=[1.0, 1.0], xy=[1.0, 1.0])
    z = predict.calc_proj_zoom(22, 31, 176, 11, 62, 88)
    assert z.xy == [2.0, 0.5] and z.xz == [2.0, 2.0] and z.yz == [0.5, 2.0]
    from radar_ml_b200.engine import mask_bits
    assert mask_bits(common.ProjMask(True, True, True)) == 7
    assert mask_bits(common.ProjMask(xz=False, yz=True, xy=False)) == 2
    assert mask_bits(common.ProjMask(xz=True, yz=False, xy=True)) == 5


def test_model_export_matches_oracle_export():
    import warnings
    from oracle import restate, synth
    from radar_ml_b200.model import from_sklearn
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cubes, y, _ = synth.make_cubes(120, seed=3)
        X = synth.features(*synth.project_max(cubes))
        cal = synth.build_svc(X[:80], y[:80], X[80:], y[80:])
        lin = synth.build_linear(X[:80], y[:80], X[80:], y[80:])
    a, b = from_sklearn(cal), restate.export_params(cal)
    assert a.kind == "svc_rbf" and a.n_classes == b.n_classes == 3 and a.n_features == 10010
    for name in ("sv", "dual_coef", "rho", "n_support", "platt_a", "platt_b"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert a.gamma == b.gamma == 0.01
    l = from_sklearn(lin)
    assert l.kind == "linear" and l.coef.shape == (3, 10010) and l.intercept.shape == (3,)
    with pytest.raises(TypeError):
        from_sklearn(object())


def test_shard_ranges_cover_batch():
    from radar_ml_b200.dist import shard_range, shard_sizes
    for total in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert sum(shard_sizes(total, world)) == total
            assert max(shard_sizes(total, world)) - min(shard_sizes(total, world)) <= 1
