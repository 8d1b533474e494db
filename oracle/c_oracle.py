"""TEST INFRASTRUCTURE — ctypes wrapper of the plain-C oracle (oracle/c/radar_oracle.c).

A second, independent restatement of the reference path next to oracle/restate.py (numpy):
the two are held to each other and to the golden vectors of the unmodified reference in
tests/test_oracle_c.py.  bench.py also times it (OpenMP over scans) as the best-effort
all-cores CPU number next to the reference-exact per-scan sklearn loop.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
LIB_PATH = os.path.join(_DIR, "libradar_oracle.so")
_lib = None


class _Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_classes", C.c_int32), ("n_features", C.c_int32),
                ("n_sv", C.c_int32), ("gamma", C.c_double), ("sv", C.c_void_p),
                ("dual_coef", C.c_void_p), ("rho", C.c_void_p), ("n_support", C.c_void_p),
                ("platt_a", C.c_void_p), ("platt_b", C.c_void_p), ("coef", C.c_void_p),
                ("intercept", C.c_void_p)]


def build(force=False):
    src = os.path.join(_DIR, "radar_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _DIR], check=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.ro_project.restype = C.c_int
        lib.ro_process_samples.restype = C.c_int
        lib.ro_predict_proba.restype = C.c_int
        lib.ro_scan_path.restype = C.c_int64
        lib.ro_matrix_indices.restype = None
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CModel:
    """oracle.restate.SvcParams -> the struct radar_oracle.c reads (keeps the arrays alive)."""

    def __init__(self, p):
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)   # noqa: E731
        self.keep = dict(sv=f64(p.sv), dual_coef=f64(p.dual_coef), rho=f64(p.rho),
                         n_support=np.ascontiguousarray(p.n_support, dtype=np.int32),
                         platt_a=f64(p.platt_a), platt_b=f64(p.platt_b),
                         coef=None if p.coef is None else f64(p.coef),
                         intercept=None if p.intercept is None else f64(p.intercept))
        k = self.keep
        linear = p.kind == "linear"
        self.n_classes = int(p.n_classes)
        self.n_features = int(k["coef"].shape[1] if linear else k["sv"].shape[1])
        self.struct = _Model(1 if linear else 0, self.n_classes, self.n_features,
                             0 if linear else int(k["sv"].shape[0]), float(p.gamma),
                             _p(k["sv"]).value, _p(k["dual_coef"]).value, _p(k["rho"]).value,
                             _p(k["n_support"]).value, _p(k["platt_a"]).value, _p(k["platt_b"]).value,
                             None if k["coef"] is None else _p(k["coef"]).value,
                             None if k["intercept"] is None else _p(k["intercept"]).value)


def project(cube, mode, ijk=None):
    """predict.py:102-107 / axis max -> (xz, yz, xy) float32."""
    lib = load()
    cube = np.ascontiguousarray(cube, dtype=np.float32)
    sx, sy, sz = cube.shape
    xz, yz, xy = (np.empty(s, np.float32) for s in ((sx, sz), (sy, sz), (sx, sy)))
    i, j, k = (0, 0, 0) if ijk is None else (int(v) for v in ijk)
    rc = lib.ro_project(_p(cube), sx, sy, sz, 1 if mode == "slice" else 0, i, j, k, _p(xz), _p(yz), _p(xy))
    if rc != 0:
        raise IndexError("slice index (%d, %d, %d) out of bounds for %s" % (i, j, k, cube.shape))
    return xz, yz, xy


def process_samples(samples, mask=(True, True, True), scale=False):
    """common.py:123-149 at zoom 1.0 for a list of (xz, yz, xy) tuples."""
    lib = load()
    bits = (1 if mask[0] else 0) | (2 if mask[1] else 0) | (4 if mask[2] else 0)
    out = []
    for xz, yz, xy in samples:
        dims = {}
        if xz is not None:
            dims["sx"], dims["sz"] = xz.shape
        if yz is not None:
            dims["sy"], dims["sz"] = yz.shape
        if xy is not None:
            dims["sx"], dims["sy"] = xy.shape
        sx, sy, sz = dims.get("sx", 0), dims.get("sy", 0), dims.get("sz", 0)
        arr = [None if a is None else np.ascontiguousarray(a, dtype=np.float32) for a in (xz, yz, xy)]
        F = (sx * sz if mask[0] else 0) + (sy * sz if mask[1] else 0) + (sx * sy if mask[2] else 0)
        row = np.empty(F, np.float32)
        got = lib.ro_process_samples(_p(arr[0]), _p(arr[1]), _p(arr[2]), sx, sy, sz, bits,
                                     1 if scale else 0, _p(row))
        assert got == F
        out.append(row)
    return np.asarray(out)


def predict_proba(X, p, want_decision=False):
    """model.predict_proba(X) (predict.py:60) -> (n, C) float64."""
    lib = load()
    m = p if isinstance(p, CModel) else CModel(p)
    X = np.ascontiguousarray(X, dtype=np.float32)
    n = X.shape[0]
    assert X.shape[1] == m.n_features
    P = np.empty((n, m.n_classes), np.float64)
    R = 1 if m.n_classes == 2 else m.n_classes
    D = np.empty((n, R), np.float64) if want_decision else None
    rc = lib.ro_predict_proba(_p(X), C.c_int64(n), C.byref(m.struct), _p(P), _p(D))
    assert rc == 0, rc
    if want_decision:
        return P, (D.ravel() if R == 1 else D)
    return P


def scan_path(cubes, p, mode="max", ijk=None, mask=(True, True, True), min_proba=0.7):
    """predict.py:90-119 for a batch: same return tuple as oracle.restate.scan_path."""
    lib = load()
    m = p if isinstance(p, CModel) else CModel(p)
    cubes = np.ascontiguousarray(cubes, dtype=np.float32)
    n, sx, sy, sz = cubes.shape
    bits = (1 if mask[0] else 0) | (2 if mask[1] else 0) | (4 if mask[2] else 0)
    ij = None if ijk is None else np.ascontiguousarray(ijk, dtype=np.int32)
    X = np.empty((n, m.n_features), np.float32)
    P = np.empty((n, m.n_classes), np.float64)
    lab = np.empty(n, np.int32)
    known = np.empty(n, np.uint8)
    rc = lib.ro_scan_path(_p(cubes), C.c_int64(n), sx, sy, sz, 1 if mode == "slice" else 0, _p(ij), bits,
                          C.byref(m.struct), C.c_double(min_proba), _p(X), _p(P), _p(lab), _p(known))
    if rc < 0:
        raise IndexError("scan %d: slice index out of bounds" % (-rc - 1))
    return X, lab, P[np.arange(n), lab], known.astype(bool), P


def matrix_indices(x, y, z, sx, sy, sz, bounds=(10, 360, -42, 42, -30, 30)):
    """common.calculate_matrix_indices (common.py:106-121)."""
    lib = load()
    out = np.empty(3, np.int32)
    lib.ro_matrix_indices(C.c_double(x), C.c_double(y), C.c_double(z), sx, sy, sz,
                          *(C.c_double(b) for b in bounds), _p(out))
    return tuple(int(v) for v in out)
