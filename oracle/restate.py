"""TEST INFRASTRUCTURE — float64 numpy restatement of the reference hot path.

Each function cites the reference (``/root/reference``) line range it follows, or the
scikit-learn 1.9.0 source (``SK/`` = site-packages/sklearn) for arithmetic that lives in
the reference's pinned third-party dependency (requirements.txt:57, scikit-learn==0.24.0;
same prediction chain in 1.9.0).

Pinned by tests/test_oracle.py against (a) the unmodified reference imported in the build
container (tests/golden/*.npz, minted by tests/golden/make_golden.py) and (b) sklearn
itself (max |dproba| ~1e-15).
"""
from __future__ import annotations

import collections
from dataclasses import dataclass, field

import numpy as np

# --- common.py:25-31, 40, 43 -------------------------------------------------------
R_MIN, R_MAX, R_RES = 10, 360, 2
THETA_MIN, THETA_MAX, THETA_RES = -42, 42, 4
PHI_MIN, PHI_MAX, PHI_RES = -30, 30, 2
RADAR_MAX = 255.0
ProjMask = collections.namedtuple("ProjMask", ["xz", "yz", "xy"])
ProjZoom = collections.namedtuple("ProjZoom", ["xz", "yz", "xy"])


def arena_dims():
    """predict.py:74-76 -> (size_x, size_y, size_z) = (22, 31, 176)."""
    sz = int((R_MAX - R_MIN) / R_RES) + 1
    sy = int((PHI_MAX - PHI_MIN) / PHI_RES) + 1
    sx = int((THETA_MAX - THETA_MIN) / THETA_RES) + 1
    return sx, sy, sz


def cartesian_to_spherical(x, y, z):
    """common.py:93-97."""
    r = np.sqrt(np.power(x, 2) + np.power(y, 2) + np.power(z, 2))
    phi = np.arctan2(y, z)
    theta = np.arcsin(x / r)
    return r, np.rad2deg(theta), np.rad2deg(phi)


def calculate_matrix_indices(x, y, z, size_x, size_y, size_z):
    """common.py:106-121 — int() truncates toward zero."""
    r, theta, phi = cartesian_to_spherical(x, y, z)
    i = int((theta - THETA_MIN) * (size_x - 1) / (THETA_MAX - THETA_MIN))
    j = int((phi - PHI_MIN) * (size_y - 1) / (PHI_MAX - PHI_MIN))
    k = int((r - R_MIN) * (size_z - 1) / (R_MAX - R_MIN))
    return i, j, k


def calc_proj_zoom(train_size_x, train_size_y, train_size_z, size_x, size_y, size_z):
    """predict.py:34-54."""
    xz_, yz_, zz_ = train_size_x / size_x, train_size_y / size_y, train_size_z / size_z
    return ProjZoom(xy=[xz_, yz_], xz=[xz_, zz_], yz=[yz_, zz_])


def project(cube: np.ndarray, mode: str, ijk=None):
    """Projection extraction.

    mode 'slice': predict.py:102-107 (= ground_truth_samples.py:413-419).
    mode 'max'  : BASELINE.json north_star axis-max projections, same shapes.
    Returns the reference sample tuple order (xz, yz, xy) (predict.py:113).
    """
    if mode == "slice":
        i, j, k = ijk
        return cube[:, j, :], cube[i, :, :], cube[:, :, k]
    if mode == "max":
        return cube.max(axis=1), cube.max(axis=0), cube.max(axis=2)
    raise ValueError(mode)


def process_samples(samples, proj_mask=ProjMask(True, True, True),
                    proj_zoom=ProjZoom([1.0, 1.0], [1.0, 1.0], [1.0, 1.0]), scale=False,
                    always_zoom=False):
    """common.py:123-149.  ndimage.zoom(order=3) is an exact identity at zoom 1.0
    (checked against the reference in tests/test_oracle.py), so the parity oracle skips it there;
    other factors go through scipy, exactly like the reference.  ``always_zoom=True`` calls
    ndimage.zoom unconditionally like common.py:143 does (the spline prefilter runs even at
    1.0): the CPU TIMING arm uses it so that the reference's cost is not understated."""
    def make(t):
        wanted = []
        for idx, p in enumerate(t):
            if not proj_mask[idx]:
                continue
            z = proj_zoom[idx]
            if not always_zoom and float(z[0]) == 1.0 and float(z[1]) == 1.0:
                wanted.append(np.asarray(p))
            else:
                from scipy import ndimage
                wanted.append(ndimage.zoom(p, z))
        cat = np.concatenate(wanted, axis=None)
        return cat / RADAR_MAX if scale else cat
    return np.array([make(t) for t in samples])


# --- model container ---------------------------------------------------------------
@dataclass
class SvcParams:
    """Flat view of CalibratedClassifierCV(FrozenEstimator(SVC)) (train.py:478-479, 723)."""
    n_classes: int
    gamma: float
    sv: np.ndarray          # (n_sv, F) float64   SVC.support_vectors_
    dual_coef: np.ndarray   # (C-1, n_sv) float64 SVC._dual_coef_  (libsvm sign)
    rho: np.ndarray         # (C(C-1)/2,) float64 = -SVC._intercept_
    n_support: np.ndarray   # (C,) int32
    platt_a: np.ndarray     # (C or 1,) float64, in estimator.classes_ order
    platt_b: np.ndarray
    classes: np.ndarray = field(default_factory=lambda: np.arange(3))
    kind: str = "svc_rbf"
    coef: np.ndarray | None = None       # linear: (C or 1, F)
    intercept: np.ndarray | None = None  # linear: (C or 1,)


def unwrap(cal):
    est = cal.calibrated_classifiers_[0].estimator
    return getattr(est, "estimator", est)  # FrozenEstimator -> inner


def export_params(cal) -> SvcParams:
    """Read the fitted sklearn object the way predict.py:224-225 unpickles it."""
    cc = cal.calibrated_classifiers_[0]
    assert len(cal.calibrated_classifiers_) == 1 and cc.method == "sigmoid"
    est = unwrap(cal)
    a = np.array([c.a_ for c in cc.calibrators], dtype=np.float64)
    b = np.array([c.b_ for c in cc.calibrators], dtype=np.float64)
    if hasattr(est, "support_vectors_"):
        assert est.kernel == "rbf"
        return SvcParams(
            n_classes=len(est.classes_), gamma=float(est._gamma),
            sv=np.asarray(est.support_vectors_, dtype=np.float64),
            dual_coef=np.asarray(est._dual_coef_, dtype=np.float64),
            rho=-np.asarray(est._intercept_, dtype=np.float64),
            n_support=np.asarray(est._n_support, dtype=np.int32),
            platt_a=a, platt_b=b, classes=np.asarray(est.classes_))
    return SvcParams(
        n_classes=len(est.classes_), gamma=0.0, sv=np.zeros((0, est.coef_.shape[1])),
        dual_coef=np.zeros((0, 0)), rho=np.zeros(0), n_support=np.zeros(0, np.int32),
        platt_a=a, platt_b=b, classes=np.asarray(est.classes_), kind="linear",
        coef=np.asarray(est.coef_, dtype=np.float64),
        intercept=np.asarray(est.intercept_, dtype=np.float64))


# --- scoring chain -----------------------------------------------------------------
def rbf_kvalues(X: np.ndarray, p: SvcParams) -> np.ndarray:
    """SK/svm/src/libsvm/svm.cpp:461-514 Kernel::k_function (RBF, dense):
    d = x - sv ; exp(-gamma * dot(d, d)), all float64 (X cast at SK/svm/_base.py:590)."""
    X = np.asarray(X, dtype=np.float64)
    out = np.empty((X.shape[0], p.sv.shape[0]), dtype=np.float64)
    for r in range(X.shape[0]):
        d = p.sv - X[r]
        out[r] = np.exp(-p.gamma * np.einsum("ij,ij->i", d, d))
    return out


def ovo_decision(kv: np.ndarray, p: SvcParams) -> np.ndarray:
    """SK/svm/src/libsvm/svm.cpp:2864-2893 svm_predict_values pair loop."""
    C = p.n_classes
    start = np.concatenate([[0], np.cumsum(p.n_support)[:-1]]).astype(int)
    dec = np.empty((kv.shape[0], C * (C - 1) // 2), dtype=np.float64)
    q = 0
    for i in range(C):
        for j in range(i + 1, C):
            si, sj, ci, cj = start[i], start[j], int(p.n_support[i]), int(p.n_support[j])
            s = kv[:, si:si + ci] @ p.dual_coef[j - 1, si:si + ci]
            s = s + kv[:, sj:sj + cj] @ p.dual_coef[i, sj:sj + cj]
            dec[:, q] = s - p.rho[q]
            q += 1
    return dec


def ovr_from_ovo(dec: np.ndarray, C: int) -> np.ndarray:
    """SK/svm/_base.py:798-828 + SK/utils/multiclass.py:557-599.
    sklearn calls _ovr_decision_function(dec < 0, -dec, C)."""
    if C == 2:
        return -dec.ravel()  # SK/svm/_base.py:_decision_function binary sign flip
    pred = dec < 0
    conf = -dec
    votes = np.zeros((dec.shape[0], C))
    soc = np.zeros((dec.shape[0], C))
    q = 0
    for i in range(C):
        for j in range(i + 1, C):
            soc[:, i] -= conf[:, q]
            soc[:, j] += conf[:, q]
            votes[pred[:, q] == 0, i] += 1
            votes[pred[:, q] == 1, j] += 1
            q += 1
    return votes + soc / (3 * (np.abs(soc) + 1))


def expit(x):
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def platt_normalise(f: np.ndarray, p: SvcParams) -> np.ndarray:
    """SK/calibration.py:781-850 (_CalibratedClassifier.predict_proba, sigmoid) and
    :1065 (_SigmoidCalibration.predict = expit(-(a*T+b)))."""
    C = p.n_classes
    n = f.shape[0]
    proba = np.zeros((n, C))
    if C == 2:
        proba[:, 1] = expit(-(p.platt_a[0] * f.reshape(n) + p.platt_b[0]))
        proba[:, 0] = 1.0 - proba[:, 1]
    else:
        for k in range(C):
            proba[:, k] = expit(-(p.platt_a[k] * f[:, k] + p.platt_b[k]))
        den = proba.sum(axis=1)[:, None]
        uniform = np.full_like(proba, 1.0 / C)
        proba = np.divide(proba, den, out=uniform, where=den != 0)
    proba[(1.0 < proba) & (proba <= 1.0 + 1e-5)] = 1.0
    return proba


def decision_function(X: np.ndarray, p: SvcParams) -> np.ndarray:
    if p.kind == "linear":
        # train.py:368-369 SGDClassifier.decision_function = X @ coef.T + intercept
        f = np.asarray(X, dtype=np.float64) @ p.coef.T + p.intercept
        return f.ravel() if f.shape[1] == 1 else f
    return ovr_from_ovo(ovo_decision(rbf_kvalues(X, p), p), p.n_classes)


def predict_proba(X: np.ndarray, p: SvcParams) -> np.ndarray:
    """What ``model.predict_proba`` returns at predict.py:60."""
    f = decision_function(X, p)
    return platt_normalise(np.asarray(f), p)


def classifier(observation, p: SvcParams, classes, min_proba=0.7):
    """predict.py:56-70."""
    preds = predict_proba(np.asarray(observation).reshape(1, -1), p)[0]
    j = int(np.argmax(preds))
    proba = preds[j]
    name = classes[j] if proba >= min_proba else "Unknown"
    return name, proba


def classify_batch(X, p: SvcParams, min_proba=0.7):
    """Batched form of predict.py:56-70: (label int32, proba f64, known bool, proba matrix)."""
    P = predict_proba(X, p)
    lab = np.argmax(P, axis=1).astype(np.int32)
    pr = P[np.arange(P.shape[0]), lab]
    return lab, pr, pr >= min_proba, P


def scan_path(cubes, p: SvcParams, mode="max", ijk=None, mask=ProjMask(True, True, True),
              min_proba=0.7):
    """End-to-end per-scan loop as predict.py:90-119 runs it (zoom 1.0)."""
    feats = []
    for s in range(cubes.shape[0]):
        t = project(cubes[s], mode, None if ijk is None else tuple(int(v) for v in ijk[s]))
        feats.append(process_samples([t], proj_mask=mask, scale=True)[0])
    X = np.asarray(feats, dtype=np.float32)
    return (X,) + classify_batch(X, p, min_proba)
