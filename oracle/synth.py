"""TEST INFRASTRUCTURE — seeded synthetic radar scans and the oracle model builder.

Synthetic cube distribution follows SURVEY.md §8(d): the real sensor emits
integer-valued float32 in [0, 255], thresholded (non-zero min 13), 65-90 % zeros
(ground_truth_samples.log:1621-39916).  One anisotropic Gaussian blob per scan whose
per-axis width depends on the class, plus noise, rint, threshold, clip.

The model builder restates train.py:478-479 (SVC) and train.py:723-724 (calibration)
in the scikit-learn >= 1.6 spelling (``base_estimator=``/``cv='prefit'`` were removed).
"""
from __future__ import annotations

import numpy as np

SX, SY, SZ = 22, 31, 176          # common.py:25-27 -> predict.py:74-76
CLASSES = ("cat", "dog", "person")  # LabelEncoder order, train_svc.log:7-9

# per-class blob widths (theta, phi, r) in voxels
_SIGMA = {
    0: (1.6, 2.2, 5.0),
    1: (2.4, 3.4, 8.0),
    2: (3.6, 6.0, 13.0),
}


def make_cubes(n: int, seed: int = 1234, n_classes: int = 3, integer: bool = True,
               threshold: float = 13.0):
    """Return (cubes float32 [n,22,31,176], y int64 [n], ijk int32 [n,3])."""
    rng = np.random.default_rng(seed)
    y = rng.integers(0, n_classes, size=n)
    gi = np.arange(SX, dtype=np.float32)[:, None, None]
    gj = np.arange(SY, dtype=np.float32)[None, :, None]
    gk = np.arange(SZ, dtype=np.float32)[None, None, :]
    cubes = np.empty((n, SX, SY, SZ), dtype=np.float32)
    ijk = np.empty((n, 3), dtype=np.int32)
    for s in range(n):
        c = int(y[s]) % 3
        sg = np.array(_SIGMA[c], dtype=np.float32) * rng.uniform(0.8, 1.25, size=3).astype(np.float32)
        ci = rng.uniform(2, SX - 3)
        cj = rng.uniform(3, SY - 4)
        ck = rng.uniform(10, SZ - 11)
        amp = rng.uniform(0.5, 1.0) * 255.0
        blob = amp * np.exp(-0.5 * (((gi - ci) / sg[0]) ** 2
                                    + ((gj - cj) / sg[1]) ** 2
                                    + ((gk - ck) / sg[2]) ** 2))
        v = blob + rng.normal(0.0, 6.0, size=blob.shape).astype(np.float32)
        if integer:
            v = np.rint(v)
            v[v < threshold] = 0.0
            v = np.clip(v, 0.0, 255.0)
        cubes[s] = v.astype(np.float32)
        ijk[s] = (int(round(ci)), int(round(cj)), int(round(ck)))
    return cubes, y.astype(np.int64), ijk


def project_max(cubes: np.ndarray):
    """north_star MAX projections: (xz [n,22,176], yz [n,31,176], xy [n,22,31])."""
    return cubes.max(axis=2), cubes.max(axis=1), cubes.max(axis=3)


def project_slice(cubes: np.ndarray, ijk: np.ndarray):
    """Reference SLICE projections, predict.py:102-107."""
    n = cubes.shape[0]
    idx = np.arange(n)
    yz = cubes[idx, ijk[:, 0], :, :]
    xz = cubes[idx, :, ijk[:, 1], :]
    xy = cubes[idx, :, :, ijk[:, 2]]
    return xz, yz, xy


def features(xz, yz, xy, scale=True):
    """common.py:141-149 at zoom 1.0: concat xz|yz|xy, C-order, float32 true divide by 255."""
    n = xz.shape[0]
    f = np.concatenate([xz.reshape(n, -1), yz.reshape(n, -1), xy.reshape(n, -1)], axis=1)
    f = f.astype(np.float32)
    return f / np.float32(255.0) if scale else f


def build_svc(X_train, y_train, X_val, y_val, C=10.0, gamma=0.01):
    """train.py:478-479 + 723-724 with train_svc.log:24-31 hyper-parameters."""
    from sklearn import svm
    from sklearn.calibration import CalibratedClassifierCV
    from sklearn.frozen import FrozenEstimator

    clf = svm.SVC(C=C, gamma=gamma, kernel="rbf", probability=True, class_weight="balanced",
                  random_state=1234, cache_size=1000)
    clf.fit(X_train, y_train)
    cal = CalibratedClassifierCV(estimator=FrozenEstimator(clf))
    cal.fit(X_val, y_val)
    return cal


def build_linear(X_train, y_train, X_val, y_val):
    """train.py:368-372 SGDClassifier(loss='log') in the >=1.3 spelling + calibration.

    scikit-learn 0.24 (requirements.txt:57) converts X to float64 inside SGD fit/predict;
    1.9 keeps float32 end to end, which changes the arithmetic.  Fitting and calibrating on
    float64 copies reproduces the pinned version's float64 chain."""
    X_train = np.asarray(X_train, dtype=np.float64)
    X_val = np.asarray(X_val, dtype=np.float64)
    from sklearn import linear_model
    from sklearn.calibration import CalibratedClassifierCV
    from sklearn.frozen import FrozenEstimator

    clf = linear_model.SGDClassifier(loss="log_loss", penalty="l2", alpha=1e-4, max_iter=200,
                                     class_weight="balanced", random_state=1234, tol=1e-4)
    clf.fit(X_train, y_train)
    cal = CalibratedClassifierCV(estimator=FrozenEstimator(clf))
    cal.fit(X_val, y_val)
    return cal


class LabelEncoderLike:
    """Duck-typed ``le`` (predict.py:65 reads only ``le.classes_``)."""

    def __init__(self, classes=CLASSES):
        self.classes_ = np.array(classes)


def standard_model(n_train=909, n_val=114, seed=1234, mode="max", n_classes=3):
    """Seeded SVC used by bench.py/tests: reference split sizes (train_svc.log:11-13)."""
    cubes, y, ijk = make_cubes(n_train + n_val, seed=seed, n_classes=n_classes)
    proj = project_max(cubes) if mode == "max" else project_slice(cubes, ijk)
    X = features(*proj, scale=True)
    return build_svc(X[:n_train], y[:n_train], X[n_train:], y[n_train:])
