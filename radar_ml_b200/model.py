"""Model export: the sklearn object predict.py:224-225 unpickles -> flat arrays for the GPU.

``predict.py`` loads ``CalibratedClassifierCV(prefit SVC | SGDClassifier)`` built at
train.py:478-479 / 368-369 and calibrated at train.py:723-724, and only ever calls
``model.predict_proba(X)`` (predict.py:60) / ``model.predict(X)`` (train.py:217).
``GpuCalibratedClassifier`` keeps that duck-typed protocol and adds batched device entry
points.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

RADAR_MAX = 255.0  # common.py:31


@dataclass
class ModelParams:
    kind: str                 # 'svc_rbf' | 'linear'
    n_classes: int
    n_features: int
    classes: np.ndarray       # estimator.classes_
    platt_a: np.ndarray       # per class, estimator.classes_ order (one entry when binary)
    platt_b: np.ndarray
    feature_scale: float = RADAR_MAX
    # svc_rbf
    gamma: float = 0.0
    sv: np.ndarray | None = None          # (n_sv, F) float64
    dual_coef: np.ndarray | None = None   # (C-1, n_sv)  libsvm sign (SVC._dual_coef_)
    rho: np.ndarray | None = None         # (C(C-1)/2,) = -SVC._intercept_
    n_support: np.ndarray | None = None   # (C,)
    # linear
    coef: np.ndarray | None = None        # (C or 1, F)
    intercept: np.ndarray | None = None

    @property
    def n_sv(self):
        return 0 if self.sv is None else int(self.sv.shape[0])


def _inner_estimator(cal):
    if not hasattr(cal, "calibrated_classifiers_"):
        raise TypeError("expected a fitted CalibratedClassifierCV (train.py:723-724), got %r"
                        % type(cal).__name__)
    if len(cal.calibrated_classifiers_) != 1:
        raise ValueError("only the prefit/frozen form (one calibrated classifier) is supported, "
                         "as built at train.py:723")
    cc = cal.calibrated_classifiers_[0]
    if getattr(cc, "method", "sigmoid") != "sigmoid":
        raise ValueError("only sigmoid (Platt) calibration is supported (train.py:723 default)")
    est = getattr(cc, "estimator", None)
    if est is None:  # scikit-learn <= 1.1 spelling
        est = cc.base_estimator
    est = getattr(est, "estimator", est)  # FrozenEstimator -> wrapped
    return cc, est


def from_sklearn(cal, feature_scale: float = RADAR_MAX) -> ModelParams:
    """Flatten a fitted calibrated classifier.  Accepts both reference model kinds."""
    cc, est = _inner_estimator(cal)
    a = np.array([c.a_ for c in cc.calibrators], dtype=np.float64)
    b = np.array([c.b_ for c in cc.calibrators], dtype=np.float64)
    classes = np.asarray(est.classes_)
    C = len(classes)
    if hasattr(est, "support_vectors_"):
        if est.kernel != "rbf":
            raise ValueError("SVC kernel %r: only 'rbf' (train_svc.log:24-25) is supported"
                             % est.kernel)
        sv = np.ascontiguousarray(est.support_vectors_, dtype=np.float64)
        return ModelParams(
            kind="svc_rbf", n_classes=C, n_features=sv.shape[1], classes=classes,
            platt_a=a, platt_b=b, feature_scale=feature_scale, gamma=float(est._gamma), sv=sv,
            dual_coef=np.ascontiguousarray(est._dual_coef_, dtype=np.float64),
            rho=np.ascontiguousarray(-np.asarray(est._intercept_), dtype=np.float64),
            n_support=np.ascontiguousarray(est._n_support, dtype=np.int32))
    if hasattr(est, "coef_"):
        coef = np.ascontiguousarray(est.coef_, dtype=np.float64)
        return ModelParams(
            kind="linear", n_classes=C, n_features=coef.shape[1], classes=classes,
            platt_a=a, platt_b=b, feature_scale=feature_scale, coef=coef,
            intercept=np.ascontiguousarray(est.intercept_, dtype=np.float64))
    raise TypeError("unsupported estimator %r" % type(est).__name__)


class GpuCalibratedClassifier:
    """Drop-in for the ``model`` argument of predict.classifier (predict.py:56).

    ``predict_proba(X)`` takes what common.process_samples(scale=True) returns — (n,F) float32
    — and returns (n,C) float64 with columns in ``le.classes_`` order, like sklearn.
    """

    def __init__(self, params: ModelParams, engine=None, device: int = 0):
        from .engine import Engine
        self.params = params
        self.engine = engine if engine is not None else Engine(device)
        self.engine.load_model(params)
        self.classes_ = params.classes

    @classmethod
    def from_sklearn(cls, cal, engine=None, device: int = 0, feature_scale: float = RADAR_MAX):
        return cls(from_sklearn(cal, feature_scale), engine=engine, device=device)

    def bind(self):
        """The native context holds ONE model: if another wrapper sharing this engine loaded its
        own since, load ours again before scoring (two models used alternately in one process
        must never score with each other's weights)."""
        if self.engine.params is not self.params:
            self.engine.load_model(self.params)
        return self.engine

    def predict_proba(self, X):
        X = np.ascontiguousarray(X, dtype=np.float32)
        if X.ndim != 2 or X.shape[1] != self.params.n_features:
            raise ValueError("X has shape %s, expected (n, %d)" % (X.shape, self.params.n_features))
        proba, _, _ = self.bind().score_features_host(X)
        return proba.astype(np.float64)

    def predict(self, X):
        return self.classes_[np.argmax(self.predict_proba(X), axis=1)]
