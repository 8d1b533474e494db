"""Host-side mirror of the reference's ``predict`` module (predict.py:34-131).

``calc_proj_zoom`` / ``classifier`` / ``predict`` keep the reference's names, argument
meaning and return conventions; the arithmetic runs on the GPU.  ``classify_cubes`` is the
batched extension (one predict.py:93-119 iteration for B scans at once).
"""
from __future__ import annotations

import logging

import numpy as np

from . import common
from .model import GpuCalibratedClassifier

logger = logging.getLogger(__name__)

# predict.py:20-27
RADAR_THRESHOLD = 5
MTI = True
R_MIN, R_MAX, R_RES = 10, 360, 2
THETA_MIN, THETA_MAX, THETA_RES = -42, 42, 4
PHI_MIN, PHI_MAX, PHI_RES = -30, 30, 2

_converted = {}  # id(sklearn model) -> GpuCalibratedClassifier


def as_gpu_model(model):
    """Accept either our GPU model or the sklearn object predict.py:224-225 unpickles."""
    if isinstance(model, GpuCalibratedClassifier):
        return model
    key = id(model)
    hit = _converted.get(key)
    if hit is None or hit[0] is not model:
        gm = GpuCalibratedClassifier.from_sklearn(model, engine=common.get_engine())
        _converted[key] = (model, gm)
        return gm
    return hit[1]


def calc_proj_zoom(train_size_x, train_size_y, train_size_z, size_x, size_y, size_z):
    """predict.py:34-54 — zoom factors that map the scan arena onto the training arena."""
    x_zoom = train_size_x / size_x
    y_zoom = train_size_y / size_y
    z_zoom = train_size_z / size_z
    logger.debug(f'zoom: {x_zoom}, {y_zoom}, {z_zoom}')
    return common.ProjZoom(xy=[x_zoom, y_zoom], xz=[x_zoom, z_zoom], yz=[y_zoom, z_zoom])


def classifier(observation, model, le, min_proba=0.7):
    """predict.py:56-70 — classify one feature vector; returns (name, proba).

    ``model`` may be the pickled sklearn CalibratedClassifierCV (it is exported to the GPU
    once and cached) or a GpuCalibratedClassifier.  'Unknown' is returned, not raised, when
    the best probability is below ``min_proba`` (predict.py:65-68).
    """
    gm = as_gpu_model(model)
    preds = gm.predict_proba(np.asarray(observation).reshape(1, -1))[0]
    j = np.argmax(preds)
    proba = preds[j]
    logger.debug('classifier proba {} name {}'.format(proba, le.classes_[j]))
    name = le.classes_[j] if proba >= min_proba else 'Unknown'
    return name, proba


def classify_cubes(cubes, model, le=None, min_proba=0.7, mode='max', ijk=None,
                   proj_mask=common.ProjMask(True, True, True)):
    """Batched predict.py:93-119: cubes [B,22,31,176] (CUDA tensor or host ndarray).

    Returns (labels int32 [B] argmax index, proba float32 [B] best probability,
    known bool [B] (= name != 'Unknown'), proba_matrix float32 [B,C]) as numpy arrays, plus
    the class names when ``le`` is given.
    """
    import torch
    from ._lib import NonIntegralInput
    gm = as_gpu_model(model)
    eng = gm.bind()
    if isinstance(cubes, np.ndarray):
        P, lab, known = eng.predict_host(np.ascontiguousarray(cubes, dtype=np.float32), mode=mode,
                                         ijk=ijk, mask=proj_mask, min_proba=min_proba)
        known = known.astype(bool)
    else:
        if ijk is not None and not isinstance(ijk, torch.Tensor):
            ijk = torch.as_tensor(np.asarray(ijk), dtype=torch.int32)
        Pd, labd, knownd = eng.predict(cubes, mode=mode, ijk=ijk, mask=proj_mask,
                                       min_proba=min_proba)
        try:
            eng.check_status()
        except NonIntegralInput:
            # real-valued float32 cubes: float32 features + the exact general-precision scorer
            if cubes.dtype != torch.float32:
                raise
            eng.set_precision(True)
            try:
                Pd, labd, knownd = eng.predict(cubes, mode=mode, ijk=ijk, mask=proj_mask,
                                               min_proba=min_proba)
                eng.check_status()
            finally:
                eng.set_precision(False)
        P, lab, known = Pd.cpu().numpy(), labd.cpu().numpy(), knownd.cpu().numpy().astype(bool)
    best = P[np.arange(P.shape[0]), lab]
    if le is None:
        return lab, best, known, P
    names = np.where(known, np.asarray(le.classes_)[lab], 'Unknown')
    return lab, best, known, P, names


def _classify_zoomed(cubes, ijk, gm, le, min_proba, proj_mask, proj_zoom, dims):
    """predict.py:102-119 when the scan arena differs from the training arena: slices at the
    scan arena's size (K1), ndimage.zoom to the training size as separable operators
    (common.py:143), /255, then the general-precision scorer (zoomed values are not integers)."""
    import torch
    eng = gm.bind()
    size_x, size_y, size_z = dims
    saved = eng.dims
    eng.set_arena(size_x, size_y, size_z)
    try:
        eng.set_affine(0.0, common.RADAR_MAX, False)
        d = torch.from_numpy(np.ascontiguousarray(cubes, dtype=np.float32)).to(eng.device)
        raw = eng.project(d, mode='slice', ijk=torch.from_numpy(np.ascontiguousarray(ijk, dtype=np.int32)),
                          mask=common.ProjMask(True, True, True))
        eng.check_status()
        shapes = [(size_x, size_z), (size_y, size_z), (size_x, size_y)]
        offs = [0, size_x * size_z, size_x * size_z + size_y * size_z]
        views = []
        for q in range(3):
            if proj_mask[q]:
                h, w = shapes[q]
                eng.set_zoom(q, common.zoom_operator(h, proj_zoom[q][0]), common.zoom_operator(w, proj_zoom[q][1]))
                views.append(raw[:, offs[q]:])
            else:
                views.append(None)
        eng.set_affine(0.0, common.RADAR_MAX, True)
        F = raw.shape[1]
        feats = eng.process_samples_zoom(views[0], views[1], views[2], mask=proj_mask, scale=True,
                                         strides=[F, F, F])
    finally:
        eng.set_affine(0.0, common.RADAR_MAX, True)
        eng.set_arena(*saved)
    from ._lib import OutOfRangeInput
    proba, label, known = eng.score(feats, None, min_proba)
    try:
        eng.check_status()
    except OutOfRangeInput:   # spline overshoot below 0 / above 255: float64 CUDA-core scorer
        proba, label, known = eng.score(feats, None, min_proba, exact=True)
    P, lab, known = proba.cpu().numpy(), label.cpu().numpy(), known.cpu().numpy().astype(bool)
    best = P[np.arange(P.shape[0]), lab]
    names = np.where(known, np.asarray(le.classes_)[lab], 'Unknown')
    return lab, best, known, P, names


def predict(min_proba, model, le, proj_mask, radar=None, max_scans=None):
    """predict.py:72-131 — the live loop: trigger, locate targets, slice, classify.

    ``radar`` is any object with the Walabot SDK surface the loop uses (Trigger,
    GetSensorTargets, GetRawImage, Stop, Disconnect, Clean); by default the real
    ``WalabotAPI`` module is imported, exactly like the reference, and its absence is an
    error.  ``max_scans`` bounds the loop for replay/tests (the reference loops until ^C).
    Returns the list of (name, proba) it logged.
    """
    if radar is None:
        import WalabotAPI as radar  # noqa: N813  (predict.py:7)
    train_size_x, train_size_y, train_size_z = common.arena_size()
    logger.debug(f'train_size: {train_size_x}, {train_size_y}, {train_size_z}')
    gm = as_gpu_model(model)
    results = []
    scans = 0
    try:
        while max_scans is None or scans < max_scans:
            radar.Trigger()
            scans += 1
            targets = radar.GetSensorTargets()
            if not targets:
                continue
            raw_image, size_x, size_y, size_z, _ = radar.GetRawImage()
            raw_image_np = np.array(raw_image, dtype=np.float32)
            proj_zoom = calc_proj_zoom(train_size_x, train_size_y, train_size_z,
                                       size_x, size_y, size_z)
            n = len(targets)
            xs = np.array([t.xPosCm for t in targets], dtype=np.float64)
            ys = np.array([t.yPosCm for t in targets], dtype=np.float64)
            zs = np.array([t.zPosCm for t in targets], dtype=np.float64)
            ijk = np.asarray(common.calculate_matrix_indices(xs, ys, zs, size_x, size_y, size_z))
            if common._unit_zoom(proj_zoom, proj_mask):
                # one cube, n index triples: the cube crosses PCIe once (rml_predict_targets_host)
                P, lab, known = gm.bind().predict_targets_host(raw_image_np, ijk.reshape(n, 3),
                                                               mask=proj_mask, min_proba=min_proba)
                known = known.astype(bool)
                best = P[np.arange(n), lab]
                names = np.where(known, np.asarray(le.classes_)[lab], 'Unknown')
            else:
                cubes = np.ascontiguousarray(np.broadcast_to(raw_image_np, (n,) + raw_image_np.shape))
                lab, best, known, P, names = _classify_zoomed(
                    cubes, ijk.reshape(n, 3), gm, le, min_proba, proj_mask, proj_zoom,
                    (size_x, size_y, size_z))
            for t, target in enumerate(targets):
                logger.info('**********')
                logger.info('Target #{}:\nx: {}\ny: {}\nz: {}\namplitude: {}\n'.format(
                    t + 1, target.xPosCm, target.yPosCm, target.zPosCm, target.amplitude))
                logger.info(f'Detected {names[t]} with probability {best[t]}')
                logger.info('**********')
                results.append((str(names[t]), float(best[t])))
    except KeyboardInterrupt:
        pass
    finally:
        radar.Stop()
        radar.Disconnect()
        radar.Clean()
        logger.info('Successful radar shutdown.')
    return results
