"""Dataset pickles and batched model evaluation.

On-disk format (datasets/README.md:8-20, written by ground_truth_samples.py:581-587):
``{'samples': [(xz, yz, xy), ...], 'labels': [str, ...]}`` with float32 projections in
[0, 255].  ``load_datasets`` / ``filter_and_encode`` follow train.py:641-674; ``evaluate_model``
follows train.py:215-228 with the predictions scored on the GPU (the confusion-matrix PNG of
train.py:221-224 is plotting and out of scope).
"""
from __future__ import annotations

import logging
import os
import pickle

import numpy as np

from . import common

logger = logging.getLogger(__name__)


def load_datasets(paths, prj_dir='./'):
    """train.py:641-654 — concatenate one or more dataset pickles."""
    samples, labels = [], []
    for dataset in paths:
        logger.info(f'Opening dataset: {dataset}')
        with open(os.path.join(prj_dir, dataset), 'rb') as fp:
            data_pickle = pickle.load(fp)
        logger.debug(f'Found class labels: {set(data_pickle["labels"])}.')
        samples.extend(data_pickle['samples'])
        labels.extend(data_pickle['labels'])
    return {'samples': samples, 'labels': labels}


def save_dataset(path, samples, labels):
    """ground_truth_samples.py:581-587 format."""
    with open(path, 'wb') as fp:
        pickle.dump({'samples': list(samples), 'labels': list(labels)}, fp)


def filter_and_encode(data, desired_labels):
    """train.py:656-674 — keep the desired classes; LabelEncoder = sorted unique names."""
    keep = [i for i, l in enumerate(data['labels']) if l in desired_labels]
    samples = [data['samples'][i] for i in keep]
    names = [data['labels'][i] for i in keep]
    class_names = sorted(set(names))
    encoded = np.array([class_names.index(n) for n in names], dtype=np.int64)
    return samples, encoded, class_names


def features(samples, proj_mask=common.ProjMask(True, True, True), batch=4096):
    """process_samples(scale=True) over a whole dataset in GPU batches -> (n, F) float32.
    (train.py:667 scales by 255 first and then calls process_samples unscaled: same values.)"""
    out = [common.process_samples(samples[lo:lo + batch], proj_mask=proj_mask, scale=True)
           for lo in range(0, len(samples), batch)]
    return np.concatenate(out, axis=0) if out else np.zeros((0, 0), np.float32)


def confusion_matrix(y_true, y_pred, n_classes):
    cm = np.zeros((n_classes, n_classes), dtype=np.int64)
    np.add.at(cm, (np.asarray(y_true, dtype=np.int64), np.asarray(y_pred, dtype=np.int64)), 1)
    return cm


def classification_report(cm, target_names):
    """precision / recall / f1 / support per class + accuracy, like sklearn's text report."""
    rows = []
    tp = np.diag(cm).astype(np.float64)
    pred, true = cm.sum(axis=0).astype(np.float64), cm.sum(axis=1).astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        prec = np.where(pred > 0, tp / pred, 0.0)
        rec = np.where(true > 0, tp / true, 0.0)
        f1 = np.where(prec + rec > 0, 2 * prec * rec / (prec + rec), 0.0)
    for k, name in enumerate(target_names):
        rows.append({'class': name, 'precision': float(prec[k]), 'recall': float(rec[k]),
                     'f1': float(f1[k]), 'support': int(true[k])})
    acc = float(tp.sum() / max(cm.sum(), 1))
    return {'per_class': rows, 'accuracy': acc, 'support': int(cm.sum())}


def evaluate_model(model, X_test, y_test, target_names, cm_name=None):
    """train.py:215-228 — accuracy, confusion matrix and classification report; ``model`` is a
    GpuCalibratedClassifier (or the sklearn object, converted once).  Returns the three."""
    from .predict import as_gpu_model
    gm = as_gpu_model(model)
    classes = np.asarray(gm.classes_)
    y_pred = gm.predict(np.asarray(X_test, dtype=np.float32))
    lut = {c: i for i, c in enumerate(classes.tolist())}
    y_pred_idx = np.array([lut[v] for v in y_pred.tolist()], dtype=np.int64)
    y_true_idx = np.array([lut[v] for v in np.asarray(y_test).tolist()], dtype=np.int64)
    cm = confusion_matrix(y_true_idx, y_pred_idx, len(classes))
    report = classification_report(cm, list(target_names))
    logger.info(f'Accuracy: {report["accuracy"]}')
    logger.info(f'Confusion matrix:\n{cm}')
    logger.info(f'Classification report:\n{report}')
    return report['accuracy'], cm, report
