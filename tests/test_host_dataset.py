"""CPU tests of radar_ml_b200/dataset.py's host logic: the on-disk format of
datasets/README.md:8-20, train.py:641-674's loading / filtering / label encoding, and the metrics of
train.py:215-228 against scikit-learn.  (Feature extraction and scoring are GPU work: -m gpu.)"""
import numpy as np
import pytest


def _samples(n, rng):
    return [(rng.integers(0, 256, (22, 176)).astype(np.float32),
             rng.integers(0, 256, (31, 176)).astype(np.float32),
             rng.integers(0, 256, (22, 31)).astype(np.float32)) for _ in range(n)]


def test_dataset_pickles_concatenate_filter_and_encode(tmp_path):
    from radar_ml_b200 import dataset
    rng = np.random.default_rng(0)
    a, b = _samples(7, rng), _samples(5, rng)
    la = ["person", "dog", "cat", "dog", "bird", "person", "cat"]
    lb = ["cat", "person", "person", "fox", "dog"]
    dataset.save_dataset(tmp_path / "a.pickle", a, la)
    dataset.save_dataset(tmp_path / "b.pickle", b, lb)
    data = dataset.load_datasets(["a.pickle", "b.pickle"], prj_dir=str(tmp_path))
    assert data["labels"] == la + lb and len(data["samples"]) == 12
    assert all(np.array_equal(x, y) for s, t in zip(data["samples"], a + b) for x, y in zip(s, t))
    smp, enc, names = dataset.filter_and_encode(data, ["person", "dog", "cat"])
    # LabelEncoder order = sorted class names (train.py:670-674; train_svc.log:7-9)
    assert names == ["cat", "dog", "person"]
    kept = [l for l in la + lb if l in ("person", "dog", "cat")]
    assert len(smp) == len(kept) == 10
    assert [names[i] for i in enc] == kept
    from sklearn.preprocessing import LabelEncoder
    assert np.array_equal(enc, LabelEncoder().fit_transform(kept))
    # a class that never occurs is simply absent
    _, enc2, names2 = dataset.filter_and_encode(data, ["dog", "zebra"])
    assert names2 == ["dog"] and (enc2 == 0).all() and len(enc2) == 3


@pytest.mark.parametrize("seed", [1, 2])
def test_confusion_matrix_and_report_match_sklearn(seed):
    from sklearn import metrics
    from radar_ml_b200 import dataset
    rng = np.random.default_rng(seed)
    y_true = rng.integers(0, 3, 500)
    y_pred = np.where(rng.random(500) < 0.7, y_true, rng.integers(0, 3, 500))
    if seed == 2:
        y_pred[y_pred == 2] = 0          # a class that is never predicted: precision 0, no division warning
    cm = dataset.confusion_matrix(y_true, y_pred, 3)
    assert np.array_equal(cm, metrics.confusion_matrix(y_true, y_pred, labels=[0, 1, 2]))
    names = ["cat", "dog", "person"]
    rep = dataset.classification_report(cm, names)
    ref = metrics.classification_report(y_true, y_pred, target_names=names, output_dict=True, zero_division=0)
    assert abs(rep["accuracy"] - metrics.accuracy_score(y_true, y_pred)) < 1e-15 and rep["support"] == 500
    for row in rep["per_class"]:
        r = ref[row["class"]]
        assert abs(row["precision"] - r["precision"]) < 1e-12 and abs(row["recall"] - r["recall"]) < 1e-12
        assert abs(row["f1"] - r["f1-score"]) < 1e-12 and row["support"] == r["support"]
