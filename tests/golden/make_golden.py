#!/usr/bin/env python
"""Mint tests/golden/*.npz by running the UNMODIFIED reference in the build container.

The reference ships no tests or golden vectors for this path, so the pins
are outputs of the reference's own code: ``common.process_samples`` and ``predict.classifier``
imported from /root/reference (with the WalabotAPI stub in oracle/refimport.py), driven by
scikit-learn 1.9.0 models built the way train.py:478-479, 723-724 builds them.

    python tests/golden/make_golden.py        # needs /root/reference; rewrites the fixtures

Fixtures (all small, stored compressed; integer data as uint8):
  svc_max.npz / svc_slice.npz  seeded cubes -> reference features + classifier outputs, the
                               fitted model's flat parameters, sklearn intermediates
  generated.npz                first samples of train-results/sgan/generated_data_0230.pickle.save
                               (non-integer, out-of-range values) -> reference process_samples
  real_xy.npz                  the 491 real 22x31 xy projections + labels printed in
                               ground_truth_samples.log:1621-39916 -> xy-only SVC -> reference
                               classifier outputs
  indices.npz                  common.calculate_matrix_indices on seeded targets
  shapes.json                  layout known-answers from the reference logs
"""
from __future__ import annotations

import json
import os
import pickle
import re
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refimport, restate, synth  # noqa: E402

REF = refimport.REF_DIR


def model_arrays(cal, prefix="m_"):
    p = restate.export_params(cal)
    d = {prefix + "n_classes": p.n_classes, prefix + "gamma": p.gamma,
         prefix + "dual_coef": p.dual_coef, prefix + "rho": p.rho, prefix + "n_support": p.n_support,
         prefix + "platt_a": p.platt_a, prefix + "platt_b": p.platt_b}
    u = np.rint(p.sv * 255.0)
    assert np.array_equal((u.astype(np.float32) / np.float32(255)).astype(np.float64), p.sv), \
        "support vectors are not float32(u/255)"
    d[prefix + "sv_u8"] = u.astype(np.uint8)
    return d


def svc_fixture(mode, seed, rc, rp):
    n_fit, n_val, n_test = 150, 40, 24
    cubes, y, ijk = synth.make_cubes(n_fit + n_val + n_test, seed=seed)
    proj = synth.project_max(cubes) if mode == "max" else synth.project_slice(cubes, ijk)
    X = synth.features(*proj)
    cal = synth.build_svc(X[:n_fit], y[:n_fit], X[n_fit:n_fit + n_val], y[n_fit:n_fit + n_val])
    le = synth.LabelEncoderLike()
    test = slice(n_fit + n_val, n_fit + n_val + n_test)
    tc, tijk = cubes[test], ijk[test]
    feats, names, probas = [], [], []
    for s in range(n_test):
        t = restate.project(tc[s], mode, tuple(int(v) for v in tijk[s]))
        obs = rc.process_samples([t], proj_mask=rc.ProjMask(True, True, True),
                                 proj_zoom=rp.calc_proj_zoom(22, 31, 176, 22, 31, 176), scale=True)
        name, proba = rp.classifier(obs, cal, le, 0.7)
        feats.append(obs[0]); names.append(name); probas.append(proba)
    feats = np.asarray(feats)
    est = restate.unwrap(cal)
    assert np.array_equal(tc, np.rint(tc)) and tc.min() >= 0 and tc.max() <= 255
    d = dict(cubes_u8=tc.astype(np.uint8), ijk=tijk, ref_features=feats,
             ref_names=np.array(names), ref_proba=np.array(probas, dtype=np.float64),
             sk_predict_proba=cal.predict_proba(feats), sk_decision=est.decision_function(feats),
             classes=np.array(le.classes_))
    d.update(model_arrays(cal))
    np.savez_compressed(os.path.join(HERE, "svc_%s.npz" % mode), **d)
    print("svc_%s: n_sv=%d names=%s" % (mode, est.support_vectors_.shape[0], sorted(set(names))))


def generated_fixture(rc):
    with open(os.path.join(REF, "train-results/sgan/generated_data_0230.pickle.save"), "rb") as f:
        data = pickle.load(f)
    samples = data["samples"][:6]
    xz = np.stack([s[0] for s in samples]); yz = np.stack([s[1] for s in samples])
    xy = np.stack([s[2] for s in samples])
    out = {}
    for tag, mask in (("all", (True, True, True)), ("xz_xy", (True, False, True)), ("yz", (False, True, False))):
        for sc in (False, True):
            out["feat_%s_%d" % (tag, sc)] = rc.process_samples(samples, proj_mask=rc.ProjMask(*mask), scale=sc)
    np.savez_compressed(os.path.join(HERE, "generated.npz"), xz=xz, yz=yz, xy=xy, **out)
    print("generated: min %.3f max %.3f" % (min(xz.min(), yz.min(), xy.min()), max(xz.max(), yz.max(), xy.max())))


def parse_real_xy():
    """ground_truth_samples.log: numpy repr of the dataset dict; every third array is a full xy."""
    with open(os.path.join(REF, "ground_truth_samples.log")) as f:
        text = f.read()
    text = text[text.index("Data dump:"):]
    text = text[:text.index("Saving data file")]
    bodies = re.findall(r"array\((.*?)dtype=float32\)", text, flags=re.S)
    xy = []
    for idx, body in enumerate(bodies):
        if idx % 3 != 2:
            continue
        nums = re.findall(r"-?\d+\.?\d*(?:e[-+]?\d+)?", body)
        arr = np.array([float(v) for v in nums], dtype=np.float32)
        assert arr.size == 22 * 31, (idx, arr.size)
        xy.append(arr.reshape(22, 31))
    labels = re.findall(r"'(person|dog|cat)'", text[text.rindex("'labels': ["):])
    xy = np.stack(xy)
    assert xy.shape[0] == len(labels) == 491, (xy.shape, len(labels))
    return xy, labels


def real_xy_fixture(rc, rp):
    xy, labels = parse_real_xy()
    assert np.array_equal(xy, np.rint(xy)) and xy.min() >= 0 and xy.max() <= 255
    classes = np.array(sorted(set(labels)))
    y = np.searchsorted(classes, np.array(labels))
    rng = np.random.default_rng(7)
    perm = rng.permutation(len(y))
    samples = [(None, None, xy[i]) for i in range(len(y))]
    mask = rc.ProjMask(False, False, True)
    X = rc.process_samples(samples, proj_mask=mask, scale=True)
    tr, va, te = perm[:380], perm[380:440], perm[440:]
    cal = synth.build_svc(X[tr], y[tr], X[va], y[va])

    class LE:
        classes_ = classes
    names, probas = [], []
    for i in te:
        obs = rc.process_samples([samples[i]], proj_mask=mask, scale=True)
        name, proba = rp.classifier(obs, cal, LE, 0.7)
        names.append(name); probas.append(proba)
    d = dict(xy_u8=xy.astype(np.uint8), labels=np.array(labels), classes=classes, test_idx=te,
             ref_features_test=X[te], ref_names=np.array(names),
             ref_proba=np.array(probas, dtype=np.float64), sk_predict_proba=cal.predict_proba(X[te]))
    d.update(model_arrays(cal))
    np.savez_compressed(os.path.join(HERE, "real_xy.npz"), **d)
    acc = np.mean(cal.predict(X[te]) == y[te])
    print("real_xy: %d samples, classes %s, test acc %.3f, names %s" % (len(y), classes, acc, sorted(set(names))))


def indices_fixture(rc):
    rng = np.random.default_rng(3)
    n = 512
    r = rng.uniform(12, 355, n)
    th = np.deg2rad(rng.uniform(-41.5, 41.5, n))
    ph = np.deg2rad(rng.uniform(-29.5, 29.5, n))
    x = r * np.sin(th); yv = r * np.cos(th) * np.sin(ph); z = r * np.cos(th) * np.cos(ph)
    xyz = np.stack([x, yv, z], axis=1)
    xyz[0] = (0.0, 0.0, 100.0)          # dead centre: theta = phi = 0 -> exact .5 boundaries
    xyz[1] = (10.0, -5.0, 50.0)
    ijk = np.array([rc.calculate_matrix_indices(*row, 22, 31, 176) for row in xyz], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "indices.npz"), xyz=xyz, ijk=ijk)
    print("indices: i %d..%d j %d..%d k %d..%d" % (ijk[:, 0].min(), ijk[:, 0].max(), ijk[:, 1].min(),
                                                    ijk[:, 1].max(), ijk[:, 2].min(), ijk[:, 2].max()))


def main():
    rc, rp = refimport.load()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        svc_fixture("max", 101, rc, rp)
        svc_fixture("slice", 202, rc, rp)
        generated_fixture(rc)
        real_xy_fixture(rc, rp)
        indices_fixture(rc)
    shapes = {"raw_image": [22, 31, 176],            # ground_truth_samples.log:10, predict.log:13
              "projection_yz": [31, 176], "projection_xz": [22, 176], "projection_xy": [22, 31],
              "feature_vector_length": 10010,         # train-results/train_svc.log:19
              "train_size": [22, 31, 176],
              "classes": ["cat", "dog", "person"]}    # train_svc.log:7-9
    with open(os.path.join(HERE, "shapes.json"), "w") as f:
        json.dump(shapes, f, indent=1)


if __name__ == "__main__":
    main()
