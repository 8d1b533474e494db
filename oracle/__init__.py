"""TEST INFRASTRUCTURE — CPU oracle for the radar-ml per-scan classification hot path.

Everything under ``oracle/`` is a checker, never a product path:

* ``restate.py``  numpy/float64 restatement of the reference algorithm
                  (projection -> process_samples -> libsvm RBF -> OvR -> Platt -> argmax).
* ``synth.py``    seeded synthetic radar cubes + the sklearn model builder that mirrors
                  train.py:478-479, 723-724 (the third-party engine the reference calls).
* ``refimport.py`` imports the UNMODIFIED reference (common.py / predict.py) from
                  /root/reference with a WalabotAPI stub; only usable in the build
                  container, used to pin the restatement and to mint tests/golden/.
* ``c/``          plain-C restatement (OpenMP) used as the CPU baseline in bench.py.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  ``radar_ml_b200`` never does.

Parity pinning: the reference ships NO tests / golden vectors for this path
(SURVEY.md §4, §8c).  The oracle is pinned against outputs of the reference itself,
run in the build container (tests/golden/make_golden.py -> tests/golden/*.npz), and
against scikit-learn 1.9.0 (the engine behind ``model.predict_proba``).
"""
