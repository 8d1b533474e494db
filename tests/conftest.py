import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def small_problem():
    """Seeded cubes + a calibrated SVC-RBF fitted on MAX-projection features (CPU, sklearn)."""
    import warnings
    from oracle import restate, synth
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cubes, y, ijk = synth.make_cubes(480, seed=2024)
        X = synth.features(*synth.project_max(cubes))
        cal = synth.build_svc(X[:300], y[:300], X[300:360], y[300:360])
    return {"cubes": cubes, "y": y, "ijk": ijk, "X": X, "cal": cal,
            "params": restate.export_params(cal), "test": slice(360, 480)}
