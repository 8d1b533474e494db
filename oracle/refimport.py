"""TEST INFRASTRUCTURE — import the UNMODIFIED reference (common.py, predict.py).

Works only where /root/reference exists (the build container).  common.py:15 and
predict.py:7 import the Walabot hardware SDK at module top; a stub module with the one
attribute read at import time (common.py:34 ``radar.PROF_SENSOR``) is injected instead.
Nothing under -m gpu tests, smoke() or bench.py may call this (the GPU box has no
/root/reference).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_DIR = os.environ.get("RADARML_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "common.py"))


def load():
    """Return (common, predict) reference modules."""
    if not available():
        raise RuntimeError("reference not present at %s" % REF_DIR)
    if "WalabotAPI" not in sys.modules:
        stub = types.ModuleType("WalabotAPI")
        stub.PROF_SENSOR = 0x10
        sys.modules["WalabotAPI"] = stub
    if REF_DIR not in sys.path:
        sys.path.append(REF_DIR)
    saved = {k: sys.modules.pop(k) for k in ("common", "predict") if k in sys.modules
             and not getattr(sys.modules[k], "__file__", "").startswith(REF_DIR)}
    try:
        ref_common = importlib.import_module("common")
        ref_predict = importlib.import_module("predict")
    finally:
        sys.modules.update(saved)
    assert ref_common.__file__.startswith(REF_DIR), ref_common.__file__
    return ref_common, ref_predict
