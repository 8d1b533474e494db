"""radar_ml_b200 — B200-native per-scan classification hot path of goruck/radar-ml.

Python host (mirrors the reference's ``common`` / ``predict`` seam) over libradarml.so:
hand-written sm_100a CUDA behind the C ABI in include/radarml.h.  No CPU fallback.
"""
from . import _lib  # noqa: F401
from .common import ProjMask, ProjZoom, RADAR_MAX  # noqa: F401

__all__ = ["common", "predict", "model", "engine", "ProjMask", "ProjZoom", "RADAR_MAX"]
__version__ = "0.1.0"
