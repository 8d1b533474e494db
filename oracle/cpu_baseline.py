"""TEST INFRASTRUCTURE — the reference's per-scan CPU path, timed on the host cores.

``scan_loop`` is what predict.py:90-119 does for one target, restated (the reference module
itself cannot travel to the GPU box): numpy projection -> process_samples(scale=True) ->
classifier() -> ``model.predict_proba`` on the sklearn CalibratedClassifierCV — i.e. the
real third-party engine (libsvm via scikit-learn) the reference calls at predict.py:60.
``run_all_cores`` runs that loop in one process per host core over contiguous chunks
(libsvm's predict is single-threaded), BLAS pinned to one thread per worker.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
import warnings

import numpy as np

from . import restate

_G = {}


def scan_loop(cubes, cal, classes, mode="max", ijk=None, min_proba=0.7):
    """Reference-exact loop, one scan at a time.  Returns (names, probas)."""
    names, probas = [], []
    mask = restate.ProjMask(True, True, True)
    zoom = restate.calc_proj_zoom(22, 31, 176, 22, 31, 176)
    for s in range(cubes.shape[0]):
        t = restate.project(cubes[s], mode, None if ijk is None else tuple(int(v) for v in ijk[s]))
        obs = restate.process_samples([t], proj_mask=mask, proj_zoom=zoom, scale=True, always_zoom=True)
        preds = cal.predict_proba(obs.reshape(1, -1))[0]     # predict.py:60
        j = int(np.argmax(preds))
        names.append(classes[j] if preds[j] >= min_proba else "Unknown")
        probas.append(preds[j])
    return names, np.asarray(probas)


def _worker(args):
    lo, hi = args
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=1)
    except Exception:  # pragma: no cover
        ctx = None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        scan_loop(_G["cubes"][lo:hi], _G["cal"], _G["classes"], _G["mode"])
        dt = time.perf_counter() - t0
    if ctx is not None:
        ctx.restore_original_limits()
    return hi - lo, dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def run_all_cores(cubes, cal, classes, mode="max", cores=None, repeats=1):
    """scans/s of the per-scan reference loop over ``cores`` worker processes (fork)."""
    cores = cores or host_cores()
    n = cubes.shape[0]
    cores = max(1, min(cores, n))
    _G.update(cubes=cubes, cal=cal, classes=classes, mode=mode)
    bounds = np.linspace(0, n, cores + 1).astype(int)
    chunks = [(int(bounds[i]), int(bounds[i + 1])) for i in range(cores) if bounds[i + 1] > bounds[i]]
    best = None
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks)) as pool:
        pool.map(_worker, [(0, 1)] * len(chunks))  # warm the workers (imports, first-call cost)
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_worker, chunks)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return n / best, best, len(chunks)


def run_one_core(cubes, cal, classes, mode="max", repeats=3):
    """The reference-exact loop on ONE core (what predict.py does); best of ``repeats`` passes.
    Returns (scans/s, seconds of the best pass, median per-scan latency in ms)."""
    best, lat = None, []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        scan_loop(cubes[:1], cal, classes, mode)
        for _ in range(repeats):
            t0 = time.perf_counter()
            for s in range(cubes.shape[0]):
                t1 = time.perf_counter()
                scan_loop(cubes[s:s + 1], cal, classes, mode)
                lat.append(time.perf_counter() - t1)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return cubes.shape[0] / best, best, 1e3 * float(np.median(lat))


def cpu_model_string():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:  # pragma: no cover
        pass
    return "unknown"
