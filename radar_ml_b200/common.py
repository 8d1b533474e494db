"""Host-side mirror of the reference's ``common`` module for the hot path.

Same public names and argument meaning as /root/reference/common.py (cited per item) so
code written against the reference imports this module unchanged; the arithmetic runs in
libradarml's CUDA kernels.  The radar SDK (``WalabotAPI``, common.py:15) is NOT imported
here: nothing on the classification path needs it.
"""
from __future__ import annotations

import collections

import numpy as np

# common.py:25-27 — radar scan arena in spherical coordinates (cm / degrees)
R_MIN, R_MAX, R_RES = 10, 360, 2
THETA_MIN, THETA_MAX, THETA_RES = -42, 42, 4
PHI_MIN, PHI_MAX, PHI_RES = -30, 30, 2

# common.py:30-31
RADAR_MIN = 0.
RADAR_MAX = 255.

# common.py:40, 43 — field order is the sample tuple order (xz, yz, xy)
ProjMask = collections.namedtuple('ProjMask', ['xz', 'yz', 'xy'])
ProjZoom = collections.namedtuple('ProjZoom', ['xz', 'yz', 'xy'])



def cartesian_to_spherical(x, y, z):
    """common.py:93-97."""
    r = np.sqrt(np.power(x, 2) + np.power(y, 2) + np.power(z, 2))
    phi = np.arctan2(y, z)
    theta = np.arcsin(x / r)
    return (r, np.rad2deg(theta), np.rad2deg(phi))


def spherical_to_cartesian(r, theta, phi):
    """common.py:99-104."""
    theta_rad, phi_rad = np.deg2rad(theta), np.deg2rad(phi)
    x = r * np.sin(theta_rad)
    y = r * np.cos(theta_rad) * np.sin(phi_rad)
    z = r * np.cos(theta_rad) * np.cos(phi_rad)
    return (x, y, z)


class DerivedTarget(collections.namedtuple('DerivedTarget',
                                           ['xPosCm', 'yPosCm', 'zPosCm', 'amplitude', 'i', 'j', 'k'])):
    """common.py:45-80 — radar targets derived from the raw cube (replaces GetSensorTargets)."""

    @staticmethod
    def get_derived_targets(radar_data, size_x, size_y, size_z, num_targets=1):
        """Axis sums + top-k on the GPU (rml_derive_targets); coordinates on the host with the
        reference's formulas (common.py:66-69).  ``radar_data`` is one (size_x,size_y,size_z)
        cube (ndarray or nested list, like GetRawImage returns)."""
        import torch
        eng = get_engine()
        cube = np.ascontiguousarray(np.asarray(radar_data, dtype=np.float32).reshape(1, size_x, size_y, size_z))
        saved = eng.dims
        if tuple(saved) != (size_x, size_y, size_z):
            eng.set_arena(size_x, size_y, size_z)
        try:
            ijk = eng.derive_targets(torch.from_numpy(cube).to(eng.device), num_targets).cpu().numpy()[0]
        finally:
            if tuple(saved) != (size_x, size_y, size_z):
                eng.set_arena(*saved)

        def make(i, j, k):
            theta = THETA_MIN + i * (THETA_MAX - THETA_MIN) / (size_x - 1)
            phi = PHI_MIN + j * (PHI_MAX - PHI_MIN) / (size_y - 1)
            r = R_MIN + k * (R_MAX - R_MIN) / (size_z - 1)
            x, y, z = spherical_to_cartesian(r, theta, phi)
            return DerivedTarget(xPosCm=x, yPosCm=y, zPosCm=z, amplitude=None, i=i, j=j, k=k)
        return [make(int(i), int(j), int(k)) for i, j, k in ijk]


_engine = None
_zoom_cache = {}


def zoom_operator(n_in, factor):
    """scipy.ndimage.zoom (order 3, as common.py:143 calls it) along one axis as a dense
    [n_out, n_in] float64 matrix, obtained by zooming the unit vectors.  Host-side, cached."""
    key = (int(n_in), float(factor))
    hit = _zoom_cache.get(key)
    if hit is None:
        from scipy import ndimage
        eye = np.eye(n_in, dtype=np.float64)
        hit = np.stack([ndimage.zoom(eye[i], float(factor)) for i in range(n_in)], axis=1)
        _zoom_cache[key] = hit
    return hit


def get_engine(device: int = 0):
    """Process-wide default Engine (one libradarml context on ``device``)."""
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine(device)
    return _engine


def set_engine(engine):
    global _engine
    _engine = engine


def arena_size():
    """(size_x, size_y, size_z) of the training arena, predict.py:74-76 -> (22, 31, 176)."""
    size_z = int((R_MAX - R_MIN) / R_RES) + 1
    size_y = int((PHI_MAX - PHI_MIN) / PHI_RES) + 1
    size_x = int((THETA_MAX - THETA_MIN) / THETA_RES) + 1
    return size_x, size_y, size_z


def calculate_matrix_indices(x, y, z, size_x, size_y, size_z):
    """common.py:106-121: target (x,y,z) in cm -> (i,j,k) voxel indices, on the GPU.

    Scalars in, tuple of python ints out (like the reference); arrays of targets give an
    (n,3) int32 array.
    """
    import torch
    eng = get_engine()
    xyz = np.stack([np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in (x, y, z)], axis=1)
    saved = eng.dims
    if tuple(saved) != (size_x, size_y, size_z):
        eng.set_arena(size_x, size_y, size_z)
    try:
        ijk = eng.matrix_indices(torch.from_numpy(xyz).to(eng.device)).cpu().numpy()
    finally:
        if tuple(saved) != (size_x, size_y, size_z):
            eng.set_arena(*saved)
    if np.ndim(x) == 0:
        return int(ijk[0, 0]), int(ijk[0, 1]), int(ijk[0, 2])
    return ijk


def _unit_zoom(proj_zoom, proj_mask):
    return all(float(z[0]) == 1.0 and float(z[1]) == 1.0
               for z, m in zip(proj_zoom, proj_mask) if m)


def process_samples(samples, proj_mask=ProjMask(xz=True, yz=True, xy=True),
                    proj_zoom=ProjZoom(xz=[1.0, 1.0], yz=[1.0, 1.0], xy=[1.0, 1.0]), scale=False):
    """common.py:123-149 — projections -> (n, F) float32 feature matrix, computed on the GPU.

    Args are the reference's: ``samples`` a list of (xz, yz, xy) arrays, ``proj_mask`` which
    projections to keep, ``proj_zoom`` per-projection zoom factors, ``scale`` divide by 255.
    Zoom factors other than 1.0 (arena mismatch, README.md:207) run as separable spline
    operators on the GPU (``zoom_operator``); results match scipy to float32 rounding.
    """
    import torch
    eng = get_engine()
    n = len(samples)
    stacked = []
    for idx in range(3):
        if proj_mask[idx]:
            arr = np.ascontiguousarray(np.stack([np.asarray(t[idx], dtype=np.float32)
                                                 for t in samples]))
            stacked.append(torch.from_numpy(arr).to(eng.device))
        else:
            stacked.append(None)
    if not _unit_zoom(proj_zoom, proj_mask):
        # arena mismatch (README.md:207): ndimage.zoom as separable operators on the GPU
        for idx in range(3):
            if stacked[idx] is not None:
                h, w = stacked[idx].shape[1:]
                eng.set_zoom(idx, zoom_operator(h, proj_zoom[idx][0]), zoom_operator(w, proj_zoom[idx][1]))
        out = eng.process_samples_zoom(stacked[0], stacked[1], stacked[2], mask=proj_mask, scale=scale)
        return out.cpu().numpy()
    # infer the arena from the projections that are present
    sx = sy = sz = None
    if stacked[0] is not None:
        sx, sz = stacked[0].shape[1:]
    if stacked[1] is not None:
        sy, sz = stacked[1].shape[1:]
    if stacked[2] is not None:
        sx, sy = stacked[2].shape[1:]
    cur = eng.dims
    dims = (sx or cur[0], sy or cur[1], sz or cur[2])
    if tuple(cur) != dims:
        eng.set_arena(*dims)
    try:
        out = eng.process_samples(stacked[0], stacked[1], stacked[2], mask=proj_mask, scale=scale)
        res = out.cpu().numpy()
    finally:
        if tuple(cur) != dims:
            eng.set_arena(*cur)
    assert res.shape[0] == n
    return res
