"""Engine — one libradarml context per GPU; torch tensors are only device containers.

Batched extensions of the reference seam: ``project`` (predict.py:102-107 +
common.py:141-149), ``score`` (predict.py:56-70) and ``predict`` (one predict.py:93-119
iteration for B scans).  Every call goes through the C ABI in include/radarml.h.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (F32, F32_EXACT, MASK_ALL, MODE_MAX, MODE_SLICE, RESERVE_HOST, RESERVE_HOST_U8,
                   RESERVE_SCORE, RESERVE_SMALL, U8, NonIntegralInput, OutOfRangeInput,
                   RadarMLError, check)

SX, SY, SZ = 22, 31, 176  # common.py:25-27 -> predict.py:74-76


def mask_bits(proj_mask) -> int:
    """ProjMask(xz, yz, xy) (common.py:40) -> bit0 xz, bit1 yz, bit2 xy."""
    if isinstance(proj_mask, int):
        return proj_mask
    xz, yz, xy = proj_mask
    return (1 if xz else 0) | (2 if yz else 0) | (4 if xy else 0)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise RadarMLError(_lib.E_CUDA, "no CUDA device: radar_ml_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.ctx = C.c_void_p()
        rc = self.lib.rml_create(device, C.byref(self.ctx))
        if rc != 0:
            msg = self.lib.rml_last_error(None)
            raise RadarMLError(rc, msg.decode() if msg else "rml_create failed")
        self.params = None
        self.dims = (SX, SY, SZ)
        self._work = None
        self._reserved = {}       # flag -> rows reserved for the currently loaded model
        self.affine_tables = None

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.lib.rml_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ configuration
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_arena(self, sx, sy, sz):
        check(self.ctx, self.lib.rml_set_arena(self.ctx, sx, sy, sz))
        self.dims = (sx, sy, sz)

    def set_arena_bounds(self, r_min, r_max, theta_min, theta_max, phi_min, phi_max):
        check(self.ctx, self.lib.rml_set_arena_bounds(self.ctx, r_min, r_max, theta_min, theta_max,
                                                      phi_min, phi_max))

    def set_affine(self, offset=0.0, scale=255.0, enabled=True):
        check(self.ctx, self.lib.rml_set_affine(self.ctx, offset, scale, 1 if enabled else 0))

    def load_affine(self, offset=None, scale=None):
        """Per-feature (x - offset[f]) / scale[f] (a fitted StandardScaler's mean_ / scale_);
        None, None restores the reference's /255 (common.py:148).  Load it BEFORE the model."""
        if offset is None and scale is None:
            check(self.ctx, self.lib.rml_load_affine(self.ctx, None, None, 0))
            self.affine_tables = None
        else:
            off = np.ascontiguousarray(offset, dtype=np.float32)
            scl = np.ascontiguousarray(scale, dtype=np.float32)
            if off.shape != scl.shape or off.ndim != 1:
                raise ValueError("offset and scale must be 1-D arrays of the same length")
            check(self.ctx, self.lib.rml_load_affine(self.ctx, _np_ptr(off), _np_ptr(scl), off.shape[0]))
            self.affine_tables = (off, scl)
        self._reserved = {}

    def set_precision(self, force_f32: bool):
        """True: rml_predict* emit float32 features (real-valued cubes); False: automatic."""
        check(self.ctx, self.lib.rml_set_precision(self.ctx, 1 if force_f32 else 0))
        self._reserved = {}
        self._work = None

    def reserve(self, rows: int, flag: int):
        """Library-owned staging is created here, never inside a hot entry point."""
        if self._reserved.get(flag, -1) < rows:
            check(self.ctx, self.lib.rml_reserve(self.ctx, int(rows), flag))
            self._reserved[flag] = rows

    def set_host_narrowing(self, enabled=True, threads=0, min_gbs=0.0):
        """``predict_host`` on float32 cubes of the sensor's integers: convert each chunk to bytes on
        the host (thread pool, exactness checked) so a quarter of the bytes crosses PCIe.  On by
        default; it switches itself off when the host converts slower than ``min_gbs`` (57 GB/s:
        a host that cannot beat the bus) or the cubes turn out not to be integral."""
        check(self.ctx, self.lib.rml_set_host_narrowing(self.ctx, int(bool(enabled)), int(threads), float(min_gbs)))
        self._reserved.pop(RESERVE_HOST, None)
        self._reserved.pop(RESERVE_HOST_U8, None)

    def last_host_transfer(self):
        """What the last ``predict_host`` moved: bus bytes, scans narrowed to bytes, conversion rate."""
        b, n, g = C.c_int64(), C.c_int64(), C.c_double()
        t, a = C.c_int(), C.c_int()
        check(self.ctx, self.lib.rml_last_host_transfer(self.ctx, C.byref(b), C.byref(n), C.byref(g), C.byref(t), C.byref(a)))
        return {"h2d_bytes": b.value, "narrowed_scans": n.value, "convert_gbs": g.value, "threads": t.value,
                "active": bool(a.value)}

    def feature_len(self, mask=MASK_ALL):
        return self.lib.rml_feature_len(self.ctx, mask_bits(mask))

    def feature_stride(self, mask=MASK_ALL, dtype=F32):
        return self.lib.rml_feature_stride(self.ctx, mask_bits(mask), dtype)

    def load_model(self, p):
        """p: radar_ml_b200.model.ModelParams."""
        if p.kind == "svc_rbf":
            ns = np.ascontiguousarray(p.n_support, dtype=np.int32)
            sv = np.ascontiguousarray(p.sv, dtype=np.float64)
            dc = np.ascontiguousarray(p.dual_coef, dtype=np.float64)
            rho = np.ascontiguousarray(p.rho, dtype=np.float64)
            a = np.ascontiguousarray(p.platt_a, dtype=np.float64)
            b = np.ascontiguousarray(p.platt_b, dtype=np.float64)
            check(self.ctx, self.lib.rml_load_svc_rbf(
                self.ctx, p.n_classes, p.n_features, sv.shape[0], _np_ptr(ns), _np_ptr(sv),
                _np_ptr(dc), _np_ptr(rho), float(p.gamma), _np_ptr(a), _np_ptr(b),
                float(p.feature_scale)))
        elif p.kind == "linear":
            coef = np.ascontiguousarray(p.coef, dtype=np.float64)
            ic = np.ascontiguousarray(p.intercept, dtype=np.float64)
            a = np.ascontiguousarray(p.platt_a, dtype=np.float64)
            b = np.ascontiguousarray(p.platt_b, dtype=np.float64)
            check(self.ctx, self.lib.rml_load_linear(self.ctx, p.n_classes, p.n_features,
                                                     _np_ptr(coef), _np_ptr(ic), _np_ptr(a),
                                                     _np_ptr(b), float(p.feature_scale)))
        else:
            raise ValueError(p.kind)
        self.params = p
        self._reserved = {}
        self._work = None

    @property
    def model_is_integral(self) -> bool:
        return bool(self.lib.rml_model_is_integral(self.ctx))

    @property
    def launch_count(self) -> int:
        return int(self.lib.rml_launch_count(self.ctx))

    def check_status(self):
        """Synchronise and raise if the u8 path met non-integral data / bad slice indices."""
        check(self.ctx, self.lib.rml_check_status(self.ctx, self._stream()))

    # ------------------------------------------------------------------ K1
    def _check_cubes(self, cubes):
        sx, sy, sz = self.dims
        if not (isinstance(cubes, torch.Tensor) and cubes.is_cuda
                and cubes.dtype in (torch.float32, torch.uint8)
                and cubes.is_contiguous() and cubes.dim() == 4
                and tuple(cubes.shape[1:]) == (sx, sy, sz)):
            raise ValueError("cubes must be a contiguous CUDA float32 (or uint8) tensor [B,%d,%d,%d]"
                             % (sx, sy, sz))
        if cubes.device != self.device:
            raise ValueError("cubes are on %s, engine on %s" % (cubes.device, self.device))
        return cubes.dtype == torch.uint8

    def project(self, cubes, mode="max", ijk=None, mask=MASK_ALL, dtype=F32, out=None, norms=None):
        """cubes [B,sx,sy,sz] (float32, or the sensor's integers as uint8) -> features
        [B,stride] (+ int32 norms for U8)."""
        cube_u8 = self._check_cubes(cubes)
        m = mask_bits(mask)
        B = cubes.shape[0]
        stride = self.feature_stride(m, dtype)
        if out is None:
            out = torch.empty((B, stride), device=self.device,
                              dtype=torch.uint8 if dtype == U8 else torch.float32)
        if dtype == U8 and norms is None:
            norms = torch.empty((B,), device=self.device, dtype=torch.int32)
        md = MODE_MAX if mode in ("max", MODE_MAX) else MODE_SLICE
        if md == MODE_SLICE:
            if ijk is None:
                raise ValueError("slice mode needs ijk [B,3] int32")
            ijk = ijk.to(device=self.device, dtype=torch.int32).contiguous()
        if B == 0:   # empty batch: nothing to launch (zero-size tensors have no storage)
            return (out, norms) if dtype == U8 else out
        fn = self.lib.rml_project_u8 if cube_u8 else self.lib.rml_project
        check(self.ctx, fn(self.ctx, _ptr(cubes), B, md, _ptr(ijk), m, dtype, _ptr(out), _ptr(norms),
                           self._stream()))
        return (out, norms) if dtype == U8 else out

    def process_samples(self, xz, yz, xy, mask=MASK_ALL, scale=False):
        """common.process_samples on device projections (zoom 1.0)."""
        m = mask_bits(mask)
        ref = next(t for t in (xz, yz, xy) if t is not None)
        B = ref.shape[0]
        F = self.feature_len(m)
        out = torch.empty((B, F), device=self.device, dtype=torch.float32)
        check(self.ctx, self.lib.rml_process_samples(self.ctx, _ptr(xz), _ptr(yz), _ptr(xy), B, m,
                                                     1 if scale else 0, _ptr(out), self._stream()))
        return out

    def derive_targets(self, cubes, num_targets=1, want_sums=False):
        """common.py:45-80 on device: cubes [B,sx,sy,sz] -> ijk int32 [B,T,3] (ascending by sum;
        last = strongest) and optionally the float32 axis sums [B, sx+sy+sz]."""
        if self._check_cubes(cubes):
            raise ValueError("derive_targets takes float32 cubes")
        B = cubes.shape[0]
        ijk = torch.empty((B, num_targets, 3), device=self.device, dtype=torch.int32)
        sums = torch.empty((B, sum(self.dims)), device=self.device, dtype=torch.float32) if want_sums else None
        if B:
            check(self.ctx, self.lib.rml_derive_targets(self.ctx, _ptr(cubes), B, num_targets, _ptr(ijk),
                                                        _ptr(sums), self._stream()))
        return (ijk, sums) if want_sums else ijk

    def project_derive(self, cubes, num_targets=1, mask=MASK_ALL, dtype=F32, want_sums=False):
        """``project(mode='max')`` and ``derive_targets`` in ONE pass over the cubes (default arena;
        other arenas run the two kernels back to back).  Returns (features[, norms], ijk[, sums])."""
        if self._check_cubes(cubes):
            raise ValueError("project_derive takes float32 cubes")
        m = mask_bits(mask)
        B = cubes.shape[0]
        out = torch.empty((B, self.feature_stride(m, dtype)), device=self.device,
                          dtype=torch.uint8 if dtype == U8 else torch.float32)
        norms = torch.empty((B,), device=self.device, dtype=torch.int32) if dtype == U8 else None
        ijk = torch.empty((B, num_targets, 3), device=self.device, dtype=torch.int32)
        sums = torch.empty((B, sum(self.dims)), device=self.device, dtype=torch.float32) if want_sums else None
        if B:
            check(self.ctx, self.lib.rml_project_derive(self.ctx, _ptr(cubes), B, m, dtype, _ptr(out), _ptr(norms),
                                                        num_targets, _ptr(ijk), _ptr(sums), self._stream()))
        res = (out, norms) if dtype == U8 else (out,)
        return res + ((ijk, sums) if want_sums else (ijk,))

    def set_zoom(self, proj, a_rows, a_cols):
        """Separable ndimage.zoom operator of projection ``proj`` (0 xz, 1 yz, 2 xy):
        a_rows [out_h, in_h], a_cols [out_w, in_w] float64 host arrays."""
        ar = np.ascontiguousarray(a_rows, dtype=np.float64)
        ac = np.ascontiguousarray(a_cols, dtype=np.float64)
        check(self.ctx, self.lib.rml_set_zoom(self.ctx, proj, ar.shape[1], ac.shape[1], ar.shape[0],
                                              ac.shape[0], _np_ptr(ar), _np_ptr(ac)))

    def process_samples_zoom(self, xz, yz, xy, mask=MASK_ALL, scale=False, strides=None):
        """common.process_samples with proj_zoom != 1 (operators set by ``set_zoom``)."""
        m = mask_bits(mask)
        ref = next(t for t in (xz, yz, xy) if t is not None)
        B = ref.shape[0]
        F = self.lib.rml_zoom_feature_len(self.ctx, m)
        out = torch.empty((B, F), device=self.device, dtype=torch.float32)
        if strides is None:
            strides = [0 if t is None else int(t[0].numel()) for t in (xz, yz, xy)]
        if B:
            check(self.ctx, self.lib.rml_process_samples_zoom(
                self.ctx, _ptr(xz), strides[0], _ptr(yz), strides[1], _ptr(xy), strides[2], B, m,
                1 if scale else 0, _ptr(out), self._stream()))
        return out

    def matrix_indices(self, xyz):
        """xyz [B,3] float64 CUDA -> ijk [B,3] int32 (common.calculate_matrix_indices)."""
        xyz = xyz.to(device=self.device, dtype=torch.float64).contiguous()
        out = torch.empty((xyz.shape[0], 3), device=self.device, dtype=torch.int32)
        check(self.ctx, self.lib.rml_matrix_indices(self.ctx, _ptr(xyz), xyz.shape[0], _ptr(out),
                                                    self._stream()))
        return out

    # ------------------------------------------------------------------ K2
    def score(self, feats, norms=None, min_proba=0.7, want_decision=False, exact=False):
        """features -> (proba [B,C] f32, label [B] i32, known [B] bool[, decision]).

        uint8 rows (+ norms) take the integer tensor-core scorer; float32 rows take the exact
        multi-digit tensor-core scorer (values in [0, 256/scale), checked by ``check_status``) or,
        with ``exact=True`` / when the model does not qualify, the float64 CUDA-core scorer."""
        if self.params is None:
            raise RadarMLError(_lib.E_NOMODEL, "no model loaded")
        B = feats.shape[0]
        Cn = self.params.n_classes
        dtype = U8 if feats.dtype == torch.uint8 else (F32_EXACT if exact else F32)
        proba = torch.empty((B, Cn), device=self.device, dtype=torch.float32)
        label = torch.empty((B,), device=self.device, dtype=torch.int32)
        known = torch.empty((B,), device=self.device, dtype=torch.uint8)
        dec = None
        if want_decision:
            dec = torch.empty((B,) if Cn == 2 else (B, Cn), device=self.device, dtype=torch.float32)
        if B == 0:
            return (proba, label, known.bool()) + ((dec,) if want_decision else ())
        if dtype == F32:
            self.reserve(B, RESERVE_SCORE)
        check(self.ctx, self.lib.rml_score(self.ctx, _ptr(feats), dtype, _ptr(norms), B,
                                           float(min_proba), _ptr(proba), _ptr(dec), _ptr(label),
                                           _ptr(known), self._stream()))
        res = (proba, label, known.bool())
        return res + (dec,) if want_decision else res

    def quantize(self, feats_f32):
        """(n,F) float32 scaled features -> (u8 rows, norms); see rml_quantize_features."""
        B, F = feats_f32.shape
        stride = (F + 127) // 128 * 128
        out = torch.empty((B, stride), device=self.device, dtype=torch.uint8)
        norms = torch.empty((B,), device=self.device, dtype=torch.int32)
        check(self.ctx, self.lib.rml_quantize_features(self.ctx, _ptr(feats_f32), B, F, _ptr(out),
                                                       _ptr(norms), self._stream()))
        return out, norms

    def score_features_host(self, X: np.ndarray, min_proba=0.7):
        """model.predict_proba(X) (+ argmax, >= min_proba) for host features as
        common.process_samples(scale=True) makes them: one C call (rml_score_host).  Integral
        rows of an integral model run on the integer tensor-core scorer, anything else on the
        exact multi-digit one or the float64 one — decided on the device, never silently."""
        if self.params is None:
            raise RadarMLError(_lib.E_NOMODEL, "no model loaded")
        X = np.ascontiguousarray(X, dtype=np.float32)
        B = X.shape[0]
        Cn = self.params.n_classes
        proba = np.empty((B, Cn), dtype=np.float32)
        label = np.empty((B,), dtype=np.int32)
        known = np.empty((B,), dtype=np.uint8)
        if B:
            self.reserve(min(B, 1024), RESERVE_SMALL)
            check(self.ctx, self.lib.rml_score_host(self.ctx, _np_ptr(X), B, float(min_proba),
                                                    _np_ptr(proba), _np_ptr(label), _np_ptr(known)))
        return proba, label, known

    def predict_targets_host(self, cube: np.ndarray, ijk, mask=MASK_ALL, min_proba=0.7):
        """One predict.py:93-119 iteration: ONE raw cube [sx,sy,sz] float32 and the (i,j,k) of its
        T targets -> (proba [T,C], label [T], known [T]).  The cube is uploaded once."""
        if self.params is None:
            raise RadarMLError(_lib.E_NOMODEL, "no model loaded")
        cube = np.ascontiguousarray(cube, dtype=np.float32)
        if cube.shape != tuple(self.dims):
            raise ValueError("cube must have shape %s" % (tuple(self.dims),))
        ij = np.ascontiguousarray(ijk, dtype=np.int32).reshape(-1, 3)
        T = ij.shape[0]
        Cn = self.params.n_classes
        proba = np.empty((T, Cn), dtype=np.float32)
        label = np.empty((T,), dtype=np.int32)
        known = np.empty((T,), dtype=np.uint8)
        if T:
            self.reserve(max(T, 8), RESERVE_SMALL)
            check(self.ctx, self.lib.rml_predict_targets_host(
                self.ctx, _np_ptr(cube), T, _np_ptr(ij), mask_bits(mask), float(min_proba),
                _np_ptr(proba), _np_ptr(label), _np_ptr(known)))
        return proba, label, known

    # ------------------------------------------------------------------ K1 -> K2
    def workspace(self, B):
        need = int(self.lib.rml_predict_workspace_bytes(self.ctx, B))
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty((need,), device=self.device, dtype=torch.uint8)
        return self._work

    def predict(self, cubes, mode="max", ijk=None, mask=MASK_ALL, min_proba=0.7, out=None):
        """B scans (float32 or uint8 cubes) -> (proba [B,C] f32, label [B] i32, known [B] u8) on
        the device, async."""
        cube_u8 = self._check_cubes(cubes)
        if self.params is None:
            raise RadarMLError(_lib.E_NOMODEL, "no model loaded")
        B = cubes.shape[0]
        Cn = self.params.n_classes
        if out is None:
            proba = torch.empty((B, Cn), device=self.device, dtype=torch.float32)
            label = torch.empty((B,), device=self.device, dtype=torch.int32)
            known = torch.empty((B,), device=self.device, dtype=torch.uint8)
        else:
            proba, label, known = out
        md = MODE_MAX if mode in ("max", MODE_MAX) else MODE_SLICE
        if md == MODE_SLICE:
            if ijk is None:
                raise ValueError("slice mode needs ijk [B,3] int32")
            ijk = ijk.to(device=self.device, dtype=torch.int32).contiguous()
        if B == 0:
            return proba, label, known
        work = self.workspace(B)
        fn = self.lib.rml_predict_u8 if cube_u8 else self.lib.rml_predict
        check(self.ctx, fn(self.ctx, _ptr(cubes), B, md, _ptr(ijk), mask_bits(mask), float(min_proba),
                           _ptr(work), _ptr(proba), _ptr(label), _ptr(known), self._stream()))
        return proba, label, known

    def predict_host(self, cubes: np.ndarray, mode="max", ijk=None, mask=MASK_ALL, min_proba=0.7,
                     out=None):
        """Host cubes [B,sx,sy,sz] (float32 as predict.py:91 makes them, or the sensor's integers
        as uint8) in, host results out (H2D/compute/D2H overlapped)."""
        sx, sy, sz = self.dims
        if not (isinstance(cubes, np.ndarray) and cubes.dtype in (np.float32, np.uint8)
                and cubes.flags.c_contiguous and cubes.shape[1:] == (sx, sy, sz)):
            raise ValueError("cubes must be a C-contiguous float32 (or uint8) ndarray [B,%d,%d,%d]"
                             % (sx, sy, sz))
        if self.params is None:
            raise RadarMLError(_lib.E_NOMODEL, "no model loaded")
        B = cubes.shape[0]
        Cn = self.params.n_classes
        if out is None:
            proba = np.empty((B, Cn), dtype=np.float32)
            label = np.empty((B,), dtype=np.int32)
            known = np.empty((B,), dtype=np.uint8)
        else:
            proba, label, known = out
        md = MODE_MAX if mode in ("max", MODE_MAX) else MODE_SLICE
        ij = None
        if md == MODE_SLICE:
            if ijk is None:
                raise ValueError("slice mode needs ijk [B,3] int32")
            ij = np.ascontiguousarray(ijk, dtype=np.int32)
        u8 = cubes.dtype == np.uint8
        fn = self.lib.rml_predict_host_u8 if u8 else self.lib.rml_predict_host
        args = (self.ctx, _np_ptr(cubes), B, md, _np_ptr(ij), mask_bits(mask), float(min_proba),
                _np_ptr(proba), _np_ptr(label), _np_ptr(known))
        self.reserve(0, RESERVE_HOST_U8 if u8 else RESERVE_HOST)
        check(self.ctx, fn(*args))
        try:
            self.check_status()
        except NonIntegralInput:
            # real-valued float32 cubes (the reference accepts any float32 cube): same call with
            # float32 features and the exact general-precision scorer
            if u8:
                raise
            self.set_precision(True)
            try:
                self.reserve(0, RESERVE_HOST)
                check(self.ctx, fn(*args))
                try:
                    self.check_status()
                except OutOfRangeInput:
                    raise
            finally:
                self.set_precision(False)
        return proba, label, known

    # ------------------------------------------------------------------ label exchange (C ABI)
    def comm_init(self, rank: int, world: int, unique_id: bytes | None = None):
        """NCCL communicator of this context.  rank 0 calls ``comm_unique_id()`` and hands the
        128 bytes to the others through any host channel (e.g. torch.distributed broadcast)."""
        buf = C.create_string_buffer(unique_id, 128)
        check(self.ctx, self.lib.rml_comm_init(self.ctx, rank, world, buf))

    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        check(self.ctx, self.lib.rml_comm_unique_id(self.ctx, buf))
        return buf.raw

    def allgather_labels(self, send, recv):
        """ONE ncclAllGather of int32 labels on the current stream; ``send`` may be the rank's
        slice of ``recv`` (in place)."""
        check(self.ctx, self.lib.rml_allgather_labels(self.ctx, _ptr(send), _ptr(recv), send.numel(),
                                                      self._stream()))
