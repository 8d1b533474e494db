"""ctypes binding of libradarml.so (include/radarml.h).

There is no CPU fallback: if the shared library is missing or no sm_100 GPU is visible,
every entry point raises.  ``build()`` compiles the library in-tree with nvcc (cross-compiles
without a GPU), the way ``__graft_entry__.build`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libradarml.so")
SRC = os.path.join(_HERE, "csrc", "radarml.cu")
HOST_SRC = os.path.join(_HERE, "csrc", "host_narrow.cpp")     # host-only code (thread pool, AVX2)
HEADER = os.path.join(os.path.dirname(_HERE), "include", "radarml.h")

OK, E_INVALID, E_CUDA, E_UNSUPPORTED, E_NOMODEL, E_NONINTEGRAL, E_RANGE = 0, -1, -2, -3, -4, -5, -6
MODE_MAX, MODE_SLICE = 0, 1
MASK_XZ, MASK_YZ, MASK_XY, MASK_ALL = 1, 2, 4, 7
F32, U8, F32_EXACT = 0, 1, 2
RESERVE_SCORE, RESERVE_HOST, RESERVE_HOST_U8, RESERVE_SMALL = 1, 2, 4, 8


class RadarMLError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libradarml error %d: %s" % (code, msg))
        self.code = code


class NonIntegralInput(RadarMLError):
    """The u8 tensor-core path saw a value that is not an integer in [0,255]."""


class OutOfRangeInput(RadarMLError):
    """The multi-digit tensor-core path saw a feature outside [0, 256/feature_scale)."""


def nvcc_command(out=LIB_PATH):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return [nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
            "-Xcompiler", "-fPIC", "-shared", "-o", out, SRC, HOST_SRC, "-ldl", "-lpthread"]


def build(force=False, verbose=False):
    """Compile libradarml.so for sm_100a if it is missing or older than its sources."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))
            if f.endswith((".cu", ".cuh", ".h", ".cpp"))] + [HEADER]
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs):
            return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None

_i64, _i32, _u32, _f64, _f32, _vp, _sz = (C.c_int64, C.c_int32, C.c_uint32, C.c_double, C.c_float,
                                           C.c_void_p, C.c_size_t)
# name -> (restype, argtypes); one entry per symbol declared in include/radarml.h
SIGNATURES = {
    "rml_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "rml_destroy": (C.c_int, [_vp]),
    "rml_last_error": (C.c_char_p, [_vp]),
    "rml_version": (C.c_int, []),
    "rml_set_arena": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "rml_set_arena_bounds": (C.c_int, [_vp, _f64, _f64, _f64, _f64, _f64, _f64]),
    "rml_feature_len": (C.c_int, [_vp, _u32]),
    "rml_feature_stride": (C.c_int, [_vp, _u32, C.c_int]),
    "rml_set_affine": (C.c_int, [_vp, _f32, _f32, C.c_int]),
    "rml_load_affine": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "rml_set_precision": (C.c_int, [_vp, C.c_int]),
    "rml_reserve": (C.c_int, [_vp, _i64, C.c_int]),
    "rml_set_host_narrowing": (C.c_int, [_vp, C.c_int, C.c_int, _f64]),
    "rml_host_narrow_f32_to_u8": (C.c_int, [_vp, _vp, _i64, C.c_int]),
    "rml_last_host_transfer": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_f64),
                                         C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rml_score_host": (C.c_int, [_vp, _vp, _i64, _f64, _vp, _vp, _vp]),
    "rml_predict_targets_host": (C.c_int, [_vp, _vp, C.c_int, _vp, _u32, _f64, _vp, _vp, _vp]),
    "rml_comm_unique_id": (C.c_int, [_vp, _vp]),
    "rml_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "rml_comm_destroy": (C.c_int, [_vp]),
    "rml_allgather_labels": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "rml_load_svc_rbf": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _f64, _vp,
                                   _vp, _f64]),
    "rml_load_linear": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _f64]),
    "rml_model_is_integral": (C.c_int, [_vp]),
    "rml_project": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _u32, C.c_int, _vp, _vp, _vp]),
    "rml_project_derive": (C.c_int, [_vp, _vp, _i64, _u32, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "rml_process_samples": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _u32, C.c_int, _vp, _vp]),
    "rml_matrix_indices": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "rml_derive_targets": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _vp, _vp]),
    "rml_set_zoom": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "rml_zoom_feature_len": (C.c_int, [_vp, _u32]),
    "rml_process_samples_zoom": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _u32, C.c_int,
                                           _vp, _vp]),
    "rml_quantize_features": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _vp, _vp]),
    "rml_score": (C.c_int, [_vp, _vp, C.c_int, _vp, _i64, _f64, _vp, _vp, _vp, _vp, _vp]),
    "rml_predict_workspace_bytes": (_sz, [_vp, _i64]),
    "rml_predict": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _u32, _f64, _vp, _vp, _vp, _vp, _vp]),
    "rml_predict_host": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _u32, _f64, _vp, _vp, _vp]),
    "rml_project_u8": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _u32, C.c_int, _vp, _vp, _vp]),
    "rml_predict_u8": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _u32, _f64, _vp, _vp, _vp, _vp, _vp]),
    "rml_predict_host_u8": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _u32, _f64, _vp, _vp, _vp]),
    "rml_net_predict_u8": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _vp, _sz, _vp, _vp, _vp]),
    "rml_set_fused_u8": (C.c_int, [_vp, C.c_int]),
    "rml_net_begin": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _f32]),
    "rml_net_set_resize_tables": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp]),
    "rml_net_add_conv": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "rml_net_set_dense": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp]),
    "rml_net_finish": (C.c_int, [_vp]),
    "rml_net_uses_igemm": (C.c_int, [_vp]),
    "rml_net_workspace_bytes": (_sz, [_vp, _i64]),
    "rml_net_forward": (C.c_int, [_vp, _vp, _i64, _vp, _sz, _vp, _vp, _vp, _vp]),
    "rml_net_resize": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "rml_net_forward_images": (C.c_int, [_vp, _vp, _i64, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "rml_net_predict": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _vp, _sz, _vp, _vp, _vp]),
    "rml_check_status": (C.c_int, [_vp, _vp]),
    "rml_set_fused": (C.c_int, [_vp, C.c_int, C.c_int, _i64]),
    "rml_enable_timing": (C.c_int, [_vp, C.c_int]),
    "rml_last_timing": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "rml_launch_count": (_i64, [_vp]),
}


def load():
    """dlopen libradarml.so and set the prototypes.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RadarMLError(E_CUDA, "%s not built; run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx, rc):
    if rc == OK:
        return
    msg = load().rml_last_error(ctx)
    msg = msg.decode() if msg else ""
    if rc == E_NONINTEGRAL:
        raise NonIntegralInput(rc, msg)
    if rc == E_RANGE:
        raise OutOfRangeInput(rc, msg)
    raise RadarMLError(rc, msg)
