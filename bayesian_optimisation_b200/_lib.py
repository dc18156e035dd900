"""ctypes binding of libbogp.so (C ABI: include/bogp.h).

There is no CPU fallback: if the shared library is missing this module raises at import
of the symbols, and `bogp_create` fails on a machine without a B200-class GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libbogp.so")

BOGP_OK = 0
BOGP_ERR_BAD_ARG = -1
BOGP_ERR_CUDA = -2
BOGP_ERR_NOT_POSDEF = -3
BOGP_ERR_NAN_SCORE = -4
BOGP_ERR_WORKSPACE = -5
BOGP_MAX_DIM = 16
ACQ_LCB = 0
ACQ_EI = 1
PATH_FP64_DMMA = 0
PATH_INT8_TCGEN05 = 1


class BogpError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libbogp error {code}: {text}")
        self.code = code


class Candidates(C.Structure):
    """struct bogp_candidates (include/bogp.h)."""
    _fields_ = [("d_points", C.c_void_p), ("d_axes", C.c_void_p), ("h_axis_len", C.POINTER(C.c_int32)),
                ("c_total", C.c_int64), ("cross_jitter", C.c_double)]


class HostCandidates(C.Structure):
    """struct bogp_host_candidates (include/bogp.h)."""
    _fields_ = [("h_points", C.c_void_p), ("h_axes", C.c_void_p), ("h_axis_len", C.POINTER(C.c_int32)),
                ("c_total", C.c_int64), ("cross_jitter", C.c_double)]


class Result(C.Structure):
    """struct bogp_result (include/bogp.h): the 24-byte device record of a sweep's winner."""
    _fields_ = [("score", C.c_double), ("index", C.c_int64), ("nan_flag", C.c_int32), ("reserved", C.c_int32)]


_vp, _i64, _i32, _dbl, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_size_t
_pd, _pi64 = C.POINTER(C.c_double), C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol include/bogp.h declares
SIGNATURES = {
    "bogp_version": (C.c_char_p, []),
    "bogp_last_error": (C.c_char_p, []),
    "bogp_device_count": (_i32, [C.POINTER(_i32)]),
    "bogp_create": (_i32, [_i32, C.POINTER(_vp)]),
    "bogp_destroy": (None, [_vp]),
    "bogp_set_stream": (_i32, [_vp, _vp]),
    "bogp_sm_count": (_i32, [_vp]),
    "bogp_launch_count": (_i64, [_vp]),
    "bogp_set_acquire_path": (_i32, [_vp, _i32]),
    "bogp_get_acquire_path": (_i32, [_vp]),
    "bogp_set_screening": (_i32, [_vp, _i32]),
    "bogp_get_screening": (_i32, [_vp]),
    "bogp_set_global_seed": (_i32, [_vp, _i32]),
    "bogp_set_fused": (_i32, [_vp, _i32, _i32]),
    "bogp_get_fused": (_i32, [_vp]),
    "bogp_screen_stats": (_i32, [_vp, _pi64, _pi64, _i32]),
    "bogp_profile": (_i32, [_vp, _i32]),
    "bogp_profile_read": (_i32, [_vp, _i32, C.POINTER(_dbl), C.POINTER(_i64)]),
    "bogp_measure_peak": (_i32, [_vp, _i32, _dbl, C.POINTER(_dbl), C.POINTER(_dbl)]),
    "bogp_kernel_matrix": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _dbl, _vp, _i64]),
    "bogp_fit_workspace_bytes": (_sz, [_i64, _i32]),
    "bogp_fit_create": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _dbl, _vp, _sz, C.POINTER(_vp), C.POINTER(_dbl)]),
    "bogp_fit_enqueue": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _dbl, _vp, _sz, C.POINTER(_vp)]),
    "bogp_fit_status": (_i32, [_vp, C.POINTER(_dbl)]),
    "bogp_fit_destroy": (None, [_vp]),
    "bogp_fit_n_pad": (_i64, [_vp]),
    "bogp_fit_chol": (_vp, [_vp]),
    "bogp_fit_linv": (_vp, [_vp]),
    "bogp_fit_alpha": (_vp, [_vp]),
    "bogp_fit_logdet": (_dbl, [_vp]),
    "bogp_cholesky": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "bogp_acquire_workspace_bytes": (_sz, [_vp, _i64]),
    "bogp_acquire": (_i32, [_vp, _vp, C.POINTER(Candidates), _i64, _i64, _i32, _dbl, _dbl, _dbl, _vp, _vp, _vp,
                            _vp, _sz, C.POINTER(_dbl), C.POINTER(_i64)]),
    "bogp_acquire_async": (_i32, [_vp, _vp, C.POINTER(Candidates), _i64, _i64, _i32, _dbl, _dbl, _dbl, _vp, _vp, _vp,
                                  _vp, _sz, _vp]),
    "bogp_reduce_results": (_i32, [_vp, _vp, _i32, _vp, _pd, _pi64]),
    "bogp_score_argmax_async": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _dbl, _dbl, _vp, _vp]),
    "bogp_score_argmax": (_i32, [_vp, _vp, _vp, _i64, _i32, _dbl, _dbl, _vp, C.POINTER(_dbl), C.POINTER(_i64)]),
    "bogp_nlml_batched_workspace_bytes": (_sz, [_i64, _i32, _i64, _i32]),
    "bogp_nlml_batched": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _i64, _dbl, _vp, _vp, _vp, _sz]),
    "bogp_session_create": (_i32, [C.POINTER(_i32), _i32, C.POINTER(_vp)]),
    "bogp_session_destroy": (None, [_vp]),
    "bogp_session_device_count": (_i32, [_vp]),
    "bogp_session_ctx": (_vp, [_vp, _i32]),
    "bogp_session_launch_count": (_i64, [_vp]),
    "bogp_session_set_acquire_path": (_i32, [_vp, _i32]),
    "bogp_session_kernel_matrix": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _vp, _dbl, _vp]),
    "bogp_session_nlml": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _i64, _dbl, _vp, _vp]),
    "bogp_session_update": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _dbl, C.POINTER(HostCandidates), _i64, _i64, _dbl,
                                   _i32, _dbl, _dbl, _vp, _vp, _vp, _pd, _pd, _pi64]),
    "bogp_session_score": (_i32, [_vp, _vp, _vp, _i64, _i32, _dbl, _dbl, _vp, _pd, _pi64]),
}

_lib = None


def load():
    """Load libbogp.so (built in-tree by `__graft_entry__.build()` / `make -C csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  bayesian_optimisation_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != BOGP_OK:
        raise BogpError(code, load().bogp_last_error().decode(errors="replace"))
