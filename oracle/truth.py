"""TEST INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/truth_ld.c: the GP posterior and log marginal
likelihood of /root/reference/point_selector.py:78-98,111-120 evaluated in x87 extended precision
(long double) with a Cholesky factorisation -- the truth both the reference's fp64 `inv` arithmetic and
the B200 path are measured against in the ill-conditioned parity tests."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtruth_ld.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def _load():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "truth_ld.c")):
            build()
        lib = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        lib.gp_truth_ld.restype = C.c_int
        lib.gp_truth_ld.argtypes = [dp, dp, C.c_long, C.c_int, dp, C.c_double, dp, C.c_long, C.c_double, C.c_double, dp, dp, dp, dp]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def posterior_truth(X, y, P, ell, jitter, prior, cross_jitter=0.0):
    """(mu, var, nlml, logdet) in long-double arithmetic, rounded to fp64 at the very end."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(-1, X.shape[1])
    ell = np.ascontiguousarray(np.broadcast_to(np.asarray(ell, dtype=np.float64).reshape(-1), (X.shape[1],)))
    mu, var = np.empty(len(P)), np.empty(len(P))
    nl, ld = np.empty(1), np.empty(1)
    rc = _load().gp_truth_ld(_p(X), _p(y), len(X), X.shape[1], _p(ell), float(jitter), _p(P), len(P), float(prior),
                             float(cross_jitter), _p(mu), _p(var), _p(nl), _p(ld))
    if rc > 0:
        raise np.linalg.LinAlgError(f"long-double Cholesky: pivot {rc} is not positive")
    if rc < 0:
        raise MemoryError("gp_truth_ld")
    return mu, var, float(nl[0]), float(ld[0])
