"""Host-buffer session: the thin ctypes layer between the reference-facing `PointSelector` and the
"session" entry points of libbogp (include/bogp.h).  numpy arrays in, numpy arrays out; device memory,
streams, copies and the sharding over the GPUs of one box live inside the library.  This module imports
neither torch nor anything else heavy, so that a one-shot `select_parameters.py` process (the reference's
deployment: one OS process per DAG node) pays only the CUDA context for its start-up.

Which devices:  `Session(devices=[0, 1, ...])`, else the environment variable BOGP_DEVICES ("all" or a
comma-separated list), else -- under torchrun, one process per GPU -- LOCAL_RANK, else device 0.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ACQ_EI, ACQ_LCB, HostCandidates  # noqa: F401

# reference jitters: kernel_rbf adds 1e-4 (point_selector.py:193), update_surrogate another 1e-6 (:78-79)
JITTER_LML = 1e-4
PRIOR_DIAG = (1.0 + 1e-4) + 1e-6            # diag of cov_pred, same rounding order as the reference
JITTER_POSTERIOR = PRIOR_DIAG - 1.0         # exact: 1.0 + JITTER_POSTERIOR == PRIOR_DIAG bit for bit


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def devices_from_environment() -> Sequence[int]:
    spec = os.environ.get("BOGP_DEVICES", "").strip()
    if spec:
        if spec.lower() == "all":
            return list(range(_cuda_device_count()))
        return [int(t) for t in spec.split(",") if t.strip() != ""]
    if "LOCAL_RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return [int(os.environ["LOCAL_RANK"])]
    return [0]


def _cuda_device_count() -> int:
    lib = _lib.load()
    n = C.c_int(0)
    lib.bogp_device_count(C.byref(n))
    return max(1, n.value)


class Session:
    def __init__(self, devices: Optional[Sequence[int]] = None):
        self.lib = _lib.load()
        devs = list(devices_from_environment() if devices is None else devices)
        arr = (C.c_int * len(devs))(*devs)
        h = C.c_void_p()
        _lib.check(self.lib.bogp_session_create(arr, len(devs), C.byref(h)))
        self._h = h
        self.devices = devs
        self.generation = 0          # bumped whenever the posterior kept on the device(s) is replaced

    def close(self):
        if getattr(self, "_h", None):
            self.lib.bogp_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.bogp_session_launch_count(self._h))

    def set_acquire_path(self, path: str):
        code = {"fp64": _lib.PATH_FP64_DMMA, "dmma": _lib.PATH_FP64_DMMA, "i8": _lib.PATH_INT8_TCGEN05, "int8": _lib.PATH_INT8_TCGEN05}[path]
        _lib.check(self.lib.bogp_session_set_acquire_path(self._h, code))

    def set_screening(self, enable: bool):
        """Screening of arg-max-only sweeps (`update(..., outputs=False)`), per device context (include/bogp.h)."""
        for i in range(len(self.devices)):
            _lib.check(self.lib.bogp_set_screening(C.c_void_p(self.lib.bogp_session_ctx(self._h, i)), 1 if enable else 0))

    def set_fused(self, enable: bool, group: int = 0):
        """One persistent fused kernel per sweep, or (default) the separate kernels, per device context (include/bogp.h)."""
        for i in range(len(self.devices)):
            _lib.check(self.lib.bogp_set_fused(C.c_void_p(self.lib.bogp_session_ctx(self._h, i)), 1 if enable else 0, int(group)))

    # ------------------------------------------------------------------ point_selector.py:166-195
    def kernel_matrix(self, a, b, ell, jitter: float = 0.0) -> np.ndarray:
        a, b = _f64(a), _f64(b)
        ell = _f64(ell, -1)
        self.generation += 1
        out = np.empty((a.shape[0], b.shape[0]))
        _lib.check(self.lib.bogp_session_kernel_matrix(self._h, _ptr(a), a.shape[0], _ptr(b), b.shape[0], a.shape[1], _ptr(ell),
                                                       float(jitter), _ptr(out)))
        return out

    # ------------------------------------------------------------------ point_selector.py:104-163
    def nlml(self, x, y, ells, jitter: float = JITTER_LML, want_grad: bool = False):
        x, y, ells = _f64(x), _f64(y, -1), _f64(ells)
        if ells.ndim != 2 or ells.shape[1] != x.shape[1]:
            raise ValueError("ells must be (R, d)")
        out = np.empty(ells.shape[0])
        grad = np.empty(ells.shape) if want_grad else None
        _lib.check(self.lib.bogp_session_nlml(self._h, _ptr(x), _ptr(y), x.shape[0], x.shape[1], _ptr(ells), ells.shape[0],
                                              float(jitter), _ptr(out), _ptr(grad)))
        return (out, grad) if want_grad else out

    # ------------------------------------------------------------------ point_selector.py:42-102
    def update(self, x, y, ell, points=None, axes=None, c_begin: int = 0, c_end: Optional[int] = None, jitter: float = JITTER_POSTERIOR,
               prior_diag: float = PRIOR_DIAG, kind: int = ACQ_LCB, explore: float = 4.0, f_best: float = 0.0,
               cross_jitter: float = 0.0, outputs: bool = True, want_acq: bool = False):
        """Fit + posterior of candidates [c_begin, c_end).  Returns dict(mu, sigma, acq, nlml, best_score, best_index);
        best_index is the GLOBAL flat index."""
        x, y, ell = _f64(x), _f64(y, -1), _f64(ell, -1)
        n, dim = x.shape
        if ell.size != dim:
            raise ValueError(f"{ell.size} length scales for {dim} features")
        cd = HostCandidates()
        cd.cross_jitter = float(cross_jitter)
        keep = []
        if axes is not None:
            ax = [_f64(a, -1) for a in axes]
            if len(ax) != dim:
                raise ValueError("candidate grid dimension does not match the measured points")
            flat = np.concatenate(ax)
            lens = (C.c_int32 * dim)(*[len(a) for a in ax])
            keep += [flat, lens]
            cd.h_points, cd.h_axes, cd.h_axis_len = None, flat.ctypes.data, lens
            cd.c_total = int(np.prod([len(a) for a in ax], dtype=np.int64))
        else:
            pts = _f64(points)
            if pts.ndim != 2 or pts.shape[1] != dim:
                raise ValueError("candidates must be (C, d)")
            keep.append(pts)
            cd.h_points, cd.h_axes, cd.h_axis_len, cd.c_total = pts.ctypes.data, None, None, pts.shape[0]
        c_end = int(cd.c_total) if c_end is None else int(c_end)
        count = c_end - int(c_begin)
        mu = np.empty(count) if outputs else None
        sigma = np.empty(count) if outputs else None
        acq = np.empty(count) if (outputs and want_acq) else None
        nl, bs, bi = C.c_double(), C.c_double(), C.c_int64()
        self.generation += 1
        code = self.lib.bogp_session_update(self._h, _ptr(x), _ptr(y), n, dim, _ptr(ell), float(jitter), C.byref(cd), int(c_begin), c_end,
                                            float(prior_diag), int(kind), float(explore), float(f_best), _ptr(mu), _ptr(sigma), _ptr(acq),
                                            C.byref(nl), C.byref(bs), C.byref(bi))
        del keep
        _raise(self.lib, code)
        return dict(mu=mu, sigma=sigma, acq=acq, nlml=nl.value, best_score=bs.value, best_index=int(bi.value), generation=self.generation)

    # ------------------------------------------------------------------ point_selector.py:197-207
    def score(self, count: int, kind: int = ACQ_LCB, explore: float = 4.0, f_best: float = 0.0, mu=None, sigma=None, want_acq: bool = True):
        """Acquisition + first arg-max on the posterior of the last `update` (kept on the device), or on host arrays."""
        if mu is not None:
            mu, sigma = _f64(mu, -1), _f64(sigma, -1)
            self.generation += 1
        acq = np.empty(int(count)) if want_acq else None
        bs, bi = C.c_double(), C.c_int64()
        code = self.lib.bogp_session_score(self._h, _ptr(mu), _ptr(sigma), int(count), int(kind), float(explore), float(f_best), _ptr(acq),
                                           C.byref(bs), C.byref(bi))
        _raise(self.lib, code)
        return dict(acq=acq, best_score=bs.value, best_index=int(bi.value))


def _raise(lib, code):
    if code == _lib.BOGP_ERR_NOT_POSDEF:
        raise np.linalg.LinAlgError(lib.bogp_last_error().decode())
    if code == _lib.BOGP_ERR_NAN_SCORE:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0 (NaN acquisition value)")
    _lib.check(code)


_default = None


def default_session() -> Session:
    """Process-wide session on the devices the environment names (module docstring)."""
    global _default
    if _default is None:
        _default = Session()
    return _default


def set_default_session(session: Optional[Session]):
    """Install (or drop, with None) the process-wide session -- tests inject a stand-in here."""
    global _default
    _default = session
