"""Drop-in replacement for the reference's `point_selector.PointSelector`
(`/root/reference/point_selector.py:13-207`), with the arithmetic on a B200.

Same attribute bag, same two methods, same outputs as the caller
`select_parameters.py:146-157,282-293` expects:

    Optimiser = PointSelector()
    Optimiser.name, .iteration, .measured_pts, .measured_vals, .feature_domain,
             .predicted_pts, .length_scales = ...
    Optimiser.update_surrogate()
    next_sample = Optimiser.lower_confidence_bound()
    Optimiser.mean_func, .cov_func, .acq_func_eval        # arrays shaped feature_domain

What runs where (every step is ONE call of the host-buffer C ABI, include/bogp.h "session"; this module imports
numpy and ctypes only):
  * tune_kernel  -> bogp_session_nlml: one batched launch over the whole length-scale grid (K3,
                    csrc/lml_batched.cu) instead of the reference's 2500-iteration Python loop (`:127-138`); the
                    table is rounded to float32 and the first row-major minimum wins, like `:126,141`.
  * update_surrogate -> bogp_session_update: Gram + Cholesky + W = L^-1 (K1/K2, csrc/fit.cu), then the
                    acquisition sweep (K4, csrc/acquire*.cu) for mu and sigma; the C x C prior and
                    posterior covariances of `:78,91` are never formed (only their diagonal is used, `:98`).
  * lower_confidence_bound -> bogp_session_score: explore*sigma - mu and the first arg-max on the device,
                    on the mu / sigma the update left there.

Multi-GPU, caller-visible API identical:
  * one process driving several devices (the reference's deployment: a plain `python3 select_parameters.py`):
    `BOGP_DEVICES=all` (or "0,1,2,3", or the class attribute `PointSelector.devices`) -- the library shards the
    candidates over the devices in contiguous flat-index slices and reduces the winners;
  * one process per GPU (torchrun): when torch.distributed is initialised with more than one rank, every rank
    scores its slice on its own GPU, mu / sigma are all-gathered so that every rank holds the full arrays, and the
    winner comes from ONE exchange of (score, index) records (sharding.allreduce_maxloc).

Reference quirks that are part of the contract and reproduced here: two different jitters
(1e-4 for the LML, 1e-4+1e-6 for the posterior), the shape-equality jitter on k(X,P) when
M == C (`:173-177`), sqrt(abs(.)) (`:98`), midpoint length scales for a single measurement
(`:63-73`), measured_pts/vals turned into lists on exit (`:101-102`), kernel_params of shape
(1, 1) after a 1-D grid search (`:161`), IndexError on a NaN acquisition (`:207`).

Extensions (not in the reference): `expected_improvement()`, `acquisition="ei"`, candidate grids
given as axes (`predicted_axes`) so that 1e8-point sweeps never materialise `predicted_pts`.
"""
from __future__ import annotations

import numpy as np

import sys

from . import session as _session
from .session import ACQ_EI, ACQ_LCB, JITTER_LML, JITTER_POSTERIOR, PRIOR_DIAG

try:  # the reference does `from plot_utils import *` and calls the ARD plots inside tune_kernel
    import plot_utils as _plot_utils  # noqa: F401
except Exception:  # pragma: no cover - plotting is optional (matplotlib may be absent)
    _plot_utils = None


class PointSelector():

    devices = None      # optional class-wide device list for the single-process multi-GPU mode (else BOGP_DEVICES)

    def __init__(self):
        # same fields as point_selector.py:22-40
        self.feature_domain = None
        self.predicted_pts = None
        self.measured_vals = []
        self.measured_pts = []
        self.mean_func = None
        self.cov_func = None
        self.acq_func_eval = None
        self.hyperparam_obj = []
        self.length_scales = None
        self.kernel_params = None
        self.gradient_steps = 0.001
        self.iteration = None
        self.name = None
        self._cov_pred = None
        self._cov_meas = None
        self._cov_meas_pred = None
        # extensions
        self.predicted_axes = None       # optional: list of axes instead of predicted_pts (grid never materialised)
        self.verbose = False
        self.nlogml = None               # float32 LML table of the last tune_kernel
        self._session = None
        self._posterior_on_device = None  # (mean_func, cov_func, shard, session generation) the device copy belongs to

    # ------------------------------------------------------------------ helpers
    def _sess(self):
        if self._session is None:
            if type(self).devices is not None:
                cached = getattr(type(self), "_class_session", None)
                if cached is None or cached.devices != list(type(self).devices):
                    cached = _session.Session(list(type(self).devices))
                    type(self)._class_session = cached
                self._session = cached
            else:
                self._session = _session.default_session()
        return self._session

    @staticmethod
    def _dist():
        """(rank, world, module) when torch.distributed is initialised with several ranks, else None.  torch is only
        looked at if the caller has imported it already -- the single-process path never imports it."""
        torch = sys.modules.get("torch")
        if torch is None:
            return None
        dist = getattr(torch, "distributed", None)
        if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
            return None
        return dist.get_rank(), dist.get_world_size(), dist

    def _ell_vector(self, dim):
        """kernel_params as a length-`dim` vector.  A single length scale broadcasts over all
        features exactly as numpy broadcasting does in point_selector.py:188 (that is what the
        reference computes when a 1-D length-scale grid meets d > 1 features)."""
        ell = np.asarray(self.kernel_params, dtype=np.float64).reshape(-1)
        return np.full(dim, ell[0]) if (ell.size == 1 and dim > 1) else ell

    def _log(self, *a):
        if self.verbose:
            print(*a)

    # inspection matrices of point_selector.py:37-40, built lazily on the device
    def _lazy_cov(self, which):
        val = getattr(self, "_" + which)
        if val is not None or self.kernel_params is None:
            return val
        sess = self._sess()
        X = np.asarray(self.measured_pts, dtype=np.float64)
        P = np.asarray(self.predicted_pts, dtype=np.float64)
        ell = self._ell_vector(X.shape[1])
        if which == "cov_pred":
            val = sess.kernel_matrix(P, P, ell, PRIOR_DIAG - 1.0)
        elif which == "cov_meas":
            val = sess.kernel_matrix(X, X, ell, JITTER_POSTERIOR)
        else:
            val = sess.kernel_matrix(P, X, ell, JITTER_LML if X.shape == P.shape else 0.0)
        setattr(self, "_" + which, val)
        return val

    cov_pred = property(lambda self: self._lazy_cov("cov_pred"), lambda self, v: setattr(self, "_cov_pred", v))
    cov_meas = property(lambda self: self._lazy_cov("cov_meas"), lambda self, v: setattr(self, "_cov_meas", v))
    cov_meas_pred = property(lambda self: self._lazy_cov("cov_meas_pred"), lambda self, v: setattr(self, "_cov_meas_pred", v))

    # ------------------------------------------------------------------ reference API
    def update_surrogate(self):
        """point_selector.py:42-102"""
        self.measured_pts = np.array(self.measured_pts, dtype=np.float64)
        self.measured_vals = np.array(self.measured_vals, dtype=np.float64)
        self._cov_pred = self._cov_meas = self._cov_meas_pred = None

        if len(self.measured_pts[:, 0]) > 1:
            self._log("Beginning ARD kernel tuning ...")
            self.tune_kernel()
        else:
            self._log("Only 1 measued point. Setting length scales to mid points of range.")
            if len(self.length_scales) == 2:
                axis1, axis2 = self.length_scales[0], self.length_scales[1]
                self.kernel_params = np.array([axis1[len(axis1) // 2], axis2[len(axis2) // 2]])
            else:
                self.kernel_params = np.array([self.length_scales[len(self.length_scales) // 2]])

        sess = self._sess()
        ell = self._ell_vector(self.measured_pts.shape[1])
        if self.predicted_axes is not None:
            kw = dict(axes=[np.asarray(a, dtype=np.float64) for a in self.predicted_axes])
            count = int(np.prod([len(a) for a in kw["axes"]], dtype=np.int64))
            quirk = False
        else:
            pts = np.ascontiguousarray(self.predicted_pts, dtype=np.float64)
            kw = dict(points=pts)
            count, quirk = len(pts), tuple(pts.shape) == self.measured_pts.shape      # jitter rule of :173-177 applied at :81
        d = self._dist()
        b, e = (0, count) if d is None else _shard_range(count, d[0], d[1])
        mu = sigma = np.empty(0)
        generation = None
        if e > b:
            res = sess.update(self.measured_pts, self.measured_vals, ell, c_begin=b, c_end=e, jitter=JITTER_POSTERIOR, prior_diag=PRIOR_DIAG,
                              kind=ACQ_LCB, explore=4.0, cross_jitter=JITTER_LML if quirk else 0.0, outputs=True, **kw)
            mu, sigma, generation = res["mu"], res["sigma"], res["generation"]
        if d is not None:
            from . import sharding
            mu, sigma = sharding.all_gather_slices([mu, sigma], count, d[0], d[1])
        self.mean_func = mu.reshape(self.feature_domain)
        self.cov_func = sigma.reshape(self.feature_domain)
        self._posterior_on_device = (self.mean_func, self.cov_func, (b, e), generation)

        self.measured_pts = self.measured_pts.tolist()
        self.measured_vals = self.measured_vals.tolist()

    def tune_kernel(self):
        """Length-scale grid search on the log marginal likelihood, point_selector.py:104-163."""
        sess = self._sess()
        X = np.asarray(self.measured_pts, dtype=np.float64)
        y = np.asarray(self.measured_vals, dtype=np.float64)
        if len(self.length_scales) == 2:
            axis1 = np.asarray(self.length_scales[0], dtype=np.float64)
            axis2 = np.asarray(self.length_scales[1], dtype=np.float64)
            ells = np.stack(np.meshgrid(axis1, axis2, indexing="ij"), axis=-1).reshape(-1, 2)
            table = self._nlml_table(sess, X, y, ells)
            nlogml = table.astype(np.float32).reshape(len(axis1), len(axis2))          # float32 table, :126
            min_idx = np.argwhere(nlogml == np.amin(nlogml))[0]                        # first row-major minimum, :141
            self.kernel_params = np.array([axis1[min_idx[0]], axis2[min_idx[1]]])
            self._log(f"Updated length scales to [{[axis1[min_idx[0]], axis2[min_idx[1]]]}]")
            self.nlogml = nlogml
            if _plot_utils is not None and hasattr(_plot_utils, "plot_ARD_LL"):
                _plot_utils.plot_ARD_LL(nlogml, self.kernel_params, self.length_scales, self.name, self.iteration)
        else:
            grid = np.asarray(self.length_scales, dtype=np.float64).reshape(-1)
            table = self._nlml_table(sess, X, y, np.repeat(grid.reshape(-1, 1), X.shape[1], axis=1))
            nlogml = table.astype(np.float32)                                          # :150
            min_idx = np.argwhere(nlogml == np.amin(nlogml))[0]                        # :159
            self.kernel_params = np.array([grid[min_idx]])                             # shape (1, 1), :161
            self._log(f"Updated length scale to [{grid[min_idx]}].")
            self.nlogml = nlogml
            if _plot_utils is not None and hasattr(_plot_utils, "plot_ARD_LL_1d"):
                _plot_utils.plot_ARD_LL_1d(nlogml, self.kernel_params, self.length_scales, self.name, self.iteration)

    def _nlml_table(self, sess, X, y, ells):
        """The whole nlml table (float64) on every rank: restarts r, r+G, ... per rank under torchrun."""
        d = self._dist()
        if d is None or len(ells) < d[1]:
            return sess.nlml(X, y, ells, JITTER_LML)
        from . import sharding
        rank, world, _ = d
        mine = np.arange(rank, len(ells), world)
        part = sess.nlml(X, y, ells[mine], JITTER_LML)
        return sharding.all_gather_strided(part, len(ells), rank, world)

    def kernel_rbf(self, x1, x2):
        """point_selector.py:166-195 (jitter iff the shapes are equal)."""
        x1 = np.asarray(x1, dtype=np.float64)
        x2 = np.asarray(x2, dtype=np.float64)
        ell = self._ell_vector(x1.shape[1])
        return self._sess().kernel_matrix(x1, x2, ell, JITTER_LML if x1.shape == x2.shape else 0.0)

    def _acquire(self, kind, explore, f_best):
        shape = np.shape(self.mean_func)
        count = int(np.prod(shape, dtype=np.int64))
        sess = self._sess()
        d = self._dist()
        held = self._posterior_on_device
        fresh = held is not None and held[0] is self.mean_func and held[1] is self.cov_func and held[3] == sess.generation
        if d is None:
            if fresh:
                res = sess.score(count, kind=kind, explore=explore, f_best=f_best)
            else:       # the caller replaced mean_func / cov_func: score what it holds now
                res = sess.score(count, kind=kind, explore=explore, f_best=f_best, mu=self.mean_func, sigma=self.cov_func)
                self._posterior_on_device = None
            acq, best = res["acq"], res["best_index"]
        else:
            from . import sharding
            rank, world, _ = d
            b, e = _shard_range(count, rank, world)
            mu = np.asarray(self.mean_func, dtype=np.float64).reshape(-1)[b:e]
            sg = np.asarray(self.cov_func, dtype=np.float64).reshape(-1)[b:e]
            if e > b:
                if fresh and held[2] == (b, e):
                    res = sess.score(e - b, kind=kind, explore=explore, f_best=f_best)
                else:
                    res = sess.score(e - b, kind=kind, explore=explore, f_best=f_best, mu=mu, sigma=sg)
                    res["best_index"] += b
                    self._posterior_on_device = None
                part, s, i = res["acq"], res["best_score"], res["best_index"]
            else:
                part, s, i = np.empty(0), float("-inf"), sharding.NO_INDEX
            (acq,) = sharding.all_gather_slices([part], count, rank, world)
            _, best = sharding.allreduce_maxloc(s, i)
        self.acq_func_eval = acq.reshape(shape)
        return np.array(np.unravel_index(best, shape), dtype=np.int64)

    def lower_confidence_bound(self, explore=4):
        """point_selector.py:197-207: maximise explore*sigma - mu; lowest flat index among ties."""
        return self._acquire(ACQ_LCB, float(explore), 0.0)

    # ------------------------------------------------------------------ extensions
    def expected_improvement(self, f_best=None):
        """EI for minimisation (docs/README.md:364 lists it as future work).  f_best defaults to min(y)."""
        if f_best is None:
            f_best = float(np.min(np.asarray(self.measured_vals, dtype=np.float64)))
        return self._acquire(ACQ_EI, 0.0, float(f_best))


def _shard_range(c_total, rank, world):
    """Flat indices [begin, end) scored by `rank`: ceil(C/G)-sized contiguous slices."""
    per = -(-c_total // world)
    b = min(c_total, rank * per)
    return b, min(c_total, b + per)
