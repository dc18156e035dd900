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

What runs where:
  * tune_kernel  -> one batched launch over the whole length-scale grid (K3, csrc/lml_batched.cu)
                    instead of the reference's 2500-iteration Python loop (`:127-138`); the table
                    is rounded to float32 and the first row-major minimum wins, like `:126,141`.
  * update_surrogate -> Gram + Cholesky + W = L^-1 (K1/K2, csrc/fit.cu), then the fused
                    acquisition sweep (K4, csrc/acquire.cu) for mu and sigma; the C x C prior and
                    posterior covariances of `:78,91` are never formed (only their diagonal is used, `:98`).
  * lower_confidence_bound -> explore*sigma - mu and the first arg-max on the device.

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

from . import engine as _engine
from .engine import ACQ_EI, ACQ_LCB, JITTER_LML, JITTER_POSTERIOR, PRIOR_DIAG, CandidateGrid

try:  # the reference does `from plot_utils import *` and calls the ARD plots inside tune_kernel
    import plot_utils as _plot_utils  # noqa: F401
except Exception:  # pragma: no cover - plotting is optional (matplotlib may be absent)
    _plot_utils = None


class PointSelector():

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
        self._engine = None
        self._dev = None                 # device-resident (mu, sigma) of the last update

    # ------------------------------------------------------------------ helpers
    def _eng(self):
        if self._engine is None:
            self._engine = _engine.default_engine()
        return self._engine

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
        eng = self._eng()
        X = np.asarray(self.measured_pts, dtype=np.float64)
        P = np.asarray(self.predicted_pts, dtype=np.float64)
        ell = self._ell_vector(X.shape[1])
        if which == "cov_pred":
            val = eng.kernel_matrix(P, P, ell, PRIOR_DIAG - 1.0).cpu().numpy()
        elif which == "cov_meas":
            val = eng.kernel_matrix(X, X, ell, JITTER_POSTERIOR).cpu().numpy()
        else:
            val = eng.kernel_matrix(P, X, ell, JITTER_LML if X.shape == P.shape else 0.0).cpu().numpy()
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

        # The candidate array starts its way to the device now, on a side stream: the copy (asynchronous when the caller's
        # buffer is pinned) runs underneath the length-scale search and the fit below.
        cand_dev = None
        if self.predicted_axes is None:
            cand_dev = self._eng().prefetch_to_device(np.ascontiguousarray(self.predicted_pts, dtype=np.float64))

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

        eng = self._eng()
        ell = self._ell_vector(self.measured_pts.shape[1])
        fit = eng.fit(self.measured_pts, self.measured_vals, ell, JITTER_POSTERIOR)
        try:
            if self.predicted_axes is not None:
                cand = CandidateGrid([np.asarray(a, dtype=np.float64) for a in self.predicted_axes])
                count, quirk = cand.size, False
            else:
                cand = eng.prefetched(cand_dev)
                count, quirk = len(cand), tuple(cand.shape) == self.measured_pts.shape   # jitter rule of :173-177 applied at :81
            res = eng.acquire(fit, cand, 0, count, kind=ACQ_LCB, explore=4.0, prior_diag=PRIOR_DIAG, outputs=True,
                              cross_jitter=JITTER_LML if quirk else 0.0)
        finally:
            fit.close()
        self._dev = (res.mu, res.sigma)
        self.mean_func = eng.to_host(res.mu).reshape(self.feature_domain)
        self.cov_func = eng.to_host(res.sigma).reshape(self.feature_domain)
        self._dev_host = (self.mean_func, self.cov_func)

        self.measured_pts = self.measured_pts.tolist()
        self.measured_vals = self.measured_vals.tolist()

    def tune_kernel(self):
        """Length-scale grid search on the log marginal likelihood, point_selector.py:104-163."""
        eng = self._eng()
        X = np.asarray(self.measured_pts, dtype=np.float64)
        y = np.asarray(self.measured_vals, dtype=np.float64)
        if len(self.length_scales) == 2:
            axis1 = np.asarray(self.length_scales[0], dtype=np.float64)
            axis2 = np.asarray(self.length_scales[1], dtype=np.float64)
            ells = np.stack(np.meshgrid(axis1, axis2, indexing="ij"), axis=-1).reshape(-1, 2)
            table = eng.nlml_batched(X, y, ells, JITTER_LML).cpu().numpy()
            nlogml = table.astype(np.float32).reshape(len(axis1), len(axis2))          # float32 table, :126
            min_idx = np.argwhere(nlogml == np.amin(nlogml))[0]                        # first row-major minimum, :141
            self.kernel_params = np.array([axis1[min_idx[0]], axis2[min_idx[1]]])
            self._log(f"Updated length scales to [{[axis1[min_idx[0]], axis2[min_idx[1]]]}]")
            self.nlogml = nlogml
            if _plot_utils is not None and hasattr(_plot_utils, "plot_ARD_LL"):
                _plot_utils.plot_ARD_LL(nlogml, self.kernel_params, self.length_scales, self.name, self.iteration)
        else:
            grid = np.asarray(self.length_scales, dtype=np.float64).reshape(-1)
            table = eng.nlml_batched(X, y, np.repeat(grid.reshape(-1, 1), X.shape[1], axis=1), JITTER_LML).cpu().numpy()
            nlogml = table.astype(np.float32)                                          # :150
            min_idx = np.argwhere(nlogml == np.amin(nlogml))[0]                        # :159
            self.kernel_params = np.array([grid[min_idx]])                             # shape (1, 1), :161
            self._log(f"Updated length scale to [{grid[min_idx]}].")
            self.nlogml = nlogml
            if _plot_utils is not None and hasattr(_plot_utils, "plot_ARD_LL_1d"):
                _plot_utils.plot_ARD_LL_1d(nlogml, self.kernel_params, self.length_scales, self.name, self.iteration)

    def kernel_rbf(self, x1, x2):
        """point_selector.py:166-195 (jitter iff the shapes are equal)."""
        x1 = np.asarray(x1, dtype=np.float64)
        x2 = np.asarray(x2, dtype=np.float64)
        ell = self._ell_vector(x1.shape[1])
        return self._eng().kernel_matrix(x1, x2, ell, JITTER_LML if x1.shape == x2.shape else 0.0).cpu().numpy()

    def _device_mu_sigma(self):
        if self._dev is not None and getattr(self, "_dev_host", None) is not None \
                and self._dev_host[0] is self.mean_func and self._dev_host[1] is self.cov_func:
            return self._dev
        return np.asarray(self.mean_func, dtype=np.float64).reshape(-1), np.asarray(self.cov_func, dtype=np.float64).reshape(-1)

    def _acquire(self, kind, explore, f_best):
        mu, sigma = self._device_mu_sigma()
        res = self._eng().score_argmax(mu, sigma, kind=kind, explore=explore, f_best=f_best)
        shape = np.shape(self.mean_func)
        self.acq_func_eval = self._eng().to_host(res.acq).reshape(shape)
        return np.array(np.unravel_index(res.best_index, shape), dtype=np.int64)

    def lower_confidence_bound(self, explore=4):
        """point_selector.py:197-207: maximise explore*sigma - mu; lowest flat index among ties."""
        return self._acquire(ACQ_LCB, float(explore), 0.0)

    # ------------------------------------------------------------------ extensions
    def expected_improvement(self, f_best=None):
        """EI for minimisation (docs/README.md:364 lists it as future work).  f_best defaults to min(y)."""
        if f_best is None:
            f_best = float(np.min(np.asarray(self.measured_vals, dtype=np.float64)))
        return self._acquire(ACQ_EI, 0.0, float(f_best))
