"""Import the *unmodified* reference `PointSelector` from /root/reference.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: `/root/reference`
does not exist on the GPU box, so nothing in the `-m gpu` tests, `smoke()` or
`bench.py` calls this.  It is used by `oracle/make_golden.py` (to generate
`tests/golden/*.npz`) and by the CPU tests that cross-check the restated oracle
against the literal class when the reference tree is present.

The reference imports `matplotlib` and `plot_utils` at module scope
(`point_selector.py:2-4`); matplotlib is not installed here, so both are replaced by
no-op stubs.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

REFERENCE_DIR = os.environ.get("BOGP_REFERENCE_DIR", "/root/reference")

_PLOT_FUNCS = ["plot_ARD_LL", "plot_ARD_LL_1d", "surrogate_uncert_acquistion",
               "surrogate_uncert_acquistion_1d", "time_residual_agreement"]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "point_selector.py"))


def install_plot_stubs():
    """Put stub `matplotlib` / `plot_utils` modules into sys.modules (idempotent)."""
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = mpl.pyplot
    if "plot_utils" not in sys.modules or not hasattr(sys.modules["plot_utils"], "_bogp_stub"):
        pu = types.ModuleType("plot_utils")
        pu._bogp_stub = True
        pu.calls = []
        for name in _PLOT_FUNCS:
            def _f(*a, _n=name, **k):
                pu.calls.append(_n)
            setattr(pu, name, _f)
        pu.__all__ = list(_PLOT_FUNCS)
        sys.modules["plot_utils"] = pu
    return sys.modules["plot_utils"]


def load_reference_class():
    """Return the reference's `PointSelector` class (module loaded under a private name so
    that it never shadows the drop-in `point_selector` module)."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_DIR}")
    install_plot_stubs()
    spec = importlib.util.spec_from_file_location(
        "_reference_point_selector", os.path.join(REFERENCE_DIR, "point_selector.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.PointSelector


def run_reference(X, y, P, feature_domain, length_scales, explore=4, name="ref", iteration=1):
    """Drive the literal class the way select_parameters.py:282-293 does; returns a dict."""
    import numpy as np
    cls = load_reference_class()
    ps = cls()
    ps.name, ps.iteration = name, iteration
    ps.measured_pts = np.array(X, dtype=np.float64)
    ps.measured_vals = np.array(y, dtype=np.float64)
    ps.feature_domain = list(feature_domain)
    ps.predicted_pts = np.array(P, dtype=np.float64)
    ps.length_scales = length_scales
    with contextlib.redirect_stdout(io.StringIO()):
        ps.update_surrogate()
        idx = ps.lower_confidence_bound(explore) if explore != 4 else ps.lower_confidence_bound()
    return dict(kernel_params=np.array(ps.kernel_params, dtype=np.float64),
                mean_func=ps.mean_func, cov_func=ps.cov_func, acq=ps.acq_func_eval,
                index=np.array(idx), measured_pts_type=type(ps.measured_pts).__name__)
