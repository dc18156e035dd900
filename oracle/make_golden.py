"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE in the build container.

    python -m oracle.make_golden          # needs /root/reference (read-only)

TEST INFRASTRUCTURE ONLY.  Every array written here comes from the reference's own
`PointSelector` (imported by `oracle/reference_loader.py` with the plotting modules
stubbed); nothing is produced by the oracle or by the CUDA path.  The domains and
length-scale grids are the reference's (`select_parameters.py:62-83`).
"""
from __future__ import annotations

import contextlib
import io
import os

import numpy as np

from . import reference_loader as rl

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

G = 50
T1, T2 = np.linspace(1, 14, G), np.linspace(10, 90, G)
T3, T4 = np.linspace(60, 150, G), np.linspace(200, 500, G)
TR = np.linspace(0.1, 2.0, G)
W59 = np.linspace(0.01, 0.9, G)
L1, L2 = np.linspace(0.5, 10, 50), np.linspace(2, 100, 50)
L3, L4 = np.linspace(10, 30, 50), np.linspace(50, 100, 50)
LTH = np.linspace(0.1, 2, 20)


def synthetic_objective(pts, scale=1e8):
    """Smooth bowl with a ripple, O(1e8) like the real objective (docs/algo_output.png)."""
    z = (pts - pts.mean(axis=0)) / (np.ptp(pts, axis=0) + 1e-12)
    return scale * (0.2 + (z ** 2).sum(axis=1) + 0.05 * np.sin(7 * z.sum(axis=1)))


def grid2(a, b):
    return np.array([[x, y] for x in a for y in b])


def native_2d(seed, m, axes, lgrids, scale=1e8):
    rng = np.random.default_rng(seed)
    P = grid2(*axes)
    idx = rng.choice(len(P), m, replace=False)
    X = P[idx]
    y = synthetic_objective(P, scale)[idx] * (1 + 0.05 * rng.standard_normal(m))
    ls = np.array([lgrids[0], lgrids[1]])
    out = rl.run_reference(X, y, P, [G, G], ls)
    return dict(X=X, y=y, P=P, feature_domain=np.array([G, G]), ls0=lgrids[0], ls1=lgrids[1],
                **{k: np.asarray(v) for k, v in out.items() if k != "measured_pts_type"})


def native_1d(seed, m, axis, lgrid, scale=1e8):
    rng = np.random.default_rng(seed)
    P = axis.reshape(-1, 1)
    idx = rng.choice(len(P), m, replace=False)
    X = P[idx]
    y = synthetic_objective(P, scale)[idx] * (1 + 0.05 * rng.standard_normal(m))
    out = rl.run_reference(X, y, P, [G], lgrid)
    return dict(X=X, y=y, P=P, feature_domain=np.array([G]), ls0=lgrid,
                **{k: np.asarray(v) for k, v in out.items() if k != "measured_pts_type"})


def direct_case(seed, n, c, d, ell_value):
    """kernel_rbf / eval_log_marginal / posterior of the literal class at d > 2 (the grid
    search only handles d <= 2, so kernel_params is supplied)."""
    cls = rl.load_reference_class()
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    y = np.sin(3.0 * X.sum(axis=1)) + 0.1 * rng.standard_normal(n)
    P = rng.random((c, d))
    ell = np.full(d, ell_value) * (1 + 0.2 * rng.random(d))
    ps = cls()
    ps.kernel_params = ell
    Kxx = ps.kernel_rbf(X, X)
    Kxp = ps.kernel_rbf(X, P)
    # literal nlml, point_selector.py:111-120 (copied semantics: inv + log(det))
    inv = np.linalg.inv(Kxx)
    with np.errstate(divide="ignore"):
        nl = 0.5 * (y.T @ inv @ y + np.log(np.linalg.det(Kxx)) + n * np.log(2 * np.pi))
    # literal posterior, point_selector.py:78-98, via the class with tune_kernel bypassed
    ps.measured_pts, ps.measured_vals = X, y
    ps.predicted_pts, ps.feature_domain = P, [c]
    ps.tune_kernel = lambda: None
    ps.length_scales = np.array([1.0])
    with contextlib.redirect_stdout(io.StringIO()):
        ps.update_surrogate()
        idx = ps.lower_confidence_bound()
    return dict(X=X, y=y, P=P, ell=ell, Kxx=Kxx, Kxp=Kxp, nlml=np.float64(nl),
                mean_func=ps.mean_func, cov_func=ps.cov_func, acq=ps.acq_func_eval,
                index=np.array(idx), feature_domain=np.array([c]))


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {}
    for m in (1, 2, 5, 10, 21):
        cases[f"native2d_t1t2_m{m}"] = native_2d(100 + m, m, (T1, T2), (L1, L2))
    cases["native2d_t3t4_m7"] = native_2d(207, 7, (T3, T4), (L3, L4))
    cases["native2d_t3t4_m15_smally"] = native_2d(215, 15, (T3, T4), (L3, L4), scale=1.0)
    for m in (1, 3, 8):
        cases[f"native1d_tr_m{m}"] = native_1d(300 + m, m, TR, L1)
    cases["native1d_a1_m6"] = native_1d(406, 6, W59, LTH)
    cases["direct_d6_n64_c300"] = direct_case(1, 64, 300, 6, 0.3)
    cases["direct_d8_n128_c512"] = direct_case(2, 128, 512, 8, 0.3)
    cases["direct_d3_n40_c40_quirk"] = direct_case(3, 40, 40, 3, 0.4)   # M == C jitter quirk
    cases["direct_d10_n200_c64"] = direct_case(4, 200, 64, 10, 0.5)
    for name, c in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **c)
        print(name, {k: np.asarray(v).shape for k, v in c.items()})


if __name__ == "__main__":
    main()
