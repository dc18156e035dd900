"""The restated CPU oracle against golden vectors produced by the unmodified reference
(`oracle/make_golden.py`) and, when /root/reference is present, against the live class."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import gp_oracle as o


def _length_scales(g):
    return np.array([g["ls0"], g["ls1"]]) if "ls1" in g else g["ls0"]


@pytest.mark.parametrize("name", golden_names("native"))
def test_native_cases_match_reference(name):
    g = load_golden(name)
    r = o.select_next(g["X"], g["y"], g["P"], list(g["feature_domain"]), _length_scales(g))
    assert r["kernel_params"].shape == g["kernel_params"].shape
    np.testing.assert_array_equal(r["kernel_params"], g["kernel_params"])
    np.testing.assert_allclose(r["mean_func"], g["mean_func"], rtol=1e-11, atol=1e-9 * np.abs(g["mean_func"]).max())
    # sigma^2 = 1.000101 - q is a cancellation: absolute error ~ cond(K)*eps of the prior
    np.testing.assert_allclose(r["cov_func"] ** 2, g["cov_func"] ** 2, rtol=1e-9, atol=2e-11)
    np.testing.assert_array_equal(r["index"], g["index"])


@pytest.mark.parametrize("name", golden_names("direct"))
def test_direct_cases_match_reference(name):
    g = load_golden(name)
    X, y, P, ell = g["X"], g["y"], g["P"], g["ell"]
    np.testing.assert_array_equal(o.kernel_rbf(X, X, ell), g["Kxx"])
    np.testing.assert_array_equal(o.kernel_rbf(X, P, ell), g["Kxp"])
    if np.isfinite(g["nlml"]):
        assert o.nlml(X, y, ell, stable=False) == g["nlml"]
        np.testing.assert_allclose(o.nlml(X, y, ell, stable=True), g["nlml"], rtol=1e-12)
    mu, sig = o.posterior_literal(X, y, P, ell)
    np.testing.assert_array_equal(mu, g["mean_func"])
    np.testing.assert_array_equal(sig, g["cov_func"])
    mu2, var2 = o.posterior_diag(X, y, P, ell, chunk=97, return_var=True)
    np.testing.assert_allclose(mu2, g["mean_func"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(var2, g["cov_func"] ** 2, rtol=1e-9, atol=2e-11)   # cancellation against the prior 1.000101
    acq = o.lcb(g["mean_func"], g["cov_func"])
    np.testing.assert_array_equal(acq, g["acq"])
    np.testing.assert_array_equal(o.first_argmax(acq), g["index"])


def test_quirk_case_really_has_cross_jitter():
    g = load_golden("direct_d3_n40_c40_quirk")
    plain = o.kernel_rbf_chunked(g["X"], g["P"], g["ell"])
    np.testing.assert_allclose(g["Kxp"] - plain, 1e-4 * np.eye(40), atol=1e-15)


def test_stable_nlml_handles_det_underflow():
    X, y, ell = o.synthetic_problem(1024, 6)
    with np.errstate(all="ignore"):
        assert o.nlml(X, y, ell, stable=False) == -np.inf
    assert np.isfinite(o.nlml(X, y, ell, stable=True))


def test_nlml_grad_matches_finite_differences():
    X, y, ell = o.synthetic_problem(60, 4, seed=3, ell=0.4)
    ell = ell * np.array([1.0, 1.3, 0.8, 1.1])
    g = o.nlml_grad(X, y, ell)
    for k in range(4):
        h = 1e-6 * ell[k]
        e1, e2 = ell.copy(), ell.copy()
        e1[k] += h
        e2[k] -= h
        fd = (o.nlml(X, y, e1) - o.nlml(X, y, e2)) / (2 * h)
        assert abs(fd - g[k]) <= 1e-6 * max(1.0, abs(g[k]))


def test_ei_against_scipy_norm_and_limits():
    from scipy.stats import norm
    rng = np.random.default_rng(0)
    mu, sig = rng.normal(size=200), rng.random(200) + 0.01
    fb = -0.3
    z = (fb - mu) / sig
    np.testing.assert_allclose(o.expected_improvement(mu, sig, fb), (fb - mu) * norm.cdf(z) + sig * norm.pdf(z), rtol=1e-12, atol=1e-300)
    np.testing.assert_array_equal(o.expected_improvement(np.array([-1.0, 1.0]), np.zeros(2), 0.0), [1.0, 0.0])


def test_first_argmax_tie_break_and_nan():
    a = np.array([[1.0, 3.0, 3.0], [3.0, 0.0, 3.0]])
    np.testing.assert_array_equal(o.first_argmax(a), [0, 1])
    with pytest.raises(IndexError):
        o.first_argmax(np.array([1.0, np.nan]))


def test_grid_points_is_row_major_axis0_slowest():
    axes = [np.linspace(0, 1, 3), np.linspace(2, 3, 4), np.linspace(5, 6, 2)]
    full = o.candidate_grid(axes)
    ref = np.array([[a, b, c] for a in axes[0] for b in axes[1] for c in axes[2]])   # select_parameters.py:273-279
    np.testing.assert_array_equal(full, ref)
    np.testing.assert_array_equal(o.grid_points(axes, 5, 17), full[5:17])


@pytest.mark.reference
def test_oracle_against_live_reference_random():
    from oracle import reference_loader as rl
    rng = np.random.default_rng(42)
    t3, t4 = np.linspace(60, 150, 50), np.linspace(200, 500, 50)
    P = o.candidate_grid([t3, t4])
    ls = np.array([np.linspace(10, 30, 50), np.linspace(50, 100, 50)])
    for m in (3, 12):
        idx = rng.choice(2500, m, replace=False)
        X, y = P[idx], rng.uniform(1e7, 1e9, m)
        ref = rl.run_reference(X, y, P, [50, 50], ls)
        r = o.select_next(X, y, P, [50, 50], ls)
        np.testing.assert_array_equal(r["kernel_params"], ref["kernel_params"])
        np.testing.assert_allclose(r["mean_func"], ref["mean_func"], rtol=1e-11, atol=1e-9 * np.abs(ref["mean_func"]).max())
        np.testing.assert_array_equal(r["index"], ref["index"])
        assert ref["measured_pts_type"] == "list"


def test_oracle_replays_the_reference_closed_loop_trace():
    """Every 4th PointSelector call of the recorded reference workflow (select_parameters.py +
    terminate_*.py + synthetic objective): same length scales, same selected index."""
    from conftest import load_closed_loop
    calls = load_closed_loop()
    assert len(calls) > 100 and {len(c["X"]) for c in calls} >= {1, 2, 5}
    assert {c["X"].shape[1] for c in calls} == {1, 2}
    for c in calls[::4]:
        r = o.select_next(c["X"], c["y"], c["P"], c["feature_domain"], c["length_scales"])
        assert np.asarray(r["kernel_params"]).shape == c["kernel_params"].shape
        np.testing.assert_array_equal(r["kernel_params"], c["kernel_params"])
        np.testing.assert_array_equal(r["index"], c["index"])
        assert abs(r["acq"].max() - c["acq_max"]) <= 1e-9 * max(1.0, abs(c["acq_max"]))
