"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden vectors
produced by the unmodified reference.  Tolerances (BASELINE.json north_star): posterior mean,
variance and log marginal likelihood within 1e-9 relative in fp64; selected index exact."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, record_error
from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu

RTOL = 1e-9
# Absolute slack of the sigma^2 comparisons.  Round 1 used max(2e-11, cond*eps); the long-double truth of
# tests/test_gpu_configs.py shows that this is the ORACLE's error (explicit inverse), not the device's (1e-14 absolute even at
# cond 4e7), so it is now what the runs actually need: see profiles/r02_parity_errors.json for the achieved values.
VAR_FLOOR = 2e-12


_ENGINE = None


@pytest.fixture(params=["i8", "fp64"], autouse=True)
def acquire_path(request):
    """Every test runs on both tensor paths of the acquisition product (include/bogp.h):
    "i8" = exact digit slices on tcgen05 kind::i8 (default), "fp64" = DMMA on the FP64 pipe."""
    from bayesian_optimisation_b200 import engine as e
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = e.GPEngine(0)
    from bayesian_optimisation_b200 import session as sm
    _ENGINE.set_acquire_path(request.param)
    sm.default_session().set_acquire_path(request.param)       # the PointSelector drop-in goes through the host-buffer session
    yield request.param
    sm.default_session().set_acquire_path("i8")


@pytest.fixture
def eng(acquire_path):
    return _ENGINE


def _consts():
    from bayesian_optimisation_b200 import engine as e
    return e


def assert_var_close(got, want, cond=None, rtol=RTOL):
    """sigma^2 = 1.000101 - k^T K^-1 k is a cancellation against the prior, so both the reference
    (explicit inverse) and this path (Cholesky) carry an ABSOLUTE error ~ cond(K)*eps of the
    prior.  Against the long-double truth the ORACLE's explicit-inverse arithmetic is off by
    0.03 .. 0.3 cond*eps (profiles/r02_parity_errors.json) while the device stays at ~1e-14, so where cond(K) is
    given the slack of the comparison WITH THE ORACLE is max(VAR_FLOOR, 0.5 cond*eps) -- the oracle's own error --
    and the 1e-9 claim of the device is checked separately against the truth (assert_var_close_to_truth)."""
    atol = VAR_FLOOR if cond is None else max(VAR_FLOOR, 0.5 * cond * np.finfo(np.float64).eps)
    diff = np.abs(np.asarray(got) - np.asarray(want))
    import inspect
    who = inspect.stack()[1].function
    record_error(who, "sigma^2: max abs err vs oracle", diff.max(), atol, note=f"max rel err {(diff / np.maximum(np.abs(want), 1e-300)).max():.2e}, min sigma^2 {np.abs(want).min():.2e}"
                 + (f", cond {cond:.2g}" if cond else ""))
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol)


def assert_var_close_to_truth(got_mu, got_var, X, y, P, ell, jitter, prior, who, max_pts=1500):
    """Device posterior against the long-double Cholesky truth (oracle/truth_ld.c) on a strided subsample of the
    candidates: 1e-9 RELATIVE on sigma^2 with no cond-dependent slack (1e-13 absolute only guards sigma^2 -> 0)."""
    from oracle import truth
    step = max(1, len(P) // max_pts)
    sel = np.arange(0, len(P), step)
    mu_t, var_t, _, _ = truth.posterior_truth(X, y, P[sel], ell, jitter, prior)
    dv = np.abs(np.asarray(got_var)[sel] - var_t)
    record_error(who, "sigma^2: max relative err vs long-double truth", (dv / np.abs(var_t)).max(), RTOL,
                 note=f"max abs err {dv.max():.2e}, min sigma^2 {np.abs(var_t).min():.2e}, {len(sel)} candidates")
    np.testing.assert_allclose(np.asarray(got_var)[sel], var_t, rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(np.asarray(got_mu)[sel], mu_t, rtol=RTOL, atol=RTOL * np.abs(mu_t).max())


def assert_grid_equals_explicit(eng, g, x):
    """The same candidates as a grid descriptor and as an explicit array.  FP64 path: bit-identical.  INT8 path: a grid
    entry is the product of d per-axis table factors (csrc/acquire_i8.cuh), an explicit entry the exp of the summed squared
    distance -- the two differ by a few ulp of k_*, far inside the 1e-9 parity bar; the selected index must not move."""
    gm, gs, xm, xs = (t.cpu().numpy() for t in (g.mu, g.sigma, x.mu, x.sigma))
    if eng.acquire_path == "fp64":
        np.testing.assert_array_equal(xm, gm)
        np.testing.assert_array_equal(xs, gs)
    else:
        np.testing.assert_allclose(xm, gm, rtol=0, atol=1e-11 * np.abs(gm).max())
        np.testing.assert_allclose(xs ** 2, gs ** 2, rtol=1e-10, atol=1e-12)


def cond_of(X, ell, jitter):
    K = o.kernel_rbf_chunked(X, X, ell)
    K[np.diag_indices_from(K)] += jitter
    return float(np.linalg.cond(K))


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("name", golden_names("direct"))
def test_kernel_matrix_matches_reference(eng, name):
    g = load_golden(name)
    e = _consts()
    Kxx = eng.kernel_matrix(g["X"], g["X"], g["ell"], e.JITTER_LML).cpu().numpy()
    np.testing.assert_allclose(Kxx, g["Kxx"], rtol=1e-13, atol=1e-300)
    quirk = g["X"].shape == g["P"].shape
    Kxp = eng.kernel_matrix(g["X"], g["P"], g["ell"], e.JITTER_LML if quirk else 0.0).cpu().numpy()
    np.testing.assert_allclose(Kxp, g["Kxp"], rtol=1e-13, atol=1e-300)


def test_kernel_matrix_odd_shapes(eng):
    rng = np.random.default_rng(5)
    for na, nb, d in [(1, 1, 1), (3, 130, 2), (65, 7, 5), (200, 257, 16)]:
        A, B, ell = rng.random((na, d)), rng.random((nb, d)), 0.2 + rng.random(d)
        K = eng.kernel_matrix(A, B, ell).cpu().numpy()
        np.testing.assert_allclose(K, o.kernel_rbf_chunked(A, B, ell), rtol=1e-13, atol=1e-300)


def test_kernel_function_extremes(eng):
    """exp_nonpos at its edges: coincident points give exactly 1 (+ jitter), huge scaled distances exactly 0
    (results below 2^-1022 are flushed, numpy may return a denormal there), a NaN coordinate gives NaN."""
    A = np.array([[0.0, 0.0], [1.0, 1.0], [0.5, 0.25]])
    K = eng.kernel_matrix(A, A, np.array([0.01, 0.01]), 1e-4).cpu().numpy()
    assert K[0, 0] == 1.0 + 1e-4 and K[1, 1] == 1.0 + 1e-4
    assert K[0, 1] == 0.0 and K[1, 0] == 0.0
    ref = o.kernel_rbf_chunked(A, A, np.array([0.01, 0.01]))
    np.testing.assert_allclose(K - 1e-4 * np.eye(3), ref, rtol=1e-13, atol=1e-300)
    # moderate arguments over the whole useful range: within 2 ulp of numpy
    rng = np.random.default_rng(11)
    P, Q = rng.random((300, 1)) * 37.0, np.zeros((1, 1))
    got = eng.kernel_matrix(P, Q, np.array([1.0])).cpu().numpy()[:, 0]
    want = np.exp(-0.5 * P[:, 0] ** 2)
    assert np.all(np.abs(got - want) <= 2.0 * np.spacing(want))
    B = A.copy(); B[2, 0] = np.nan
    Kn = eng.kernel_matrix(B, A, np.array([0.3, 0.3])).cpu().numpy()
    assert np.isnan(Kn[2]).all() and not np.isnan(Kn[:2]).any()


def test_single_large_lml_uses_the_pipelined_fit_and_matches_oracle(eng):
    n, d = 2100, 5
    X, y, _ = o.synthetic_problem(n, d, seed=3)
    ells = np.array([[0.35, 0.4, 0.3, 0.45, 0.5]])
    got = eng.nlml_batched(X, y, ells).cpu().numpy()
    ref = o.nlml(X, y, ells[0], stable=True)
    assert abs(got[0] - ref) <= RTOL * abs(ref)
    both = eng.nlml_batched(X, y, np.repeat(ells, 2, axis=0)).cpu().numpy()     # batched driver on the same system
    np.testing.assert_allclose(both, [ref, ref], rtol=RTOL)
    assert both[0] == both[1]


# ------------------------------------------------------------------ K2 / fit
@pytest.mark.parametrize("n,d", [(64, 3), (256, 4), (1024, 6)])
def test_cholesky_matches_numpy(eng, n, d):
    import torch
    X, y, ell = o.synthetic_problem(n, d, seed=n)
    K = o.kernel_rbf(X, X, ell)
    a = torch.from_numpy(K).cuda()
    logdet, info = eng.cholesky(a)
    assert info == 0
    L = np.tril(a.cpu().numpy())
    Lref = np.linalg.cholesky(K)
    np.testing.assert_allclose(L, Lref, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(logdet, np.linalg.slogdet(K)[1], rtol=1e-12)


def test_cholesky_reports_non_positive_definite(eng):
    import torch
    a = torch.eye(128, dtype=torch.float64).cuda()
    a[70, 70] = -1.0
    _, info = eng.cholesky(a)
    assert info == 71
    X = np.zeros((3, 2))
    with pytest.raises(np.linalg.LinAlgError):
        eng.fit(X, np.ones(3), np.ones(2), jitter=-2.0)


# 1024 / 1300 / 2100 / 4500: interleaved right-looking triangular inverse with 4 / 6 / 9 / 18 panels (the last one
# beyond 4096 rows with a panel count that is not a power of two, so not the recursive-doubling schedule)
@pytest.mark.parametrize("n,d", [(5, 2), (200, 3), (300, 4), (1024, 6), (1300, 5), (2100, 4), (4500, 6)])
def test_fit_state_matches_numpy(eng, n, d):
    e = _consts()
    X, y, ell = o.synthetic_problem(n, d, seed=7 + n)
    fit = eng.fit(X, y, ell, e.JITTER_LML)
    K = o.kernel_rbf(X, X, ell)
    Lref = np.linalg.cholesky(K)
    L = np.tril(fit.chol().cpu().numpy())[:n, :n]
    np.testing.assert_allclose(L, Lref, rtol=1e-9, atol=1e-12)
    W = np.tril(fit.linv().cpu().numpy())
    np.testing.assert_allclose(W[:n, :n] @ Lref, np.eye(n), atol=1e-9)
    if fit.n_pad > n:   # identity padding
        np.testing.assert_array_equal(W[n:, n:], np.eye(fit.n_pad - n))
        assert not W[n:, :n].any()
    alpha = fit.alpha().cpu().numpy()
    ref_alpha = np.linalg.inv(K) @ y
    np.testing.assert_allclose(alpha, ref_alpha, rtol=1e-7, atol=1e-9 * np.abs(ref_alpha).max())
    ref = o.nlml(X, y, ell, stable=True)
    assert abs(fit.nlml - ref) <= RTOL * abs(ref)
    np.testing.assert_allclose(fit.logdet, np.linalg.slogdet(K)[1], rtol=1e-11)


# ------------------------------------------------------------------ K4
@pytest.mark.parametrize("name", golden_names("direct"))
def test_acquire_explicit_matches_reference(eng, name):
    g = load_golden(name)
    e = _consts()
    X, y, P, ell = g["X"], g["y"], g["P"], g["ell"]
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    quirk = X.shape == P.shape
    res = eng.acquire(fit, P, outputs=True, cross_jitter=e.JITTER_LML if quirk else 0.0)
    mu, sig, acq = res.mu.cpu().numpy(), res.sigma.cpu().numpy(), res.acq.cpu().numpy()
    np.testing.assert_allclose(mu, g["mean_func"], rtol=RTOL, atol=RTOL * np.abs(g["mean_func"]).max())
    assert_var_close(sig ** 2, g["cov_func"] ** 2)
    np.testing.assert_array_equal(acq, 4 * sig - mu)              # two roundings, exactly numpy's
    assert res.best_index == int(g["index"][0])
    assert res.best_score == acq[res.best_index]


@pytest.mark.parametrize("n,d,G,chunk", [(1024, 6, 5, 4096), (300, 3, 21, 1000), (700, 8, 3, 8192)])
def test_acquire_grid_matches_oracle(eng, n, d, G, chunk):
    from bayesian_optimisation_b200.engine import CandidateGrid
    e = _consts()
    X, y, ell = o.synthetic_problem(n, d, seed=n + d)
    axes = [np.linspace(0, 1, G + (k % 2)) for k in range(d)]      # ragged axis lengths
    P = o.candidate_grid(axes)
    mu_ref, var_ref = o.posterior_diag(X, y, P, ell, return_var=True)
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    res = eng.acquire(fit, CandidateGrid(axes), outputs=True, chunk=chunk)
    mu, sig = res.mu.cpu().numpy(), res.sigma.cpu().numpy()
    cond = cond_of(X, ell, e.JITTER_POSTERIOR)
    print(f"cond(K) = {cond:.3g}")
    np.testing.assert_allclose(mu, mu_ref, rtol=RTOL, atol=RTOL * np.abs(mu_ref).max())
    assert_var_close(sig ** 2, var_ref, cond)
    assert_var_close_to_truth(mu, sig ** 2, X, y, P, ell, e.JITTER_POSTERIOR, e.PRIOR_DIAG, f"grid n={n} d={d}")
    acq_ref = o.lcb(mu_ref, np.sqrt(np.abs(var_ref)))
    assert res.best_index == int(o.first_argmax(acq_ref)[0])
    # the same sweep on an explicit copy of the grid
    res2 = eng.acquire(fit, P, outputs=True, chunk=chunk)
    assert_grid_equals_explicit(eng, res, res2)
    assert res2.best_index == res.best_index


def test_expected_improvement_matches_oracle(eng):
    from bayesian_optimisation_b200.engine import ACQ_EI, CandidateGrid
    e = _consts()
    X, y, ell = o.synthetic_problem(512, 4, seed=11)
    axes = [np.linspace(0, 1, 9)] * 4
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    fb = float(y.min())
    res = eng.acquire(fit, CandidateGrid(axes), kind=ACQ_EI, f_best=fb, outputs=True)
    mu, sig = res.mu.cpu().numpy(), res.sigma.cpu().numpy()
    ei_ref = o.expected_improvement(mu, sig, fb)             # oracle formula on the device mu/sigma
    np.testing.assert_allclose(res.acq.cpu().numpy(), ei_ref, rtol=1e-9, atol=1e-300)
    mu_ref, var_ref = o.posterior_diag(X, y, o.candidate_grid(axes), ell, return_var=True)
    ei_full = o.expected_improvement(mu_ref, np.sqrt(np.abs(var_ref)), fb)
    assert res.best_index == int(np.flatnonzero(ei_full == ei_full.max())[0])


def test_the_two_tensor_paths_agree(eng):
    """INT8 digit-slice product vs FP64 DMMA product of the same fit: sigma^2 within 1e-12 absolute
    (both are within rounding of the exact product), identical winner."""
    from bayesian_optimisation_b200.engine import CandidateGrid
    e = _consts()
    X, y, ell = o.synthetic_problem(1500, 7, seed=9)
    grid = CandidateGrid([np.linspace(0, 1, 4)] * 7)
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    out = {}
    for path in ("fp64", "i8"):
        eng.set_acquire_path(path)
        assert eng.acquire_path == path
        out[path] = eng.acquire(fit, grid, outputs=True)
    np.testing.assert_allclose(out["i8"].sigma.cpu().numpy() ** 2, out["fp64"].sigma.cpu().numpy() ** 2, rtol=0, atol=1e-12)
    np.testing.assert_allclose(out["i8"].mu.cpu().numpy(), out["fp64"].mu.cpu().numpy(), rtol=1e-12, atol=1e-12)
    assert out["i8"].best_index == out["fp64"].best_index


@pytest.mark.parametrize("d,G", [(8, 4), (7, 4), (6, 5), (5, 6)])
def test_shared_prefix_of_grid_ordered_candidates_changes_nothing(eng, d, G):
    """The panel kernel computes the squared-distance part of coordinates shared by a whole 64-candidate tile once
    per row (grid-ordered candidates share all but the last two to four axes).  Same operations, same order: every
    candidate gets bit-identical mu and sigma whether the block is grid-ordered (shared prefix found) or randomly
    permuted (no shared coordinates).  The same candidates as a grid DESCRIPTOR are bit-identical on the FP64 path and
    within a few ulp of k_* on the INT8 path (per-axis factor tables, assert_grid_equals_explicit)."""
    from bayesian_optimisation_b200.engine import CandidateGrid
    e = _consts()
    X, y, ell = o.synthetic_problem(700, d, seed=21 + d)
    axes = [np.linspace(0, 1, G)] * d
    P = o.candidate_grid(axes)[: 3 * 64 * 7 + 13]
    perm = np.random.default_rng(d).permutation(len(P))
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    a = eng.acquire(fit, P, outputs=True)
    b = eng.acquire(fit, np.ascontiguousarray(P[perm]), outputs=True)
    c = eng.acquire(fit, CandidateGrid(axes), 0, len(P), outputs=True)
    for name in ("mu", "sigma", "acq"):
        ordered = getattr(a, name).cpu().numpy()
        np.testing.assert_array_equal(getattr(b, name).cpu().numpy(), ordered[perm])
    assert_grid_equals_explicit(eng, c, a)
    assert perm[b.best_index] == a.best_index or a.acq.cpu().numpy()[perm[b.best_index]] == a.best_score
    fit.close()


def test_sharded_ranges_are_bit_identical_and_pick_the_same_index(eng):
    from bayesian_optimisation_b200.engine import CandidateGrid
    from bayesian_optimisation_b200.sharding import reduce_pairs, shard_range
    e = _consts()
    X, y, ell = o.synthetic_problem(600, 5, seed=3)
    axes = [np.linspace(0, 1, 6)] * 5                           # 7776 candidates
    grid = CandidateGrid(axes)
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    full = eng.acquire(fit, grid, outputs=True, chunk=2048)
    for world in (2, 3, 8):
        parts, pairs = [], []
        for r in range(world):
            b, en = shard_range(grid.size, r, world)
            res = eng.acquire(fit, grid, b, en, outputs=True, chunk=512)
            parts.append(res.acq.cpu().numpy())
            pairs.append((res.best_score, res.best_index))
        np.testing.assert_array_equal(np.concatenate(parts), full.acq.cpu().numpy())
        assert reduce_pairs(pairs) == (full.best_score, full.best_index)


def test_exact_ties_resolve_to_lowest_flat_index(eng):
    e = _consts()
    X = np.array([[0.0, 0.0], [1.0, 1.0]])
    y = np.array([1.0, 2.0])
    P = np.full((500, 2), 1e3) + np.arange(500)[:, None]        # k_* underflows to exactly 0 everywhere
    fit = eng.fit(X, y, np.array([0.5, 0.5]), e.JITTER_POSTERIOR)
    res = eng.acquire(fit, P, outputs=True)
    assert np.all(res.mu.cpu().numpy() == 0.0)
    assert np.all(res.sigma.cpu().numpy() == np.sqrt(e.PRIOR_DIAG))
    assert res.best_index == 0
    res = eng.acquire(fit, P, 123, 400)
    assert res.best_index == 123


def test_nan_acquisition_raises_index_error(eng):
    import torch
    mu = torch.tensor([0.0, float("nan"), 1.0], dtype=torch.float64).cuda()
    sig = torch.ones(3, dtype=torch.float64).cuda()
    with pytest.raises(IndexError):
        eng.score_argmax(mu, sig)


# ------------------------------------------------------------------ K3
@pytest.mark.parametrize("n,d", [(2, 1), (7, 2), (21, 2), (64, 5)])
def test_nlml_batched_small_matches_literal_reference_formula(eng, n, d):
    rng = np.random.default_rng(n)
    X, y = rng.random((n, d)) * 10, rng.uniform(1e7, 1e9, n)
    ells = 0.5 + 5 * rng.random((40, d))
    got, grad = eng.nlml_batched(X, y, ells, want_grad=True)
    got, grad = got.cpu().numpy(), grad.cpu().numpy()
    for r in range(len(ells)):
        lit = o.nlml(X, y, ells[r], stable=False)
        if np.isfinite(lit):
            assert abs(got[r] - lit) <= RTOL * abs(lit)
        gref = o.nlml_grad(X, y, ells[r])
        np.testing.assert_allclose(grad[r], gref, rtol=1e-6, atol=1e-6 * np.abs(gref).max())


@pytest.mark.parametrize("n,d,R", [(100, 3, 5), (512, 8, 6), (130, 2, 3), (300, 4, 1), (256, 3, 1), (200, 3, 1), (700, 5, 1)])
def test_nlml_batched_large_matches_oracle(eng, n, d, R):
    rng = np.random.default_rng(n + R)
    X, y, _ = o.synthetic_problem(n, d, seed=n)
    ells = np.exp(rng.uniform(np.log(0.2), np.log(1.0), (R, d)))
    got, grad = eng.nlml_batched(X, y, ells, want_grad=True)
    got, grad = got.cpu().numpy(), grad.cpu().numpy()
    for r in range(R):
        ref = o.nlml(X, y, ells[r], stable=True)
        assert abs(got[r] - ref) <= RTOL * abs(ref), (r, got[r], ref)
        gref = o.nlml_grad(X, y, ells[r])
        np.testing.assert_allclose(grad[r], gref, rtol=1e-6, atol=1e-7 * np.abs(gref).max())
    only = eng.nlml_batched(X, y, ells).cpu().numpy()
    np.testing.assert_array_equal(only, got)


# ------------------------------------------------------------------ drop-in class on the reference's own cases
def _length_scales(g):
    return np.array([g["ls0"], g["ls1"]]) if "ls1" in g else g["ls0"]


@pytest.mark.parametrize("name", golden_names("native"))
def test_point_selector_dropin_matches_reference(name):
    from bayesian_optimisation_b200.point_selector import PointSelector
    g = load_golden(name)
    ps = PointSelector()
    ps.name, ps.iteration = "golden", 1
    ps.measured_pts, ps.measured_vals = g["X"].copy(), g["y"].copy()
    ps.feature_domain = list(g["feature_domain"])
    ps.predicted_pts = g["P"].copy()
    ps.length_scales = _length_scales(g)
    ps.update_surrogate()
    idx = ps.lower_confidence_bound()
    assert isinstance(ps.measured_pts, list) and isinstance(ps.measured_vals, list)     # point_selector.py:101-102
    assert np.asarray(ps.kernel_params).shape == g["kernel_params"].shape
    np.testing.assert_array_equal(ps.kernel_params, g["kernel_params"])
    scale = np.abs(g["mean_func"]).max()
    np.testing.assert_allclose(ps.mean_func, g["mean_func"], rtol=RTOL, atol=RTOL * scale)
    assert_var_close(ps.cov_func ** 2, g["cov_func"] ** 2)
    np.testing.assert_array_equal(idx, g["index"])
    assert idx.dtype == np.int64 and ps.acq_func_eval.shape == g["acq"].shape
    np.testing.assert_array_equal(ps.acq_func_eval, 4 * ps.cov_func - ps.mean_func)


def test_point_selector_inspection_matrices():
    from bayesian_optimisation_b200.point_selector import PointSelector
    g = load_golden("native1d_tr_m8")
    ps = PointSelector()
    ps.measured_pts, ps.measured_vals = g["X"], g["y"]
    ps.feature_domain, ps.predicted_pts, ps.length_scales = [50], g["P"], g["ls0"]
    ps.update_surrogate()
    ell = np.asarray(ps.kernel_params).reshape(-1)
    np.testing.assert_allclose(ps.cov_meas, o.kernel_rbf(g["X"], g["X"], ell) + 1e-6 * np.eye(8), rtol=1e-13)
    np.testing.assert_allclose(ps.cov_meas_pred, o.kernel_rbf(g["X"], g["P"], ell).T, rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(ps.cov_pred, o.kernel_rbf(g["P"], g["P"], ell) + 1e-6 * np.eye(50), rtol=1e-13, atol=1e-300)


def test_point_selector_dropin_replays_the_reference_closed_loop():
    """all PointSelector calls made by the unmodified select_parameters.py while the
    unmodified terminate_opto/block/algo.py scripts drive two full algorithm iterations (trace
    recorded by oracle/make_closed_loop.py).  The drop-in must choose the same length scales and the
    same next point at every step -- otherwise the workflow would diverge from the reference."""
    from conftest import load_closed_loop
    from bayesian_optimisation_b200.point_selector import PointSelector
    calls = load_closed_loop()
    for n, c in enumerate(calls):
        ps = PointSelector()
        ps.name, ps.iteration = "trace", n
        ps.measured_pts, ps.measured_vals = c["X"].copy(), c["y"].copy()
        ps.feature_domain, ps.predicted_pts, ps.length_scales = list(c["feature_domain"]), c["P"], c["length_scales"]
        ps.update_surrogate()
        idx = ps.lower_confidence_bound()
        assert np.asarray(ps.kernel_params).shape == c["kernel_params"].shape, n
        np.testing.assert_array_equal(ps.kernel_params, c["kernel_params"], err_msg=f"call {n}")
        np.testing.assert_array_equal(idx, c["index"], err_msg=f"call {n}")
        assert abs(ps.acq_func_eval.max() - c["acq_max"]) <= 1e-9 * max(1.0, abs(c["acq_max"])), n
        assert abs(ps.mean_func.min() - c["mu_min"]) <= 1e-9 * max(1.0, abs(c["mu_min"])), n
        # grid given as axes (never materialised) selects the same point
        if n % 10 == 0:
            ps2 = PointSelector()
            ps2.measured_pts, ps2.measured_vals = c["X"].copy(), c["y"].copy()
            ps2.feature_domain, ps2.predicted_axes, ps2.length_scales = list(c["feature_domain"]), c["axes"], c["length_scales"]
            ps2.update_surrogate()
            np.testing.assert_array_equal(ps2.lower_confidence_bound(), c["index"])


# ------------------------------------------------------------------ BASELINE.json shapes
def test_baseline_config_n4096_d8_slice_matches_oracle_and_shards(eng):
    """configs[2] shapes (N=4096, d=8, 10^8-point grid): a slice against the numpy oracle, plus the
    size-independent properties on a larger slice (8-way sharding == single sweep, EI winner stable)."""
    from bayesian_optimisation_b200.engine import ACQ_EI, CandidateGrid
    from bayesian_optimisation_b200.sharding import reduce_pairs, shard_range
    e = _consts()
    X, y, ell = o.synthetic_problem(4096, 8)
    axes = [np.linspace(0, 1, 10)] * 8
    grid = CandidateGrid(axes)
    assert grid.size == 10 ** 8
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    start = 31_415_926
    res = eng.acquire(fit, grid, start, start + 6000, outputs=True)
    P = o.grid_points(axes, start, start + 6000)
    mu_ref, var_ref = o.posterior_diag(X, y, P, ell, chunk=1024, return_var=True)
    cond = cond_of(X, ell, e.JITTER_POSTERIOR)
    print(f"cond(K) = {cond:.3g}")
    np.testing.assert_allclose(res.mu.cpu().numpy(), mu_ref, rtol=RTOL, atol=RTOL * np.abs(mu_ref).max())
    assert_var_close(res.sigma.cpu().numpy() ** 2, var_ref, cond)
    acq_ref = o.lcb(mu_ref, np.sqrt(np.abs(var_ref)))
    assert res.best_index == start + int(np.flatnonzero(acq_ref == acq_ref.max())[0])
    nl = eng.nlml_batched(X, y, ell.reshape(1, -1)).cpu().numpy()[0]
    ref = o.nlml(X, y, ell, stable=True)
    assert abs(nl - ref) <= RTOL * abs(ref)
    # properties on 2^17 candidates
    fb = float(y.min())
    b0, count = 50_000_000, 1 << 17
    full = eng.acquire(fit, grid, b0, b0 + count, kind=ACQ_EI, f_best=fb, outputs=True)
    parts, pairs = [], []
    for r in range(8):
        b, en = shard_range(count, r, 8)
        part = eng.acquire(fit, grid, b0 + b, b0 + en, kind=ACQ_EI, f_best=fb, outputs=True, chunk=4096)
        parts.append(part.acq.cpu().numpy())
        pairs.append((part.best_score, part.best_index))
    np.testing.assert_array_equal(np.concatenate(parts), full.acq.cpu().numpy())
    assert reduce_pairs(pairs) == (full.best_score, full.best_index)


def test_baseline_config_n16384_d10_properties(eng):
    """configs[4] shapes (N=16384, d=10, 8^10-point grid): the fit must succeed and the two tensor paths,
    grid vs explicit candidates and chunkings must agree (the oracle would need a 16384^3 CPU inverse)."""
    from bayesian_optimisation_b200.engine import CandidateGrid
    e = _consts()
    X, y, ell = o.synthetic_problem(16384, 10)
    axes = [np.linspace(0, 1, 8)] * 10
    grid = CandidateGrid(axes)
    assert grid.size == 8 ** 10
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    assert np.isfinite(fit.nlml)
    start, count = 777_000_000, 3000
    a = eng.acquire(fit, grid, start, start + count, outputs=True, chunk=1024)
    b = eng.acquire(fit, o.grid_points(axes, start, start + count), outputs=True, chunk=4096)
    assert_grid_equals_explicit(eng, a, b)
    assert a.best_index - start == b.best_index
    other = "fp64" if eng.acquire_path == "i8" else "i8"
    eng.set_acquire_path(other)
    c = eng.acquire(fit, grid, start, start + count, outputs=True)
    np.testing.assert_allclose(c.sigma.cpu().numpy() ** 2, a.sigma.cpu().numpy() ** 2, rtol=0, atol=5e-12)
    assert c.best_index == a.best_index
    # against a direct fp64 evaluation with the device factor: sigma^2 = prior - |L^-1 k|^2
    import torch
    Lt = torch.tril(fit.chol()[:16384, :16384])
    Pd = torch.from_numpy(o.grid_points(axes, start, start + 64)).cuda()
    Xd = torch.from_numpy(X).cuda()
    ks = torch.exp(-0.5 * (((Pd[:, None, :] - Xd[None, :, :]) ** 2) / torch.from_numpy(ell ** 2).cuda()).sum(-1))   # (64, N)
    v = torch.linalg.solve_triangular(Lt, ks.T.contiguous(), upper=False)
    var = e.PRIOR_DIAG - (v * v).sum(0)
    np.testing.assert_allclose(a.sigma.cpu().numpy()[:64] ** 2, var.cpu().numpy(), rtol=1e-9, atol=5e-11)


# ------------------------------------------------------------------ edge shapes
@pytest.mark.parametrize("n,d,c", [(1, 1, 1), (1, 2, 63), (2, 16, 65), (3, 5, 64), (257, 3, 129), (40, 7, 1000)])
def test_edge_shapes_match_oracle(eng, n, d, c):
    e = _consts()
    rng = np.random.default_rng(n * 100 + d)
    X, y, P = rng.random((n, d)), rng.standard_normal(n), rng.random((c, d))
    ell = 0.3 + rng.random(d)
    mu_ref, var_ref = o.posterior_diag(X, y, P, ell, return_var=True)
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    quirk = X.shape == P.shape        # the reference's shape-equality jitter on k(X, P) (point_selector.py:173-177)
    res = eng.acquire(fit, P, outputs=True, cross_jitter=e.JITTER_LML if quirk else 0.0)
    np.testing.assert_allclose(res.mu.cpu().numpy(), mu_ref, rtol=RTOL, atol=RTOL * max(1e-300, np.abs(mu_ref).max()))
    assert_var_close(res.sigma.cpu().numpy() ** 2, var_ref, cond_of(X, ell, e.JITTER_POSTERIOR))
    acq = o.lcb(mu_ref, np.sqrt(np.abs(var_ref)))
    assert res.best_index == int(np.flatnonzero(acq == acq.max())[0])


def test_bad_arguments_are_rejected(eng):
    from bayesian_optimisation_b200.engine import CandidateGrid
    from bayesian_optimisation_b200 import BogpError
    e = _consts()
    X, y, ell = o.synthetic_problem(10, 3, seed=1)
    with pytest.raises(ValueError):
        eng.fit(X, y, ell[:2])
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    with pytest.raises(ValueError):
        eng.acquire(fit, np.zeros((5, 2)))
    with pytest.raises(ValueError):
        eng.acquire(fit, CandidateGrid([np.linspace(0, 1, 3)] * 2))
    with pytest.raises(BogpError):
        eng.acquire(fit, np.zeros((5, 3)), 3, 9)          # range beyond the candidate count
    with pytest.raises(ValueError):
        eng.fit(np.zeros((4, 17)), np.zeros(4), np.ones(17))   # more than BOGP_MAX_DIM features


def test_fit_enqueue_is_graph_capturable(eng):
    """bogp_fit_enqueue issues device work only, so the whole fit (~290 launches on two streams with
    look-ahead) can be recorded once into a CUDA graph and replayed; the replay reproduces the eager
    nlml bit for bit."""
    import ctypes as C
    import torch
    from bayesian_optimisation_b200 import _lib
    e = _consts()
    X, y, ell = o.synthetic_problem(1024, 6, seed=4)
    dX, dy, dl = eng.to_device(X), eng.to_device(y), eng.to_device(ell)
    nbytes = eng.lib.bogp_fit_workspace_bytes(1024, 6)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")

    def enqueue():
        eng._sync_stream()
        h = C.c_void_p()
        _lib.check(eng.lib.bogp_fit_enqueue(eng._ctx, dX.data_ptr(), dy.data_ptr(), 1024, 6, dl.data_ptr(), e.JITTER_LML,
                                            ws.data_ptr(), nbytes, C.byref(h)))
        return h

    h = enqueue()
    eager = C.c_double()
    _lib.check(eng.lib.bogp_fit_status(h, C.byref(eager)))
    eng.lib.bogp_fit_destroy(h)
    ref = o.nlml(X, y, ell)
    assert abs(eager.value - ref) <= RTOL * abs(ref)
    g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            hg = enqueue()
    ws.zero_()
    g.replay()
    torch.cuda.synchronize()
    eng._sync_stream()
    replay = C.c_double()
    _lib.check(eng.lib.bogp_fit_status(hg, C.byref(replay)))
    eng.lib.bogp_fit_destroy(hg)
    assert replay.value == eager.value


# ------------------------------------------------------------------ workspace bounds (compute-sanitizer is closed on this pool)
@pytest.mark.parametrize("n,d", [(1, 1), (64, 2), (200, 3), (256, 4), (257, 3), (300, 4), (512, 5), (600, 6), (1100, 3)])
def test_kernels_stay_inside_their_declared_workspaces(eng, n, d):
    """Every entry point gets a workspace of exactly the size it asks for, followed by a guard region
    filled with a pattern; the guard must be intact afterwards."""
    import ctypes as C
    import torch
    from bayesian_optimisation_b200 import _lib
    from bayesian_optimisation_b200.engine import Candidates
    e = _consts()
    GUARD = 1 << 16
    X, y, ell = o.synthetic_problem(n, d, seed=n)
    dX, dy, dl = eng.to_device(X), eng.to_device(y), eng.to_device(ell)
    eng._sync_stream()

    def guarded(nbytes):
        nbytes = (nbytes + 255) // 256 * 256
        buf = torch.full((nbytes + GUARD,), 0xA5, dtype=torch.uint8, device="cuda")
        return buf, nbytes

    def intact(buf, nbytes):
        torch.cuda.synchronize()
        return bool((buf[nbytes:] == 0xA5).all().item())

    # fit
    ws, nb = guarded(eng.lib.bogp_fit_workspace_bytes(n, d))
    h, nl = C.c_void_p(), C.c_double()
    _lib.check(eng.lib.bogp_fit_create(eng._ctx, dX.data_ptr(), dy.data_ptr(), n, d, dl.data_ptr(), e.JITTER_POSTERIOR,
                                       ws.data_ptr(), nb, C.byref(h), C.byref(nl)))
    assert intact(ws, nb), "fit wrote past its workspace"
    ref = o.nlml(X, y, ell) if n > 1 else None
    # acquire (both buffer sets of the two-stream path)
    P = eng.to_device(np.random.default_rng(1).random((700, d)))
    cd = Candidates(); cd.d_points, cd.d_axes, cd.h_axis_len, cd.c_total, cd.cross_jitter = P.data_ptr(), None, None, 700, 0.0
    aws, anb = guarded(eng.lib.bogp_acquire_workspace_bytes(h, 512))
    bs, bi = C.c_double(), C.c_int64()
    _lib.check(eng.lib.bogp_acquire(eng._ctx, h, C.byref(cd), 0, 700, 0, 4.0, 0.0, e.PRIOR_DIAG, None, None, None,
                                    aws.data_ptr(), anb, C.byref(bs), C.byref(bi)))
    assert intact(aws, anb), "acquire wrote past its workspace"
    eng.lib.bogp_fit_destroy(h)
    # batched LML, single and several restarts, with gradient
    for R in (1, 3):
        ells = eng.to_device(np.tile(ell, (R, 1)) * (1 + 0.1 * np.arange(R))[:, None])
        out = torch.empty(R, dtype=torch.float64, device="cuda"); grad = torch.empty((R, d), dtype=torch.float64, device="cuda")
        lws, lnb = guarded(max(256, eng.lib.bogp_nlml_batched_workspace_bytes(n, d, R, 1)))
        _lib.check(eng.lib.bogp_nlml_batched(eng._ctx, dX.data_ptr(), dy.data_ptr(), n, d, ells.data_ptr(), R, e.JITTER_LML,
                                             out.data_ptr(), grad.data_ptr(), lws.data_ptr(), lnb))
        assert intact(lws, lnb), f"nlml_batched (R={R}) wrote past its workspace"
        if ref is not None:
            assert abs(out[0].item() - ref) <= RTOL * abs(ref)


@pytest.mark.parametrize("n,lens,chunk", [(300, (6, 5, 7, 9, 8), 512), (1100, (6,) * 7, 2048), (700, (9, 8), 64), (520, (4,) * 9, 4096)])
def test_round2_sweep_paths_stay_inside_their_declared_workspace(eng, n, lens, chunk):
    """The same guard-region check for the paths added in round 2, each with exactly `bogp_acquire_workspace_bytes(fit, chunk)`
    bytes: grid sweeps with per-axis factor tables (separate kernels and the fused persistent kernel with its ring), and
    screened arg-max-only grid sweeps (mean GEMM in stored or generated mode, max-times bound, survivor lists, exact passes)."""
    import ctypes as C
    import torch
    from bayesian_optimisation_b200 import _lib
    from bayesian_optimisation_b200.engine import Candidates
    if eng.acquire_path != "i8":
        pytest.skip("the round-2 sweep paths belong to the INT8 tensor path")
    e = _consts()
    GUARD = 1 << 16
    d = len(lens)
    X, y, ell = o.synthetic_problem(n, d, seed=n)
    fit = eng.fit(X, y, ell, e.JITTER_POSTERIOR)
    axes = eng.to_device(np.concatenate([np.linspace(0, 1, L) for L in lens]))
    alen = (C.c_int32 * d)(*lens)
    total = int(np.prod(lens))
    cd = Candidates(); cd.d_points, cd.d_axes, cd.h_axis_len, cd.c_total, cd.cross_jitter = None, axes.data_ptr(), alen, total, 0.0
    nb = (eng.lib.bogp_acquire_workspace_bytes(fit._h, chunk) + 255) // 256 * 256
    eng._sync_stream()
    results = []
    try:
        for fused, screening, kind in ((False, False, 0), (True, False, 0), (False, True, 0), (False, True, 1)):
            eng.set_fused(fused); eng.set_screening(screening)
            ws = torch.full((nb + GUARD,), 0xA5, dtype=torch.uint8, device="cuda")
            bs, bi = C.c_double(), C.c_int64()
            _lib.check(eng.lib.bogp_acquire(eng._ctx, fit._h, C.byref(cd), 0, total, kind, 4.0, float(y.min()), e.PRIOR_DIAG, None, None, None,
                                            ws.data_ptr(), nb, C.byref(bs), C.byref(bi)))
            torch.cuda.synchronize()
            assert bool((ws[nb:] == 0xA5).all().item()), f"sweep (fused={fused}, screening={screening}, kind={kind}) wrote past its workspace"
            results.append((kind, bs.value, bi.value))
    finally:
        eng.set_fused(False); eng.set_screening(True)
    assert results[0] == results[1] == results[2]              # separate == fused == screened, bitwise
    fit.close()
