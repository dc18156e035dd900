"""Arg-max-only sweeps are screened by the posterior-mean bound (csrc/acquire.cu, screen_kernel): the (score, index)
they return must be EXACTLY the one of the unscreened sweep -- on the BASELINE shapes, on explicit candidates, on
sub-ranges, on exact ties (lowest flat index), on flat landscapes where nothing can be dropped, and NaNs must still
raise.  The bound itself is also checked on the host: U(c) >= exact score for every candidate."""
import numpy as np
import pytest

from conftest import record_error
from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu
_ENG = None


@pytest.fixture
def eng():
    from bayesian_optimisation_b200.engine import GPEngine
    global _ENG
    if _ENG is None:
        _ENG = GPEngine(0)
    _ENG.set_acquire_path("i8")
    _ENG.set_screening(True)
    yield _ENG
    _ENG.set_screening(True)


def _both(eng, fit, cand, b, e, **kw):
    eng.set_screening(False)
    full = eng.acquire(fit, cand, b, e, **kw)
    eng.set_screening(True)
    eng.screen_stats()
    scr = eng.acquire(fit, cand, b, e, **kw)
    screened, survived = eng.screen_stats()
    return full, scr, screened, survived


@pytest.mark.parametrize("kind", ["lcb", "ei"])
@pytest.mark.parametrize("n,d,G", [(1024, 6, 10), (4096, 8, 10)])
def test_screened_sweep_returns_the_exact_winner_on_baseline_shapes(eng, n, d, G, kind):
    from bayesian_optimisation_b200.engine import ACQ_EI, ACQ_LCB, CandidateGrid, JITTER_POSTERIOR
    X, y, ell = o.synthetic_problem(n, d)
    grid = CandidateGrid([np.linspace(0, 1, G)] * d)
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    kw = dict(kind=ACQ_EI, f_best=float(y.min())) if kind == "ei" else dict(kind=ACQ_LCB, explore=4.0)
    b, e = (0, grid.size) if n == 1024 else (37_000_000, 37_000_000 + (1 << 19))
    full, scr, screened, survived = _both(eng, fit, grid, b, e, **kw)
    assert (scr.best_score, scr.best_index) == (full.best_score, full.best_index)
    assert screened == e - b
    record_error(f"screen N={n} d={d} {kind}", "survivor fraction of the screen", survived / screened, note=f"{survived} of {screened}")
    assert survived < screened
    fit.close()


def test_screened_sweep_on_explicit_candidates_ties_and_flat_landscapes(eng):
    from bayesian_optimisation_b200.engine import ACQ_EI, JITTER_POSTERIOR
    X, y, ell = o.synthetic_problem(600, 4, seed=2)
    rng = np.random.default_rng(3)
    P = rng.random((70_001, 4))
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    full, scr, screened, survived = _both(eng, fit, P, 0, len(P))
    assert (scr.best_score, scr.best_index) == (full.best_score, full.best_index)
    # exact tie between two far-apart copies of the winner: the lower flat index must win, screened or not
    P2 = P.copy()
    w = full.best_index
    lo, hi = (5, 69_000) if w not in (5, 69_000) else (6, 68_000)
    P2[lo] = P[w]; P2[hi] = P[w]
    if w < lo:
        lo = w
    full2, scr2, _, _ = _both(eng, fit, P2, 0, len(P2))
    assert full2.best_index == min(lo, w) and (scr2.best_score, scr2.best_index) == (full2.best_score, full2.best_index)
    # a sub-range that excludes the global winner
    full3, scr3, _, _ = _both(eng, fit, P, 20_000, 60_000, kind=ACQ_EI, f_best=float(y.min()))
    assert (scr3.best_score, scr3.best_index) == (full3.best_score, full3.best_index) and 20_000 <= scr3.best_index < 60_000
    # flat landscape: every candidate is the same point, nothing can be dropped, index = first of the range
    Pf = np.repeat(P[:1], 20_000, axis=0)
    full4, scr4, screened, survived = _both(eng, fit, Pf, 100, 20_000)
    assert scr4.best_index == full4.best_index == 100 and survived == screened
    fit.close()


def test_screened_sweep_still_raises_on_nan(eng):
    from bayesian_optimisation_b200.engine import JITTER_POSTERIOR
    X, y, ell = o.synthetic_problem(300, 3, seed=4)
    P = np.random.default_rng(5).random((40_000, 3))
    P[31_337, 1] = np.nan
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    with pytest.raises(IndexError):
        eng.acquire(fit, P)
    fit.close()


def test_the_bound_dominates_the_exact_score(eng):
    """U(c) = A(mu_c, sqrt(prior)) >= A(mu_c, sigma_c) for every candidate, for the device's own mu / sigma: LCB without
    any slack, EI with the 1e-12 (|f_best - mu| + sigma_max) slack of the kernel -- and how much of that slack is used."""
    from bayesian_optimisation_b200.engine import ACQ_EI, CandidateGrid, JITTER_POSTERIOR, PRIOR_DIAG
    X, y, ell = o.synthetic_problem(800, 5, seed=6)
    grid = CandidateGrid([np.linspace(0, 1, 9)] * 5)
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    fb = float(y.min())
    lcb = eng.acquire(fit, grid, outputs=True)
    ei = eng.acquire(fit, grid, kind=ACQ_EI, f_best=fb, outputs=True)
    mu, sig = lcb.mu.cpu().numpy(), lcb.sigma.cpu().numpy()
    smax = np.sqrt(PRIOR_DIAG)
    assert sig.max() <= smax
    assert np.all(4.0 * smax - mu >= lcb.acq.cpu().numpy())
    bound = o.expected_improvement(mu, np.full_like(mu, smax), fb)
    gap = ei.acq.cpu().numpy() - bound
    record_error("screen bound EI", "max (exact EI - bound) / slack", (gap / (1e-12 * (np.abs(fb - mu) + smax))).max())
    assert np.all(gap <= 1e-12 * (np.abs(fb - mu) + smax))
    fit.close()


# ------------------------------------------------------------------ grid sweeps: means from GEMMs (csrc/screen_gemm.cu)
@pytest.mark.parametrize("n,lens,rng_", [
    (700, (9, 7, 5, 9, 11), (123, 30000)),           # odd number of trailing settings (padded row stride), range cut inside rows
    (1200, (6,) * 6, (0, 6 ** 6)),                # balanced 3 + 3 split, stored operands
    (300, (13, 50, 40), (777, 25999)),            # long axes: no factor tables, the screen falls back to the mean-only panel pass
    (2100, (4,) * 9, (1000, 260_000)),            # nine axes
])
def test_gemm_screen_returns_the_exact_winner_on_ragged_grids_and_ranges(eng, n, lens, rng_):
    from bayesian_optimisation_b200.engine import ACQ_EI, CandidateGrid, JITTER_POSTERIOR
    d = len(lens)
    X, y, ell = o.synthetic_problem(n, d, seed=n)
    grid = CandidateGrid([np.linspace(0, 1, L) for L in lens])
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    b, e = rng_
    for kw in (dict(), dict(kind=ACQ_EI, f_best=float(y.min()))):
        full, scr, screened, survived = _both(eng, fit, grid, b, e, **kw)
        assert (scr.best_score, scr.best_index) == (full.best_score, full.best_index)
        assert b <= scr.best_index < e and screened == e - b
    fit.close()


def test_gemm_screen_generated_operand_mode_when_the_stored_operands_do_not_fit(eng):
    """A small workspace (chunk = 64 candidates) leaves no room for the two stored operand matrices of the balanced split:
    the GEMM then forms its A tiles from composite table rows (gemm_f64.cuh, A_GEN).  Same winner."""
    from bayesian_optimisation_b200.engine import ACQ_EI, CandidateGrid, JITTER_POSTERIOR
    X, y, ell = o.synthetic_problem(1024, 8, seed=8)
    grid = CandidateGrid([np.linspace(0, 1, 6)] * 8)            # 1.68 M candidates; balanced split 1296 + 1296 rows of 8 KB = 21 MB
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    fb = float(y.min())
    eng.set_screening(False)
    full = eng.acquire(fit, grid, kind=ACQ_EI, f_best=fb)
    eng.set_screening(True)
    eng._acq_ws = None                                          # force the minimal workspace of this chunk size
    scr = eng.acquire(fit, grid, kind=ACQ_EI, f_best=fb, chunk=4096)
    assert (scr.best_score, scr.best_index) == (full.best_score, full.best_index)
    eng._acq_ws = None
    fit.close()


def test_gemm_screen_exact_tie_between_grid_duplicates_keeps_the_lowest_index(eng):
    """An axis with a repeated grid value makes pairs of candidates with bit-identical scores: the lower flat index wins,
    screened or not."""
    from bayesian_optimisation_b200.engine import CandidateGrid, JITTER_POSTERIOR
    X, y, ell = o.synthetic_problem(500, 4, seed=12)
    ax = np.linspace(0, 1, 8)
    dup = np.concatenate([ax, ax[::-1]])                        # 16 values, every one twice
    grid = CandidateGrid([dup, ax, ax, dup])
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    full, scr, screened, survived = _both(eng, fit, grid, 0, grid.size)
    assert (scr.best_score, scr.best_index) == (full.best_score, full.best_index)
    acq = eng.acquire(fit, grid, outputs=True).acq.cpu().numpy()
    assert full.best_index == int(np.flatnonzero(acq == acq.max())[0]) and (acq == acq.max()).sum() >= 4
    fit.close()


def test_gemm_mean_slack_covers_the_rounding_difference(eng):
    """The screen trusts mu_gemm - eps <= mu_exact with eps = (2 n_pad + 16) 2^-53 1.0002 |alpha|_1.  Host check on the
    device's own numbers: a float64 GEMM of the same factor tables (numpy, another summation order again) against the mu
    of the exact kernels stays far inside eps."""
    from bayesian_optimisation_b200.engine import CandidateGrid, JITTER_POSTERIOR
    n, d, G = 1500, 5, 7
    X, y, ell = o.synthetic_problem(n, d, seed=14)
    axes = [np.linspace(0, 1, G)] * d
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    mu = eng.acquire(fit, CandidateGrid(axes), outputs=True).mu.cpu().numpy()
    alpha = fit.alpha().cpu().numpy()[:n]
    f = [np.exp(-0.5 * (axes[k][None, :] - X[:, k:k + 1]) ** 2 / ell[k] ** 2) for k in range(d)]      # (n, G) per axis
    lead = np.einsum("ja,jb,jc->abcj", f[0], f[1], f[2]).reshape(-1, n)
    trail = np.einsum("j,ja,jb->abj", alpha, f[3], f[4]).reshape(-1, n)
    mu_gemm = (lead @ trail.T).reshape(-1)
    n_pad = fit.n_pad
    eps = (2 * n_pad + 16) * 2.0 ** -53 * 1.0002 * np.abs(alpha).sum()
    err = np.abs(mu_gemm - mu).max()
    record_error("gemm screen", "max |mu_gemm - mu_exact| / eps", err / eps, 1.0, note=f"eps = {eps:.2e}, |alpha|_1 = {np.abs(alpha).sum():.3g}")
    assert err <= eps
    fit.close()


def test_global_seed_shards_pick_the_full_winner(eng):
    """bogp_set_global_seed: every shard of a sharded arg-max seeds its screen with the same strided sample of the WHOLE grid.
    The winner over the shards is the winner of the full sweep (a shard may report a seed candidate outside its range), and
    shards far from the winner no longer score their plateau exactly."""
    from bayesian_optimisation_b200.engine import CandidateGrid, JITTER_POSTERIOR
    from bayesian_optimisation_b200.sharding import reduce_pairs, shard_range
    X, y, ell = o.synthetic_problem(900, 6, seed=31)
    X = 0.25 * X                                               # all measurements in one corner: most of the grid is a flat plateau
    grid = CandidateGrid([np.linspace(0, 1, 8)] * 6)           # 262144 candidates
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    eng.set_screening(False)
    full = eng.acquire(fit, grid)
    eng.set_screening(True)
    tot = {}
    for glob in (False, True):
        eng.set_global_seed(glob)
        eng.screen_stats()
        pairs = []
        for r in range(8):
            b, e = shard_range(grid.size, r, 8)
            res = eng.acquire(fit, grid, b, e)
            pairs.append((res.best_score, res.best_index))
        eng.set_global_seed(False)
        assert reduce_pairs(pairs) == (full.best_score, full.best_index)
        tot[glob] = eng.screen_stats()
    record_error("global seed", "survivors of 8 shards with the global seed / with shard-local seeds", tot[True][1] / max(1, tot[False][1]),
                 note=f"{tot[True][1]} vs {tot[False][1]} of {grid.size}")
    assert tot[True][1] < grid.size // 4            # (which of the two seeds leaves fewer survivors depends on the landscape)
    fit.close()


def test_nearest_measurement_variance_bound_dominates_sigma(eng):
    """gs_kmax_kernel: sigma^2(c) <= prior - max_j k_j(c)^2 / K_jj for the device's own sigma, with the packed-half max-times
    arithmetic restated on the host (fp16-rounded values, (1 - 2^-9) m - 2^-22 as the lower bound of the maximum) -- and how
    much tighter than sqrt(prior) it is for the candidates that matter."""
    from bayesian_optimisation_b200.engine import CandidateGrid, JITTER_POSTERIOR, PRIOR_DIAG
    n, d, G = 1100, 6, 7
    X, y, ell = o.synthetic_problem(n, d, seed=17)
    axes = [np.linspace(0, 1, G)] * d
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    res = eng.acquire(fit, CandidateGrid(axes), outputs=True)
    sig = res.sigma.cpu().numpy()
    P = o.candidate_grid(axes)
    K = o.kernel_rbf_chunked(X, P, ell)                            # (n, C)
    m = np.maximum(0.0, K.astype(np.float16).max(axis=0).astype(np.float64) * (1.0 - 2.0 ** -9) - 2.0 ** -22)
    s_ub = np.sqrt(PRIOR_DIAG - m * m / (1.0 + JITTER_POSTERIOR) + 1e-8)
    assert np.all(sig <= s_ub)
    record_error("kmax bound", "mean sigma_ub / sigma_max (mean sigma / sigma_max)", float(s_ub.mean() / np.sqrt(PRIOR_DIAG)),
                 note=f"{sig.mean() / np.sqrt(PRIOR_DIAG):.3f}")
    # screened LCB sweep with the bound == unscreened, and far fewer survivors than with sigma_max alone
    full, scr, screened, survived = _both(eng, fit, CandidateGrid(axes), 0, len(P))
    assert (scr.best_score, scr.best_index) == (full.best_score, full.best_index)
    lcb = res.acq.cpu().numpy()
    loose = int((4.0 * np.sqrt(PRIOR_DIAG) - res.mu.cpu().numpy() >= lcb.max()).sum())
    record_error("kmax bound", "survivors with the bound / survivors a sigma_max screen against the FINAL best would leave", survived / max(1, loose),
                 note=f"{survived} vs {loose} of {len(P)}")
    fit.close()
