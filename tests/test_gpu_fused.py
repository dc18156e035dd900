"""The fused persistent sweep kernel (csrc/acquire_fused.cu: grid index -> k_* digits -> tcgen05 product -> sigma^2, mu ->
acquisition -> max-loc in ONE launch) against the separate panel / product / finalize / merge kernels of the same INT8
path.  Digits, integer level sums and every reduction order are the same, so the comparison is ==, not a tolerance; the
separate kernels (the default) are in turn checked against the oracle and the reference's golden vectors in
tests/test_gpu_parity.py.  Replaces point_selector.py:81,90-98,204-207."""
import numpy as np
import pytest

from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from bayesian_optimisation_b200.engine import GPEngine
    e = GPEngine(0)
    e.set_acquire_path("i8")
    e.set_screening(False)
    yield e
    e.set_fused(False)
    e.set_screening(True)


def both(eng, fit, cands, **kw):
    eng.set_fused(True, kw.pop("group", 0))
    l0 = eng.launches
    a = eng.acquire(fit, cands, **kw)
    fused_launches = eng.launches - l0
    eng.set_fused(False)
    b = eng.acquire(fit, cands, **kw)
    return a, b, fused_launches


def same(a, b, outputs=True):
    assert a.best_index == b.best_index
    assert a.best_score == b.best_score
    if outputs:
        for x, y in ((a.mu, b.mu), (a.sigma, b.sigma), (a.acq, b.acq)):
            np.testing.assert_array_equal(x.cpu().numpy(), y.cpu().numpy())


@pytest.mark.parametrize("n,d,G,group", [(300, 3, 21, 0), (1024, 6, 5, 0), (700, 8, 3, 1), (1500, 4, 9, 3), (64, 2, 50, 0), (5, 1, 50, 0)])
def test_fused_grid_sweep_is_bit_identical_to_the_separate_kernels(eng, n, d, G, group):
    from bayesian_optimisation_b200.engine import CandidateGrid, JITTER_POSTERIOR, ACQ_EI
    X, y, ell = o.synthetic_problem(n, d, seed=n + d)
    axes = [np.linspace(0, 1, G + (k % 2)) for k in range(d)]
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    a, b, nl = both(eng, fit, CandidateGrid(axes), outputs=True, chunk=2048, group=group)
    assert nl <= 2, f"a fused grid sweep is the table kernel + ONE sweep kernel, got {nl} launches"
    same(a, b)
    a, b, _ = both(eng, fit, CandidateGrid(axes), outputs=True, kind=ACQ_EI, f_best=float(y.min()), chunk=4096, group=group)
    same(a, b)
    a, b, _ = both(eng, fit, CandidateGrid(axes), outputs=False, chunk=1024, group=group)      # arg-max only
    same(a, b, outputs=False)
    fit.close()


def test_fused_explicit_ranges_ragged_tail_and_cross_jitter(eng):
    from bayesian_optimisation_b200.engine import JITTER_POSTERIOR
    rng = np.random.default_rng(5)
    X, y, ell = o.synthetic_problem(900, 5, seed=3)
    P = rng.random((10007, 5))                       # not a multiple of the 64-candidate tile
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    a, b, nl = both(eng, fit, P, outputs=True)
    assert nl == 1, f"a fused sweep over explicit candidates is one kernel launch, got {nl}"
    same(a, b)
    a, b, _ = both(eng, fit, P, c_begin=1234, c_end=7777, outputs=True)      # a shard of the range
    same(a, b)
    # the reference's shape-equality quirk: M == C puts the kernel jitter on the diagonal of K(X, P) (point_selector.py:192-194)
    Xq, yq, ellq = o.synthetic_problem(320, 3, seed=9)
    Pq = rng.random((320, 3))
    fq = eng.fit(Xq, yq, ellq, JITTER_POSTERIOR)
    a, b, _ = both(eng, fq, Pq, outputs=True, cross_jitter=1e-4)
    same(a, b)
    fit.close(); fq.close()


def test_fused_large_system_signed_digits_and_many_tiles(eng):
    """n_pad > 8192 takes the balanced signed panel digits; 2^17 candidates = 2048 tiles cycle the ring many times."""
    from bayesian_optimisation_b200.engine import CandidateGrid, JITTER_POSTERIOR, ACQ_EI
    X, y, ell = o.synthetic_problem(4096, 8)
    grid = CandidateGrid([np.linspace(0, 1, 10)] * 8)
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    a, b, nl = both(eng, fit, grid, c_begin=777, c_end=777 + (1 << 17), kind=ACQ_EI, f_best=float(y.min()), outputs=True, chunk=32768)
    assert nl <= 2
    same(a, b)
    fit.close()
    X, y, ell = o.synthetic_problem(8300, 10, seed=2)
    grid = CandidateGrid([np.linspace(0, 1, 8)] * 10)
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    a, b, _ = both(eng, fit, grid, c_begin=10 ** 6, c_end=10 ** 6 + 5000, outputs=True)
    same(a, b)
    fit.close()


def test_fused_nan_score_raises_like_the_reference(eng):
    from bayesian_optimisation_b200.engine import JITTER_POSTERIOR
    X, y, ell = o.synthetic_problem(100, 2, seed=1)
    P = np.random.default_rng(0).random((500, 2))
    P[77, 1] = np.nan
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    eng.set_fused(True)
    with pytest.raises(IndexError):
        eng.acquire(fit, P)
    eng.set_fused(False)
    fit.close()
