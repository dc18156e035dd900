"""CPU restatements of the host-checkable arithmetic behind the round-2 grid machinery (no GPU):

  * per-axis kernel-factor tables (csrc/acquire_i8.cuh): a grid entry is the ORDERED product of d correctly rounded factors
    exp(-0.5 (p_k - x_k)^2 / ell_k^2) instead of the exp of the summed squared distance -- (d + 4 + |ln k|) ulp apart
    relatively, a few ulp of 1 absolutely, far inside the 1e-9 parity bar of point_selector.py:166-195;
  * the rounding slack of the GEMM-based posterior mean of the screen (csrc/screen_gemm.cu): any summation order and any
    association of the factors stays within eps = (2 n_pad + 16) 2^-53 1.0002 |alpha|_1 of the ordered evaluation;
  * the golden-ratio seed sample (csrc/common.cuh seed_index): inside the range, and -- unlike a plain stride -- not stuck
    on a face of the 8^10 grid."""
import numpy as np

from oracle import gp_oracle as o


def ordered_product_kstar(X, P, ell):
    """k_*[j, c] the way the table mode forms it: (((1 f_0) f_1) ... f_{d-1}), every factor rounded once."""
    v = np.ones((len(X), len(P)))
    for k in range(X.shape[1]):
        d = P[None, :, k] - X[:, k:k + 1]
        v = v * np.exp(-0.5 * ((d * d) * (1.0 / ell[k] ** 2)))
    return v


def test_factor_product_is_within_a_few_ulp_of_the_exp_of_the_sum():
    rng = np.random.default_rng(0)
    for d in (2, 6, 8, 10, 16):
        X, P, ell = rng.random((300, d)), rng.random((200, d)), 0.2 + rng.random(d)
        ref = o.kernel_rbf_chunked(X, P, ell)                       # exp of the summed squared distance (the reference's formula)
        got = ordered_product_kstar(X, P, ell)
        rel = np.abs(got - ref) / ref
        # (d + 1) roundings of the product, those of the sum, and the rounding of each exponent argument t, which exp turns
        # into |t| ulp: relative (d + 4 + |ln k|) ulp, i.e. an ABSOLUTE difference of a few ulp of 1 (k <= 1)
        assert np.all(rel <= (d + 4 + np.abs(np.log(ref))) * 2.0 ** -52), (d, rel.max())
        assert np.abs(got - ref).max() <= 8 * 2.0 ** -52


def test_gemm_mean_slack_dominates_reassociation_and_reordering():
    rng = np.random.default_rng(1)
    n, d, G = 1500, 6, 5
    X = rng.random((n, d)); ell = np.full(d, 0.3)
    for scale in (1.0, 1e6):                                        # |alpha|_1 from O(1e3) to O(1e9)
        alpha = scale * rng.standard_normal(n)
        axes = [np.linspace(0, 1, G)] * d
        f = [np.exp(-0.5 * ((axes[k][None, :] - X[:, k:k + 1]) ** 2) * (1.0 / ell[k] ** 2)) for k in range(d)]     # (n, G)
        # exact-kernel order: ordered product per entry, then rows ascending
        P = o.candidate_grid(axes)
        dig = np.stack(np.unravel_index(np.arange(len(P)), (G,) * d), axis=1)
        k_ord = np.ones((n, len(P)))
        for k in range(d):
            k_ord = k_ord * f[k][:, dig[:, k]]
        mu_exact = np.zeros(len(P))
        for j in range(n):                                          # ascending rows, one rounding per term and per addition
            mu_exact = mu_exact + alpha[j] * k_ord[j]
        # GEMM order: G = product of the leading half, F = alpha * product of the trailing half, BLAS summation
        lead = np.ones((n, G ** 3)); trail = np.ones((n, G ** 3))
        dl = np.stack(np.unravel_index(np.arange(G ** 3), (G,) * 3), axis=1)
        for k in range(3):
            lead = lead * f[k][:, dl[:, k]]
            trail = trail * f[3 + k][:, dl[:, k]]
        mu_gemm = (lead.T @ (alpha[:, None] * trail)).reshape(-1)
        n_pad = (n + 255) // 256 * 256
        eps = (2 * n_pad + 16) * 2.0 ** -53 * 1.0002 * np.abs(alpha).sum()
        err = np.abs(mu_gemm - mu_exact).max()
        assert err <= eps, (scale, err, eps)
        assert err <= 0.05 * eps                                    # the bound is worst case; real sums are random walks


def seed_index(i, begin, total):
    return begin + ((((i + 1) * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)) * total >> 64)


def test_golden_ratio_seed_sample_covers_the_grid():
    total = 8 ** 10
    idx = np.array([seed_index(i, 0, total) for i in range(4096)], dtype=np.int64)
    assert idx.min() >= 0 and idx.max() < total and len(np.unique(idx)) == 4096
    digits = np.stack(np.unravel_index(idx, (8,) * 10), axis=1)
    for k in range(10):                                             # every axis sees all of its 8 grid points, roughly evenly
        counts = np.bincount(digits[:, k], minlength=8)
        assert counts.min() > 4096 / 8 * 0.8, (k, counts)
    # the plain stride that it replaced: total / 4096 = 8^6, the six trailing digits of every seed are 0
    plain = np.arange(4096, dtype=np.int64) * (total // 4096)
    assert np.all(np.stack(np.unravel_index(plain, (8,) * 10), axis=1)[:, 4:] == 0)
    # a shard: indices stay inside [begin, begin + total)
    sh = np.array([seed_index(i, 5 * 2 ** 27, 2 ** 27) for i in range(4096)])
    assert sh.min() >= 5 * 2 ** 27 and sh.max() < 6 * 2 ** 27
