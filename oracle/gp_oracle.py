"""CPU oracle for the GP surrogate + acquisition hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy, what the reference class
`PointSelector` (`/root/reference/point_selector.py:13-207`) computes.  It is the
checker for the CUDA path -- never the thing shipped or measured (the one
exception, sanctioned by the task contract, is `bench.py`'s `cpu_baseline` /
`--impl reference` leg, which *times* it as the CPU baseline).

Pinning status
--------------
The reference ships no tests, fixtures or golden vectors, so the
oracle is pinned by *executing the reference itself* in the build container:
`oracle/make_golden.py` imports the reference class from `/root/reference`
(plotting modules stubbed, `oracle/reference_loader.py`), runs it on seeded
inputs and stores inputs + outputs under `tests/golden/`.  `tests/test_oracle.py`
checks every function below against those vectors.  The two extensions the
reference does not implement -- expected improvement and the LML gradient -- have
no reference vectors: **parity unpinned** for those two; they are checked
against scipy and against central finite differences instead.

Where the literal reference code cannot run at the benchmark sizes the oracle
uses an algebraically identical restatement and the golden tests prove the two
agree wherever the literal code can run:
  * `np.log(np.linalg.det(K))` underflows to -inf once det K < 2^-1074: with the 1e-4
    jitter from about M = 80-90 points on the reference's own grids, earlier for long
    length scales (`point_selector.py:118`) -> `nlml(..., stable=True)` uses `slogdet`;
  * the full C x C prior/posterior covariance (`point_selector.py:78,91`) of which
    only the diagonal is used (`:98`) -> `posterior_diag` computes the diagonal
    in chunks.  Like the reference it uses an explicit `np.linalg.inv`.
"""
from __future__ import annotations

import numpy as np

JITTER_KERNEL = 1e-4   # point_selector.py:193  (added iff x1.shape == x2.shape)
JITTER_EXTRA = 1e-6    # point_selector.py:78-79
PRIOR_DIAG = 1.0 + JITTER_KERNEL + JITTER_EXTRA   # diag of cov_pred, point_selector.py:78


# --------------------------------------------------------------------------------------
# kernel                                                         point_selector.py:166-195
# --------------------------------------------------------------------------------------
def kernel_rbf(x1, x2, ell):
    """ARD squared-exponential Gram matrix; follows point_selector.py:166-195.

    Same evaluation order as the reference: subtract, square, divide by ell**2,
    sum over the feature axis, times -0.5, exp (`:187-189`).  `+1e-4*I` is added iff
    the two inputs have the same *shape* (`:173-177,191-193`) -- a shape test, not
    an identity test.
    """
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    ell = np.asarray(ell, dtype=np.float64)
    jitter = x1.shape == x2.shape
    a = x1[:, None, :]
    b = x2[None, :, :]
    rbf = np.exp(-0.5 * np.sum((a - b) ** 2 / ell ** 2, axis=2))
    if jitter:
        return rbf + JITTER_KERNEL * np.eye(len(x1))
    return rbf


def kernel_rbf_chunked(x1, x2, ell, out=None, chunk=2048):
    """`kernel_rbf` without the jitter rule, built in row chunks (memory-bounded)."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    ell2 = np.asarray(ell, dtype=np.float64) ** 2
    if out is None:
        out = np.empty((len(x1), len(x2)))
    for s in range(0, len(x1), chunk):
        a = x1[s:s + chunk, None, :]
        out[s:s + chunk] = np.exp(-0.5 * np.sum((a - x2[None, :, :]) ** 2 / ell2, axis=2))
    return out


# --------------------------------------------------------------------------------------
# log marginal likelihood                                        point_selector.py:111-120
# --------------------------------------------------------------------------------------
def nlml(X, y, ell, stable=True):
    """Negative log marginal likelihood, point_selector.py:111-120.

    `K = kernel_rbf(X, X)` (jitter 1e-4 only); `0.5*(y^T K^-1 y + log det K + M log 2pi)`.
    `stable=False` is the literal code (`inv` + `log(det)`); `stable=True` replaces
    `log(det)` by `slogdet` (identical where det is finite, see tests).
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    K = kernel_rbf(X, X, ell) if len(X) <= 4096 else _gram_big(X, ell, JITTER_KERNEL)
    inv = np.linalg.inv(K)
    if stable:
        sign, logdet = np.linalg.slogdet(K)
        if sign <= 0:
            logdet = np.nan
    else:
        with np.errstate(divide="ignore", invalid="ignore"):
            logdet = np.log(np.linalg.det(K))
    return 0.5 * (y.T @ inv @ y + logdet + len(X) * np.log(2 * np.pi))


def _gram_big(X, ell, jitter):
    K = kernel_rbf_chunked(X, X, ell)
    K[np.diag_indices_from(K)] += jitter
    return K


def nlml_grad(X, y, ell):
    """Analytic d nlml / d ell_k (jitter not differentiated).  EXTENSION: the reference
    has no gradient (`gradient_steps`, point_selector.py:33, is never read) -- parity
    unpinned; checked against central finite differences of `nlml` in the tests.

    With alpha = K^-1 y:  d/d ell_k = -0.5 * sum_ij (alpha_i alpha_j - K^-1_ij) k_ij (x_ik-x_jk)^2 / ell_k^3
    where k_ij excludes the jitter.
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    ell = np.asarray(ell, dtype=np.float64)
    K = kernel_rbf(X, X, ell)
    Kn = K - JITTER_KERNEL * np.eye(len(X))
    inv = np.linalg.inv(K)
    alpha = inv @ y
    G = np.outer(alpha, alpha) - inv
    g = np.empty(len(ell))
    for k in range(len(ell)):
        D2 = (X[:, None, k] - X[None, :, k]) ** 2
        g[k] = -0.5 * np.sum(G * Kn * D2) / ell[k] ** 3
    return g


def tune_kernel(X, y, length_scales):
    """Length-scale grid search, point_selector.py:104-163.

    Returns (kernel_params, nlogml_table_float32).  The table is float32
    (`:126,150`) and the winner is the first row-major entry equal to `np.amin`
    (`:141,159`).  Two grids -> 2 features (`len(length_scales) == 2`, `:122`),
    otherwise a single 1-D grid (`:148`).
    """
    if len(length_scales) == 2:
        ax1, ax2 = length_scales[0], length_scales[1]
        table = np.zeros((len(ax1), len(ax2)), dtype=np.float32)
        with np.errstate(all="ignore"):
            for i in range(len(ax1)):
                for j in range(len(ax2)):
                    table[i, j] = nlml(X, y, np.array([ax1[i], ax2[j]]), stable=False)
        idx = np.argwhere(table == np.amin(table))[0]
        return np.array([ax1[idx[0]], ax2[idx[1]]]), table
    table = np.zeros(len(length_scales), dtype=np.float32)
    with np.errstate(all="ignore"):
        for i in range(len(length_scales)):
            table[i] = nlml(X, y, np.array([length_scales[i]]), stable=False)
    idx = np.argwhere(table == np.amin(table))[0]
    return np.array([length_scales[idx]]), table   # shape (1, 1), like the reference (`:161`)


def midpoint_length_scales(length_scales):
    """Single-measurement fallback, point_selector.py:63-73."""
    if len(length_scales) == 2:
        a1, a2 = length_scales[0], length_scales[1]
        return np.array([a1[len(a1) // 2], a2[len(a2) // 2]])
    return np.array([length_scales[len(length_scales) // 2]])


# --------------------------------------------------------------------------------------
# posterior                                                        point_selector.py:78-98
# --------------------------------------------------------------------------------------
def posterior_literal(X, y, P, ell):
    """mu and sigma exactly as point_selector.py:78-98 (builds the C x C matrices)."""
    cov_pred = kernel_rbf(P, P, ell) + JITTER_EXTRA * np.eye(len(P))
    cov_meas = kernel_rbf(X, X, ell) + JITTER_EXTRA * np.eye(len(X))
    cov_meas_pred = kernel_rbf(X, P, ell).T
    inv = np.linalg.inv(cov_meas)
    mu = cov_meas_pred @ (inv @ y)
    cov = cov_pred - cov_meas_pred @ (inv @ cov_meas_pred.T)
    return mu, np.sqrt(np.abs(np.diag(cov)))


def posterior_diag(X, y, P, ell, chunk=4096, return_var=False, c_offset=0):
    """Chunked, diagonal-only restatement of point_selector.py:78-98.

    `inv`-based like the reference.  Prior diagonal = 1 + 1e-4 + 1e-6 (`:78` with the
    jitter rule of `:193`).  Reproduces the M == C quirk (jitter on `kernel_rbf(X,P)`
    when the shapes are equal): the caller says which global candidate offset this
    block starts at via `c_offset` so that the quirk can be placed on the diagonal.
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    P = np.asarray(P, dtype=np.float64)
    M = len(X)
    if M <= 4096:
        K = kernel_rbf(X, X, ell) + JITTER_EXTRA * np.eye(M)
    else:
        K = _gram_big(X, ell, JITTER_KERNEL + JITTER_EXTRA)
    inv = np.linalg.inv(K)
    alpha = inv @ y
    quirk = (X.shape == P.shape) and c_offset == 0
    mu = np.empty(len(P))
    var = np.empty(len(P))
    ell2 = np.asarray(ell, dtype=np.float64) ** 2
    for s in range(0, len(P), chunk):
        Pc = P[s:s + chunk]
        Ks = np.exp(-0.5 * np.sum((Pc[:, None, :] - X[None, :, :]) ** 2 / ell2, axis=2))  # (c, M)
        if quirk:
            r = np.arange(s, min(s + chunk, len(P)))
            Ks[r - s, r] += JITTER_KERNEL
        mu[s:s + chunk] = Ks @ alpha
        var[s:s + chunk] = PRIOR_DIAG - np.einsum("cm,cm->c", Ks @ inv, Ks)
    if return_var:
        return mu, var
    return mu, np.sqrt(np.abs(var))


# --------------------------------------------------------------------------------------
# acquisition                                                    point_selector.py:197-207
# --------------------------------------------------------------------------------------
def lcb(mu, sigma, explore=4):
    """`explore*sigma - mu` (point_selector.py:204); two roundings, no FMA."""
    return explore * np.asarray(sigma) - np.asarray(mu)


def expected_improvement(mu, sigma, f_best):
    """EI for minimisation.  EXTENSION (docs/README.md:364 lists EI as future work):
    parity unpinned by the reference; restated with scipy (a reference dependency,
    time_residuals.py:4).  z=(f_best-mu)/sigma; EI=(f_best-mu)*Phi(z)+sigma*phi(z);
    sigma == 0 -> max(f_best-mu, 0)."""
    from scipy.special import ndtr
    mu = np.asarray(mu, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    imp = f_best - mu
    with np.errstate(divide="ignore", invalid="ignore"):
        z = imp / sigma
        ei = imp * ndtr(z) + sigma * np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi)
    return np.where(sigma > 0, ei, np.maximum(imp, 0.0))


def first_argmax(a, shape=None):
    """`np.argwhere(a == np.amax(a))[0]` (point_selector.py:207): lowest row-major index
    among exact maxima, as a multi-index.  Raises IndexError when a NaN is present
    (amax is NaN, nothing compares equal) -- same as the reference."""
    a = np.asarray(a)
    if shape is not None:
        a = a.reshape(shape)
    return np.argwhere(a == np.amax(a))[0]


# --------------------------------------------------------------------------------------
# candidate grid                                             select_parameters.py:273-279
# --------------------------------------------------------------------------------------
def candidate_grid(axes):
    """Row-major Cartesian product, axis 0 slowest (select_parameters.py:273-279)."""
    mesh = np.meshgrid(*[np.asarray(a, dtype=np.float64) for a in axes], indexing="ij")
    return np.stack([m.reshape(-1) for m in mesh], axis=1)


def grid_points(axes, start, stop):
    """Rows [start, stop) of `candidate_grid(axes)` without materialising the whole grid."""
    sizes = [len(a) for a in axes]
    flat = np.arange(start, stop, dtype=np.int64)
    out = np.empty((len(flat), len(axes)))
    for k in range(len(axes) - 1, -1, -1):
        out[:, k] = np.asarray(axes[k], dtype=np.float64)[flat % sizes[k]]
        flat = flat // sizes[k]
    return out


# --------------------------------------------------------------------------------------
# whole-path restatement of update_surrogate + lower_confidence_bound
# --------------------------------------------------------------------------------------
def select_next(X, y, P, feature_domain, length_scales, explore=4):
    """point_selector.py:42-102 + 197-207 end to end (diag-only posterior).

    Returns dict(kernel_params, mean_func, cov_func, acq, index, nlml_table)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if len(X) > 1:
        ell, table = tune_kernel(X, y, length_scales)
    else:
        ell, table = midpoint_length_scales(length_scales), None
    mu, sigma = posterior_diag(X, y, P, ell)
    mean_func = mu.reshape(feature_domain)
    cov_func = sigma.reshape(feature_domain)
    acq = lcb(mean_func, cov_func, explore)
    return dict(kernel_params=ell, mean_func=mean_func, cov_func=cov_func, acq=acq,
                index=first_argmax(acq), nlml_table=table)


# --------------------------------------------------------------------------------------
# synthetic workloads of BASELINE.json / SURVEY.md section 8(d)
# --------------------------------------------------------------------------------------
def synthetic_problem(n, d, seed=0, ell=0.3):
    """X ~ U[0,1]^{n x d}; y = sin(3*sum x) + 0.1*N(0,1); ell_k = 0.3."""
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    y = np.sin(3.0 * X.sum(axis=1)) + 0.1 * rng.standard_normal(n)
    return X, y, np.full(d, float(ell))
