"""TEST INFRASTRUCTURE.  Golden vectors for the FULL-SIZE BASELINE.json configurations, computed with the
reference's own arithmetic restated in numpy (oracle/gp_oracle.py: inv-based posterior of
/root/reference/point_selector.py:78-98, nlml of :111-120 with slogdet).  These runs take minutes of host
time (a 16384^3 `np.linalg.inv`, a 10^6-point numpy sweep), so their results are committed as small fixtures
under tests/golden/ and the GPU parity tests compare against them:

    python oracle/make_config_goldens.py [c2] [c4] [c5]

  config_c2.npz  N=1024, d=6, whole 10^6-point grid: global LCB / EI arg-max (exact, first row-major maximum) and
                 their scores, + mu / sigma^2 on every 50th candidate (the tests recompute every 10th live)
  config_c4.npz  1024 restarts x N=512, d=8: nlml of every restart, gradient of every 16th
  config_c5.npz  N=16384, d=10: mu / sigma^2 on 2048 candidates of the 8^10 grid, nlml; CPU time of the fit
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gp_oracle as o  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
C5_START, C5_COUNT = 777_000_000, 2048


def make_c2():
    X, y, ell = o.synthetic_problem(1024, 6)
    axes = [np.linspace(0, 1, 10)] * 6
    P = o.candidate_grid(axes)
    t0 = time.time()
    mu, var = o.posterior_diag(X, y, P, ell, chunk=8192, return_var=True)
    sig = np.sqrt(np.abs(var))
    acq = o.lcb(mu, sig)
    fb = float(y.min())
    ei = o.expected_improvement(mu, sig, fb)
    np.savez_compressed(os.path.join(GOLDEN, "config_c2.npz"),
                        lcb_index=np.int64(np.flatnonzero(acq == acq.max())[0]), lcb_max=acq.max(),
                        ei_index=np.int64(np.flatnonzero(ei == ei.max())[0]), ei_max=ei.max(), f_best=fb,
                        lcb_runner_up=np.sort(acq)[-2], ei_runner_up=np.sort(ei)[-2],
                        stride=np.int64(50), mu_sub=mu[::50], var_sub=var[::50], seconds=time.time() - t0)
    print(f"c2: {time.time() - t0:.1f} s, lcb argmax {np.flatnonzero(acq == acq.max())[0]}, ei argmax {np.flatnonzero(ei == ei.max())[0]}")


def c4_problem():
    X, y, _ = o.synthetic_problem(512, 8, seed=4)
    ells = np.exp(np.random.default_rng(44).uniform(np.log(0.1), np.log(1.0), (1024, 8)))
    return X, y, ells


def make_c4():
    X, y, ells = c4_problem()
    t0 = time.time()
    ref = np.array([o.nlml(X, y, ells[r], stable=True) for r in range(len(ells))])
    sub = np.arange(0, len(ells), 16)
    gref = np.array([o.nlml_grad(X, y, ells[r]) for r in sub])
    np.savez_compressed(os.path.join(GOLDEN, "config_c4.npz"), nlml=ref, grad_rows=sub, grad=gref, seconds=time.time() - t0)
    print(f"c4: {time.time() - t0:.1f} s, argmin {np.argmin(ref)}")


def make_c5():
    X, y, ell = o.synthetic_problem(16384, 10)
    axes = [np.linspace(0, 1, 8)] * 10
    t0 = time.time()
    K = o.kernel_rbf_chunked(X, X, ell)
    K[np.diag_indices_from(K)] += o.JITTER_KERNEL + o.JITTER_EXTRA
    t1 = time.time()
    sign, logdet = np.linalg.slogdet(K)
    assert sign > 0
    inv = np.linalg.inv(K)                              # point_selector.py:89
    fit_s = time.time() - t1
    del K
    alpha = inv @ y
    nl = 0.5 * (y @ alpha + logdet + len(X) * np.log(2 * np.pi))
    P = o.grid_points(axes, C5_START, C5_START + C5_COUNT)
    mu, var = np.empty(C5_COUNT), np.empty(C5_COUNT)
    for s in range(0, C5_COUNT, 256):
        Ks = np.exp(-0.5 * np.sum((P[s:s + 256, None, :] - X[None, :, :]) ** 2 / ell ** 2, axis=2))
        mu[s:s + 256] = Ks @ alpha
        var[s:s + 256] = o.PRIOR_DIAG - np.einsum("cm,cm->c", Ks @ inv, Ks)
    try:
        from threadpoolctl import threadpool_info
        threads = max(p.get("num_threads", 1) for p in threadpool_info())
    except Exception:
        threads = os.cpu_count()
    np.savez_compressed(os.path.join(GOLDEN, "config_c5.npz"), start=np.int64(C5_START), mu=mu, var=var, nlml=nl, logdet=logdet,
                        inv_slogdet_seconds=fit_s, threads=np.int64(threads), seconds=time.time() - t0)
    print(f"c5: {time.time() - t0:.1f} s (inv + slogdet {fit_s:.1f} s on {threads} threads), nlml {nl!r}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c4", "c5"]
    for w in which:
        {"c2": make_c2, "c4": make_c4, "c5": make_c5}[w]()
