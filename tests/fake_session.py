"""CPU stand-in for `bayesian_optimisation_b200.session.Session` (TEST INFRASTRUCTURE): the same method surface, the
arithmetic done by the numpy oracle.  Lets the host logic of the drop-in -- `PointSelector`, the `dropin/` shim, the
sharded modes -- run on a box without a GPU (the product itself has no CPU path; this class lives under tests/)."""
import numpy as np

from oracle import gp_oracle as o


class OracleSession:
    devices = [0]

    def __init__(self):
        self.generation = 0
        self.calls = []
        self._post = None

    launches = 0

    def set_acquire_path(self, path):
        pass

    def kernel_matrix(self, a, b, ell, jitter=0.0):
        self.generation += 1
        K = o.kernel_rbf_chunked(a, b, np.asarray(ell, dtype=np.float64).reshape(-1))
        m = min(K.shape)
        K[np.arange(m), np.arange(m)] += jitter
        return K

    def nlml(self, x, y, ells, jitter=1e-4, want_grad=False):
        assert jitter == o.JITTER_KERNEL
        self.calls.append(("nlml", len(ells)))
        out = np.array([o.nlml(x, y, e, stable=True) for e in np.asarray(ells)])
        if want_grad:
            return out, np.array([o.nlml_grad(x, y, e) for e in np.asarray(ells)])
        return out

    def update(self, x, y, ell, points=None, axes=None, c_begin=0, c_end=None, jitter=None, prior_diag=None, kind=0, explore=4.0,
               f_best=0.0, cross_jitter=0.0, outputs=True, want_acq=False):
        P = o.candidate_grid(axes) if axes is not None else np.asarray(points, dtype=np.float64)
        c_end = len(P) if c_end is None else c_end
        self.calls.append(("update", c_begin, c_end))
        if cross_jitter:                       # the M == C quirk lives on the global diagonal: evaluate everything, slice
            mu, sigma = o.posterior_diag(x, y, P, ell)
            mu, sigma = mu[c_begin:c_end], sigma[c_begin:c_end]
        else:
            mu, sigma = o.posterior_diag(x, y, P[c_begin:c_end], ell, c_offset=1)
        acq = o.lcb(mu, sigma, explore) if kind == 0 else o.expected_improvement(mu, sigma, f_best)
        if np.isnan(acq).any():
            raise IndexError("NaN acquisition value")
        i = int(np.flatnonzero(acq == acq.max())[0])
        self.generation += 1
        self._post = (mu, sigma, c_begin)
        return dict(mu=mu, sigma=sigma, acq=acq if want_acq else None, nlml=float("nan"), best_score=float(acq[i]), best_index=c_begin + i,
                    generation=self.generation)

    def score(self, count, kind=0, explore=4.0, f_best=0.0, mu=None, sigma=None, want_acq=True):
        off = 0
        if mu is None:
            mu, sigma, off = self._post
        else:
            self.generation += 1
            self._post = (np.asarray(mu, dtype=np.float64).reshape(-1), np.asarray(sigma, dtype=np.float64).reshape(-1), 0)
            mu, sigma, off = self._post
        assert len(mu) == count
        acq = o.lcb(mu, sigma, explore) if kind == 0 else o.expected_improvement(mu, sigma, f_best)
        if np.isnan(acq).any():
            raise IndexError("NaN acquisition value")
        i = int(np.flatnonzero(acq == acq.max())[0])
        return dict(acq=acq if want_acq else None, best_score=float(acq[i]), best_index=off + i)
