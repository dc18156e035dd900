"""The host-buffer C ABI ("session", include/bogp.h) that `PointSelector` binds: numpy in, numpy out, device memory /
copies / multi-device sharding inside the library.  Checked against the device-resident engine path (bit-identical: same
kernels, same orders) and the numpy oracle."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, record_error
from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu

_S = {}


def _session(devs=(0,)):
    from bayesian_optimisation_b200.session import Session
    if devs not in _S:
        _S[devs] = Session(list(devs))
    return _S[devs]


def _engine():
    from bayesian_optimisation_b200.engine import GPEngine
    if "eng" not in _S:
        _S["eng"] = GPEngine(0)
    return _S["eng"]


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("path", ["i8", "fp64"])
def test_session_update_is_bit_identical_to_the_device_resident_path(path):
    from bayesian_optimisation_b200.engine import ACQ_EI, CandidateGrid, JITTER_POSTERIOR
    s, eng = _session(), _engine()
    s.set_acquire_path(path); eng.set_acquire_path(path)
    X, y, ell = o.synthetic_problem(700, 5, seed=3)
    axes = [np.linspace(0, 1, 7)] * 5
    P = o.candidate_grid(axes)
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    # the session against the device-resident engine on the same kind of candidates (a grid descriptor and an explicit array
    # of the same points differ by a few ulp of k_* on the INT8 path: per-axis factor tables, csrc/acquire_i8.cuh)
    for kw in (dict(points=P), dict(axes=axes)):
        ref = eng.acquire(fit, CandidateGrid(axes) if "axes" in kw else P, outputs=True)
        got = s.update(X, y, ell, want_acq=True, **kw)
        np.testing.assert_array_equal(got["mu"], ref.mu.cpu().numpy())
        np.testing.assert_array_equal(got["sigma"], ref.sigma.cpu().numpy())
        np.testing.assert_array_equal(got["acq"], ref.acq.cpu().numpy())
        assert (got["best_score"], got["best_index"]) == (ref.best_score, ref.best_index)
        assert got["nlml"] == fit.nlml
    # a sub-range, EI, winner only
    fb = float(y.min())
    ref2 = eng.acquire(fit, CandidateGrid(axes), 1000, 9000, kind=ACQ_EI, f_best=fb)
    got2 = s.update(X, y, ell, axes=axes, c_begin=1000, c_end=9000, kind=ACQ_EI, f_best=fb, outputs=False)
    assert got2["mu"] is None and (got2["best_score"], got2["best_index"]) == (ref2.best_score, ref2.best_index)
    # score on the posterior the update left on the device, then on arrays the caller changed
    got = s.update(X, y, ell, points=P)
    sc = s.score(len(P), explore=4.0)
    np.testing.assert_array_equal(sc["acq"], 4 * got["sigma"] - got["mu"])
    assert sc["best_index"] == ref.best_index
    sc2 = s.score(len(P), explore=2.5, mu=got["mu"] + 1.0, sigma=got["sigma"])
    acq2 = 2.5 * got["sigma"] - (got["mu"] + 1.0)
    np.testing.assert_array_equal(sc2["acq"], acq2)
    assert sc2["best_index"] == int(np.flatnonzero(acq2 == acq2.max())[0])
    fit.close()
    s.set_acquire_path("i8"); eng.set_acquire_path("i8")


def test_session_streams_large_candidate_arrays_in_pieces():
    """600k explicit candidates: copied in several pieces under the sweep, outputs copied back piece by piece -- same
    numbers as one device-resident sweep."""
    from bayesian_optimisation_b200.engine import JITTER_POSTERIOR
    s, eng = _session(), _engine()
    X, y, ell = o.synthetic_problem(100, 2, seed=5)
    P = np.random.default_rng(1).random((600_011, 2))
    fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
    ref = eng.acquire(fit, P, outputs=True)
    got = s.update(X, y, ell, points=P, want_acq=True)
    np.testing.assert_array_equal(got["mu"], ref.mu.cpu().numpy())
    np.testing.assert_array_equal(got["sigma"], ref.sigma.cpu().numpy())
    assert (got["best_score"], got["best_index"]) == (ref.best_score, ref.best_index)
    # 100 points in the unit square at ell = 0.3 are a badly conditioned system (cond ~ 1e5..1e6): compare with the
    # long-double truth, against which the inv-based oracle itself is only good to ~1e-11 absolute
    from oracle import truth
    from bayesian_optimisation_b200.engine import PRIOR_DIAG
    mu_t, var_t, _, _ = truth.posterior_truth(X, y, P[::997], ell, JITTER_POSTERIOR, PRIOR_DIAG)
    np.testing.assert_allclose(got["mu"][::997], mu_t, rtol=1e-9, atol=1e-9 * np.abs(mu_t).max())
    np.testing.assert_allclose(got["sigma"][::997] ** 2, var_t, rtol=1e-9, atol=2e-13)
    fit.close()


def test_session_nlml_table_chunks_by_free_memory():
    """the reference's 50 x 50 length-scale grid at M = 1024 needs a 42 GB batched workspace; the session (and
    GPEngine.nlml_batched) work through it in chunks.  Same numbers as restart-by-restart evaluation."""
    s, eng = _session(), _engine()
    X, y, _ = o.synthetic_problem(1024, 2, seed=8)
    ls = [np.linspace(0.05, 1.0, 50), np.linspace(0.05, 1.0, 50)]
    ells = np.stack(np.meshgrid(*ls, indexing="ij"), axis=-1).reshape(-1, 2)
    table = s.nlml(X, y, ells)
    t2 = eng.nlml_batched(X, y, ells).cpu().numpy()
    np.testing.assert_array_equal(table, t2)
    for r in (0, 777, 2499):
        ref = o.nlml(X, y, ells[r], stable=True)
        one = eng.nlml_batched(X, y, ells[r:r + 1].repeat(2, axis=0)).cpu().numpy()[0]
        assert table[r] == one
        if np.isfinite(ref):
            assert abs(table[r] - ref) <= 1e-9 * abs(ref), (r, table[r], ref)
    sub, g = s.nlml(X, y, ells[100:140], want_grad=True)
    np.testing.assert_array_equal(sub, table[100:140])
    gref = o.nlml_grad(X, y, ells[120])
    np.testing.assert_allclose(g[20], gref, rtol=1e-8, atol=1e-9 * np.abs(gref).max())


def test_session_kernel_matrix_and_errors():
    s = _session()
    g = load_golden("direct_d6_n64_c300")
    K = s.kernel_matrix(g["X"], g["P"], g["ell"])
    np.testing.assert_allclose(K, g["Kxp"], rtol=1e-13, atol=1e-300)
    with pytest.raises(np.linalg.LinAlgError):
        s.update(np.zeros((3, 2)), np.ones(3), np.ones(2), points=np.zeros((4, 2)), jitter=-2.0)
    X = np.array([[0.0, 0.0], [1.0, 1.0]])
    with pytest.raises(IndexError):
        s.update(X, np.array([np.nan, 1.0]), np.ones(2), points=np.random.default_rng(0).random((70, 2)))
    with pytest.raises(IndexError):
        s.score(3, mu=np.array([0.0, np.nan, 1.0]), sigma=np.ones(3))


@pytest.mark.parametrize("devs", [(0, 0), (0, 0, 0)], ids=["2-way", "3-way"])
def test_session_shards_over_several_contexts_bit_identically(devs):
    """The single-process multi-device mode, exercised on ONE GPU by naming device 0 several times (each entry gets its
    own context, streams and buffers): contiguous shards, replicated fit, host-side fold of the winners.  Bit-identical
    to the one-context session, ties resolved to the lowest flat index."""
    from bayesian_optimisation_b200.engine import ACQ_EI
    one, many = _session(), _session(devs)
    X, y, ell = o.synthetic_problem(300, 3, seed=4)
    P = np.random.default_rng(2).random((10_001, 3))
    P[7000] = P[123]                                   # an exact tie across shards: the lower index must win if it is the maximum
    a = one.update(X, y, ell, points=P, want_acq=True)
    b = many.update(X, y, ell, points=P, want_acq=True)
    for k in ("mu", "sigma", "acq"):
        np.testing.assert_array_equal(a[k], b[k])
    assert (a["best_score"], a["best_index"]) == (b["best_score"], b["best_index"])
    sa, sb = one.score(len(P), kind=ACQ_EI, f_best=float(y.min())), many.score(len(P), kind=ACQ_EI, f_best=float(y.min()))
    np.testing.assert_array_equal(sa["acq"], sb["acq"])
    assert (sa["best_score"], sa["best_index"]) == (sb["best_score"], sb["best_index"])
    # constant scores everywhere: index 0 wins on every layout
    flat = many.score(len(P), mu=np.zeros(len(P)), sigma=np.ones(len(P)))
    assert flat["best_index"] == 0
    axes = [np.linspace(0, 1, 11)] * 3
    ga, gb = one.update(X, y, ell, axes=axes), many.update(X, y, ell, axes=axes)
    np.testing.assert_array_equal(ga["sigma"], gb["sigma"])
    assert ga["best_index"] == gb["best_index"]
    ells = np.exp(np.random.default_rng(3).uniform(np.log(0.2), np.log(1.0), (17, 3)))
    np.testing.assert_array_equal(one.nlml(X, y, ells), many.nlml(X, y, ells))
    # more shards than candidates
    tiny = many.update(X, y, ell, points=P[:2])
    np.testing.assert_array_equal(tiny["mu"], a["mu"][:2])


def test_point_selector_class_level_devices():
    from bayesian_optimisation_b200.point_selector import PointSelector
    g = load_golden("native2d_t1t2_m10")

    def run(cls):
        ps = cls()
        ps.measured_pts, ps.measured_vals = g["X"].copy(), g["y"].copy()
        ps.feature_domain, ps.predicted_pts, ps.length_scales = list(g["feature_domain"]), g["P"].copy(), np.array([g["ls0"], g["ls1"]])
        ps.update_surrogate()
        return ps, ps.lower_confidence_bound()

    class Sharded(PointSelector):
        devices = [0, 0]
    a, ia = run(PointSelector)
    b, ib = run(Sharded)
    np.testing.assert_array_equal(ia, g["index"]); np.testing.assert_array_equal(ib, g["index"])
    np.testing.assert_array_equal(a.mean_func, b.mean_func)
    np.testing.assert_array_equal(a.cov_func, b.cov_func)
    np.testing.assert_array_equal(a.acq_func_eval, b.acq_func_eval)
    # two selectors interleaved on one session: the second update replaces the device copy, the first one must notice
    c, _ = run(PointSelector)
    i_again = a.lower_confidence_bound()
    np.testing.assert_array_equal(i_again, ia)


_NCCL_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
rank = int(sys.argv[1])
os.environ.update(RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT="{port}")
import numpy as np, torch, torch.distributed as dist
from conftest import load_golden
from oracle import gp_oracle as o
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from bayesian_optimisation_b200.point_selector import PointSelector
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
from bayesian_optimisation_b200 import sharding
g = load_golden("native2d_t1t2_m21")
ps = PointSelector()
ps.measured_pts, ps.measured_vals = g["X"].copy(), g["y"].copy()
ps.feature_domain, ps.predicted_pts, ps.length_scales = list(g["feature_domain"]), g["P"].copy(), np.array([g["ls0"], g["ls1"]])
ps.update_surrogate()
idx = ps.lower_confidence_bound()
np.testing.assert_array_equal(ps.kernel_params, g["kernel_params"])
np.testing.assert_array_equal(idx, g["index"])
np.testing.assert_allclose(ps.mean_func, g["mean_func"], rtol=1e-9, atol=1e-9 * np.abs(g["mean_func"]).max())
np.testing.assert_array_equal(ps.acq_func_eval, 4 * ps.cov_func - ps.mean_func)
# device-side max-loc of a sharded sweep: all_gather of the 24-byte records, fold on the device, one host read
eng = GPEngine(rank)
X, y, ell = o.synthetic_problem(500, 4, seed=1)
grid = CandidateGrid([np.linspace(0, 1, 9)] * 4)
fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
s, i, _ = sharding.sharded_acquire(eng, fit, grid, grid.size, rank, 2, kind=ACQ_EI, f_best=float(y.min()))
full = eng.acquire(fit, grid, kind=ACQ_EI, f_best=float(y.min()))
assert (s, i) == (full.best_score, full.best_index), (s, i, full.best_score, full.best_index)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_nccl_point_selector_and_device_maxloc(tmp_path):
    if _device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out
        assert "ok" in out


def test_session_on_two_real_devices():
    if _device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    one, two = _session(), _session((0, 1))
    X, y, ell = o.synthetic_problem(600, 4, seed=6)
    axes = [np.linspace(0, 1, 12)] * 4
    a, b = one.update(X, y, ell, axes=axes, want_acq=True), two.update(X, y, ell, axes=axes, want_acq=True)
    for k in ("mu", "sigma", "acq"):
        np.testing.assert_array_equal(a[k], b[k])
    assert (a["best_score"], a["best_index"]) == (b["best_score"], b["best_index"])
