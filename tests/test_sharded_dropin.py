"""The one-process-per-GPU mode behind the UNCHANGED caller-visible API, on CPU: two gloo ranks, the
oracle-backed stand-in session of tests/fake_session.py in place of the GPU.  Every rank must end up with the same
full `mean_func` / `cov_func` / `acq_func_eval`, the same `kernel_params` and the same selected index as a single
process -- through all-gathers of the slices and ONE (score, index, nan) record exchange."""
import os
import socket
import subprocess
import sys

from conftest import ROOT

_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from conftest import load_golden
from fake_session import OracleSession
from bayesian_optimisation_b200 import session as sm
from bayesian_optimisation_b200.point_selector import PointSelector
from bayesian_optimisation_b200 import sharding

rank = int(sys.argv[1])
def run(name, ei=False):
    g = load_golden(name)
    ps = PointSelector()
    ps.name, ps.iteration = "gloo", 1
    ps.measured_pts, ps.measured_vals = g["X"].copy(), g["y"].copy()
    ps.feature_domain, ps.predicted_pts = list(g["feature_domain"]), g["P"].copy()
    ps.length_scales = np.array([g["ls0"], g["ls1"]]) if "ls1" in g else g["ls0"]
    ps.update_surrogate()
    idx = ps.expected_improvement() if ei else ps.lower_confidence_bound()
    return ps, idx

# single process first (no process group yet)
sm.set_default_session(OracleSession())
single = {{n: run(n) for n in ("native2d_t1t2_m10", "native1d_tr_m8", "native2d_t1t2_m1")}}
single_ei = run("native2d_t1t2_m10", ei=True)

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)
fake = OracleSession()
sm.set_default_session(fake)
for n, (ref, ref_idx) in single.items():
    ps, idx = run(n)
    C = int(np.prod(ps.feature_domain))
    b, e = sharding.shard_range(C, rank, 2)
    assert ("update", b, e) in fake.calls, (n, fake.calls)          # this rank scored only its slice
    np.testing.assert_array_equal(ps.kernel_params, ref.kernel_params)
    # (the stand-in's BLAS products depend on the slice shape in the last bit; the CUDA path is bit-identical across
    #  shards by construction and is held to that in tests/test_gpu_parity.py)
    np.testing.assert_allclose(ps.mean_func, ref.mean_func, rtol=1e-14)
    np.testing.assert_allclose(ps.cov_func, ref.cov_func, rtol=1e-14)
    np.testing.assert_allclose(ps.acq_func_eval, ref.acq_func_eval, rtol=1e-13)
    np.testing.assert_array_equal(idx, ref_idx)
    np.testing.assert_array_equal(ps.nlogml, ref.nlogml) if ref.nlogml is not None else None
    assert isinstance(ps.measured_pts, list)
ps, idx = run("native2d_t1t2_m10", ei=True)
np.testing.assert_array_equal(idx, single_ei[1]); np.testing.assert_allclose(ps.acq_func_eval, single_ei[0].acq_func_eval, rtol=1e-12)
# the restart table was sharded r, r+2, ...: each rank evaluated half of the 2500 cells
assert ("nlml", 1250) in fake.calls, [c for c in fake.calls if c[0] == "nlml"][:3]
# caller replaced the arrays: the acquisition is evaluated on what it holds now
ps.mean_func = ps.mean_func + 1.0
i2 = ps.lower_confidence_bound()
np.testing.assert_array_equal(ps.acq_func_eval, 4 * ps.cov_func - ps.mean_func)
np.testing.assert_array_equal(i2, np.argwhere(ps.acq_func_eval == ps.acq_func_eval.max())[0])

# NaN semantics (ADVICE r1): a NaN on ONE rank raises IndexError on EVERY rank, for the sweep and for the restart table
try:
    sharding.allreduce_maxloc(1.0, 5, nan_flag=(rank == 1)); raise SystemExit("no IndexError")
except IndexError:
    pass
class _E:
    device = None
    def nlml_batched(self, x, y, ells, **kw):
        t = np.arange(len(ells), dtype=np.float64) + 10.0 * rank
        if rank == 0: t[1] = np.nan
        return t
try:
    sharding.sharded_nlml_argmin(_E(), None, None, np.zeros((9, 2)), rank, 2); raise SystemExit("no IndexError")
except IndexError:
    pass
class _F(_E):
    def nlml_batched(self, x, y, ells, **kw):
        return np.array([5.0, 3.0, 3.0, 9.0, 7.0])[np.arange(rank, 5, 2)]
gv, gi, _ = sharding.sharded_nlml_argmin(_F(), None, None, np.zeros((5, 2)), rank, 2)
assert (gv, gi) == (3.0, 1), (gv, gi)
# strided gather helper
full = sharding.all_gather_strided(np.arange(rank, 7, 2, dtype=np.float64), 7, rank, 2)
np.testing.assert_array_equal(full, np.arange(7.0))
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_gloo_point_selector_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out
        assert "ok" in out
