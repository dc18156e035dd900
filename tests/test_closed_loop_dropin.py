"""the UNMODIFIED reference entry point against the drop-in.

`/root/reference/select_parameters.py` is executed (runpy) with `dropin/point_selector.py` answering its
`from point_selector import PointSelector`, for full 20-iteration sample loops over all five parameter groups, with the
unmodified `terminate_opto.py` / `terminate_block.py` / `terminate_algo.py` chained between the iterations the way the
DAG templates chain them and the synthetic objective of SURVEY f-3 standing in for RAT.  Every file the workflow leaves
behind -- `measured_points/*.npy`, `opto_log.JSON`, `macros/*.mac`, `submit_files/simulate.submit` -- must be byte for
byte what the same run with the reference's own class leaves behind (tests/golden/closed_loop_files.npz, written by
oracle/make_closed_loop.py; BOGP_LIVE_REFERENCE_ARM=1 re-runs the reference arm in the session instead).

There is no GPU on this box, so the drop-in's session is the oracle-backed stand-in of tests/fake_session.py: what is
tested here is everything between the caller and the C ABI -- the shim, the attribute bag, list/array conversions,
kernel_params shapes, the float32 table arg-min, the index conventions.  The arithmetic of the CUDA path is tested by
replaying the same recorded calls on the GPU (tests/test_gpu_parity.py::test_point_selector_dropin_replays_...)."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.reference


def _load_shim():
    spec = importlib.util.spec_from_file_location("_dropin_point_selector", os.path.join(ROOT, "dropin", "point_selector.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.PointSelector


def _golden_files():
    with np.load(os.path.join(GOLDEN, "closed_loop_files.npz"), allow_pickle=False) as z:
        names, sizes, blob = [str(n) for n in z["names"]], z["sizes"], z["blob"].tobytes()
        conf = tuple(int(v) for v in z["sizes_arg"]), int(z["seed"])
    out, off = {}, 0
    for n, s in zip(names, sizes):
        out[n] = blob[off:off + int(s)]
        off += int(s)
    return out, conf


def test_unmodified_select_parameters_with_the_dropin_writes_the_same_files():
    from bayesian_optimisation_b200 import session as sm
    from fake_session import OracleSession
    from oracle import reference_loader as rl
    from oracle.workflow import run_workflow
    want, (sizes, seed) = _golden_files()
    if os.environ.get("BOGP_LIVE_REFERENCE_ARM") == "1":
        want = run_workflow(rl.load_reference_class(), sizes=sizes, seed=seed)["files"]
    fake = OracleSession()
    sm.set_default_session(fake)
    try:
        cls = _load_shim()
        assert cls.__module__ == "bayesian_optimisation_b200.point_selector"
        got = run_workflow(cls, sizes=sizes, seed=seed)
    finally:
        sm.set_default_session(None)
    files = got["files"]
    assert sorted(files) == sorted(want)
    kinds = {"npy": 0, "JSON": 0, "mac": 0, "submit": 0}
    for name in sorted(want):
        assert files[name] == want[name], f"{name} differs from the reference run"
        for k in kinds:
            kinds[k] += name.endswith(k)
    assert kinds["npy"] >= 5 and kinds["JSON"] == 1 and kinds["mac"] >= 1 and kinds["submit"] == 1, kinds
    n_calls = len(got["calls"])
    assert n_calls >= 100 and max(len(c["X"]) for c in got["calls"]) >= 20           # full 20-iteration sample loops
    assert sum(1 for c in fake.calls if c[0] == "update") == n_calls
