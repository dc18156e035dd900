"""CPU-side checks: the C-ABI library loads and exports every symbol include/bogp.h declares
(no compute without a GPU), fails loudly without a device, and the host-side sharding logic."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bogp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bogp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    from bayesian_optimisation_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        g.build()
    return _lib.load()


def test_header_declares_what_the_binding_binds(lib):
    from bayesian_optimisation_b200 import _lib
    declared = _declared_symbols()
    assert declared, "no symbols parsed from include/bogp.h"
    assert sorted(_lib.SIGNATURES) == declared


def test_library_exports_every_declared_symbol(lib):
    from bayesian_optimisation_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (bogp_[a-z0-9_]+)", out))
    for name in _declared_symbols():
        assert name in exported, name
        assert getattr(lib, name) is not None


def test_library_contains_sm100a_tensor_and_tma_code(lib):
    from bayesian_optimisation_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    assert "DMMA.8x8x4" in sass          # FP64 tensor path
    assert "UBLKCP" in sass              # bulk-TMA staging
    assert "UTCIMMA" in sass             # tcgen05.mma kind::i8 (acquisition product)
    assert "LDTM" in sass                # tcgen05.ld: TMEM accumulators read back in the epilogue
    assert "UTMALDG" in sass             # tensor-map TMA loads (trailing SYRK)
    assert "UCGABAR_ARV" in sass         # thread-block cluster barrier (in-block factorisation kernel)


def test_workspace_queries_are_pure_host_functions(lib):
    assert lib.bogp_fit_workspace_bytes(4096, 8) > 2 * 4096 * 4096 * 8
    assert lib.bogp_fit_workspace_bytes(0, 8) == 0
    assert lib.bogp_fit_workspace_bytes(10, 17) == 0
    small = lib.bogp_nlml_batched_workspace_bytes(21, 2, 2500, 0)
    big = lib.bogp_nlml_batched_workspace_bytes(512, 8, 1024, 1)
    assert 0 < small < 1 << 20 and big > 2 * 1024 * 512 * 512 * 8
    assert lib.bogp_version().startswith(b"bogp")


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert lib.bogp_create(0, C.byref(h)) == -2
    assert b"no CPU fallback" in lib.bogp_last_error()
    from bayesian_optimisation_b200.engine import GPEngine
    with pytest.raises(RuntimeError):
        GPEngine(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bayesian_optimisation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/", "").lower() or f == "__none__", f


def test_jitter_constants_reproduce_the_reference_rounding():
    from bayesian_optimisation_b200 import engine as e
    assert e.PRIOR_DIAG == (1.0 + 1e-4) + 1e-6
    assert 1.0 + e.JITTER_POSTERIOR == e.PRIOR_DIAG


def test_shard_ranges_cover_everything_in_order():
    from bayesian_optimisation_b200.sharding import shard_range, restart_slice
    for total in (1, 7, 2500, 10 ** 8, 8 ** 10):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            assert all(e - b <= -(-total // world) for b, e in edges)
    assert sorted(i for r in range(4) for i in restart_slice(1024, r, 4)) == list(range(1024))


def test_reduce_pairs_tie_rule():
    from bayesian_optimisation_b200.sharding import reduce_pairs, NO_INDEX
    assert reduce_pairs([(1.0, 5), (2.0, 9), (2.0, 3), (float("nan"), 0)]) == (2.0, 3)
    assert reduce_pairs([(float("-inf"), NO_INDEX)]) == (float("-inf"), NO_INDEX)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from oracle import gp_oracle as o
from bayesian_optimisation_b200.sharding import shard_range, allreduce_maxloc
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
X, y, ell = o.synthetic_problem(40, 3, seed=2)
axes = [np.linspace(0, 1, 6)] * 3
P = o.candidate_grid(axes)
b, e = shard_range(len(P), rank, world)
mu, sig = o.posterior_diag(X, y, P[b:e], ell)            # the oracle stands in for the GPU scorer here
# restart sharding (sharded_nlml_argmin) with a CPU stand-in for the batched GPU kernel
class _CpuEngine:
    device = None
    def nlml_batched(self, x, y, ells, **kw):
        return torch.tensor([o.nlml(x, y, e) for e in ells], dtype=torch.float64)
from bayesian_optimisation_b200.sharding import sharded_nlml_argmin
ells = np.exp(np.random.default_rng(5).uniform(np.log(0.2), np.log(1.0), (13, 3)))
gv, gi, _ = sharded_nlml_argmin(_CpuEngine(), X, y, ells, rank, world)
full_table = np.array([o.nlml(X, y, e) for e in ells]).astype(np.float32)
assert gi == int(np.flatnonzero(full_table == full_table.min())[0]) and np.float32(gv) == full_table.min(), (gi, gv)
acq = o.lcb(mu, sig)
acq[:] = np.round(acq, 1)                                 # force exact ties across ranks
li = int(np.flatnonzero(acq == acq.max())[0])
s, i = allreduce_maxloc(float(acq[li]), b + li)
mu_f, sig_f = o.posterior_diag(X, y, P, ell)
full = np.round(o.lcb(mu_f, sig_f), 1)
want = int(np.flatnonzero(full == full.max())[0])
assert i == want and s == full[want], (rank, s, i, want)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok", i)
"""


def test_two_rank_gloo_maxloc_matches_single_process(tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out
        assert "ok" in out


# ---------------------------------------------------------------------------------------------
# property tests (hypothesis): the sharded reduction equals the reference's first-arg-max on the whole array
# ---------------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st


@settings(max_examples=200, deadline=None)
@given(st.integers(min_value=1, max_value=10 ** 10), st.integers(min_value=1, max_value=16))
def test_shard_range_partitions_any_count(total, world):
    from bayesian_optimisation_b200.sharding import shard_range
    edges = [shard_range(total, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
    assert all(0 <= e - b <= -(-total // world) for b, e in edges)


_scores = st.lists(st.one_of(st.sampled_from([0.0, 1.0, -1.0, 2.5, float("-inf")]), st.floats(allow_nan=False, width=16)),
                   min_size=1, max_size=60)


@settings(max_examples=300, deadline=None)
@given(_scores, st.integers(min_value=1, max_value=9))
def test_sharded_maxloc_equals_numpy_first_argmax(scores, world):
    """Every rank reports (its maximum, its FIRST position of it); reduce_pairs over the ranks must give what
    np.argwhere(a == np.amax(a))[0] gives on the whole array (point_selector.py:207), ties included."""
    import numpy as np
    from bayesian_optimisation_b200.sharding import reduce_pairs, shard_range, NO_INDEX
    a = np.array(scores, dtype=np.float64)
    pairs = []
    for r in range(world):
        b, e = shard_range(len(a), r, world)
        if e > b:
            loc = int(np.flatnonzero(a[b:e] == a[b:e].max())[0])
            pairs.append((float(a[b + loc]), b + loc))
        else:
            pairs.append((float("-inf"), NO_INDEX))
    s, i = reduce_pairs(pairs)
    want = int(np.argwhere(a == np.amax(a))[0][0])
    if np.isneginf(a).all():
        assert s == float("-inf")           # nothing beats the initial value: no index is reported
    else:
        assert (s, i) == (float(a[want]), want)


@settings(max_examples=100, deadline=None)
@given(st.integers(min_value=1, max_value=5000), st.integers(min_value=1, max_value=16))
def test_restart_slices_partition_the_restarts(total, world):
    from bayesian_optimisation_b200.sharding import restart_slice
    seen = sorted(i for r in range(world) for i in restart_slice(total, r, world))
    assert seen == list(range(total))


# ---------------------------------------------------------------------------------------------
# the kernel function's exp: the constants of csrc/common.cuh, restated in C (same operations, libm fma) and
# checked against long-double expl on the host -- pins the table and the coefficients without a GPU
# ---------------------------------------------------------------------------------------------
_EXP_C = r"""
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static const double tab[64] = { %(table)s };
static const double C[8] = { %(coeffs)s };
static double exp_nonpos(double t) {                 /* csrc/common.cuh: exp_nonpos, operation for operation */
    const double kd = fma(t, C[0], 6755399441055744.0);
    int64_t bits; memcpy(&bits, &kd, 8);
    const int ki = (int)(uint32_t)bits;
    const double kf = kd - 6755399441055744.0;
    double r = fma(kf, C[1], t);
    r = fma(kf, C[2], r);
    const double T = tab[ki & 63];
    double q = fma(C[3], r, C[4]);
    q = fma(q, r, C[5]);
    q = fma(q, r, C[6]);
    q = fma(q, r, 0.5);
    const double p = fma(r * r, q, r);
    const double m = fma(T, p, T);
    int64_t mb; memcpy(&mb, &m, 8);
    const uint32_t hi = (uint32_t)(mb >> 32) + (((uint32_t)ki << 14) & 0xfff00000u);
    mb = ((int64_t)hi << 32) | (mb & 0xffffffffLL);
    double res; memcpy(&res, &mb, 8);
    return t < -708.0 ? 0.0 : res;
}
int main(void) {
    double worst = 0.0; srand(7);
    for (long i = 0; i < 3000000; i++) {
        const double u = rand() / (double)RAND_MAX, v = rand() / (double)RAND_MAX;
        const double t = (i %% 3 == 0) ? -708.0 * u : ((i %% 3 == 1) ? -45.0 * u * v : -1e-3 * u);
        const long double ref = expl((long double)t);
        const double ulp = nextafter((double)ref, INFINITY) - (double)ref;
        const double e = (double)(fabsl((long double)exp_nonpos(t) - ref) / ulp);
        if (e > worst) worst = e;
    }
    printf("%%.4f %%.17g %%.17g %%.17g %%d\n", worst, exp_nonpos(0.0), exp_nonpos(-0.0), exp_nonpos(-709.0), isnan(exp_nonpos(NAN)) ? 1 : 0);
    return 0;
}
"""


def test_kernel_function_exp_constants_give_one_ulp(tmp_path):
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = open(os.path.join(ROOT, "bayesian_optimisation_b200", "csrc", "common.cuh")).read()
    table = re.search(r"kExp2Tab\[64\] = \{(.*?)\};", src, re.S).group(1)
    coeffs = re.search(r"kExpC\[8\] = \{(.*?)\};", src, re.S).group(1)
    coeffs = re.sub(r"//[^\n]*", "", coeffs)
    assert len(re.findall(r"0x1\.[0-9a-f]+p\+0", table)) == 64
    c = tmp_path / "exp_nonpos_check.c"
    c.write_text(_EXP_C % {"table": table, "coeffs": coeffs})
    exe = tmp_path / "exp_nonpos_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), str(c), "-lm"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    worst, at0, atm0, below, nan_ok = float(out[0]), float(out[1]), float(out[2]), float(out[3]), int(out[4])
    assert worst <= 1.1, worst                    # maximum error in ulp against long-double expl
    assert at0 == 1.0 and atm0 == 1.0 and below == 0.0 and nan_ok == 1
