"""Per-kernel CUDA-event timing of the acquisition sweep on either path."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
count = int(sys.argv[3]) if len(sys.argv) > 3 else 262144
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
eng = GPEngine(0)
X, y, ell = o.synthetic_problem(n, d)
grid = CandidateGrid([np.linspace(0, 1, 10 if d < 10 else 8)] * d)
fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
for path in ("fp64", "i8"):
    eng.set_acquire_path(path)
    eng.acquire(fit, grid, 0, count, kind=ACQ_EI, f_best=float(y.min()), chunk=chunk)
    eng.profile(True)
    eng.acquire(fit, grid, 0, count, kind=ACQ_EI, f_best=float(y.min()), chunk=chunk)
    pr = eng.profile_read(); eng.profile(False)
    tot = sum(v[0] for v in pr.values())
    print(path, {k: (round(v[0], 2), v[1]) for k, v in pr.items() if v[1]}, f"total {tot:.2f} ms -> {count/tot*1e3:.3e} cand/s (serialised)")
    tri = pr["trigemm"][0]
    print(f"   trigemm: {count * float(n)**2 / tri * 1e-9:.1f} fp64-equivalent TFLOP/s" + (f", {34 * count * float(fit.n_pad) * (fit.n_pad + 128) / tri * 1e-9:.0f} int8 TOP/s executed" if path == "i8" else ""))
