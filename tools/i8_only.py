"""One fit + a few INT8-path acquisition chunks (ncu target)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
count = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
eng = GPEngine(0)
eng.set_acquire_path(sys.argv[5] if len(sys.argv) > 5 else "i8")
eng.set_screening(False)
eng.set_fused(len(sys.argv) > 6 and sys.argv[6] == "fused")      # one persistent fused kernel per sweep instead of the separate kernels      # every candidate through the N^2 product (the screen would leave almost nothing to profile)
X, y, ell = o.synthetic_problem(n, d)
grid = CandidateGrid([np.linspace(0, 1, 10 if d < 10 else 8)] * d)
fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = eng.acquire(fit, grid, 0, count, kind=ACQ_EI, f_best=float(y.min()), chunk=chunk); b.record(); torch.cuda.synchronize()
    print(f"{a.elapsed_time(b):.2f} ms best={r.best_index}")
