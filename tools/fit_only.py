"""Three fits at (N, d) -- target for an ncu launch list of the fit kernels."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200.engine import GPEngine, JITTER_POSTERIOR
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = GPEngine(0)
X, y, ell = o.synthetic_problem(n, d)
dX, dy = eng.to_device(X), eng.to_device(y)
for rep in range(3):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); fit = eng.fit(dX, dy, ell, JITTER_POSTERIOR); b.record(); torch.cuda.synchronize()
    print(f"fit {a.elapsed_time(b):.3f} ms launches {eng.launches}")
    fit.close()
