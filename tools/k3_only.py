"""configs[3]: 1024 restarts x N=512, d=8 batched LML (+ gradient) -- ncu target / timing."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200.engine import GPEngine, JITTER_LML
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
grad = (sys.argv[3] == "1") if len(sys.argv) > 3 else True
eng = GPEngine(0)
X, y, _ = o.synthetic_problem(n, 8)
ells = np.exp(np.random.default_rng(3).uniform(np.log(0.1), np.log(1.0), (R, 8)))
dX, dy, dE = eng.to_device(X), eng.to_device(y), eng.to_device(ells)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = eng.nlml_batched(dX, dy, dE, JITTER_LML, want_grad=grad)
    torch.cuda.synchronize(); print(f"{(time.perf_counter()-t0)*1e3:.3f} ms launches {eng.launches}")
