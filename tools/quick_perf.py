"""Quick device timing of fit + acquisition at a given (N, d) -- development aid, not the bench."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
G = int(sys.argv[3]) if len(sys.argv) > 3 else 10
count = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 18
chunk = int(sys.argv[5]) if len(sys.argv) > 5 else 8192
eng = GPEngine(0)
X, y, ell = o.synthetic_problem(n, d)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for rep in range(3):
    a = ev(); fit = eng.fit(X, y, ell, JITTER_POSTERIOR); b = ev(); torch.cuda.synchronize()
    print(f"fit n={n} d={d}: {a.elapsed_time(b):.3f} ms  nlml={fit.nlml:.6f}")
    if rep < 2: fit.close()
grid = CandidateGrid([np.linspace(0, 1, G)] * d)
for rep in range(3):
    a = ev(); res = eng.acquire(fit, grid, 0, count, kind=ACQ_EI, f_best=float(y.min()), chunk=chunk); b = ev(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    npad = fit.n_pad
    print(f"acquire {count} cands chunk={chunk}: {ms:.2f} ms  {count/ms*1e3:.3e} cand/s  {count*npad*(npad+256)/ms*1e-9:.2f} TFLOP/s(tri)  best={res.best_index}")
