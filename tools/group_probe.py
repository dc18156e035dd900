import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
eng = GPEngine(0); eng.set_screening(False)
X, y, ell = bench.synthetic()
grid = CandidateGrid([np.linspace(0, 1, 10)] * 8)
fit = eng.fit(X, y, ell, JITTER_POSTERIOR); fb = float(y.min())
count = 1 << 20
def run(g):
    eng.set_fused(False, g)
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = eng.acquire(fit, grid, 0, count, kind=ACQ_EI, f_best=fb, chunk=65536); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"group {g or 'auto(17)':>9}: {min(ts):8.2f} ms  {count / min(ts) * 1e3:.4e} cand/s  best={r.best_index}", flush=True)
for g in (0, 8, 12, 24, 34, 48, 64, 0):
    run(g)
