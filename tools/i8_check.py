"""Compare the INT8/tcgen05 acquisition path with the FP64/DMMA path (and time both)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
d = int(sys.argv[2]) if len(sys.argv) > 2 else 6
count = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
eng = GPEngine(0)
X, y, ell = o.synthetic_problem(n, d)
grid = CandidateGrid([np.linspace(0, 1, 10)] * d)
fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
res = {}
for path in ("fp64", "i8"):
    eng.set_acquire_path(path)
    r = eng.acquire(fit, grid, 0, count, kind=ACQ_EI, f_best=float(y.min()), outputs=True)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.acquire(fit, grid, 0, count, kind=ACQ_EI, f_best=float(y.min())); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res[path] = r
    print(f"{path}: {min(ts):.2f} ms  {count/min(ts)*1e3:.3e} cand/s  best={r.best_index} score={r.best_score:.15g}")
s0, s1 = res["fp64"].sigma.cpu().numpy(), res["i8"].sigma.cpu().numpy()
m0, m1 = res["fp64"].mu.cpu().numpy(), res["i8"].mu.cpu().numpy()
print("max |sigma^2 diff|", np.abs(s0**2 - s1**2).max(), " max |mu diff|", np.abs(m0 - m1).max(), " same index:", res["fp64"].best_index == res["i8"].best_index)
if n <= 2048:
    P = o.grid_points(grid.axes, 0, min(count, 20000))
    mu, var = o.posterior_diag(X, y, P, ell, return_var=True)
    k = len(P)
    print("vs oracle: i8 max|dvar|", np.abs(s1[:k]**2 - var).max(), " fp64 max|dvar|", np.abs(s0[:k]**2 - var).max())
