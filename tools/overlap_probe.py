"""How much of the panel kernel is hidden behind the tensor-core kernel?  Times the INT8 sweep
(a) as shipped (two streams, panel of chunk s+1 concurrent with the product of chunk s) and
(b) per kernel with the profiling hooks (serialised), for grid and explicit candidates."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
count = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
eng = GPEngine(0)
X, y, ell = o.synthetic_problem(n, d)
grid = CandidateGrid([np.linspace(0, 1, 10 if d < 10 else 8)] * d)
fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
fb = float(y.min())
pts = eng.to_device(o.grid_points(grid.axes, 0, count))


def timed(cand, chunk, reps=3):
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = eng.acquire(fit, cand, 0, count, kind=ACQ_EI, f_best=fb, chunk=chunk); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, r.best_index


for name, cand in (("grid", grid), ("explicit", pts)):
    for chunk in ([int(c) for c in sys.argv[4].split(',')] if len(sys.argv) > 4 else (16384, 32768, 65536, 131072)):
        ms, bi = timed(cand, chunk)
        eng.profile(True)
        eng.acquire(fit, cand, 0, count, kind=ACQ_EI, f_best=fb, chunk=chunk)
        pr = eng.profile_read(); eng.profile(False)
        parts = {k: round(v[0], 2) for k, v in pr.items() if v[1] and k in ("panel", "trigemm", "finalize", "merge")}
        print(f"{name:8s} chunk {chunk:6d}: shipped {ms:7.2f} ms ({count/ms*1e3:.3e} cand/s) best={bi} | serialised {parts} sum {sum(parts.values()):.2f} ms", flush=True)
