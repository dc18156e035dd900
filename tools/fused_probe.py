"""Fused persistent sweep kernel vs the separate kernels: timing at the benchmark's shape, for a few work-group sizes.
    python tools/fused_probe.py [N] [d] [candidates] [groups, comma separated]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
count = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
groups = [int(g) for g in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
eng = GPEngine(0)
eng.set_screening(False)
if (n, d) == (bench.N_OBS, bench.DIM):
    X, y, ell = bench.synthetic()
else:
    from oracle import gp_oracle as o
    X, y, ell = o.synthetic_problem(n, d)
grid = CandidateGrid([np.linspace(0, 1, 10 if d < 10 else 8)] * d)
fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
fb = float(y.min())

def run(label, reps=3):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = eng.acquire(fit, grid, 0, count, kind=ACQ_EI, f_best=fb, chunk=65536); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{label:28s} {min(ts):9.2f} ms  {count / min(ts) * 1e3:.4e} cand/s  best={r.best_index} score={r.best_score!r}", flush=True)

def prof(label):
    eng.profile(True)
    eng.acquire(fit, grid, 0, min(count, 1 << 18), kind=ACQ_EI, f_best=fb, chunk=65536)
    pr = eng.profile_read(); eng.profile(False)
    print(f"   {label}: " + ", ".join(f"{k} {v[0]:.2f} ms / {v[1]}" for k, v in pr.items() if v[1]), flush=True)

eng.set_fused(False); run("separate kernels"); prof("separate, tables")
for g in groups:
    eng.set_fused(True, g); run(f"fused, group={g or 'auto'}")
eng.set_fused(False); run("separate kernels (again)")
