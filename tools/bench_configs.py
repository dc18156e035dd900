"""Secondary configurations of BASELINE.json (configs[0], [1], [3], [4]) on one B200: timings with
CUDA events / wall clock, written as JSON lines.  Development/measurement aid; the contract bench is bench.py."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, JITTER_LML, ACQ_EI, ACQ_LCB
from bayesian_optimisation_b200.point_selector import PointSelector

eng = GPEngine(0)
out = []

def synth(n, d, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d)); y = np.sin(3.0 * X.sum(axis=1)) + 0.1 * rng.standard_normal(n)
    return X, y, np.full(d, 0.3)

def ev_time(fn, reps=3):
    best = 1e30; r = None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, r

# ---- configs[0]: the reference's native shapes through the drop-in class (M=21, 50x50 candidates, 50x50 length-scale grid)
rng = np.random.default_rng(1)
t1, t2 = np.linspace(1, 14, 50), np.linspace(10, 90, 50)
P = np.stack([m.reshape(-1) for m in np.meshgrid(t1, t2, indexing="ij")], axis=1)
idx = rng.choice(2500, 21, replace=False)
X, y = P[idx], rng.uniform(1e7, 1e9, 21)
def native():
    ps = PointSelector(); ps.name, ps.iteration = "c0", 1
    ps.measured_pts, ps.measured_vals, ps.feature_domain, ps.predicted_pts = X, y, [50, 50], P
    ps.length_scales = np.array([np.linspace(0.5, 10, 50), np.linspace(2, 100, 50)])
    ps.update_surrogate(); return ps.lower_confidence_bound()
native(); torch.cuda.synchronize()
t0 = time.perf_counter(); 
for _ in range(5): r = native()
torch.cuda.synchronize()
out.append({"config": "configs[0] native: M=21, C=2500 (50x50), 50x50 length-scale grid, update_surrogate+lower_confidence_bound via drop-in (host arrays in/out)",
            "ms_per_call": (time.perf_counter() - t0) / 5 * 1e3, "reference_cpu_ms_per_call": "~1000 (BASELINE.md section 2)", "index": [int(v) for v in r]})

# ---- configs[1]: N=1024, d=6, 1e6-point grid
X, y, ell = synth(1024, 6)
grid = CandidateGrid([np.linspace(0, 1, 10)] * 6)
fit_ms, fit = ev_time(lambda: eng.fit(X, y, ell, JITTER_POSTERIOR))
for kind, name in ((ACQ_LCB, "LCB"), (ACQ_EI, "EI")):
    ms, res = ev_time(lambda: eng.acquire(fit, grid, kind=kind, f_best=float(y.min()), chunk=16384))
    out.append({"config": f"configs[1] N=1024, d=6, 1e6-point grid, {name}", "fit_ms": fit_ms, "sweep_ms": ms,
                "candidates_per_s": grid.size / ms * 1e3, "tflops_n2": grid.size * 1024.0 ** 2 / ms * 1e-9, "best_index": res.best_index})

# ---- configs[3]: 1024 restarts x N=512, d=8, LML + gradients
X, y, _ = synth(512, 8)
ells = np.exp(np.random.default_rng(3).uniform(np.log(0.1), np.log(1.0), (1024, 8)))
dX, dy, dE = eng.to_device(X), eng.to_device(y), eng.to_device(ells)
for grad in (False, True):
    t = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = eng.nlml_batched(dX, dy, dE, JITTER_LML, want_grad=grad)
        torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
    nl = (r[0] if grad else r).cpu().numpy()
    flops = 1024 * (512 ** 3 / 3 * (3 if grad else 2))
    out.append({"config": f"configs[3] 1024 restarts x N=512, d=8, nlml{' + gradient' if grad else ''}", "ms": min(t) * 1e3,
                "restarts_per_s": 1024 / min(t), "approx_tflops": flops / min(t) * 1e-12, "argmin": int(np.argmin(nl)), "nlml_min": float(nl.min())})

# ---- configs[4]: N=16384, d=10, UCB sweep over a slice of the 8^10 grid
X, y, ell = synth(16384, 10)
grid = CandidateGrid([np.linspace(0, 1, 8)] * 10)
fit_ms, fit = ev_time(lambda: eng.fit(X, y, ell, JITTER_POSTERIOR), reps=2)
count = 1 << 17
ms, res = ev_time(lambda: eng.acquire(fit, grid, 0, count, kind=ACQ_LCB, chunk=8192), reps=2)
out.append({"config": "configs[4] N=16384, d=10, LCB sweep, 2^17-candidate slice of the 8^10 grid, 1 GPU", "fit_ms": fit_ms, "sweep_ms": ms,
            "candidates_per_s": count / ms * 1e3, "tflops_n2": count * 16384.0 ** 2 / ms * 1e-9, "nlml": fit.nlml,
            "full_1e9_sweep_estimate_s_8gpu": grid.size / (count / ms * 1e3) / 8})
for o_ in out:
    print(json.dumps(o_))
