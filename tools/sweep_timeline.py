"""Kernel time of one screened full-grid sweep by kernel name (CUPTI via torch.profiler; concurrency preserved).
    python tools/sweep_timeline.py [N] [d] [candidates]"""
import sys, json, tempfile, os, collections
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
import bench
from bayesian_optimisation_b200.engine import GPEngine, CandidateGrid, JITTER_POSTERIOR, ACQ_EI
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
count = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 8
begin = int(sys.argv[4]) if len(sys.argv) > 4 else 0
eng = GPEngine(0)
if (n, d) == (bench.N_OBS, bench.DIM):
    X, y, ell = bench.synthetic()
else:
    from oracle import gp_oracle as o
    X, y, ell = o.synthetic_problem(n, d)
grid = CandidateGrid([np.linspace(0, 1, 10 if d < 10 else 8)] * d)
count = min(count, grid.size - begin)
fit = eng.fit(X, y, ell, JITTER_POSTERIOR)
fb = float(y.min())
eng.acquire(fit, grid, 0, min(count, 1 << 22), kind=ACQ_EI, f_best=fb)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    kind = 0 if len(sys.argv) > 5 and sys.argv[5] == "lcb" else ACQ_EI
    r = eng.acquire(fit, grid, begin, begin + count, kind=kind, f_best=fb)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "sweep_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
span = max(e["ts"] + e["dur"] for e in ev) - ev[0]["ts"]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    k = e["name"].replace("bogp::", "").split("(")[0].split("<")[0][:48]
    agg[k][0] += 1; agg[k][1] += e["dur"]
print(f"sweep of {count} candidates: span {span / 1e3:.2f} ms, {len(ev)} device activities, best {r.best_index}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:48s} n={v[0]:5d} total {v[1] / 1e3:8.2f} ms  avg {v[1] / v[0]:8.1f} us")
