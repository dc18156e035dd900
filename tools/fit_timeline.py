"""Kernel timeline of one fit (concurrency preserved) through CUPTI, via torch.profiler -- the stand-in for nsys.
Prints every kernel of the last fit with its start offset, duration and stream, plus the busy/idle split of the
critical (high-priority) stream."""
import sys, json, tempfile, os
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
from oracle import gp_oracle as o
from bayesian_optimisation_b200.engine import GPEngine, JITTER_POSTERIOR
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = GPEngine(0)
X, y, ell = o.synthetic_problem(n, d)
dX, dy = eng.to_device(X), eng.to_device(y)
for _ in range(3):
    eng.fit(dX, dy, ell, JITTER_POSTERIOR).close()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    f = eng.fit(dX, dy, ell, JITTER_POSTERIOR)
    torch.cuda.synchronize()
f.close()
path = os.path.join(tempfile.gettempdir(), "fit_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
end = max(e["ts"] + e["dur"] for e in ev)
print(f"fit span {(end - t0) / 1e3:.3f} ms, {len(ev)} device activities")
streams = sorted({e["args"].get("stream") for e in ev})
for e in ev:
    name = e["name"].replace("bogp::", "").split("(")[0][:44]
    print(f"{(e['ts'] - t0):9.1f} us  +{e['dur']:8.1f} us  s{streams.index(e['args'].get('stream'))}  {name}  grid={e['args'].get('grid')}")
for s in streams:
    se = [e for e in ev if e["args"].get("stream") == s]
    busy = sum(e["dur"] for e in se)
    print(f"stream s{streams.index(s)}: {len(se)} activities, busy {busy / 1e3:.3f} ms")
